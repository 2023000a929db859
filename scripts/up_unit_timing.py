"""Per-unit clock64 timeline of CTA 0 of the MelGAN up-sampling GEMMs in the "fp16s" precision (profiling aid)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops, packing


def run(c_in, c_out, r, B, L):
    torch.manual_seed(0)
    wt = torch.randn(c_in, c_out, 2 * r) / (c_in * 2) ** 0.5
    w3 = packing.conv_transpose_as_conv(wt, r, r // 2 + r % 2)
    layer = ops.ConvGemm(*packing.pack_conv(w3, torch.randn(c_out).repeat(r), "fp16s"), tap_t0=[-1], act="none").to("cuda")
    x = packing.to_act(torch.randn(B, L, c_in), "fp16s").cuda()
    out = ops.alloc_act(B, r * L + 2, c_out, "fp16s", "cuda")
    call = lambda: layer(x, B, L, out=out, out_row0=1, phases=r)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    dbg = torch.zeros(4 * 148 + 8 * 256, dtype=torch.int64, device="cuda")
    layer.debug_clk = dbg
    call()
    torch.cuda.synchronize()
    u = dbg[4 * 148:].view(256, 8).double().cpu()[20:100]
    nxt = dbg[4 * 148:].view(256, 8).double().cpu()[21:101]
    f = lambda a: f"{a.mean():.0f}"
    print(f"up {c_in}->{c_out} r={r} B={B} L={L}: {ms:.3f} ms, N={layer.meta['N']} block_n={layer.meta['block_n']} k-blocks={layer.meta['k_pad'] // 64}; "
          f"per unit (cycles): period {f(nxt[:, 7] - u[:, 7])}; producer issue span {f(u[:, 1] - u[:, 0])}, producer period {f(nxt[:, 0] - u[:, 0])}; "
          f"MMA: wait free acc {f(u[:, 3] - u[:, 2])}, wait first k-block {f(u[:, 4] - u[:, 3])}, issue span {f(u[:, 5] - u[:, 4])}; "
          f"issued -> acc ready {f(u[:, 6] - u[:, 5])}; epilogue {f(u[:, 7] - u[:, 6])}; epilogue idle {f(nxt[:, 6] - u[:, 7])}")


run(64, 32, 2, 32, 128000)
run(128, 64, 2, 32, 64000)
run(256, 128, 8, 32, 8000)
