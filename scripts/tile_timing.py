"""Per-CTA clock64 breakdown of one conv_gemm launch: setup / mainloop / epilogue cycles (profiling aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

def run(precision, c_in, c_out, k, B=512, T=128, cta=2, out2=False):
    torch.manual_seed(0)
    w = torch.randn(c_out, c_in, k) / (c_in * k) ** 0.5
    layer = ops.ConvGemm(*packing.pack_conv(w, torch.randn(c_out), precision), act="relu").to("cuda")
    layer.cta_group = cta
    x = packing.to_act(torch.randn(B, T, c_in), precision).cuda()
    out = ops.alloc_act(B, T, c_out, precision, "cuda") if not out2 else None
    o2 = torch.empty(B * T, c_out, device="cuda") if out2 else None
    for _ in range(2):
        layer(x, B, T, out=out, out2=o2)
    n_cta = 8192
    dbg = torch.zeros(n_cta * 4, dtype=torch.int64, device="cuda")
    layer.debug_clk = dbg
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); layer(x, B, T, out=out, out2=o2); e1.record()
    torch.cuda.synchronize()
    d = dbg.view(-1, 4).cpu()
    d = d[d[:, 0] > 0].double()
    setup, main, epi = (d[:, 1] - d[:, 0]), (d[:, 2] - d[:, 1]), (d[:, 3] - d[:, 2])
    span = (d[:, 3].max() - d[:, 0].min())
    print(f"{precision} {c_in}->{c_out} k{k} cta_group={cta} out2={out2}: {e0.elapsed_time(e1)*1e3:.0f} us, CTAs {len(d)}, kblocks {layer.meta['k_pad']//packing.KC[precision]}, "
          f"cycles/CTA: setup {setup.mean():.0f}, mainloop(wait for acc) {main.mean():.0f}, epilogue {epi.mean():.0f}; whole-kernel span {span:.0f} cyc")

if len(sys.argv) > 1 and sys.argv[1] == "one":
    run("bf16", 512, 512, 5, cta=2)
    sys.exit(0)
for prec in ("bf16", "fp32"):
    for cta in (1, 2):
        run(prec, 512, 512, 5, cta=cta)
run("fp32", 512, 4096, 1, out2=True)
run("bf16", 512, 4096, 1, out2=True)
