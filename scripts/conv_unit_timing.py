"""Per-unit clock64 timeline of CTA 0 of one narrow convolution (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing


def run(precision, c, k, B=8, T=256000, dil=1):
    torch.manual_seed(0)
    w = torch.randn(c, c, k) / (c * k) ** 0.5
    layer = ops.ConvGemm(*packing.pack_conv(w, torch.randn(c), precision), act="lrelu", tap_t0=[-dil * (k // 2)],
                         tap_dt=[dil]).to("cuda")
    x = packing.to_act(torch.randn(B, T, c), precision).cuda()
    out = ops.alloc_act(B, T, c, precision, "cuda")
    for _ in range(2):
        layer(x, B, T, out=out)
    dbg = torch.zeros(4 * 148 + 8 * 256, dtype=torch.int64, device="cuda")
    layer.debug_clk = dbg
    layer(x, B, T, out=out)
    torch.cuda.synchronize()
    u = dbg[4 * 148:].view(256, 8).double().cpu()[20:100]
    nxt = dbg[4 * 148:].view(256, 8).double().cpu()[21:101]
    f = lambda a: f"{a.mean():.0f}"
    print(f"{precision} C={c} k{k}: per unit (cycles): unit period {f(nxt[:, 7] - u[:, 7])}; producer issue span {f(u[:, 1] - u[:, 0])}, "
          f"producer unit period {f(nxt[:, 0] - u[:, 0])}; MMA: wait for free accumulator {f(u[:, 3] - u[:, 2])}, "
          f"wait for first k-block {f(u[:, 4] - u[:, 3])}, issue span {f(u[:, 5] - u[:, 4])}; "
          f"MMAs issued -> accumulator ready {f(u[:, 6] - u[:, 5])}; epilogue {f(u[:, 7] - u[:, 6])}; "
          f"epilogue idle before next accumulator {f(nxt[:, 6] - u[:, 7])}; loads issued -> first k-block landed (same unit) {f(u[:, 4] - u[:, 0])}")


run("fp32", 32, 3, dil=3)
run("fp32", 32, 1)
run("fp32", 64, 3, dil=3)
run("fp32", 128, 3, dil=3)
run("bf16", 32, 3, dil=3)
