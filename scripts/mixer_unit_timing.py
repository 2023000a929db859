"""Per-unit clock64 timeline of CTA 0 for the MLP-Mixer channel GEMMs (dim -> 4 dim with GELU, 4 dim -> dim) at the
decoder mixer's shape (B = 512, 1856 tokens, dim 344): what bounds the low-K GEMMs (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing


def run(precision, c_in, c_out, act, B=512, T=1856, out_f32=False):
    torch.manual_seed(0)
    w = torch.randn(c_out, c_in) / c_in ** 0.5
    layer = ops.ConvGemm(*packing.pack_linear(w, torch.randn(c_out), precision), act=act).to("cuda")
    x = packing.to_act(torch.randn(B, T, c_in), precision).cuda()
    out = ops.alloc_act(B, T, c_out, precision, "cuda") if not out_f32 else None
    out2 = torch.empty(B * T, c_out, device="cuda") if out_f32 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        layer(x, B, T, out=out, out2=out2)
    e0.record(); layer(x, B, T, out=out, out2=out2); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    dbg = torch.zeros(4 * 148 + 8 * 256, dtype=torch.int64, device="cuda")
    layer.debug_clk = dbg
    layer(x, B, T, out=out, out2=out2)
    torch.cuda.synchronize()
    u = dbg[4 * 148:].view(256, 8).double().cpu()[20:100]
    nxt = dbg[4 * 148:].view(256, 8).double().cpu()[21:101]
    f = lambda a: f"{a.mean():.0f}"
    print(f"{precision} {c_in}->{c_out} {act} ({'fp32 out' if out_f32 else 'operand out'}): {ms:.2f} ms = "
          f"{2.0 * c_in * c_out * B * T / ms / 1e9:.0f} TFLOP/s; per unit (cycles): period {f(nxt[:, 7] - u[:, 7])}; "
          f"MMA issue span {f(u[:, 5] - u[:, 4])}, wait for first k-block {f(u[:, 4] - u[:, 3])}, wait for free accumulator "
          f"{f(u[:, 3] - u[:, 2])}; epilogue {f(u[:, 7] - u[:, 6])}; epilogue idle before next accumulator {f(nxt[:, 6] - u[:, 7])}")


if __name__ == "__main__":
    for prec in ("fp16x2", "fp32"):
        run(prec, 344, 1376, "gelu")
        run(prec, 344, 1376, "none")
        run(prec, 1376, 344, "none", out_f32=True)
