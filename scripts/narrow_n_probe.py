"""The two N = 80 layers of the headline step (postnet 512 -> 80 k5 with the residual, Linear 1024 -> 80 with two outputs) at
B = 512, T = 128, fp16x2: CUDA-event time and the per-unit clock64 timeline of CTA 0 (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

B, T, P = 512, 128, "fp16x2"


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def units(layer, fn):
    dbg = torch.zeros(4 * 148 + 8 * 256, dtype=torch.int64, device="cuda")
    layer.debug_clk = dbg
    fn()
    torch.cuda.synchronize()
    layer.debug_clk = None
    u = dbg[4 * 148:].view(256, 8).cpu()
    n = int((u[:, 7] > 0).sum())
    base = int(u[0, 0])
    rows = []
    for i in range(n):
        r = u[i]
        rows.append(f"unit {i}: producer start +{int(r[0]) - base}, first k-block +{int(r[4]) - base}, MMAs issued +{int(r[5]) - base}, "
                    f"accumulator ready +{int(r[6]) - base}, epilogue done +{int(r[7]) - base}")
    return rows


torch.manual_seed(0)
c_in, c_out, k = 512, 80, 5
conv = ops.ConvGemm(*packing.pack_conv(torch.randn(c_out, c_in, k) / (c_in * k) ** 0.5, torch.randn(c_out), P)).to("cuda")
x = packing.to_act(torch.randn(B, T, c_in), P).cuda()
mel = torch.randn(B * T, c_out, device="cuda")
post = torch.empty(B * T, c_out, device="cuda")
cases = {"postnet 512->80 k5, out2 + residual after (as in the model)": lambda: conv(x, B, T, out2=post, residual=mel, res_after=True),
         "postnet 512->80 k5, out2 only": lambda: conv(x, B, T, out2=post)}
for name, fn in cases.items():
    print(f"{name}: {timed(fn):.1f} us")
    for r in units(conv, fn):
        print("   ", r)
lin = ops.ConvGemm(*packing.pack_linear(torch.randn(80, 1024) / 32, torch.randn(80), P)).to("cuda")
h = packing.to_act(torch.randn(B, T, 1024), P).cuda()
mel_op = ops.alloc_act(B, T, 80, P, "cuda")
cases = {"Linear 1024->80, out + out2 (as in the model)": lambda: lin(h, B, T, out=mel_op, out2=mel),
         "Linear 1024->80, out only": lambda: lin(h, B, T, out=mel_op)}
for name, fn in cases.items():
    print(f"{name}: {timed(fn):.1f} us")
    for r in units(lin, fn):
        print("   ", r)
