"""In-kernel clock64 stamps of the wide fused ResnetBlock (C = 128) on one GPU: where a tile's time goes."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops, packing

C, d = 128, int(os.environ.get("D", 3))
B, L = int(os.environ.get("AVC_B", 32)), 64000
torch.manual_seed(0)
w3, w1, wsc = torch.randn(C, C, 3) / (3 * C) ** 0.5, torch.randn(C, C, 1) / C ** 0.5, torch.randn(C, C, 1) / C ** 0.5
blk = ops.Resblock2(*packing.pack_resblock2(w3, torch.randn(C), w1, torch.randn(C), wsc, torch.randn(C)), dilation=d).to("cuda")
x = (torch.randn(B, L + 2 * d, 2 * C, device="cuda") * 0.5).half()
y = torch.empty(B, L + 2 * 9, 2 * C, dtype=torch.float16, device="cuda")
for _ in range(2):
    blk(x, B, L, y=y, y_row0=9, y_reflect=9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    blk(x, B, L, y=y, y_row0=9, y_reflect=9)
e1.record()
torch.cuda.synchronize()
print(f"C={C} d={d} B={B} L={L}: {e0.elapsed_time(e1) / 5:.3f} ms per block")
clk = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
blk.debug_clk = clk
blk(x, B, L, y=y, y_row0=9, y_reflect=9)
torch.cuda.synchronize()
c = clk.view(64, 16).cpu()
names = ["win_req", "w_req_done", "xa_ready", "gemm1_issued", "mid_ready", "gemm2_issued", "win_landed", "-",
         "xa_written", "d1_full", "mid_written", "d2_full", "staged_hi", "staged_lo"]
base = int(c[2, 0])
for it in (2, 3, 4, 20):
    row = c[it]
    t0 = int(row[0])
    print(f"tile {it}: start +{t0 - base:7d} | " + "  ".join(f"{n}={int(row[i]) - t0}" for i, n in enumerate(names)))
print("tile period (cycles):", [int(c[i + 1, 0] - c[i, 0]) for i in range(2, 12)])
