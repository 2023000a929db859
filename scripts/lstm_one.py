import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing
B, T, H = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1024
persistent = (sys.argv[2] != "step") if len(sys.argv) > 2 else True
G = ops.choose_gate_group(B, H, True)
hh = packing.pack_lstm_hh(torch.randn(4 * H, H) * 0.03, "fp32", G).cuda()
xp = torch.randn(B * T, 4 * H, device="cuda")
for _ in range(2):
    ops.lstm_seq(xp, hh, B, T, H, "fp32", G, persistent=persistent)
torch.cuda.synchronize()
print("ok", B, T, H, G, persistent)
