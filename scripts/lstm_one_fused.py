"""Two launches of the fused persistent LSTM kernel at the decoder's lstm2-layer-1 shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

B, T, H, I = 512, 128, 1024, 1024
torch.manual_seed(0)
G = ops.choose_gate_group(B, H, True)
hh = packing.pack_lstm_hh(torch.randn(4 * H, H) * 0.03, "fp32", G).cuda()
wih, bias = packing.pack_lstm_ih_fused(torch.randn(4 * H, I) * 0.03, torch.zeros(4 * H), torch.zeros(4 * H), "fp32", G)
x = packing.to_act(torch.randn(B, T, I), "fp32").cuda()
for _ in range(2):
    ops.lstm_seq(None, hh, B, T, H, "fp32", G, xin=x, w_ih=wih.cuda(), bias=bias.cuda(), c_in=I, persistent=True)
torch.cuda.synchronize()
print("ok")
