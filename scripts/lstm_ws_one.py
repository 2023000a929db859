"""Launches of the weight-stationary small-batch LSTM recurrence at the decoder's lstm2 shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

B, T, H = 32, 256, 1024
torch.manual_seed(0)
hh = packing.pack_lstm_hh(torch.randn(4 * H, H) * 0.03, "fp32", packing.WS_GROUP).cuda()
xp = (torch.randn(B * T, 4 * H) * 0.5).cuda()
for _ in range(3):
    assert ops.lstm_seq_ws(xp, hh, B, T, H) is not None
torch.cuda.synchronize()
print("ok")
