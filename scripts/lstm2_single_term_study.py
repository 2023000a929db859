"""CPU study (no GPU): could AutoVC's decoder LSTM (lstm2, two layers of 1024) run on the wavefront stack kernel, whose
weights are ONE fp16 term?  scripts/precision_study.py's fp16x2 emulation ("f16a": fp16 activations, two-term weights,
fp64 accumulation) with lstm2's W_hh (both layers) and layer 1's W_ih rounded to one fp16 term.

    python scripts/lstm2_single_term_study.py > profiles/r02_lstm2_single_term_study.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import precision_study as ps
from oracle import rel_l2, templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker

torch.set_grad_enabled(False)
torch.set_num_threads(16)


class OneTermLstm2(ps.Scheme):
    def __init__(self):
        super().__init__("f16a")
        self.single = False

    def wgt(self, w):
        return w.float().half().double() if self.single else super().wgt(w)


_lstm = ps.lstm


def lstm(sd, prefix, x, layers, s, bidir=False):
    if prefix != "decoder.lstm2" or not isinstance(s, OneTermLstm2):
        return _lstm(sd, prefix, x, layers, s, bidir)
    for layer in range(layers):
        sfx = f"_l{layer}"
        s.single = layer > 0                       # layer 0's input projection stays a dense two-term GEMM
        w_ih = s.wgt(sd[f"{prefix}.weight_ih{sfx}"])
        s.single = True
        w_hh = s.wgt(sd[f"{prefix}.weight_hh{sfx}"])
        s.single = False
        bias = sd[f"{prefix}.bias_ih{sfx}"] + sd[f"{prefix}.bias_hh{sfx}"]
        B, T, _ = x.shape
        H = w_hh.shape[1]
        h, c = x.new_zeros(B, H), x.new_zeros(B, H)
        out = x.new_empty(B, T, H)
        xp = s.act(x) @ w_ih.t() + bias
        for t in range(T):
            z = xp[:, t] + s.act(h) @ w_hh.t()
            zi, zf, zg, zo = z.split(H, dim=1)
            c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
            h = torch.sigmoid(zo) * torch.tanh(c)
            out[:, t] = h
        x = out
    return x


ps.lstm = lstm
args = (32, 256, 512, 32)
print("# AutoVC(32,256,512,32), rel-L2 against the fp64 run; gate 1e-3")
for wseed, B, T in ((0, 4, 128), (11, 2, 256), (5, 2, 256)):
    sd = {k: v.double() for k, v in seeded_state_dict(templates.autovc_template(*args), wseed).items()}
    x, co, ct = synthetic_mel(B, T, 21).double(), synthetic_speaker(B, 21, "org").double(), synthetic_speaker(B, 21, "trg").double()
    ref = ps.forward(sd, x, co, ct, 32, 32, ps.Scheme("exact"))
    for name, s in (("fp16x2 (two-term weights)", ps.Scheme("f16a")), ("fp16x2 + one-term lstm2", OneTermLstm2())):
        out = ps.forward(sd, x, co, ct, 32, 32, s)
        print(f"weights seed {wseed:2d} B={B} T={T} {name:28s}: mel {rel_l2(out[0], ref[0]):.2e}  mel_postnet {rel_l2(out[1], ref[1]):.2e}  "
              f"codes {rel_l2(out[2], ref[2]):.2e}")
