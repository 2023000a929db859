"""CPU study (no GPU): what ONE fp16 term per recurrent / inter-layer weight costs the wavefront stack kernel
(avc_lstm_stack_ws) on LstmDV, against keeping two terms (the layer-by-layer kernels), and which layers matter.
fp64 accumulation; operands rounded exactly where the kernels round them: h to fp16 every frame, the layer-0 dense input
projection with two-term weights in both cases, the chosen layers' W_hh (and W_ih above the first layer) to one fp16 term.
The reference is the same recurrence without any rounding.

    python scripts/lstm_stack_precision.py > profiles/r02_lstm_stack_precision.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel

torch.set_grad_enabled(False)


def f16(t):
    return t.half().double()


def run(sd, x, single_layers, h_round=True):
    B, T, _ = x.shape
    inp = f16(x.double()) if h_round else x.double()
    for l in range(3):
        wih, whh = sd[f"lstm.weight_ih_l{l}"].double(), sd[f"lstm.weight_hh_l{l}"].double()
        b = (sd[f"lstm.bias_ih_l{l}"] + sd[f"lstm.bias_hh_l{l}"]).double()
        if l in single_layers:
            whh = f16(whh)
            if l > 0:
                wih = f16(wih)
        H = whh.shape[1]
        xp = inp @ wih.t() + b
        h, c = torch.zeros(B, H, dtype=torch.float64), torch.zeros(B, H, dtype=torch.float64)
        outs = []
        for t in range(T):
            z = xp[:, t] + h @ whh.t()
            i, f, g, o = z.chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            hf = torch.sigmoid(o) * torch.tanh(c)
            h = f16(hf) if h_round else hf
            outs.append(h)
        inp = torch.stack(outs, 1)
        last = hf
    e = last @ sd["embedding.weight"].double().t() + sd["embedding.bias"].double()
    return e / e.norm(dim=-1, keepdim=True)


if __name__ == "__main__":
    torch.set_num_threads(16)
    print("# rel-L2 of the LstmDV embedding (B = 4) against the un-rounded recurrence; 'one term' = which layers keep ONE fp16 term per weight")
    print(f"{'weights seed':>12s} {'lstm gain':>9s} {'T':>5s} {'two terms':>10s} {'one term: all (kernel)':>22s} {'layers 1,2':>11s} {'layer 2':>9s} {'layer 1':>9s}")
    for seed, gain, T in ((3, 1.5, 256), (3, 1.5, 1000), (5, 1.5, 256), (7, 1.0, 256), (1, 3.0, 256), (1, 3.0, 1000), (2, 3.0, 256)):
        sd = seeded_state_dict(templates.lstmdv_template(), seed, lstm_gain=gain)
        x = synthetic_mel(4, T, seed)
        ref = run(sd, x, (), h_round=False)
        err = [float((run(sd, x, sl) - ref).norm() / ref.norm()) for sl in ((), (0, 1, 2), (1, 2), (2,), (1,))]
        print(f"{seed:12d} {gain:9.1f} {T:5d} {err[0]:10.2e} {err[1]:22.2e} {err[2]:11.2e} {err[3]:9.2e} {err[4]:9.2e}")
    print("# lstm gain 1.5 is the bench's / the tests' LstmDV initialisation; 3.0 is the stress initialisation of the golden fixture;")
    print("# the gate is 1e-3.  The weight rounding is the same every frame, so it accumulates coherently through the recurrence,")
    print("# unlike the activations' rounding: one term per weight costs about as much again as all of the activation rounding.")
