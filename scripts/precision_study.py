"""CPU study (no GPU): end-to-end error of AutoVC for candidate tensor-core operand schemes, against an fp64 run of
the oracle.  Activations / weights are rounded at every matmul input exactly where a kernel would round them
(conv inputs, LSTM x_t and h_{t-1}, Linear input); accumulation stays fp64, so this isolates operand rounding.
  split   : both operands as bf16 hi + lo, products hi*hi + lo*hi + hi*lo        (3 MMA passes; the current default)
  tf32    : both operands rounded to TF32                                         (1 pass at half rate = 2 units)
  f16a    : activations rounded to fp16, weights exact (fp16 hi + lo)             (2 passes)
  bf16a   : activations rounded to bf16, weights exact                            (2 passes)
  bf16    : both operands bf16                                                    (1 pass)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import rel_l2, templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker


def tf32(x):
    i = x.float().view(torch.int32)
    return (((i + 0x1000) & ~0x1FFF).view(torch.float32)).double()


def split_bf16(x):
    hi = x.float().bfloat16().float()
    lo = (x.float() - hi).bfloat16().float()
    return hi.double(), lo.double()


class Scheme:
    def __init__(self, name):
        self.name = name

    def act(self, x):
        n = self.name
        if n in ("f16a",):
            return x.float().half().double()
        if n in ("bf16a", "bf16"):
            return x.float().bfloat16().double()
        if n == "tf32":
            return tf32(x)
        if n == "split":
            hi, lo = split_bf16(x)
            return hi + lo
        return x

    def wgt(self, w):
        n = self.name
        if n == "bf16":
            return w.float().bfloat16().double()
        if n == "tf32":
            return tf32(w)
        if n == "split":
            hi, lo = split_bf16(w)
            return hi + lo              # (the dropped lo*lo term is ~2^-18 relative: below this study's resolution)
        if n == "f16a":
            hi = w.float().half().float()
            lo = (w.float() - hi).half().float()
            return hi.double() + lo.double()
        return w


def conv_bn(sd, prefix, x, act, s):
    mean, var = sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"]
    g, b = sd[prefix + ".1.weight"], sd[prefix + ".1.bias"]
    scale = g / torch.sqrt(var + 1e-5)
    w = sd[prefix + ".0.conv.weight"] * scale.view(-1, 1, 1)           # BN folded like the kernels do
    bias = (sd[prefix + ".0.conv.bias"] - mean) * scale + b
    y = F.conv1d(s.act(x), s.wgt(w), bias, padding=2)
    return torch.relu(y) if act == "relu" else torch.tanh(y) if act == "tanh" else y


def lstm(sd, prefix, x, layers, s, bidir=False):
    for layer in range(layers):
        outs = []
        for d in (["", "_reverse"] if bidir else [""]):
            sfx = f"_l{layer}{d}"
            w_ih, w_hh = s.wgt(sd[f"{prefix}.weight_ih{sfx}"]), s.wgt(sd[f"{prefix}.weight_hh{sfx}"])
            bias = sd[f"{prefix}.bias_ih{sfx}"] + sd[f"{prefix}.bias_hh{sfx}"]
            B, T, _ = x.shape
            H = w_hh.shape[1]
            h, c = x.new_zeros(B, H), x.new_zeros(B, H)
            out = x.new_empty(B, T, H)
            xp = s.act(x) @ w_ih.t() + bias
            for t in (range(T - 1, -1, -1) if d else range(T)):
                z = xp[:, t] + s.act(h) @ w_hh.t()
                zi, zf, zg, zo = z.split(H, dim=1)
                c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
                h = torch.sigmoid(zo) * torch.tanh(c)
                out[:, t] = h
            outs.append(out)
        x = torch.cat(outs, dim=-1)
    return x


def forward(sd, x, c_org, c_trg, dim_neck, freq, s):
    B, T, _ = x.shape
    h = torch.cat((x.transpose(2, 1), c_org.unsqueeze(-1).expand(-1, -1, T)), dim=1)
    for i in range(3):
        h = conv_bn(sd, f"encoder.convolutions.{i}", h, "relu", s)
    out = lstm(sd, "encoder.lstm", h.transpose(1, 2), 2, s, bidir=True)
    codes = torch.cat((out[:, freq - 1::freq, :dim_neck], out[:, ::freq, dim_neck:]), dim=-1)
    dec_in = torch.cat((codes.repeat_interleave(freq, dim=1), c_trg.unsqueeze(1).expand(-1, T, -1)), dim=-1)
    h = lstm(sd, "decoder.lstm1", dec_in, 1, s)
    h = h.transpose(1, 2)
    for i in range(3):
        h = conv_bn(sd, f"decoder.convolutions.{i}", h, "relu", s)
    h = lstm(sd, "decoder.lstm2", h.transpose(1, 2), 2, s)
    mel = s.act(h) @ s.wgt(sd["decoder.linear_projection.linear_layer.weight"]).t() + sd["decoder.linear_projection.linear_layer.bias"]
    p = mel.transpose(2, 1)
    for i in range(4):
        p = conv_bn(sd, f"postnet.convolutions.{i}", p, "tanh", s)
    p = conv_bn(sd, "postnet.convolutions.4", p, "none", s)
    return mel, mel + p.transpose(2, 1), codes.reshape(B, -1)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    args = (32, 256, 512, 32)
    for wseed, B, T in ((0, 4, 128), (11, 2, 256)):
        sd = {k: v.double() for k, v in seeded_state_dict(templates.autovc_template(*args), wseed).items()}
        x, co, ct = synthetic_mel(B, T, 21).double(), synthetic_speaker(B, 21, "org").double(), synthetic_speaker(B, 21, "trg").double()
        ref = forward(sd, x, co, ct, 32, 32, Scheme("exact"))
        for name in ("split", "tf32", "f16a", "bf16a", "bf16"):
            out = forward(sd, x, co, ct, 32, 32, Scheme(name))
            print(f"weights seed {wseed} B={B} T={T} {name:6s}: rel-L2 mel {rel_l2(out[0], ref[0]):.2e}  mel_postnet "
                  f"{rel_l2(out[1], ref[1]):.2e}  codes {rel_l2(out[2], ref[2]):.2e}")
