"""CPU study (no GPU): waveform error of the MelGAN generator for candidate storage / operand schemes, against an fp64
run.  Accumulation is fp64; operands are rounded exactly where a kernel would round them, per TENSOR ROLE:

  x    the residual stream as stored in HBM between kernels (ConvTranspose output, ResnetBlock output)
  xa   LeakyReLU(x) as the k3 / ConvTranspose MMA operand (re-rounded when it is formed in-kernel from the stored x)
  mid  the ResnetBlock intermediate LeakyReLU(conv3(xa) + b) (on-chip in the fused kernel)
  w    weights (always two terms: exact to ~2^-22)

Formats: "x"  exact (fp64), "s" split bf16 hi+lo (~2^-17), "h" one fp16 value (2^-12), "H" two fp16 terms (~2^-22)
A scheme is four letters per stage group, e.g. all-split = x:s xa:s mid:s.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import rel_l2, templates
from oracle.melgan import wn_weight, RATIOS
from oracle.seeded import seeded_state_dict, synthetic_mel


def rnd(t, f):
    if f == "x":
        return t
    if f == "s":
        hi = t.float().bfloat16().float()
        lo = (t.float() - hi).bfloat16().float()
        return hi.double() + lo.double()
    if f == "h":
        return t.float().half().double()
    if f == "H":
        hi = t.float().half().float()
        lo = (t.float() - hi).half().float()
        return hi.double() + lo.double()
    raise ValueError(f)


def lrelu(x):
    return F.leaky_relu(x, 0.2)


def run(sd, mel, scheme):
    """scheme: dict stage -> dict(x=, xa=, mid=) with stage in 0..3 plus 'stem'."""
    W = lambda p: wn_weight(sd, p)
    b = lambda p: sd[p + ".bias"]
    x = F.pad(mel, (3, 3), mode="reflect")
    x = F.conv1d(rnd(x, scheme["stem"]["x"]), W("model.1"), b("model.1"))
    idx = 2
    for si, r in enumerate(RATIOS):
        s = scheme[si]
        prev = scheme["stem"] if si == 0 else scheme[si - 1]
        a = rnd(lrelu(x), prev.get("up", prev["xa"]))     # operand of the ConvTranspose (stored by the producer)
        x = F.conv_transpose1d(a, W(f"model.{idx + 1}"), b(f"model.{idx + 1}"), stride=r, padding=r // 2 + r % 2,
                               output_padding=r % 2)
        for j in range(3):
            p = f"model.{idx + 2 + j}"
            d = 3 ** j
            xs = rnd(x, s["x"])                            # the stored residual stream (HBM)
            xa = rnd(lrelu(xs), s["xa"])                   # LeakyReLU formed from the STORED value, then operand format
            y = F.conv1d(F.pad(xa, (d, d), mode="reflect"), W(p + ".block.2"), b(p + ".block.2"), dilation=d)
            mid = rnd(lrelu(y), s["mid"])
            x = F.conv1d(mid, W(p + ".block.4"), b(p + ".block.4")) + F.conv1d(xs, W(p + ".shortcut"), b(p + ".shortcut"))
        idx += 5
    last = scheme[3]
    a = rnd(lrelu(rnd(x, last["x"])), last.get("out", "x"))
    a = F.pad(a, (3, 3), mode="reflect")
    return torch.tanh(F.conv1d(a, W("model.24"), b("model.24")))


def scheme(default, **over):
    sc = {k: dict(default) for k in ("stem", 0, 1, 2, 3)}
    for k, v in over.items():
        key = k if k == "stem" else int(k[1:])
        sc[key] = dict(sc[key], **v)
    return sc


SCHEMES = {
    "all split (r01 default)": scheme(dict(x="s", xa="s", mid="s")),
    "all fp16 1-term (r01 fp16x2)": scheme(dict(x="h", xa="h", mid="h")),
    "x fp16, xa fp16, mid 2-term": scheme(dict(x="h", xa="h", mid="H")),
    "x fp16, xa 2-term, mid 2-term": scheme(dict(x="h", xa="H", mid="H")),
    "stages 0-1 fp16 (mid 2-term), 2-3 split": scheme(dict(x="s", xa="s", mid="s"), stem=dict(x="h", xa="h"),
                                                       s0=dict(x="h", xa="h", mid="H"), s1=dict(x="h", xa="h", mid="H")),
    "stages 0-1 all fp16 1-term, 2-3 split": scheme(dict(x="s", xa="s", mid="s"), stem=dict(x="h", xa="h"),
                                                     s0=dict(x="h", xa="h", mid="h"), s1=dict(x="h", xa="h", mid="h")),
    "stages 0-2 fp16 (mid 2-term), 3 split": scheme(dict(x="h", xa="h", mid="H"), s3=dict(x="s", xa="s", mid="s")),
    "x split, xa fp16, mid fp16": scheme(dict(x="s", xa="h", mid="h")),
    "x split, xa fp16, mid 2-term": scheme(dict(x="s", xa="h", mid="H")),
    # the round-2 MelGAN format: residual stream as two fp16 terms in HBM, LeakyReLU formed in-kernel as ONE fp16 value
    "x 2xfp16, xa fp16, mid fp16": scheme(dict(x="H", xa="h", mid="h")),
    "x 2xfp16, xa fp16, mid 2-term": scheme(dict(x="H", xa="h", mid="H")),
    "x 2xfp16, xa fp16, mid 2-term, up 2-term": scheme(dict(x="H", xa="h", mid="H", up="H")),
    "x 2xfp16, xa fp16, mid fp16, up 2-term": scheme(dict(x="H", xa="h", mid="h", up="H")),
    "same, stage 0 xa 2-term": scheme(dict(x="H", xa="h", mid="H", up="H"), s0=dict(xa="H")),
}


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    B, T = 2, int(os.environ.get("T", 48))
    rows = {}
    seeds = [int(s) for s in os.environ.get("SEEDS", "4,5,6,7").split(",")]
    for seed in seeds:
        sd = {k: v.double() for k, v in seeded_state_dict(templates.melgan_template(), seed).items()}
        mel = synthetic_mel(B, T, 100 + seed).transpose(1, 2).contiguous().double()
        exact = run(sd, mel, scheme(dict(x="x", xa="x", mid="x")))
        for name, sc in SCHEMES.items():
            rows.setdefault(name, []).append(rel_l2(run(sd, mel, sc), exact))
    print(f"MelGAN waveform rel-L2 vs fp64, B={B} T={T}, weight seeds {seeds}")
    for name, v in rows.items():
        print(f"  {name:<46} " + "  ".join(f"{e:.2e}" for e in v) + f"   max {max(v):.2e}")
