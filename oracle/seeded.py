"""Seeded weights and synthetic inputs shared by the oracle, the tests and the bench.

Test infrastructure (see ``oracle/__init__.py``).  A 28 M-parameter state_dict is
too large to commit, so golden fixtures store only the seed; both the reference
(in ``make_golden.py``) and the CUDA path (in ``tests/``) regenerate identical
weights from ``seeded_state_dict``.  The generator is keyed per tensor name
(crc32) so it does not depend on iteration order.

The init is "signal preserving" (protocol item 2): BatchNorm
running statistics and affine terms are non-trivial so that folding them is
actually exercised, and LSTM matrices are scaled x3 so the output depends on the
input strongly enough for a parity number to mean something.
"""
import math
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _normal(shape, mean, std, g):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def seeded_state_dict(template, seed: int = 0, lstm_gain: float = 3.0):
    """Return {name: fp32 tensor} for every entry of ``template`` ({name: tensor}).

    Rules are decided from the key name and rank only:
      *.running_mean ~ N(0, .5)   *.running_var ~ U(.5, 2)
      BatchNorm/GroupNorm/LayerNorm weight ~ U(.5, 1.5), bias ~ N(0, .2)
      LSTM weight_* ~ U(-k, k) * lstm_gain, bias_* ~ U(-k, k), k = 1/sqrt(H)
      weight_g (MelGAN weight norm) = ||v|| * U(.8, 1.25)  (about unit gain)
      conv / linear weights: xavier-uniform bound sqrt(6 / (fan_in + fan_out))
      conv / linear biases ~ N(0, .05)
    """
    out = {}
    keys = list(template.keys())
    for key in keys:
        ref = template[key]
        shape = tuple(ref.shape)
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=ref.dtype)
            continue
        if leaf == "running_mean":
            out[key] = _normal(shape, 0.0, 0.5, g)
        elif leaf == "running_var":
            out[key] = _uniform(shape, 0.5, 2.0, g)
        elif leaf.startswith("weight_ih") or leaf.startswith("weight_hh"):
            hidden = shape[0] // 4
            k = 1.0 / math.sqrt(hidden)
            out[key] = _uniform(shape, -k, k, g) * lstm_gain
        elif leaf.startswith("bias_ih") or leaf.startswith("bias_hh"):
            hidden = shape[0] // 4
            k = 1.0 / math.sqrt(hidden)
            out[key] = _uniform(shape, -k, k, g)
        elif leaf == "weight_v":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = math.sqrt(3.0 / fan_in) * 0.95
            out[key] = _uniform(shape, -bound, bound, g)
        elif leaf == "weight_g":
            out[key] = None  # filled below from the matching weight_v
        elif leaf == "weight" and len(shape) == 1:
            out[key] = _uniform(shape, 0.5, 1.5, g)  # norm-layer gamma
        elif leaf == "bias" and (key[: -len(".bias")] + ".weight") in template and \
                len(tuple(template[key[: -len(".bias")] + ".weight"].shape)) == 1:
            out[key] = _normal(shape, 0.0, 0.2, g)   # norm-layer beta
        elif leaf == "weight":
            fan_out = shape[0]
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            rf = 1
            for s in shape[2:]:
                rf *= s
            bound = math.sqrt(6.0 / (fan_in + fan_out * rf))
            out[key] = _uniform(shape, -bound, bound, g)
        elif leaf == "bias":
            out[key] = _normal(shape, 0.0, 0.05, g)
        else:
            raise KeyError(f"seeded_state_dict: no rule for {key} {shape}")
    for key in keys:
        if key.endswith("weight_g"):
            v = out[key[: -len("weight_g")] + "weight_v"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(template[key].shape)
            out[key] = norm * _uniform(tuple(template[key].shape), 0.8, 1.25, _gen(seed, key))
    return out


def synthetic_mel(batch: int, frames: int, seed: int = 1234) -> torch.Tensor:
    """log10-mel-like input in [-5, 1] (clamp floor of melgan/modules.py:68), (B, T, 80) fp32."""
    g = _gen(seed, f"mel/{batch}x{frames}")
    return _uniform((batch, frames, 80), -5.0, 1.0, g)


def synthetic_speaker(batch: int, seed: int = 1234, tag: str = "org", dim: int = 256) -> torch.Tensor:
    """L2-normalised N(0,1) speaker codes, what LstmDV emits (factory/LstmDV.py:22-24)."""
    g = _gen(seed, f"spk/{tag}/{batch}")
    e = _normal((batch, dim), 0.0, 1.0, g)
    return e / e.norm(dim=-1, keepdim=True)
