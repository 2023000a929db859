"""CPU oracle for the AdaIN "2" variants (factory/AutoVC2.py, MetaPool2.py, MetaConv2.py).  Test infrastructure.

Encoder: three [Conv1d(80->80,k5) + BN] layers on the mel before the speaker concat, recording the batch-global scalar
``x.mean()`` / ``x.std()`` after each (AutoVC2.py:54-60); Postnet: after the five convolutions, three times
``x = combine_i(AdaIN(x, mu_i, std_i))`` (AutoVC2.py:192-203, factory/Norm.py:86-94).
"""
import torch

from .autovc import decoder_mel, encoder_codes, postnet_residual, upsample_codes
from .layers import batchnorm_eval, cast_state_dict, conv1d_k


def pre_extract(sd, x):
    """x (B,T,80) -> (features output (B,T,80), [[mean, std]] x 3)   (AutoVC2.py:55-60)."""
    if x.dim() == 4:
        x = x.squeeze(1)
    h = x.transpose(2, 1)
    feats = []
    for i in range(3):
        h = conv1d_k(sd, f"encoder.feature_pre_extract.{i}.0.conv", h, padding=2)
        h = batchnorm_eval(sd, f"encoder.feature_pre_extract.{i}.1", h)
        feats.append([h.mean(), h.std()])
    return h.transpose(2, 1), feats


def adain(content, mu, std):
    """factory/Norm.py:86-94: scalar statistics over the WHOLE tensor, Bessel-corrected std."""
    return (content - content.mean()) / content.std() * std + mu


def postnet_adain(sd, mel, feats):
    """``Postnet.forward(x, features)`` (AutoVC2.py:192-203).  mel (B,T,80) -> (B,T,80)."""
    h = postnet_residual(sd, mel).transpose(2, 1)                 # the five convolutions
    for i in range(3):
        h = adain(h, feats[i][0], feats[i][1])
        h = conv1d_k(sd, f"postnet.feature_last_combine.{i}.0.conv", h, padding=2)
    return h.transpose(2, 1)


def _finish(sd, codes, feats, x, c_trg, target_feature, decode):
    flat = codes.reshape(codes.shape[0], -1)
    if c_trg is None and target_feature is None:
        return flat, feats
    T = x.shape[-2]
    mel = decode(upsample_codes(codes, T, c_trg))
    style = target_feature if target_feature is not None else feats
    post = mel + postnet_adain(sd, mel, style)
    return mel.unsqueeze(1), post.unsqueeze(1), flat


@torch.no_grad()
def autovc2_forward(sd, x, c_org, c_trg, dim_neck, freq, target_feature=None, dtype=torch.float32, lstm_impl="aten"):
    """``AutoVC2.forward`` in eval mode (factory/AutoVC2.py:213-243)."""
    sd = cast_state_dict(sd, dtype)
    x = x.to(dtype)
    xf, feats = pre_extract(sd, x)
    codes = encoder_codes(sd, xf, c_org.to(dtype), dim_neck, freq, None, lstm_impl)
    c_trg = c_trg.to(dtype) if c_trg is not None else None
    return _finish(sd, codes, feats, xf, c_trg, target_feature, lambda d: decoder_mel(sd, d, None, lstm_impl))


@torch.no_grad()
def meta2_forward(sd, kind, x, c_org, c_trg, dim_neck, freq, target_feature=None, dtype=torch.float32):
    """``MetaPool2.forward`` / ``MetaConv2.forward`` (factory/MetaPool2.py:303-330)."""
    from .meta import meta_decoder, meta_encoder
    sd = cast_state_dict(sd, dtype)
    x = x.to(dtype)
    xf, feats = pre_extract(sd, x)
    codes = meta_encoder(sd, xf, c_org.to(dtype), dim_neck, freq, kind)
    c_trg = c_trg.to(dtype) if c_trg is not None else None
    return _finish(sd, codes, feats, xf, c_trg, target_feature, lambda d: meta_decoder(sd, d, kind))
