"""CPU oracle for ``factory/Adjust.py:7-43`` and the ``*_Adjust`` model variants.  Test infrastructure.

``Adjust`` re-estimates a speaker embedding from (mel, embedding): speaker-code concat, 3x [Conv1d k5 + BN + ReLU],
3x LSTM(512 -> 768), last step, Linear(768 -> 256), L2 normalisation.  ``AutoVC_Adjust.forward``
(factory/AutoVC_Adjust.py:177-205) and ``MetaConv_Adjust.forward`` (factory/MetaConv_Adjust.py:255-280) adjust BOTH
speaker codes; ``factory/MetaPool_Adjust.py:258-283`` (whose class is still named ``MetaPool``) adjusts only the
target code and returns c_org unchanged.  All return the 4-tuple ``(c_org, mel, mel_postnet, codes)``.
"""
import torch

from .autovc import decoder_mel, encoder_codes, postnet_residual, upsample_codes
from .layers import cast_state_dict, conv_bn, lstm_stack


def adjust_forward(sd, x, emb, prefix="adjust", lstm_impl="aten", taps=None):
    """``Adjust.forward(x, emb)`` (factory/Adjust.py:29-43).  x (B,T,80) or (B,1,T,80); emb (B,E) -> (B,256)."""
    if x.dim() == 4:
        x = x.squeeze(1)                                               # Adjust.py:30
    T = x.shape[1]
    h = torch.cat((x.transpose(2, 1), emb.unsqueeze(-1).expand(-1, -1, T)), dim=1)   # :30-32
    for i in range(3):                                                 # :36-37
        h = conv_bn(sd, f"{prefix}.convolutions.{i}", h, "relu")
        if taps is not None:
            taps[f"adjust_conv{i}"] = h.transpose(1, 2)
    out = lstm_stack(sd, f"{prefix}.lstm", h.transpose(1, 2), num_layers=3, impl=lstm_impl)   # :38-39
    last = out[:, -1, :]
    e = last @ sd[f"{prefix}.embedding.linear_layer.weight"].t() + sd[f"{prefix}.embedding.linear_layer.bias"]  # :40
    return e / e.norm(p=2, dim=-1, keepdim=True)                       # :41-42


@torch.no_grad()
def autovc_adjust_forward(sd, x, c_org, c_trg, dim_neck, freq, is_convert=False, x_target=None,
                          dtype=torch.float32, lstm_impl="aten"):
    """``AutoVC_Adjust.forward`` in eval mode (factory/AutoVC_Adjust.py:177-205)."""
    sd = cast_state_dict(sd, dtype)
    x = x.to(dtype)
    c_org = adjust_forward(sd, x, c_org.to(dtype), lstm_impl=lstm_impl)              # :179
    codes = encoder_codes(sd, x, c_org, dim_neck, freq, None, lstm_impl)            # :180
    flat = codes.reshape(codes.shape[0], -1)
    if c_trg is None:
        return flat                                                                  # :182-183
    src = x_target.to(dtype) if is_convert else x                                    # :184-189
    c_trg = adjust_forward(sd, src, c_trg.to(dtype), lstm_impl=lstm_impl)
    xs = x.squeeze(1) if x.dim() == 4 else x
    mel = decoder_mel(sd, upsample_codes(codes, xs.shape[1], c_trg), None, lstm_impl)
    post = mel + postnet_residual(sd, mel)
    return c_org, mel.unsqueeze(1), post.unsqueeze(1), flat


@torch.no_grad()
def meta_adjust_forward(sd, kind, x, c_org, c_trg, dim_neck, freq, is_convert=False, x_target=None,
                        dtype=torch.float32):
    """Meta ``*_Adjust`` forward.  kind "conv" (factory/MetaConv_Adjust.py:255-280): c_org is adjusted before the
    encoder, as in AutoVC_Adjust.  kind "pool" (factory/MetaPool_Adjust.py:258-283): the encoder sees the UN-adjusted
    c_org, only the target code goes through ``Adjust`` and c_org is returned unchanged."""
    from .meta import meta_forward
    sd = cast_state_dict(sd, dtype)
    x = x.to(dtype)
    if kind == "conv":
        c_org = adjust_forward(sd, x, c_org.to(dtype))
    if c_trg is None:
        return meta_forward(sd, x, c_org.to(dtype), None, dim_neck, freq, kind, dtype=dtype)
    src = x_target.to(dtype) if is_convert else x
    c_trg = adjust_forward(sd, src, c_trg.to(dtype))
    mel, post, codes = meta_forward(sd, x, c_org.to(dtype), c_trg, dim_neck, freq, kind, dtype=dtype)
    return c_org.to(dtype), mel, post, codes
