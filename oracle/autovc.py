"""CPU oracle for ``factory/AutoVC.py`` (eval-mode forward).  Test infrastructure.

``autovc_forward`` follows ``AutoVC.forward`` (factory/AutoVC.py:191-211) and returns the
same 3-tuple; ``taps`` (optional dict) is filled with every intermediate the parity
tests gate on, all in channels-last (B, T, C) layout.
"""
import torch

from .layers import conv_bn, lstm_stack, cast_state_dict


def encoder_codes(sd, x, c_org, dim_neck, freq, taps=None, lstm_impl="aten"):
    """``Encoder.forward`` (factory/AutoVC.py:45-68).  x (B,T,80) or (B,1,T,80); c_org (B,E).

    Returns the down-sampled content code as one tensor (B, T/freq, 2*dim_neck):
    code_j = [h_fwd[j*freq + freq-1] || h_bwd[j*freq]].
    """
    if x.dim() == 4:
        x = x.squeeze(1)                                   # AutoVC.py:46
    B, T, _ = x.shape
    if T % freq != 0:
        raise IndexError(f"T={T} is not a multiple of freq={freq} (AutoVC.py:60-66 indexes i+freq-1)")
    h = torch.cat((x.transpose(2, 1), c_org.unsqueeze(-1).expand(-1, -1, T)), dim=1)  # :46-48
    for i in range(3):                                     # :50-51
        h = conv_bn(sd, f"encoder.convolutions.{i}", h, "relu")
        if taps is not None:
            taps[f"enc_conv{i}"] = h.transpose(1, 2)
    h = h.transpose(1, 2)
    out = lstm_stack(sd, "encoder.lstm", h, num_layers=2, bidirectional=True, impl=lstm_impl)  # :54
    if taps is not None:
        taps["enc_lstm"] = out
    fwd = out[:, :, :dim_neck]
    bwd = out[:, :, dim_neck:]
    codes = torch.cat((fwd[:, freq - 1::freq, :], bwd[:, ::freq, :]), dim=-1)  # :56-66
    return codes


def decoder_mel(sd, dec_in, taps=None, lstm_impl="aten"):
    """``Decoder.forward`` (factory/AutoVC.py:100-114).  dec_in (B,T,2H+E) -> (B,T,80)."""
    h = lstm_stack(sd, "decoder.lstm1", dec_in, num_layers=1, impl=lstm_impl)
    if taps is not None:
        taps["dec_lstm1"] = h
    h = h.transpose(1, 2)
    for i in range(3):
        h = conv_bn(sd, f"decoder.convolutions.{i}", h, "relu")
        if taps is not None:
            taps[f"dec_conv{i}"] = h.transpose(1, 2)
    h = h.transpose(1, 2)
    h = lstm_stack(sd, "decoder.lstm2", h, num_layers=2, impl=lstm_impl)
    if taps is not None:
        taps["dec_lstm2"] = h
    w = sd["decoder.linear_projection.linear_layer.weight"]
    b = sd["decoder.linear_projection.linear_layer.bias"]
    return h @ w.t() + b


def postnet_residual(sd, mel, taps=None, prefix="postnet"):
    """``Postnet.forward`` (factory/AutoVC.py:173-179): 4x conv+BN+tanh, conv+BN.  (B,T,80)->(B,T,80)."""
    h = mel.transpose(2, 1)
    for i in range(4):
        h = conv_bn(sd, f"{prefix}.convolutions.{i}", h, "tanh")
        if taps is not None:
            taps[f"post_conv{i}"] = h.transpose(1, 2)
    h = conv_bn(sd, f"{prefix}.convolutions.4", h, "none")
    return h.transpose(2, 1)


def upsample_codes(codes, frames, c_trg):
    """Code up-sampling + target speaker concat (factory/AutoVC.py:197-204)."""
    B, n_codes, _ = codes.shape
    rep = int(frames / n_codes)
    code_exp = codes.repeat_interleave(rep, dim=1)
    return torch.cat((code_exp, c_trg.unsqueeze(1).expand(-1, frames, -1)), dim=-1)


@torch.no_grad()
def autovc_forward(sd, x, c_org, c_trg, dim_neck, freq, dtype=torch.float32, taps=None,
                   lstm_impl="aten"):
    """``AutoVC.forward(x, c_org, c_trg)`` in eval mode (factory/AutoVC.py:191-211).

    Returns ``(mel (B,1,T,80), mel_postnet (B,1,T,80), codes (B, 2H*T/freq))`` or, when
    ``c_trg is None``, the codes tensor alone (:194-195).
    """
    sd = cast_state_dict(sd, dtype)
    x = x.to(dtype)
    c_org = c_org.to(dtype)
    codes = encoder_codes(sd, x, c_org, dim_neck, freq, taps, lstm_impl)
    flat_codes = codes.reshape(codes.shape[0], -1)
    if taps is not None:
        taps["codes"] = flat_codes
    if c_trg is None:
        return flat_codes
    T = x.shape[1]
    dec_in = upsample_codes(codes, T, c_trg.to(dtype))
    mel = decoder_mel(sd, dec_in, taps, lstm_impl)
    post = mel + postnet_residual(sd, mel, taps)
    if taps is not None:
        taps["mel"] = mel
        taps["mel_postnet"] = post
    return mel.unsqueeze(1), post.unsqueeze(1), flat_codes
