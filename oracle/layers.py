"""Layer-level CPU restatements shared by the oracle models.  Test infrastructure.

Every function works on plain tensors and a flat ``state_dict`` (name -> tensor);
no ``nn.Module`` from the reference is involved.  ``dtype`` may be float32 (the
reference's precision) or float64 (ground truth for tolerance studies).
"""
import torch
import torch.nn.functional as F


def conv1d_k(sd, prefix, x, padding, dilation=1):
    """``nn.Conv1d`` cross-correlation with zero padding (factory/Norm.py:21-29,35-37).

    x: (B, C_in, T).  y[b,o,t] = bias[o] + sum_i sum_k w[o,i,k] x[b,i,t + k*dil - pad].
    """
    return F.conv1d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], stride=1,
                    padding=padding, dilation=dilation)


def batchnorm_eval(sd, prefix, x, eps=1e-5):
    """``nn.BatchNorm1d`` in eval mode: running statistics, per-channel affine."""
    mean = sd[prefix + ".running_mean"].view(1, -1, 1)
    var = sd[prefix + ".running_var"].view(1, -1, 1)
    w = sd[prefix + ".weight"].view(1, -1, 1)
    b = sd[prefix + ".bias"].view(1, -1, 1)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def conv_bn(sd, prefix, x, act):
    """``nn.Sequential(ConvNorm(k=5,p=2), BatchNorm1d)`` + activation
    (factory/AutoVC.py:26-39,50-51 and :127-179).  Keys: ``{prefix}.0.conv.*``, ``{prefix}.1.*``."""
    y = conv1d_k(sd, prefix + ".0.conv", x, padding=2)
    y = batchnorm_eval(sd, prefix + ".1", y)
    if act == "relu":
        return torch.relu(y)
    if act == "tanh":
        return torch.tanh(y)
    if act == "none":
        return y
    raise ValueError(act)


def lstm_explicit(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """One LSTM direction spelled out gate by gate (PyTorch ``nn.LSTM`` semantics).

    x: (B, T, I).  Rows of w_ih (4H, I) / w_hh (4H, H) are ordered [i; f; g; o];
    z = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh; c_t = s(z_f) c_{t-1} + s(z_i) tanh(z_g);
    h_t = s(z_o) tanh(c_t); h_0 = c_0 = 0.  The reverse direction runs t = T-1 .. 0 and
    its output at index t is aligned with time t.
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_empty(B, T, H)
    xp = x @ w_ih.t() + (b_ih + b_hh)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        z = xp[:, t] + h @ w_hh.t()
        zi, zf, zg, zo = z.split(H, dim=1)
        c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
        h = torch.sigmoid(zo) * torch.tanh(c)
        out[:, t] = h
    return out


def lstm_stack(sd, prefix, x, num_layers, bidirectional=False, impl="aten"):
    """``nn.LSTM(batch_first=True)`` forward, zero initial state
    (factory/AutoVC.py:43,54; :77,103; :96,110; factory/LstmDV.py:12-20).

    ``impl="aten"`` calls the same ATen kernel the reference reaches through
    ``nn.LSTM`` (used for CPU timing); ``impl="explicit"`` uses ``lstm_explicit``.
    """
    dirs = ["", "_reverse"] if bidirectional else [""]
    if impl == "explicit":
        for layer in range(num_layers):
            outs = []
            for d in dirs:
                sfx = f"_l{layer}{d}"
                outs.append(lstm_explicit(
                    x, sd[f"{prefix}.weight_ih{sfx}"], sd[f"{prefix}.weight_hh{sfx}"],
                    sd[f"{prefix}.bias_ih{sfx}"], sd[f"{prefix}.bias_hh{sfx}"],
                    reverse=(d != "")))
            x = torch.cat(outs, dim=-1)
        return x
    flat = []
    for layer in range(num_layers):
        for d in dirs:
            sfx = f"_l{layer}{d}"
            flat += [sd[f"{prefix}.weight_ih{sfx}"], sd[f"{prefix}.weight_hh{sfx}"],
                     sd[f"{prefix}.bias_ih{sfx}"], sd[f"{prefix}.bias_hh{sfx}"]]
    H = flat[1].shape[1]
    n = num_layers * len(dirs)
    h0 = x.new_zeros(n, x.shape[0], H)
    out, _, _ = torch.lstm(x, (h0, h0.clone()), flat, True, num_layers, 0.0, False,
                           bidirectional, True)
    return out


def cast_state_dict(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
