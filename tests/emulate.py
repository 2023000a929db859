"""CPU emulation of the kernels' INDEXING (k-block schedule, gate interleave) on packed weights, in fp64.
Used by CPU tests to validate autoformer_b200.packing against torch's own conv/LSTM ops before any GPU run."""
import torch

from autoformer_b200.packing import KC


def emulate_conv_gemm(w, bias, meta, srcs, B, T, tap_t0, tap_dt):
    """Mirror of avc_conv_gemm's k-block walk: srcs are channels-last [B][rows][C] tensors."""
    kc = KC[meta["precision"]]
    n = meta["N"]
    W = w.double()[:n]
    out = bias.double()[:n].view(1, 1, n).expand(B, T, n).clone()
    col = 0
    for s, a in enumerate(srcs):
        a = a.double()
        C = meta["channels"][s]
        chunks = (C + kc - 1) // kc
        rows_avail = a.shape[1]
        for tap in range(meta["taps"][s]):
            rows = tap_t0[s] + torch.arange(T) + tap * tap_dt[s]
            ok = (rows >= 0) & (rows < rows_avail)
            A = torch.zeros(B, T, chunks * kc, dtype=torch.float64)
            A[:, ok, :C] = a[:, rows[ok], :]
            out += A @ W[:, col:col + chunks * kc].t()
            col += chunks * kc
    assert col == meta["k_pad"]
    return out


def emulate_lstm_seq(xproj, w_hh_packed, B, T, H, G):
    """Mirror of lstm_step_kernel: packed gate order, tile width 4G."""
    xp = xproj.double().view(B, T, 4 * H)
    W = w_hh_packed.double()
    h = torch.zeros(B, H, dtype=torch.float64)
    c = torch.zeros(B, H, dtype=torch.float64)
    out = torch.zeros(B, T, H, dtype=torch.float64)
    u = torch.arange(H)
    base = (u // G) * 4 * G + u % G
    for t in range(T):
        z = xp[:, t] + h @ W.t()
        zi, zf, zg, zo = (z[:, base + g * G] for g in range(4))
        c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
        h = torch.sigmoid(zo) * torch.tanh(c)
        out[:, t] = h
    return out
