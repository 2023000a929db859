"""CPU oracle for ``Audio2Mel`` (melgan/modules.py:26-69).  Test infrastructure.

Pinned (round 2) against the reference's own ``Audio2Mel`` class run through two shims
(``oracle/make_golden.py::golden_audio2mel`` -> tests/golden/audio2mel_*.npz): as written the class cannot run on this
container's stack -- it builds ``mel_basis`` with librosa (absent, no network) and calls ``torch.stft`` without
``return_complex``, which torch >= 2.0 rejects -- so (1) ``librosa.filters.mel`` is supplied by the restated filter
bank below and (2) ``torch.stft`` is wrapped to return the pre-2.0 real view.  Padding, framing, window, magnitude, mel
projection and log10 clamp are the reference's code.  NOT pinned by the reference: the filter bank itself (librosa
0.8's published ``filters.mel`` algorithm, restated as plain loops); it agrees to 1e-9 with the librosa-compatible
``transformers.audio_utils.mel_filter_bank`` (tests/test_audio2mel.py).
"""
import math

import torch
import torch.nn.functional as F


def _hz_to_mel(f):
    f_sp = 200.0 / 3.0
    if f >= 1000.0:
        return 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f / f_sp


def _mel_to_hz(m):
    f_sp = 200.0 / 3.0
    min_log_mel = 1000.0 / f_sp
    if m >= min_log_mel:
        return 1000.0 * math.exp((math.log(6.4) / 27.0) * (m - min_log_mel))
    return f_sp * m


def librosa_mel(sr=22050, n_fft=1024, n_mels=80, fmin=0.0, fmax=None):
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with its defaults htk=False, norm='slaney'."""
    fmax = sr / 2.0 if fmax is None else fmax
    bins = 1 + n_fft // 2
    fft_f = [i * (sr / 2.0) / (bins - 1) for i in range(bins)]
    m_lo, m_hi = _hz_to_mel(fmin), _hz_to_mel(fmax)
    mel_f = [_mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    w = torch.zeros(n_mels, bins, dtype=torch.float64)
    for i in range(n_mels):
        enorm = 2.0 / (mel_f[i + 2] - mel_f[i])
        for k, f in enumerate(fft_f):
            lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i])
            upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1])
            w[i, k] = max(0.0, min(lower, upper)) * enorm
    return w.float()


@torch.no_grad()
def audio2mel_forward(audio, mel_basis=None, n_fft=1024, hop_length=256, win_length=1024, dtype=torch.float32):
    """``Audio2Mel.forward`` (melgan/modules.py:54-69).  audio (B, 1, L) -> (B, 80, frames)."""
    if mel_basis is None:
        mel_basis = librosa_mel(22050, n_fft, 80, 0.0, None)
    window = torch.hann_window(win_length).to(dtype)
    p = (n_fft - hop_length) // 2
    a = F.pad(audio.to(dtype), (p, p), "reflect").squeeze(1)                                  # :55-56
    fft = torch.stft(a, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=False,
                     return_complex=True)                                                     # :57-64
    magnitude = torch.sqrt(fft.real ** 2 + fft.imag ** 2)                                     # :65-66
    mel_output = torch.matmul(mel_basis.to(dtype), magnitude)                                 # :67
    return torch.log10(torch.clamp(mel_output, min=1e-5))                                     # :68
