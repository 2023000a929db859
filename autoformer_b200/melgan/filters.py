"""Mel filter bank of ``librosa.filters.mel`` (Slaney scale, area normalisation), restated from the published
algorithm because librosa is not a dependency here.  The reference builds its ``Audio2Mel.mel_basis`` buffer with
``librosa_mel_fn(sampling_rate, n_fft, n_mel_channels, mel_fmin, mel_fmax)`` (melgan/modules.py:42-44, librosa's
defaults htk=False, norm="slaney"); a reference checkpoint that carries ``mel_basis`` overrides this on load."""
import numpy as np


def hz_to_mel(f):
    """Slaney's auditory-toolbox scale: linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per x6.4)."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """(n_mels, 1 + n_fft // 2) float32 triangular filters with unit area per band (norm="slaney")."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fftfreqs = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)
