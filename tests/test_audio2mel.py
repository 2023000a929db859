"""Audio2Mel front end (melgan/modules.py:26-69): filter bank pinning, host logic on CPU stand-ins,
GPU parity.  The reference's own Audio2Mel cannot run here (librosa absent, pre-2.0 torch.stft call): see the header of
oracle/audio2mel.py for what is and is not pinned."""
import numpy as np
import pytest
import torch

from oracle import rel_l2
from oracle.audio2mel import audio2mel_forward, librosa_mel
from tests.emulate import install_cpu_kernels


def _audio(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L) / 22050.0
    tones = sum(a * torch.sin(2 * torch.pi * f * t + p) for a, f, p in ((0.4, 220.0, 0.1), (0.2, 1760.0, 1.0), (0.1, 5200.0, 2.0)))
    return (tones.unsqueeze(0) + 0.05 * torch.randn(B, L, generator=g)).unsqueeze(1)      # (B, 1, L)


def test_filter_bank_matches_librosa_compatible_implementation():
    from autoformer_b200.melgan.filters import mel_filterbank
    ours = mel_filterbank(22050, 1024, 80, 0.0, None)
    oracle = librosa_mel(22050, 1024, 80, 0.0, None).numpy()
    assert ours.shape == (80, 513) and ours.dtype == np.float32
    assert np.abs(ours - oracle).max() < 1e-8                   # product restatement vs oracle restatement
    try:
        from transformers.audio_utils import mel_filter_bank
    except Exception:                                            # pragma: no cover
        pytest.skip("transformers not importable")
    tf = mel_filter_bank(num_frequency_bins=513, num_mel_filters=80, min_frequency=0.0, max_frequency=11025.0,
                         sampling_rate=22050, norm="slaney", mel_scale="slaney").T
    assert np.abs(ours - tf).max() < 1e-8                        # third implementation, written to match librosa
    assert (ours >= 0).all() and ((ours > 0).sum(axis=1) >= 1).all()
    # Slaney normalisation: every triangle has unit area in Hz
    df = 22050 / 1024
    assert np.allclose(ours[10:].sum(axis=1) * df, 1.0, atol=0.05)


@pytest.mark.parametrize("B,L", [(2, 256 * 20), (1, 256 * 7 + 100)])
def test_audio2mel_host_logic(monkeypatch, B, L):
    install_cpu_kernels(monkeypatch)
    from autoformer_b200.melgan.modules import Audio2Mel
    audio = _audio(B, L, 3)
    ref = audio2mel_forward(audio, dtype=torch.float64)
    m = Audio2Mel()
    assert set(m.state_dict()) == {"mel_basis", "window"}       # the reference's buffer names
    out = m(audio)
    assert out.shape == ref.shape == (B, 80, (L + 768 - 1024) // 256 + 1)
    assert rel_l2(out, ref) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,L", [(3, 256 * 40), (1, 256 * 9 + 77), (2, 22050)])
def test_audio2mel_gpu_parity(precision, tol, B, L):
    from autoformer_b200.melgan.modules import Audio2Mel
    audio = _audio(B, L, 5)
    ref = audio2mel_forward(audio, dtype=torch.float64)
    m = Audio2Mel().cuda()
    m.precision = precision
    out = m(audio.cuda())
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < tol, rel_l2(out, ref)
    if precision == "fp32":
        assert rel_l2(m(_audio(B, L, 6).cuda()), ref) > 1e-3     # negative control: another signal must not pass


@pytest.mark.gpu
def test_melvocoder_round_trip_shapes():
    """MelVocoder surface: inverse (mel -> wav) then __call__ (wav -> mel) returns the frame count it started from."""
    from autoformer_b200.melgan.interface import MelVocoder
    from oracle import templates
    from oracle.seeded import seeded_state_dict, synthetic_mel
    voc = MelVocoder(device="cuda", state_dict=seeded_state_dict(templates.melgan_template(), 4))
    mel = synthetic_mel(2, 24, 9).transpose(1, 2).contiguous()
    wav = voc.inverse(mel)
    assert wav.shape == (2, 24 * 256)
    back = voc(wav)
    assert back.shape == (2, 80, 24)
    assert rel_l2(back, audio2mel_forward(wav.cpu().unsqueeze(1), dtype=torch.float64)) < 2e-4
