"""Round-2 boundary pieces, each pinned to outputs of the unmodified reference (tests/golden, oracle/make_golden.py):
the ``make_data`` LstmDV twin with its ``(predictions, d_vec)`` contract (make_data/factory/LstmDV.py:4-25,
util/evaluate.py:104), the Evaluator's crop_mel / get_trans_mel recipe (util/evaluate.py:36-98), the shim-run reference
Audio2Mel (melgan/modules.py:26-69), plus the host-side fixes (plan-cache invalidation, hyper-parameter validation)."""
import os
import types
import warnings

import numpy as np
import pytest
import torch

from oracle import rel_l2, templates
from oracle.audio2mel import audio2mel_forward
from oracle.autovc import autovc_forward
from oracle.lstmdv import lstmdv_twin_forward
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker
from tests.emulate import install_cpu_kernels

warnings.filterwarnings("ignore", category=FutureWarning)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture
def cpu_kernels(monkeypatch):
    install_cpu_kernels(monkeypatch)


# ----------------------------------------------------------------------------------------------- oracle pinning (CPU)
def test_lstmdv_twin_oracle_matches_reference():
    g = _load("lstmdv_twin_b2_t100")
    sd = seeded_state_dict(templates.lstmdv_twin_template(), int(g["wseed"]), lstm_gain=float(g["lstm_gain"]))
    x = synthetic_mel(int(g["B"]), int(g["T"]), int(g["xseed"]))
    pred, dv = lstmdv_twin_forward(sd, x)
    assert rel_l2(pred, torch.from_numpy(g["predictions"])) < TOL
    assert rel_l2(dv, torch.from_numpy(g["d_vec"])) < TOL


@pytest.mark.parametrize("name", ["audio2mel_b2_l5120", "audio2mel_b1_l2381"])
def test_audio2mel_oracle_matches_shimmed_reference(name):
    g = _load(name)
    ours = audio2mel_forward(torch.from_numpy(g["audio"]))
    assert ours.shape == g["mel"].shape
    assert rel_l2(ours, torch.from_numpy(g["mel"])) < TOL


def _evaluator_oracle(g, trim):
    """crop_mel (zero-pad at the end to len_crop = 176) -> AutoVC(44,256,512,22) -> optional trim, restated."""
    sd = seeded_state_dict(templates.autovc_template(44, 256, 512, 22), int(g["wseed"]))
    src = torch.from_numpy(g["src"])
    x = torch.nn.functional.pad(src, (0, 0, 0, 176 - src.shape[0])).unsqueeze(0)
    _, post, _ = autovc_forward(sd, x, torch.from_numpy(g["e_src"])[None], torch.from_numpy(g["e_trg"])[None], 44, 22)
    post = post.squeeze(1)
    return post[:, :src.shape[0]] if trim else post


def test_evaluator_recipe_oracle_matches_reference():
    g = _load("evaluator_autovcR_t100_t150")
    assert rel_l2(_evaluator_oracle(g, True), torch.from_numpy(g["mel_trans_play"])) < TOL
    assert rel_l2(_evaluator_oracle(g, False), torch.from_numpy(g["mel_trans_full"])) < TOL
    # padding to the next multiple of freq (110 frames) instead of len_crop is a DIFFERENT mel: the recipe matters
    sd = seeded_state_dict(templates.autovc_template(44, 256, 512, 22), int(g["wseed"]))
    x110 = torch.nn.functional.pad(torch.from_numpy(g["src"]), (0, 0, 0, 10)).unsqueeze(0)
    _, post110, _ = autovc_forward(sd, x110, torch.from_numpy(g["e_src"])[None], torch.from_numpy(g["e_trg"])[None], 44, 22)
    assert rel_l2(post110.squeeze(1)[:, :100], torch.from_numpy(g["mel_trans_play"])) > 1e-3


# ----------------------------------------------------------------------------------------------- host logic (CPU stand-ins)
def test_crop_mel_matches_reference_semantics():
    from autoformer_b200.pipeline import crop_mel
    short = np.random.RandomState(0).randn(100, 80).astype(np.float32)
    m, pad = crop_mel(short, 176)
    assert m.shape == (1, 176, 80) and pad == 76
    assert torch.equal(m[0, :100], torch.from_numpy(short)) and float(m[0, 100:].abs().max()) == 0.0
    m, pad = crop_mel(short[:176 - 76 + 76][:100], 100)
    assert m.shape == (1, 100, 80) and pad == 0
    long = np.random.RandomState(1).randn(300, 80).astype(np.float32)
    np.random.seed(5)
    left = np.random.randint(0, 300 - 176)                        # the draw the reference makes (evaluate.py:48)
    np.random.seed(5)
    m, pad = crop_mel(long, 176)
    assert pad == 0 and torch.equal(m[0], torch.from_numpy(long[left:left + 176]))


def _make_evaluator(g, root, embedder=None, vocoder=None):
    from autoformer_b200.util.evaluate import Evaluator
    np.save(os.path.join(root, "a.npy"), g["src"])
    np.save(os.path.join(root, "b.npy"), g["trg"])
    cfg = types.SimpleNamespace(root=str(root), num_speaker=2, batch_size=1, max_uttr_idx=4, erroment_num=1,
                                len_crop=176, device="cpu", all_speaker=["p1", "p2"], embedder=embedder,
                                metadata=[["p1", g["e_src"], "a.npy"], ["p2", g["e_trg"], "b.npy"]], vocoder=vocoder)
    return Evaluator, cfg


def test_evaluator_get_trans_mel_host_logic(cpu_kernels, tmp_path, monkeypatch):
    from autoformer_b200.factory.AutoVC import AutoVC
    from autoformer_b200.melgan import interface
    g = _load("evaluator_autovcR_t100_t150")
    Evaluator, cfg = _make_evaluator(g, tmp_path)
    monkeypatch.setattr(interface.MelVocoder, "__init__", lambda self, **kw: None)     # no checkpoint on disk
    E = Evaluator(cfg)
    model = AutoVC(44, 256, 512, 22)
    model.load_state_dict(seeded_state_dict(templates.autovc_template(44, 256, 512, 22), int(g["wseed"])))
    model.eval()
    ms, mt, trans = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=True)
    assert ms.shape == (1, 100, 80) and mt.shape == (1, 150, 80) and trans.shape == (1, 100, 80)
    assert torch.equal(ms, torch.from_numpy(g["mel_source"])) and torch.equal(mt, torch.from_numpy(g["mel_target"]))
    assert rel_l2(trans, torch.from_numpy(g["mel_trans_play"])) < 1e-4
    _, _, full = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=False)
    assert full.shape == (1, 176, 80) and rel_l2(full, torch.from_numpy(g["mel_trans_full"])) < 1e-4


def test_lstmdv_twin_host_logic(cpu_kernels, monkeypatch):
    from autoformer_b200 import ops
    from autoformer_b200.make_data.factory.LstmDV import LstmDV

    def emu_linear_rows(h, w, bias, want_raw=True, want_normed=False):
        e = h.double() @ w.double().t() + bias.double()
        return (e.float() if want_raw else None), ((e / e.norm(dim=-1, keepdim=True)).float() if want_normed else None)
    monkeypatch.setattr(ops, "linear_rows", emu_linear_rows)
    g = _load("lstmdv_twin_b2_t100")
    sd = seeded_state_dict(templates.lstmdv_twin_template(), int(g["wseed"]), lstm_gain=float(g["lstm_gain"]))
    m = LstmDV()
    assert set(m.state_dict()) == set(sd)                          # strict load of a twin checkpoint
    m.load_state_dict(sd, strict=True)
    out = m(synthetic_mel(int(g["B"]), int(g["T"]), int(g["xseed"])))
    assert isinstance(out, tuple) and len(out) == 2                # Evaluator indexes [1] (util/evaluate.py:104)
    assert rel_l2(out[0], torch.from_numpy(g["predictions"])) < 1e-4
    assert rel_l2(out[1], torch.from_numpy(g["d_vec"])) < 1e-4


def test_plan_cache_sees_data_writes_and_invalidate(cpu_kernels):
    """`.data` writes do not bump the version counter; the content digest must catch them."""
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    sd = seeded_state_dict(templates.autovc_template(*args), 0)
    x, c_org, c_trg = synthetic_mel(1, 32, 5), synthetic_speaker(1, 5, "org"), synthetic_speaker(1, 5, "trg")
    m = AutoVC(*args)
    m.load_state_dict(sd)
    m.eval()
    base = m(x, c_org, c_trg)[1].clone()
    plan0 = m._plan()
    assert m._plan() is plan0                                      # nothing changed: no re-pack
    w = m.postnet.convolutions[4][0].conv.weight
    v0 = w._version
    w.data.mul_(2.0)                                               # invisible to the version counter
    assert w._version == v0
    changed = m(x, c_org, c_trg)[1]
    assert m._plan() is not plan0 and rel_l2(changed, base) > 1e-3
    w.data.mul_(0.5)
    assert rel_l2(m(x, c_org, c_trg)[1], base) < 1e-6
    # load_state_dict and .to() re-pack as well
    plan1 = m._plan()
    m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 1))
    assert m._plan() is not plan1 and rel_l2(m(x, c_org, c_trg)[1], base) > 1e-3
    plan2 = m._plan()
    m.double().float()                                             # .to(): parameters are re-created
    assert m._plan() is not plan2
    # frozen: the digest is skipped, invalidate() is the explicit way
    m.freeze_weights()
    plan3 = m._plan()
    w = m.postnet.convolutions[4][0].conv.weight
    w.data.mul_(2.0)
    assert m._plan() is plan3                                      # documented: frozen models do not look
    m.invalidate()
    assert m._plan() is not plan3
    m.freeze_weights(False)


def test_hyper_parameter_validation():
    from autoformer_b200.factory.AutoVC import AutoVC
    AutoVC(32, 256, 512, 32)
    AutoVC(44, 256, 512, 22)
    for bad in [(65, 256, 512, 32), (32, 256, 500, 32), (32, 250, 512, 32), (33, 256, 512, 32)]:
        with pytest.raises(ValueError):
            AutoVC(*bad)
    from autoformer_b200.factory.LstmDV import LstmDV
    LstmDV()
    with pytest.raises(ValueError):
        LstmDV(dim_cell=700)


# ----------------------------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16x2", 1e-3)])
def test_lstmdv_twin_gpu_matches_reference_golden(precision, tol):
    from autoformer_b200.make_data.factory.LstmDV import LstmDV
    g = _load("lstmdv_twin_b2_t100")
    sd = seeded_state_dict(templates.lstmdv_twin_template(), int(g["wseed"]), lstm_gain=float(g["lstm_gain"]))
    m = LstmDV()
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    x = synthetic_mel(int(g["B"]), int(g["T"]), int(g["xseed"])).cuda()
    pred, dv = m(x)
    assert rel_l2(pred, torch.from_numpy(g["predictions"])) < tol, rel_l2(pred, torch.from_numpy(g["predictions"]))
    assert rel_l2(dv, torch.from_numpy(g["d_vec"])) < tol
    assert rel_l2(m(x)[1], m(x.flip(1))[1]) > 1e-2                 # negative control


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["audio2mel_b2_l5120", "audio2mel_b1_l2381"])
def test_audio2mel_gpu_matches_shimmed_reference(name):
    from autoformer_b200.melgan.modules import Audio2Mel
    g = _load(name)
    m = Audio2Mel().cuda()
    out = m(torch.from_numpy(g["audio"]).cuda())
    assert out.shape == g["mel"].shape
    assert rel_l2(out, torch.from_numpy(g["mel"])) < 2e-4, rel_l2(out, torch.from_numpy(g["mel"]))


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16x2", 1e-3)])
def test_evaluator_gpu_matches_reference_golden(precision, tol, tmp_path):
    from autoformer_b200.factory.AutoVC import AutoVC
    from autoformer_b200.melgan.interface import MelVocoder
    g = _load("evaluator_autovcR_t100_t150")
    voc = MelVocoder(device="cuda", state_dict=seeded_state_dict(templates.melgan_template(), 4))
    Evaluator, cfg = _make_evaluator(g, tmp_path, vocoder=voc)
    cfg.device = "cuda"
    E = Evaluator(cfg)
    model = AutoVC(44, 256, 512, 22)
    model.load_state_dict(seeded_state_dict(templates.autovc_template(44, 256, 512, 22), int(g["wseed"])))
    model = model.cuda().eval()
    model.precision = precision
    ms, mt, trans = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=True)
    assert trans.shape == (1, 100, 80)
    assert rel_l2(trans, torch.from_numpy(g["mel_trans_play"])) < tol, rel_l2(trans, torch.from_numpy(g["mel_trans_play"]))
    _, _, full = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=False)
    assert rel_l2(full, torch.from_numpy(g["mel_trans_full"])) < tol
    wav = E.get_wavs(trans.transpose(2, 1))                        # conversion.ipynb cell 14
    assert wav.shape == (1, 100 * 256) and bool(torch.isfinite(wav).all())


@pytest.mark.gpu
def test_plan_cache_data_write_on_gpu():
    from autoformer_b200.melgan.modules import Generator
    gen = Generator(80, 32, 3)
    gen.load_state_dict(seeded_state_dict(templates.melgan_template(), 4))
    gen = gen.cuda().eval()
    mel = synthetic_mel(1, 16, 3).transpose(1, 2).contiguous().cuda()
    a = gen(mel).clone()
    gen.model[1].weight_g.data.mul_(1.5)                           # what weights_init-style code does (modules.py:9-15)
    b = gen(mel)
    assert rel_l2(b, a) > 1e-3
    gen.model[1].weight_g.data.div_(1.5)
    assert rel_l2(gen(mel), a) < 1e-5


@pytest.mark.gpu
def test_model_on_second_device_runs_there():
    """a model on cuda:1 must launch on cuda:1 whatever the current device is."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    sd = seeded_state_dict(templates.autovc_template(*args), 11)
    x, c_org, c_trg = synthetic_mel(2, 64, 21), synthetic_speaker(2, 21, "org"), synthetic_speaker(2, 21, "trg")
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = AutoVC(*args)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        torch.cuda.set_device(0)                                   # current device stays 0
        outs.append([t.cpu() for t in m(x.to(dev), c_org.to(dev), c_trg.to(dev))])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
