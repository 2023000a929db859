"""Generate tests/golden/*.npz by running the UNMODIFIED reference.  Build-container only.

    python -m oracle.make_golden

Each fixture records the constructor arguments, the weight seed (weights are
regenerated with ``oracle.seeded.seeded_state_dict`` from the model's own state_dict
template), the input seed/shape and the reference outputs (eval mode, no_grad, fp32,
torch CPU).  Stage taps are stored sub-sampled ([:, ::8, ::8]) to keep fixtures small.
"""
import os
import sys

import numpy as np
import torch

from . import ref_import
from .seeded import seeded_state_dict, synthetic_mel, synthetic_speaker

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sub(t):
    t = t.detach()
    if t.dim() == 3:
        return t[:, ::8, ::8].contiguous().numpy()
    return t.numpy()


def _hook_taps(model, names):
    taps = {}
    handles = []
    for tap_name, module, post in names:
        def fn(_m, _i, o, tap_name=tap_name, post=post):
            o = o[0] if isinstance(o, tuple) else o
            taps[tap_name] = post(o)
        handles.append(module.register_forward_hook(fn))
    return taps, handles


def golden_autovc(ref, name, cls_name, args, B, T, wseed, xseed):
    torch.manual_seed(0)
    model = getattr(ref, cls_name)(*args).eval()
    sd = seeded_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    x = synthetic_mel(B, T, xseed)
    c_org = synthetic_speaker(B, xseed, "org")
    c_trg = synthetic_speaker(B, xseed, "trg")
    tl = lambda o: o.transpose(1, 2)
    ident = lambda o: o
    hooks = []
    if cls_name == "AutoVC":
        for i in range(3):
            hooks.append((f"enc_conv{i}", model.encoder.convolutions[i], lambda o: torch.relu(o).transpose(1, 2)))
            hooks.append((f"dec_conv{i}", model.decoder.convolutions[i], lambda o: torch.relu(o).transpose(1, 2)))
        hooks.append(("enc_lstm", model.encoder.lstm, ident))
        hooks.append(("dec_lstm1", model.decoder.lstm1, ident))
        hooks.append(("dec_lstm2", model.decoder.lstm2, ident))
    for i in range(4):
        hooks.append((f"post_conv{i}", model.postnet.convolutions[i], lambda o: torch.tanh(o).transpose(1, 2)))
    taps, handles = _hook_taps(model, hooks)
    with torch.no_grad():
        mel, post, codes = model(x, c_org, c_trg)
    for h in handles:                                  # taps belong to the first call only
        h.remove()
    with torch.no_grad():
        codes_only = model(mel, c_org, None)          # 4-D input path, train.py:90-92
    rec = dict(cls=cls_name, args=np.array(args), B=B, T=T, wseed=wseed, xseed=xseed,
               mel=mel.numpy(), mel_postnet=post.numpy(), codes=codes.numpy(),
               codes_of_mel=codes_only.numpy())
    for k, v in taps.items():
        rec["tap_" + k] = _sub(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "mel", tuple(mel.shape), "|mel|", float(mel.norm()), "taps", sorted(taps))


def golden_lstmdv(ref, name, B, T, wseed, xseed):
    model = ref.LstmDV().eval()
    sd = seeded_state_dict(model.state_dict(), wseed, lstm_gain=1.5)
    model.load_state_dict(sd)
    x = synthetic_mel(B, T, xseed)
    with torch.no_grad():
        e = model(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cls="LstmDV", B=B, T=T, wseed=wseed,
                        xseed=xseed, lstm_gain=1.5, emb=e.numpy())
    print(name, tuple(e.shape))


def golden_melgan(ref, name, B, T, wseed, xseed):
    model = ref.Generator(80, 32, 3).eval()
    sd = seeded_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    mel = synthetic_mel(B, T, xseed).transpose(1, 2).contiguous()
    with torch.no_grad():
        wav = model(mel)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cls="Generator", B=B, T=T, wseed=wseed,
                        xseed=xseed, wav=wav.numpy())
    print(name, tuple(wav.shape), "std", float(wav.std()), "mean", float(wav.mean()))


def golden_adjust(ref, name, cls_name, args, B, T, wseed, xseed):
    """*_Adjust models: 4-tuple return, training-style call (c_trg adjusted with x) and conversion-style call
    (isConvert=True with a target utterance), plus the codes-only path."""
    torch.manual_seed(0)
    model = getattr(ref, cls_name)(*args).eval()
    sd = seeded_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    x = synthetic_mel(B, T, xseed)
    x_target = synthetic_mel(B, T, xseed + 1000)
    c_org = synthetic_speaker(B, xseed, "org")
    c_trg = synthetic_speaker(B, xseed, "trg")
    with torch.no_grad():
        a = model(x, c_org, c_trg)
        b = model(x, c_org, c_trg, True, x_target)
        codes_only = model(x, c_org, None)
        adj = model.adjust(x, c_org)
    rec = dict(cls=cls_name, args=np.array(args), B=B, T=T, wseed=wseed, xseed=xseed, adjust_of_c_org=adj.numpy(),
               codes_only=codes_only.numpy())
    for tag, out in (("train", a), ("convert", b)):
        for key, t in zip(("c_org", "mel", "mel_postnet", "codes"), out):
            rec[f"{tag}_{key}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "mel", tuple(a[1].shape), "|mel|", float(a[1].norm()), "|mel_convert|", float(b[1].norm()))


def golden_adain(ref, name, cls_name, args, B, T, wseed, xseed):
    """AdaIN "2" variants: features-only call, self-styled conversion, conversion styled with another batch's features
    (the Evaluator's isAdain recipe, util/evaluate.py:81-82)."""
    torch.manual_seed(0)
    model = getattr(ref, cls_name)(*args).eval()
    sd = seeded_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    x = synthetic_mel(B, T, xseed)
    x_other = synthetic_mel(B, T, xseed + 1000)
    c_org = synthetic_speaker(B, xseed, "org")
    c_trg = synthetic_speaker(B, xseed, "trg")
    with torch.no_grad():
        codes, feats = model(x, c_org, None, None)
        _, feats_other = model(x_other, c_org, None, None)
        own = model(x, c_org, c_trg)
        styled = model(x, c_org, c_trg, feats_other)
    rec = dict(cls=cls_name, args=np.array(args), B=B, T=T, wseed=wseed, xseed=xseed, codes_only=codes.numpy(),
               features=np.array([[float(m), float(s)] for m, s in feats], dtype=np.float64),
               features_other=np.array([[float(m), float(s)] for m, s in feats_other], dtype=np.float64))
    for tag, out in (("own", own), ("styled", styled)):
        for key, t in zip(("mel", "mel_postnet", "codes"), out):
            rec[f"{tag}_{key}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "|post own|", float(own[1].norm()), "|post styled|", float(styled[1].norm()), rec["features"].tolist())


def golden_lstmdv_twin(name, B, T, wseed, xseed):
    """make_data/factory/LstmDV.py (the classifier twin the Evaluator calls): (predictions, d_vec)."""
    import importlib.util
    path = os.path.join(ref_import.REFERENCE_ROOT, "make_data", "factory", "LstmDV.py")
    spec = importlib.util.spec_from_file_location("_ref_make_data_lstmdv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = mod.LstmDV().eval()
    sd = seeded_state_dict(model.state_dict(), wseed, lstm_gain=1.5)
    model.load_state_dict(sd)
    x = synthetic_mel(B, T, xseed)
    with torch.no_grad():
        pred, dv = model(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cls="make_data.LstmDV", B=B, T=T, wseed=wseed, xseed=xseed,
                        lstm_gain=1.5, predictions=pred.numpy(), d_vec=dv.numpy())
    print(name, tuple(pred.shape), tuple(dv.shape))


def golden_audio2mel(ref, name, B, L, seed):
    """The reference's own Audio2Mel (melgan/modules.py:26-69) run through two shims, because as written it cannot
    run on this container's stack: (1) ``librosa.filters.mel`` (librosa absent) is supplied by the restated filter bank
    ``oracle.audio2mel.librosa_mel`` -- itself cross-checked against transformers' librosa-compatible implementation;
    (2) ``torch.stft`` is wrapped to pass ``return_complex=True`` and hand back ``view_as_real`` (the pre-2.0 return
    convention the reference unbinds, :65).  Everything else -- padding, framing, window, magnitude, mel projection,
    log10 clamp -- is the reference's code."""
    import importlib
    from .audio2mel import librosa_mel
    mods = importlib.import_module("melgan.modules")
    mods.librosa_mel_fn = lambda sr, n_fft, n_mels, fmin, fmax: librosa_mel(sr, n_fft, n_mels, fmin, fmax).numpy()
    real_stft = torch.stft

    def stft_pre2(*a, **k):
        k["return_complex"] = True
        return torch.view_as_real(real_stft(*a, **k))
    model = mods.Audio2Mel().eval()
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L) / 22050.0
    tones = sum(a * torch.sin(2 * torch.pi * f * t + ph) for a, f, ph in ((0.4, 220.0, 0.1), (0.2, 1760.0, 1.0),
                                                                            (0.1, 5200.0, 2.0)))
    audio = (tones.unsqueeze(0) + 0.05 * torch.randn(B, L, generator=g)).unsqueeze(1)
    torch.stft = stft_pre2
    try:
        with torch.no_grad():
            mel = model(audio)
    finally:
        torch.stft = real_stft
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cls="Audio2Mel", B=B, L=L, seed=seed,
                        audio=audio.numpy().astype(np.float32), mel=mel.numpy())
    print(name, tuple(mel.shape), "range", float(mel.min()), float(mel.max()))


def golden_evaluator(ref, name, T_src, T_trg, wseed, xseed):
    """The reference's Evaluator (util/evaluate.py:36-98) on synthetic .npy utterances shorter than ``len_crop``:
    ``get_trans_mel(..., isPlay=True)`` = crop_mel (zero-pad to 176) -> AutoVC(44,256,512,22) -> trim."""
    import importlib
    import tempfile
    import types
    ev = importlib.import_module("util.evaluate")
    model = ref.AutoVC(44, 256, 512, 22).eval()
    sd = seeded_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd)
    src = synthetic_mel(1, T_src, xseed)[0].numpy()
    trg = synthetic_mel(1, T_trg, xseed + 1)[0].numpy()
    e_src = synthetic_speaker(1, xseed, "org")[0].numpy()
    e_trg = synthetic_speaker(1, xseed, "trg")[0].numpy()
    with tempfile.TemporaryDirectory() as root:
        np.save(os.path.join(root, "a.npy"), src)
        np.save(os.path.join(root, "b.npy"), trg)
        cfg = types.SimpleNamespace(root=root, num_speaker=2, batch_size=1, max_uttr_idx=4, erroment_num=1, len_crop=176,
                                    device="cpu", all_speaker=["p1", "p2"], embedder=None,
                                    metadata=[["p1", e_src, "a.npy"], ["p2", e_trg, "b.npy"]])
        E = ev.Evaluator.__new__(ev.Evaluator)                    # __init__ loads a vocoder checkpoint from disk (:23)
        for k, v in vars(cfg).items():
            setattr(E, k, v)
        E.metadata = E.build_metadata(cfg.metadata)
        with torch.no_grad():
            ms, mt, trans_play = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=True)
            _, _, trans_full = E.get_trans_mel(model, 0, 1, 2, False, False, isPlay=False)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cls="Evaluator", T_src=T_src, T_trg=T_trg, wseed=wseed,
                        xseed=xseed, src=src, trg=trg, e_src=e_src, e_trg=e_trg, mel_source=ms.numpy(),
                        mel_target=mt.numpy(), mel_trans_play=trans_play.numpy(), mel_trans_full=trans_full.numpy())
    print(name, tuple(ms.shape), tuple(mt.shape), tuple(trans_play.shape), tuple(trans_full.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden_autovc(ref, "autovc_A_b2_t128", "AutoVC", (32, 256, 512, 32), 2, 128, 0, 1234)
    golden_autovc(ref, "autovc_A_b3_t64", "AutoVC", (32, 256, 512, 32), 3, 64, 1, 77)
    golden_autovc(ref, "autovc_R_b2_t176", "AutoVC", (44, 256, 512, 22), 2, 176, 2, 99)
    golden_lstmdv(ref, "lstmdv_b2_t100", 2, 100, 3, 5)
    golden_melgan(ref, "melgan_b1_t40", 1, 40, 4, 6)
    golden_melgan(ref, "melgan_b2_t17", 2, 17, 5, 7)
    golden_autovc(ref, "metapool_b1_t176", "MetaPool", (44, 256, 512, 22), 1, 176, 6, 8)
    golden_autovc(ref, "metaconv_b1_t176", "MetaConv", (44, 256, 512, 22), 1, 176, 7, 9)
    golden_adjust(ref, "autovc_adjust_b2_t64", "AutoVC_Adjust", (32, 256, 512, 32), 2, 64, 10, 11)
    golden_adjust(ref, "metapool_adjust_b1_t176", "MetaPool_Adjust", (44, 256, 512, 22), 1, 176, 12, 13)
    golden_adjust(ref, "metaconv_adjust_b1_t176", "MetaConv_Adjust", (44, 256, 512, 22), 1, 176, 14, 15)
    golden_adain(ref, "autovc2_b2_t64", "AutoVC2", (32, 256, 512, 32), 2, 64, 16, 17)
    golden_adain(ref, "metapool2_b1_t176", "MetaPool2", (44, 256, 512, 22), 1, 176, 18, 19)
    golden_adain(ref, "metaconv2_b1_t176", "MetaConv2", (44, 256, 512, 22), 1, 176, 20, 21)
    round2(ref)


def round2(ref):
    golden_lstmdv_twin("lstmdv_twin_b2_t100", 2, 100, 22, 23)
    golden_audio2mel(ref, "audio2mel_b2_l5120", 2, 5120, 24)
    golden_audio2mel(ref, "audio2mel_b1_l2381", 1, 2381, 25)
    golden_evaluator(ref, "evaluator_autovcR_t100_t150", 100, 150, 26, 27)


if __name__ == "__main__":
    if "--round2" in sys.argv:          # only the fixtures added in round 2 (the earlier ones are unchanged)
        os.makedirs(OUT, exist_ok=True)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        sys.exit(round2(ref_import.load()))
    sys.exit(main())
