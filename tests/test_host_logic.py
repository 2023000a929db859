"""CPU tests of the drop-in models' host logic: the real model classes run end to end with the C-ABI entry points
replaced by CPU stand-ins that mirror the kernels' indexing (tests/emulate.py), and are compared with the oracle.
This validates weight folding/packing, buffer formats, halo rows, poly-phase up-sampling and gate interleave
without a GPU; the arithmetic of the real kernels is checked by the `-m gpu` tests."""
import warnings

import pytest
import torch

from oracle import rel_l2, templates
from oracle.autovc import autovc_forward
from oracle.lstmdv import lstmdv_forward
from oracle.melgan import melgan_forward
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker
from tests.emulate import install_cpu_kernels

warnings.filterwarnings("ignore", category=FutureWarning)


@pytest.fixture
def cpu_kernels(monkeypatch):
    install_cpu_kernels(monkeypatch)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tf32", 3e-3), ("bf16", 2e-2), ("fp16x2", 1e-3)])
def test_autovc_host_logic(cpu_kernels, precision, tol):
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    sd = seeded_state_dict(templates.autovc_template(*args), 0)
    B, T = 2, 64
    x, c_org, c_trg = synthetic_mel(B, T, 5), synthetic_speaker(B, 5, "org"), synthetic_speaker(B, 5, "trg")
    ref = autovc_forward(sd, x, c_org, c_trg, 32, 32)
    m = AutoVC(*args)
    m.load_state_dict(sd)
    m.eval()
    m.precision = precision
    mel, post, codes = m(x, c_org, c_trg)
    assert rel_l2(mel, ref[0]) < tol and rel_l2(post, ref[1]) < tol and rel_l2(codes, ref[2]) < tol


def test_convert_batches_order_and_unpaired_fallback(cpu_kernels, monkeypatch):
    """pipeline.convert_batches returns one result per batch in the order given; batches it may not pair (<= 64 utterances:
    the weight-stationary kernels fill the device; any batch when the LSTM kernels run without the cooperative launch
    attribute) simply run one after the other -- on the CPU stand-ins that is the whole path, no CUDA stream is touched."""
    from autoformer_b200 import ops, pipeline
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    m = AutoVC(*args)
    m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 0))
    m.eval()
    monkeypatch.setenv("AVC_LSTM_NO_COOP", "1")
    batches = [(synthetic_mel(B, T, 3 + i), synthetic_speaker(B, i, "org"), synthetic_speaker(B, i, "trg"))
               for i, (B, T) in enumerate([(3, 32), (70, 32), (2, 64)])]
    calls = []
    sums = pipeline.convert_batches(m, batches, reduce=lambda out: calls.append(out[1].shape[0]) or out[1].double().sum())
    assert calls == [3, 70, 2] and ops.LSTM_CTA_BUDGET is None
    for b, s in zip(batches, sums):
        assert float(s) == float(m(*b)[1].double().sum())
    outs = pipeline.convert_batches(m, batches[:1])
    assert len(outs) == 1 and len(outs[0]) == 3 and outs[0][1].shape == (3, 1, 32, 80)


def test_autovc_config_r_host_logic(cpu_kernels):
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (44, 256, 512, 22)
    sd = seeded_state_dict(templates.autovc_template(*args), 2)
    x, c_org, c_trg = synthetic_mel(1, 44, 6), synthetic_speaker(1, 6, "org"), synthetic_speaker(1, 6, "trg")
    ref = autovc_forward(sd, x, c_org, c_trg, 44, 22)
    m = AutoVC(*args)
    m.load_state_dict(sd)
    m.eval()
    out = m(x, c_org, c_trg)
    for a, b in zip(out, ref):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16x2", 5e-4)])
def test_lstmdv_host_logic(cpu_kernels, precision, tol):
    from autoformer_b200.factory.LstmDV import LstmDV
    sd = seeded_state_dict(templates.lstmdv_template(), 3, lstm_gain=1.5)
    x = synthetic_mel(2, 20, 5)
    ref = lstmdv_forward(sd, x)
    m = LstmDV()
    m.load_state_dict(sd)
    m.precision = precision
    e = m(x)
    assert e.shape == (2, 256) and rel_l2(e, ref) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tf32", 3e-3), ("bf16", 3e-2), ("fp16x2", 2e-3), ("fp16s", 1e-3)])
@pytest.mark.parametrize("B,T", [(1, 12), (2, 17)])
def test_melgan_host_logic(cpu_kernels, precision, tol, B, T):
    from autoformer_b200.melgan.modules import Generator
    sd = seeded_state_dict(templates.melgan_template(), 4)
    mel = synthetic_mel(B, T, 6).transpose(1, 2).contiguous()
    rt = {}
    ref = melgan_forward(sd, mel, taps=rt)
    g = Generator(80, 32, 3)
    g.load_state_dict(sd)
    g.precision = precision
    g.collect_taps = True
    wav = g(mel)
    assert wav.shape == ref.shape == (B, 1, 256 * T)
    for k in ("up0", "stage0", "up1", "stage1", "up2", "stage2", "up3", "stage3"):
        assert rel_l2(g.taps[k], rt[k].transpose(1, 2)) < tol, k
    assert rel_l2(wav, ref) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16x2", 1.5e-3)])
@pytest.mark.parametrize("kind", ["pool", "conv"])
def test_meta_host_logic(cpu_kernels, kind, precision, tol):
    """MetaPool / MetaConv wiring (GroupNorm, pooling mixer, patchify, LN+transpose, padded token GEMMs, the
    channels=time decoder) against the oracle, stage by stage."""
    from autoformer_b200.factory.MetaConv import MetaConv
    from autoformer_b200.factory.MetaPool import MetaPool
    from oracle.meta import meta_forward
    args = (44, 256, 512, 22)
    sd = seeded_state_dict(templates.meta_template(kind, *args), 6)
    x, c_org, c_trg = synthetic_mel(1, 176, 8), synthetic_speaker(1, 8, "org"), synthetic_speaker(1, 8, "trg")
    rt = {}
    ref = meta_forward(sd, x, c_org, c_trg, 44, 22, kind, taps=rt)
    m = (MetaPool if kind == "pool" else MetaConv)(*args)
    m.load_state_dict(sd)
    m.eval()
    m.precision = precision
    m.collect_taps = True
    mel, post, codes = m(x, c_org, c_trg)
    cl = lambda t: t.transpose(1, 2)                       # oracle taps are channels-first
    assert rel_l2(m.taps["enc_embed"], cl(rt["enc_embed"])) < tol
    for i in range(3):
        assert rel_l2(m.taps[f"encoder.metablock.{i}"], cl(rt[f"encoder.metablock.{i}"])) < tol, i
    assert rel_l2(m.taps["enc_out"], rt["enc_out"]) < tol
    assert rel_l2(m.taps["decoder.metablock.0"], cl(rt["decoder.metablock.0"])) < tol
    assert rel_l2(m.taps["dec_conv2"], cl(rt["dec_conv2"])) < tol
    assert rel_l2(codes, ref[2]) < tol and rel_l2(mel, ref[0]) < tol and rel_l2(post, ref[1]) < tol
    assert rel_l2(m(x, c_org, None), ref[2]) < tol
    with pytest.raises(RuntimeError):
        m(synthetic_mel(1, 128, 8), c_org, c_trg)          # the reference's hard-wired T = 176


def _adjust_model(kind, args, sd):
    from autoformer_b200.factory.AutoVC_Adjust import AutoVC_Adjust
    from autoformer_b200.factory.MetaConv_Adjust import MetaConv_Adjust
    from autoformer_b200.factory.MetaPool_Adjust import MetaPool as MetaPool_Adjust
    m = {None: AutoVC_Adjust, "pool": MetaPool_Adjust, "conv": MetaConv_Adjust}[kind](*args)
    m.load_state_dict(sd)                      # the reference's key names, incl. adjust.*
    return m.eval()


@pytest.mark.parametrize("name,kind", [("autovc_adjust_b2_t64", None), ("metapool_adjust_b1_t176", "pool"),
                                       ("metaconv_adjust_b1_t176", "conv")])
def test_adjust_models_host_logic(cpu_kernels, name, kind):
    """*_Adjust drop-ins: 4-tuple return, isConvert / x_target routing, codes-only path, checked
    against the reference's own outputs (golden) through the CPU stand-ins."""
    from tests.test_oracle_golden import adjust_case
    g, args, sd, i, _ = adjust_case(name, kind)
    m = _adjust_model(kind, args, sd)
    G = lambda k: torch.from_numpy(g[k])
    out = m(i["x"], i["c_org"], i["c_trg"])
    conv = m(i["x"], i["c_org"], i["c_trg"], True, i["x_target"])
    for tag, o in (("train", out), ("convert", conv)):
        assert len(o) == 4
        for key, t in zip(("c_org", "mel", "mel_postnet", "codes"), o):
            assert t.shape == g[f"{tag}_{key}"].shape and rel_l2(t, G(f"{tag}_{key}")) < 2e-4, (tag, key)
    assert rel_l2(m(i["x"], i["c_org"], None), G("codes_only")) < 2e-4
    assert rel_l2(m.adjust(i["x"], i["c_org"]), G("adjust_of_c_org")) < 2e-4


def _adain_model(kind, args, sd):
    from autoformer_b200.factory.AutoVC2 import AutoVC2
    from autoformer_b200.factory.MetaConv2 import MetaConv2
    from autoformer_b200.factory.MetaPool2 import MetaPool2
    m = {None: AutoVC2, "pool": MetaPool2, "conv": MetaConv2}[kind](*args)
    m.load_state_dict(sd)                      # the reference's key names, incl. feature_pre_extract / feature_last_combine
    return m.eval()


@pytest.mark.parametrize("name,kind", [("autovc2_b2_t64", None), ("metapool2_b1_t176", "pool"),
                                       ("metaconv2_b1_t176", "conv")])
def test_adain_models_host_logic(cpu_kernels, name, kind):
    """AdaIN "2" drop-ins through the CPU stand-ins against the reference's own outputs."""
    from tests.test_oracle_golden import adain_case, check_adain_outputs
    g, args, sd, i, _ = adain_case(name, kind)
    m = _adain_model(kind, args, sd)
    check_adain_outputs(g, lambda xk, conv, tf: m(i[xk], i["c_org"], i["c_trg"] if conv else None, tf), 2e-4, 1e-4)
    with pytest.raises(AttributeError):
        m(i["x"], i["c_org"], None, [[0.0, 1.0]] * 3)      # the reference dereferences c_trg here


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
@pytest.mark.parametrize("B,H", [(2, 512), (32, 1024), (64, 768)])
def test_lstm_layer_small_batch_takes_weight_stationary_packing(cpu_kernels, B, H, precision):
    """B <= 64 in split precision: dense projection + avc_lstm_seq_ws with the gates of a unit adjacent
    (p = 128 (u//32) + 4 (u%32) + gate); the packing of W_ih, the biases and W_hh must agree with that order."""
    from autoformer_b200 import layers, ops, packing
    from oracle.layers import lstm_explicit
    assert ops.ws_supported(B, H, "fp32") and not ops.ws_supported(65, H, "fp32") and not ops.ws_supported(B, H, "bf16")
    assert ops.ws_supported(64, 1024, "fp32") and not ops.ws_supported(32, 2048, "fp32")   # 256 CTAs: more than one wave
    perm = packing.gate_permutation(H, packing.WS_GROUP)
    assert sorted(perm.tolist()) == list(range(4 * H))
    assert perm[:8].tolist() == [0, H, 2 * H, 3 * H, 1, H + 1, 2 * H + 1, 3 * H + 1] and perm[128].item() == 32
    torch.manual_seed(B)
    T, I = 4, 80
    k = 1.0 / H ** 0.5
    w_ih, w_hh = (torch.rand(4 * H, I) * 2 - 1) * k, (torch.rand(4 * H, H) * 2 - 1) * k
    b_ih, b_hh = (torch.rand(4 * H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k
    x = torch.randn(B, T, I)
    ref = lstm_explicit(x.double(), w_ih.double(), w_hh.double(), b_ih.double(), b_hh.double())
    layer = layers.LstmLayer(w_ih, w_hh, b_ih, b_hh, precision)
    f32 = torch.full((B, T, H), float("nan"))
    last = torch.full((B, H), float("nan"))
    layer(packing.to_act(x, precision), B, T, hseq_f32=f32, h_last=last, persistent=True)
    assert packing.WS_GROUP in layer._packs and not layer._fused_packs
    tol = 1e-4 if precision == "fp32" else 1e-3
    assert rel_l2(f32, ref) < tol and rel_l2(last, ref[:, -1]) < tol


@pytest.mark.parametrize("B,H,L", [(2, 768, 3), (40, 512, 2), (64, 256, 4)])
def test_lstm_stack_wavefront_packing_and_fallback(cpu_kernels, B, H, L):
    """Small-batch fp16x2 stacks take the wavefront kernel (avc_lstm_stack_ws): one fp16 term of every W_hh and of W_ih
    above the first layer, biases in the same packed gate order; `wavefront=False`, other precisions and unsupported shapes fall back to
    the layer-by-layer kernels with the same result."""
    from autoformer_b200 import layers, ops, packing
    from oracle.layers import lstm_explicit
    assert ops.stack_supported(B, H, L, "fp16x2") and not ops.stack_supported(B, H, L, "fp32")
    assert not ops.stack_supported(65, H, L, "fp16x2") and not ops.stack_supported(B, H, 1, "fp16x2")
    assert ops.stack_supported(64, 768, 3, "fp16x2") and not ops.stack_supported(64, 896, 2, "fp16x2")    # TMEM columns
    assert not ops.stack_supported(64, 768, 4, "fp16x2")                                                  # 192 CTAs
    torch.manual_seed(B)
    T, I = 5, 80
    k = 1.0 / H ** 0.5
    ws = []
    x = torch.randn(B, T, I)
    ref = x.double()
    for l in range(L):
        w_ih = (torch.rand(4 * H, I if l == 0 else H) * 2 - 1) * k
        w_hh, b_ih, b_hh = (torch.rand(4 * H, H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k
        ws.append((w_ih, w_hh, b_ih, b_hh))
        ref = lstm_explicit(ref, w_ih.double(), w_hh.double(), b_ih.double(), b_hh.double())
    calls = []
    real = ops.lstm_stack_ws
    ops.lstm_stack_ws = lambda *a, **kw: calls.append(1) or real(*a, **kw)
    try:
        outs = {}
        for name, precision, wavefront in (("stack", "fp16x2", True), ("layers", "fp16x2", False), ("fp32", "fp32", True)):
            stack = layers.LstmStack([layers.LstmLayer(*w, precision) for w in ws], precision, wavefront)
            last = torch.full((B, H), float("nan"))
            n0 = len(calls)
            stack.last_hidden(packing.to_act(x, precision), B, T, last, persistent=True)
            assert (len(calls) - n0 == 1) == (name == "stack")
            outs[name] = last
        # persistent=False (one launch per frame) never takes the wavefront
        stack = layers.LstmStack([layers.LstmLayer(*w, "fp16x2") for w in ws], "fp16x2", True)
        n0 = len(calls)
        stack.last_hidden(packing.to_act(x, "fp16x2"), B, T, torch.empty(B, H), persistent=False)
        assert len(calls) == n0
    finally:
        ops.lstm_stack_ws = real
    assert rel_l2(outs["fp32"], ref[:, -1]) < 1e-4
    assert rel_l2(outs["layers"], ref[:, -1]) < 1e-3 and rel_l2(outs["stack"], ref[:, -1]) < 1e-3
    assert rel_l2(outs["stack"], outs["layers"]) < 1e-3


def test_lstm_layer_sub_batches_when_persistent_grid_exceeds_one_wave(cpu_kernels):
    """B = 600 at H = 1024 needs 192 CTAs > 148 SMs: the layer must split the batch (utterances are independent) and
    still fill every output the caller asked for."""
    from autoformer_b200 import layers, ops, packing
    from oracle.layers import lstm_explicit
    assert ops.persistent_batch_cap(1024) == 512 and ops.persistent_batch_cap(512) == 1024
    torch.manual_seed(0)
    B, T, I, H = 600, 3, 64, 1024
    k = 1.0 / H ** 0.5
    w_ih, w_hh = (torch.rand(4 * H, I) * 2 - 1) * k, (torch.rand(4 * H, H) * 2 - 1) * k
    b_ih, b_hh = (torch.rand(4 * H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k
    x = torch.randn(B, T, I)
    ref = lstm_explicit(x.double(), w_ih.double(), w_hh.double(), b_ih.double(), b_hh.double())
    layer = layers.LstmLayer(w_ih, w_hh, b_ih, b_hh, "fp32")
    f32 = torch.full((B, T, H), float("nan"))
    last = torch.full((B, H), float("nan"))
    hseq = layer(packing.to_act(x, "fp32"), B, T, hseq_f32=f32, h_last=last, persistent=True)
    assert rel_l2(f32, ref) < 1e-4 and rel_l2(last, ref[:, -1]) < 1e-4
    assert rel_l2(packing.act_to_float(hseq, "fp32"), ref) < 1e-4
