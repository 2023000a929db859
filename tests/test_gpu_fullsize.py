"""Full-size GPU parity for BASELINE configs 3 and 4 (the sizes bench.py measures), the 2-rank GPU equality test of the
sharded run, and the fp16x2 dynamic-range stress case.  Modelled on test_autovc_fp16x2_full_bench_config: the big batch
runs once, the oracle checks a handful of picked utterances (tile-boundary rows included), a shuffled input must fail."""
import os
import subprocess
import sys
import warnings

import pytest
import torch

from oracle import rel_l2, templates
from oracle.autovc import autovc_forward
from oracle.lstmdv import lstmdv_forward
from oracle.melgan import melgan_forward
from oracle.meta import meta_forward
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore", category=FutureWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
META_ARGS = (44, 256, 512, 22)


def _load(cls, args, sd, precision):
    m = cls(*args)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = precision
    return m


# ------------------------------------------------------------------------------------------------ config 3
@pytest.mark.parametrize("kind", ["pool", "conv"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16x2", 1e-3)])
def test_meta_full_size_512x176(kind, precision, tol):
    """MetaPool / MetaConv at 512 x 176 (multi-wave persistent GEMM grids, every tile shape of the mixers): sampled
    utterances against the oracle, a slice of the batch against the same utterances alone, negative control."""
    from autoformer_b200.factory.MetaConv import MetaConv
    from autoformer_b200.factory.MetaPool import MetaPool
    sd = seeded_state_dict(templates.meta_template(kind, *META_ARGS), 30 + (kind == "conv"))
    B, T = 512, 176
    x, c_org, c_trg = synthetic_mel(B, T, 41), synthetic_speaker(B, 41, "org"), synthetic_speaker(B, 41, "trg")
    m = _load(MetaPool if kind == "pool" else MetaConv, META_ARGS, sd, precision)
    big = m(x.cuda(), c_org.cuda(), c_trg.cuda())
    idx = [0, 255, 256, 511]
    ref = meta_forward(sd, x[idx], c_org[idx], c_trg[idx], META_ARGS[0], META_ARGS[3], kind)
    errs = [rel_l2(u[idx], v) for u, v in zip(big, ref)]
    print(f"Meta{kind} {precision} rel-L2 (mel, mel_postnet, codes) at 512 x 176:", errs)
    assert all(e < tol for e in errs), errs
    small = m(x[idx].cuda(), c_org[idx].cuda(), c_trg[idx].cuda())
    for u, v in zip(small, big):
        assert rel_l2(u, v[idx]) < tol                                    # other tile shapes, same utterances
    wrong = m(x[idx].flip(0).cuda(), c_org[idx].cuda(), c_trg[idx].cuda())
    assert rel_l2(wrong[1], ref[1]) > 1e-2


# ------------------------------------------------------------------------------------------------ config 4
def test_melgan_full_size_b32_t1000():
    """MelGAN at B = 32, T = 1000 (256,000 samples per utterance, thousands of tiles): waveform and stage taps of two
    utterances against the oracle, batch independence, negative control."""
    from autoformer_b200.melgan.modules import Generator
    sd = seeded_state_dict(templates.melgan_template(), 4)
    B, T = 32, 1000
    mel = synthetic_mel(B, T, 51).transpose(1, 2).contiguous()
    gen = _load(Generator, (80, 32, 3), sd, "fp32")
    gen.collect_taps = True
    wav = gen(mel.cuda())
    assert wav.shape == (B, 1, 256 * T) and bool(torch.isfinite(wav).all())
    got_taps = {k: v[[0, 31]].cpu() for k, v in gen.taps.items()}
    gen.collect_taps = False
    idx = [0, 31]
    taps = {}
    ref = melgan_forward(sd, mel[idx], taps=taps)
    err = rel_l2(wav[idx], ref)
    print("MelGAN fp32 waveform rel-L2 at B=32, T=1000:", err)
    assert err < 2e-4, err
    for name in ("up0", "up1", "up2", "up3", "stage0", "stage1", "stage2", "stage3"):
        assert name in got_taps and name in taps, (name, sorted(got_taps), sorted(taps))
        e = rel_l2(got_taps[name], taps[name].transpose(1, 2))
        assert e < 3e-4, (name, e)
    alone = gen(mel[idx].cuda())
    assert rel_l2(alone, wav[idx]) < 2e-5
    assert rel_l2(gen(mel[idx].flip(2).cuda()), ref) > 1e-2


def test_lstmdv_full_size_b64_t1000():
    from autoformer_b200.factory.LstmDV import LstmDV
    sd = seeded_state_dict(templates.lstmdv_template(), 3, lstm_gain=1.5)
    B, T = 64, 1000
    x = synthetic_mel(B, T, 52)
    idx = [0, 31, 32, 63]
    ref = lstmdv_forward(sd, x[idx])
    for precision, tol in (("fp32", 1e-4), ("fp16x2", 1e-3)):
        dv = _load(LstmDV, (), sd, precision)
        e = dv(x.cuda())
        err = rel_l2(e[idx], ref)
        print(f"LstmDV {precision} rel-L2 at B=64, T=1000:", err)
        assert err < tol, err
        assert rel_l2(dv(x[idx].cuda()), e[idx]) < tol
    assert rel_l2(dv(x[idx].flip(1).cuda()), ref) > 1e-2


def test_convert_and_vocode_full_size_b32_t1000():
    """The whole config-4 pipeline at B = 32, T = 1000 with the 1000 -> 1024 pad / trim, in the precision mix bench.py
    measures (embedder / AutoVC fp16x2, MelGAN split): embeddings, converted mel and waveform of two utterances against
    the oracle pipeline; the vocoder alone on the oracle's mel (isolates its own error from the upstream one)."""
    from autoformer_b200 import pipeline
    from autoformer_b200.factory.AutoVC import AutoVC
    from autoformer_b200.factory.LstmDV import LstmDV
    from autoformer_b200.melgan.modules import Generator
    args = (32, 256, 512, 32)
    B, T = 32, 1000
    sd_dv = seeded_state_dict(templates.lstmdv_template(), 3, lstm_gain=1.5)
    sd_vc = seeded_state_dict(templates.autovc_template(*args), 0)
    sd_g = seeded_state_dict(templates.melgan_template(), 4)
    src, trg = synthetic_mel(B, T, 61), synthetic_mel(B, T, 62)
    dv, vc, gen = _load(LstmDV, (), sd_dv, "fp16x2"), _load(AutoVC, args, sd_vc, "fp16x2"), \
        _load(Generator, (80, 32, 3), sd_g, "fp32")
    mel, wav, eo, et = pipeline.convert_and_vocode(dv, vc, gen, src.cuda(), trg.cuda())
    assert mel.shape == (B, T, 80) and wav.shape == (B, 256 * T)
    idx = [0, 31]
    r_eo, r_et = lstmdv_forward(sd_dv, src[idx]), lstmdv_forward(sd_dv, trg[idx])
    xpad = torch.nn.functional.pad(src[idx], (0, 0, 0, 24))                  # util/evaluate.py:38-43
    r_mel = autovc_forward(sd_vc, xpad, r_eo, r_et, 32, 32)[1].squeeze(1)[:, :T]     # :85-92
    r_wav = melgan_forward(sd_g, r_mel.transpose(1, 2)).squeeze(1)
    errs = dict(emb_org=rel_l2(eo[idx], r_eo), emb_trg=rel_l2(et[idx], r_et), mel=rel_l2(mel[idx], r_mel),
                wav=rel_l2(wav[idx], r_wav))
    print("config-4 pipeline rel-L2 at B=32, T=1000:", errs)
    assert errs["emb_org"] < 1e-3 and errs["emb_trg"] < 1e-3 and errs["mel"] < 1e-3, errs
    # the waveform error includes the upstream mel error amplified by the vocoder; its own error is gated below
    assert errs["wav"] < 3e-3, errs
    alone = gen(r_mel.transpose(1, 2).contiguous().cuda()).squeeze(1)
    assert rel_l2(alone, r_wav) < 2e-4
    # padding recipe: converting WITHOUT the pad must refuse (1000 is not a multiple of 32), as the reference does
    with pytest.raises(IndexError):
        vc(src[idx].cuda(), eo[idx], et[idx])


# ------------------------------------------------------------------------------------------------ fp16 dynamic range
@pytest.mark.parametrize("scale", [4.0, 8.0])
def test_fp16x2_scaled_weights_stay_finite_and_accurate(scale):
    """fp16 activations saturate at +-65504 (csrc/avc_pipe.cuh: sat_f16) instead of overflowing to inf.  Random-init
    parity never comes near that range, so scale the encoder's three conv layers (BatchNorm is folded with its running
    statistics, so the activations grow by scale^3 through the ReLU stack: hundreds of times the random-init range) and
    require: every output finite, the overflow probe reports head-room, and the stage taps of the scaled stack still
    track the fp32-grade (split) run of the same weights -- i.e. nothing saturated silently.  (The FINAL outputs are not
    compared: the LSTMs behind the scaled stack run deep in gate saturation, where any two arithmetics diverge.)"""
    from autoformer_b200 import packing
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    sd = seeded_state_dict(templates.autovc_template(*args), 5)
    for k in sd:
        if k.startswith("encoder.convolutions") and k.endswith("conv.weight"):
            sd[k] = sd[k] * scale
    B, T = 8, 64
    x, c_org, c_trg = synthetic_mel(B, T, 71).cuda(), synthetic_speaker(B, 71, "org").cuda(), synthetic_speaker(B, 71, "trg").cuda()
    hi = _load(AutoVC, args, sd, "fp32")
    lo = _load(AutoVC, args, sd, "fp16x2")
    hi.collect_taps = lo.collect_taps = True
    a, b = hi(x, c_org, c_trg), lo(x, c_org, c_trg)
    peak = max(float(v.abs().max()) for v in lo.taps.values())
    errs = {k: rel_l2(lo.taps[k], hi.taps[k]) for k in ("enc_conv0", "enc_conv1", "enc_conv2")}
    print(f"scale {scale}: largest activation {peak:.1f}; fp16x2 vs split on the scaled stack", errs)
    for u in b:
        assert bool(torch.isfinite(u).all())
    assert all(e < 2e-3 for e in errs.values()), errs
    assert packing.fp16_overflow_margin(lo.taps) > 1.0, "an activation reached the fp16 range: saturation would be silent"
    assert peak > 50.0                                         # the stress really left the random-init range


def test_fp16x2_overflow_is_flagged():
    """A weight scale that pushes an activation past the fp16 range must be visible: the store saturates (finite), and
    ``packing.fp16_overflow_margin`` on the stage taps reports a margin <= 1 so callers can refuse the precision."""
    from autoformer_b200 import packing
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    sd = seeded_state_dict(templates.autovc_template(*args), 5)
    for k in sd:
        if k.startswith("decoder.convolutions") and k.endswith("conv.weight"):
            sd[k] = sd[k] * 400.0
    B, T = 2, 64
    x, c_org, c_trg = synthetic_mel(B, T, 72).cuda(), synthetic_speaker(B, 72, "org").cuda(), synthetic_speaker(B, 72, "trg").cuda()
    lo = _load(AutoVC, args, sd, "fp16x2")
    lo.collect_taps = True
    out = lo(x, c_org, c_trg)
    assert all(bool(torch.isfinite(v).all()) for v in lo.taps.values())      # saturated, never inf / NaN
    assert bool(torch.isfinite(out[1]).all())
    assert packing.fp16_overflow_margin(lo.taps) <= 1.0


# ------------------------------------------------------------------------------------------------ multi-GPU equality
WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["AVC_ROOT"])
from autoformer_b200 import sharding
from autoformer_b200.factory.AutoVC import AutoVC
from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
args = (32, 256, 512, 32)
m = AutoVC(*args); m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 21)); m = m.cuda().eval()
m.precision = "fp16x2"
lengths = [64 if i % 3 else 96 for i in range(40)]
def convert(batches):
    out = {}
    for T, ids in batches:
        x = torch.stack([synthetic_mel(1, T, 1000 + i)[0] for i in ids]).cuda()
        co = torch.stack([synthetic_speaker(1, 1000 + i, "org")[0] for i in ids]).cuda()
        ct = torch.stack([synthetic_speaker(1, 1000 + i, "trg")[0] for i in ids]).cuda()
        post = m(x, co, ct)[1]
        for j, i in enumerate(ids):
            out[i] = post[j, 0, :64].clone()             # equal-shape rows for the gather
    return out
plan = sharding.plan(lengths, world, max_batch=16)
mine = convert(plan[rank])
ids = sorted(mine)
full = sharding.gather_outputs(torch.stack([mine[i] for i in ids]), len(lengths), ids, dst=0)
if rank == 0:
    # the same batches, all on ONE GPU: every utterance must come out bit-identical whatever rank converted it
    single = {}
    for r in range(world):
        single.update(convert(plan[r]))
    want = torch.stack([single[i] for i in range(len(lengths))])
    assert torch.equal(full, want), float((full - want).abs().max())
    print("EQUAL", full.shape, float(full.double().sum()))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_gpu_outputs_equal_single_gpu(tmp_path):
    """rank r's outputs of the sharded run bit-equal the same batches converted on one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, AVC_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "EQUAL" in out.stdout


@pytest.mark.gpu
def test_convert_batches_pairs_underfilled_batches_bit_equal():
    """pipeline.convert_batches runs batches of <= 256 utterances two at a time on two streams with the persistent LSTM
    grids capped at half the SMs; every output must be bit-equal to the same batch converted alone, whatever the mix of
    batch sizes (full, under-filled, odd count of under-filled ones, weight-stationary-sized)."""
    from autoformer_b200 import ops, pipeline
    from autoformer_b200.factory.AutoVC import AutoVC
    args = (32, 256, 512, 32)
    m = AutoVC(*args)
    m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 0))
    m = m.cuda().eval()
    m.precision = "fp16x2"
    shapes = [(300, 64), (200, 96), (150, 64), (130, 128), (70, 96), (40, 64), (256, 32)]
    batches = [(synthetic_mel(B, T, 7 + i).cuda(), synthetic_speaker(B, i, "org").cuda(), synthetic_speaker(B, i, "trg").cuda())
               for i, (B, T) in enumerate(shapes)]
    alone = [tuple(t.clone() for t in m(*b)) for b in batches]
    assert ops.LSTM_CTA_BUDGET is None
    for _ in range(2):
        got = pipeline.convert_batches(m, batches)
        torch.cuda.synchronize()
        assert ops.LSTM_CTA_BUDGET is None
        for (B, T), a, g in zip(shapes, alone, got):
            for x, y in zip(a, g):
                assert torch.equal(x, y), (B, T)
    sums = pipeline.convert_batches(m, batches, reduce=lambda out: out[1].double().sum())
    assert all(float(s) == float(a[1].double().sum()) for s, a in zip(sums, alone))
