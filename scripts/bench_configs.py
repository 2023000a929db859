"""Throughput of the other BASELINE configs on one GPU: config 3 (MetaPool / MetaConv, 512 x 176 frames) and
config 4 (LstmDV embedding + AutoVC + MelGAN on 1000-frame utterances).  Prints one JSON line per measurement."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import _lib, ops, pipeline
from autoformer_b200.factory.AutoVC import AutoVC
from autoformer_b200.factory.LstmDV import LstmDV
from autoformer_b200.factory.MetaConv import MetaConv
from autoformer_b200.factory.MetaPool import MetaPool
from autoformer_b200.melgan.modules import Generator

PEAK = 1374.3


def timed(fn, steps=3, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (_lib.launch_count() - n0) // steps


def families():
    ops.PROFILER.reset()
    ops.PROFILER.enabled = True


def fam_summary():
    torch.cuda.synchronize()
    ops.PROFILER.enabled = False
    return {k: dict(ms=round(v["ms"], 3), tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None)
            for k, v in ops.PROFILER.summary().items()}


def meta(cls, name, flop_per_utt, precision, B):
    torch.manual_seed(0)
    m = cls(44, 256, 512, 22).cuda().eval()
    m.precision = precision
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(B, 176, 80, device="cuda", generator=g) * 6 - 5
    co = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
    ct = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
    ms, launches = timed(lambda: m(x, co, ct))
    families(); m(x, co, ct); fam = fam_summary()
    print(json.dumps({"config": f"{name}(44,256,512,22) B={B} T=176 {precision}", "ms": round(ms, 2),
                      "frames_per_s": round(B * 176 / ms * 1e3), "tflops": round(flop_per_utt * B / ms / 1e9, 1),
                      "frac_of_bf16_sustained": round(flop_per_utt * B / ms / 1e9 / PEAK, 3), "launches": launches,
                      "families": fam}))
    del m
    torch.cuda.empty_cache()


def full_pipeline(precision, B, T=1000):
    torch.manual_seed(0)
    dv, vc, gen = LstmDV().cuda().eval(), AutoVC(32, 256, 512, 32).cuda().eval(), Generator(80, 32, 3).cuda().eval()
    for m in (dv, vc, gen):
        m.precision = precision
    if precision == "mixed":      # fp16x2 where it keeps a margin inside the 1e-3 gate, the split format for MelGAN
        dv.precision = vc.precision = "fp16x2"
        gen.precision = "fp32"
    dv.persistent_lstm = vc.persistent_lstm = True
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.rand(B, T, 80, device="cuda", generator=g) * 6 - 5
    trg = torch.rand(B, T, 80, device="cuda", generator=g) * 6 - 5
    res = {}
    eo, et = dv(src), dv(trg)
    ms, _ = timed(lambda: dv(src), steps=2, warmup=1)
    res["lstmdv_ms"] = round(ms, 2)
    ms, _ = timed(lambda: pipeline.convert(vc, src, eo, et), steps=2, warmup=1)
    res["autovc_ms"] = round(ms, 2)
    mel = pipeline.convert(vc, src, eo, et).transpose(2, 1).contiguous()
    ms, _ = timed(lambda: gen(mel), steps=2, warmup=1)
    res["melgan_ms"] = round(ms, 2)
    families(); gen(mel); res["melgan_families"] = fam_summary()
    families(); dv(src); res["lstmdv_families"] = fam_summary()
    families(); pipeline.convert(vc, src, eo, et); res["autovc_families"] = fam_summary()
    ms, launches = timed(lambda: pipeline.convert_and_vocode(dv, vc, gen, src, trg), steps=2, warmup=1)
    res.update(config=f"LstmDV x2 + AutoVC(pad 1000->1024, trim) + MelGAN, B={B} T={T} {precision}", ms=round(ms, 2),
               frames_per_s=round(B * T / ms * 1e3), launches=launches,
               melgan_tflops=round(90_341_376 * B * T / res["melgan_ms"] / 1e9, 1),
               lstmdv_tflops=round(24_084_480 * B * T / res["lstmdv_ms"] / 1e9, 1))
    print(json.dumps(res))


if __name__ == "__main__":
    what = sys.argv[1:] or ["meta", "pipe"]
    precs = os.environ.get("AVC_BENCH_PRECS", "fp32,bf16").split(",")
    if "pipe" in what:
        for prec in precs:
            for B in (1, 32):
                full_pipeline(prec, B)
    if "meta" in what:
        for prec in precs:
            meta(MetaPool, "MetaPool", 53.441e9, prec, 512)
            meta(MetaConv, "MetaConv", 55.727e9, prec, 512)
