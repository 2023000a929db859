"""MelGAN generator timing by precision on one GPU (B = 32 x 1000 frames unless AVC_B / AVC_T say otherwise)."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import _lib, ops
from autoformer_b200.melgan.modules import Generator
from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel

B, T = int(os.environ.get("AVC_B", 32)), int(os.environ.get("AVC_T", 1000))
sd = seeded_state_dict(templates.melgan_template(), 4)
mel = synthetic_mel(B, T, 2).transpose(1, 2).contiguous().cuda()
for prec in sys.argv[1:] or ["fp32", "fp16s"]:
    g = Generator(80, 32, 3)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    g.precision = prec
    g(mel); g(mel)
    g.freeze_weights()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(5):
        g(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ops.PROFILER.reset(); ops.PROFILER.enabled = True
    g(mel); torch.cuda.synchronize(); ops.PROFILER.enabled = False
    rows = []
    for fam, s, e, work in ops.PROFILER.records:
        rows.append((fam, round(s.elapsed_time(e), 3), round(work.get("flops", 0) / max(s.elapsed_time(e), 1e-6) / 1e9, 1),
                     round(work.get("bytes", 0) / max(s.elapsed_time(e), 1e-6) / 1e6, 1)))
    print(json.dumps({"precision": prec, "B": B, "T": T, "ms": round(ms, 3), "launches": (_lib.launch_count() - n0) // 5,
                      "tflops": round(90_341_376 * B * T / ms / 1e9, 1), "per_launch(fam, ms, TFLOP/s, GB/s)": rows}))
    del g
    torch.cuda.empty_cache()
