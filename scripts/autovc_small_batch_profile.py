"""Per-kernel-family device time of the AutoVC forward at a small batch (config 4's conversion stage: B = 32 or 1, T = 1024;
profiling aid).  CUDA events around every launch (ops.PROFILER)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops
from autoformer_b200.factory.AutoVC import AutoVC
from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker

args = (32, 256, 512, 32)
m = AutoVC(*args)
m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 0))
m = m.cuda().eval()
m.precision = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
for B in (32, 1):
    T = 1024
    x, co, ct = synthetic_mel(B, T, 1).cuda(), synthetic_speaker(B, 1, "org").cuda(), synthetic_speaker(B, 1, "trg").cuda()
    m(x, co, ct); m(x, co, ct)
    m.freeze_weights()
    ops.PROFILER.reset(); ops.PROFILER.enabled = True
    m(x, co, ct)
    torch.cuda.synchronize()
    ops.PROFILER.enabled = False
    tot = 0.0
    print(f"AutoVC B={B} T={T} {m.precision}")
    for name, d in ops.PROFILER.summary().items():
        tot += d["ms"]
        print(f"  {name:14s} launches {d['launches']:3d}  {d['ms']:8.3f} ms")
    print(f"  sum of launches {tot:.3f} ms")
