"""Per-launch CUDA-event times of one AutoVC forward at the headline configuration (B = 512, T = 128, fp16x2; profiling aid)."""
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
B, T = 512, 128
x, co, ct = synthetic_mel(B, T, 1).cuda(), synthetic_speaker(B, 1, "org").cuda(), synthetic_speaker(B, 1, "trg").cuda()
for _ in range(3):
    m(x, co, ct)
m.freeze_weights()
ops.PROFILER.reset(); ops.PROFILER.enabled = True
for _ in range(5):
    m(x, co, ct)
torch.cuda.synchronize()
ops.PROFILER.enabled = False
recs = ops.PROFILER.records
n = len(recs) // 5
tot = 0.0
for i in range(n):
    fam, _, _, work = recs[i]
    ms = sum(recs[k * n + i][1].elapsed_time(recs[k * n + i][2]) for k in range(5)) / 5
    tot += ms
    fl = work.get("flops", 0)
    print(f"{i:2d} {fam:14s} {ms * 1e3:8.1f} us  {fl / ms / 1e9 if fl else 0:7.1f} TFLOP/s")
print(f"sum {tot:.3f} ms")
