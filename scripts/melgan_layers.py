"""Per-layer device time of the MelGAN generator (profiling aid)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops
from autoformer_b200.melgan.modules import Generator

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B, T = 32, 1000
gen = Generator(80, 32, 3).cuda().eval()
gen.precision = prec
mel = torch.rand(B, 80, T, device="cuda") * 6 - 5
gen(mel)
plan = gen._plan()
plan.stem.tag = "stem 80->512 k7"
for si, (r, C, up, blocks) in enumerate(plan.stages):
    up.tag = f"s{si} up x{r} ->{C}"
    for j, (d, c3, k1, fused) in enumerate(blocks):
        c3.tag = f"s{si} b{j} k3 d{d} C{C}"
        k1.tag = f"s{si} b{j} k1+sc C{C}"
        if fused is not None:
            fused.tag = f"s{si} b{j} fused d{d} C{C}"
ops.PROFILER.reset(); ops.PROFILER.enabled = True
gen(mel)
torch.cuda.synchronize()
ops.PROFILER.enabled = False
tot = 0
L = T
for fam, s, e, work in ops.PROFILER.records:
    ms = s.elapsed_time(e)
    tot += ms
    fl = work.get("flops", 0)
    print(f"{fam:22s} {ms*1e3:8.0f} us  {fl/ms/1e9 if fl else 0:7.1f} TFLOP/s")
print("total", tot, "ms")
