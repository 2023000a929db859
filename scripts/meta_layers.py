"""Per-GEMM device time of the MetaPool forward at B=512 (profiling aid)."""
import os, sys, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops
from autoformer_b200.factory.MetaPool import MetaPool

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = 512
m = MetaPool(44, 256, 512, 22).cuda().eval()
m.precision = prec
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(B, 176, 80, device="cuda", generator=g) * 6 - 5
co = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
ct = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
m(x, co, ct)
plan = m._plan()


def tag_mixer(mp, name):
    for k in ("embed", "tok1", "tok2", "ch1", "ch2", "conv"):
        getattr(mp, k).tag = f"{name}.{k} np={mp.np} dim={mp.dim}"


for i, blk in enumerate(plan.enc_blocks):
    tag_mixer(blk.mlp, f"enc_blk{i}")
tag_mixer(plan.dec_block.mlp, "dec_blk")
tag_mixer(plan.enc_mlp, "enc_mlp")
tag_mixer(plan.dec_mlp, "dec_mlp")
ops.PROFILER.reset(); ops.PROFILER.enabled = True
m(x, co, ct)
torch.cuda.synchronize()
ops.PROFILER.enabled = False
agg = collections.OrderedDict()
for fam, s, e, work in ops.PROFILER.records:
    d = agg.setdefault(fam, [0.0, 0.0, 0])
    d[0] += s.elapsed_time(e); d[1] += work.get("flops", 0); d[2] += 1
tot = sum(v[0] for v in agg.values())
for k, (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"{k:40s} n={n:2d} {ms:8.3f} ms {100 * ms / tot:5.1f}%  {fl / ms / 1e9 if fl else 0:7.1f} TFLOP/s")
print("total", tot)
