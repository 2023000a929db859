"""Small forwards of every kernel family (finite-output check; also the target for a memcheck run where the tool is available): AutoVC in the four precisions (persistent and
per-frame LSTM launches, the weight-stationary small-batch kernel), LstmDV, MelGAN (fused ResnetBlocks), MetaPool."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200.factory.AutoVC import AutoVC
from autoformer_b200.factory.LstmDV import LstmDV
from autoformer_b200.factory.MetaPool import MetaPool
from autoformer_b200.melgan.modules import Generator

torch.manual_seed(0)
g = torch.Generator(device="cuda").manual_seed(1)
for B, T in ((3, 64), (130, 32)):
    x = torch.rand(B, T, 80, device="cuda", generator=g) * 6 - 5
    co = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
    ct = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda", generator=g), dim=-1)
    vc = AutoVC(32, 256, 512, 32).cuda().eval()
    for prec in ("fp32", "fp16x2", "tf32", "bf16"):
        for persistent in (True, False):
            vc.precision, vc.persistent_lstm = prec, persistent
            out = vc(x, co, ct)
            assert all(torch.isfinite(o).all() for o in out), (prec, persistent)
dv = LstmDV().cuda().eval()
for prec in ("fp32", "fp16x2"):
    dv.precision = prec
    assert torch.isfinite(dv(torch.rand(5, 40, 80, device="cuda", generator=g) * 6 - 5)).all()
gen = Generator(80, 32, 3).cuda().eval()
for prec in ("fp32", "fp16x2"):
    gen.precision = prec
    assert torch.isfinite(gen(torch.rand(2, 80, 16, device="cuda", generator=g) * 6 - 5)).all()
mp = MetaPool(44, 256, 512, 22).cuda().eval()
for prec in ("fp32", "fp16x2"):
    mp.precision = prec
    x = torch.rand(1, 176, 80, device="cuda", generator=g) * 6 - 5
    c = torch.nn.functional.normalize(torch.randn(1, 256, device="cuda", generator=g), dim=-1)
    assert all(torch.isfinite(o).all() for o in mp(x, c, c))
torch.cuda.synchronize()
print("sanitize_small: ok")
