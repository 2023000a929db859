"""Two MelGAN generator forwards at B=8, T=1000 (ncu target for the narrow late-stage convolutions)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200.melgan.modules import Generator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen = Generator(80, 32, 3).cuda().eval()
mel = torch.rand(B, 80, 1000, device="cuda") * 6 - 5
for _ in range(2):
    gen(mel)
torch.cuda.synchronize()
print("ok")
