"""One low-K GEMM launch (64 -> 1376 at the decoder mixer's row count: the kernel is its epilogue) for a source-level ncu
capture of conv_gemm_kernel's epilogue:  ncu --set full --import-source on -k regex:conv_gemm -s 2 -c 1 ... """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

act = sys.argv[1] if len(sys.argv) > 1 else "none"
B, T, c_in, c_out = 128, 1856, 64, 1376
torch.manual_seed(0)
layer = ops.ConvGemm(*packing.pack_linear(torch.randn(c_out, c_in) / 8, torch.randn(c_out), "fp16x2"), act=act).to("cuda")
x = packing.to_act(torch.randn(B, T, c_in), "fp16x2").cuda()
out = ops.alloc_act(B, T, c_out, "fp16x2", "cuda")
for _ in range(3):
    layer(x, B, T, out=out)
torch.cuda.synchronize()
print("ok")
