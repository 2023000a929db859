"""Device time of the narrow late-stage MelGAN convolutions by variant (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing


def run(precision, c, k, B=8, T=256000, cta=2, dil=1, raw=False):
    torch.manual_seed(0)
    w = torch.randn(c, c, k) / (c * k) ** 0.5
    layer = ops.ConvGemm(*packing.pack_conv(w, torch.randn(c), precision), act="lrelu", tap_t0=[-dil * (k // 2)],
                         tap_dt=[dil]).to("cuda")
    layer.cta_group = cta
    x = packing.to_act(torch.randn(B, T, c), precision).cuda()
    out = ops.alloc_act(B, T, c, precision, "cuda")
    out_raw = ops.alloc_act(B, T, c, precision, "cuda") if raw else None
    for _ in range(2):
        layer(x, B, T, out=out, out_raw=out_raw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        layer(x, B, T, out=out, out_raw=out_raw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 3
    gb = (x.numel() + out.numel() * (2 if raw else 1)) * x.element_size() / 1e9
    kb = layer.meta["k_pad"] // packing.KC[precision]
    units = B * T / 256
    print(f"{precision} C={c} k{k} d{dil} cta={cta} raw={raw}: {us:7.0f} us  {gb / us * 1e6:6.0f} GB/s  k-blocks {kb}  "
          f"{us * 1e3 / (units / 74) :6.0f} ns per 256-row unit")


for c in (32, 64, 128):
    for prec in ("fp32", "bf16"):
        for cta in (2, 1):
            run(prec, c, 3, cta=cta, dil=3)
run("fp32", 32, 1)
run("fp32", 32, 3, raw=True)
run("fp32", 32, 3, B=2)
