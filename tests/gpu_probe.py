"""First-contact GPU probe: runs each kernel family in its own subprocess (a device trap poisons the context),
from the smallest case up, and prints error statistics.  Usage: python tests/gpu_probe.py [case ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _gemm_case(precision, B, T, c_in, c_out, k, act="none", block_n=None, seed=0):
    import torch
    import torch.nn.functional as F
    from autoformer_b200 import ops, packing
    from oracle import rel_l2
    torch.manual_seed(seed)
    w = torch.randn(c_out, c_in, k) / (c_in * k) ** 0.5
    b = torch.randn(c_out)
    x = torch.randn(B, T, c_in)
    layer = ops.ConvGemm(*packing.pack_conv(w, b, precision, block_n), act=act).to("cuda")
    xo = packing.to_act(x, precision)
    if precision == "fp32":
        ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=k // 2).transpose(1, 2)
    else:
        ref = F.conv1d(xo.double().transpose(1, 2), packing.to_operand(w, precision).double(), b.double(),
                       padding=k // 2).transpose(1, 2)
    if act == "relu":
        ref = ref.relu()
    elif act == "tanh":
        ref = ref.tanh()
    out = ops.alloc_act(B, T, c_out, precision, "cuda")
    out2 = torch.full((B * T, c_out), float("nan"), dtype=torch.float32, device="cuda")
    layer(xo.cuda(), B, T, out=out, out2=out2, round_tf32=False)
    torch.cuda.synchronize()
    got = out2.view(B, T, c_out).cpu()
    err = rel_l2(got, ref)
    err_out = rel_l2(packing.act_to_float(out, precision).cpu(), ref)
    print(f"  (operand-format output rel_l2={err_out:.3e})")
    nan = int(torch.isnan(got).sum())
    print(f"  gemm {precision} B={B} T={T} Cin={c_in} Cout={c_out} k={k} act={act} bn={layer.meta['block_n']}: "
          f"rel_l2={err:.3e} nan={nan} |ref|={float(ref.norm()):.3f} |got|={float(got.nan_to_num().norm()):.3f}")
    if not (err < 1e-4):
        d = (got.double() - ref).abs()
        print("   max abs diff", float(d.nan_to_num().max()), "per-row-block err:",
              [f"{float(d[:, i:i + 8].nan_to_num().mean()):.2e}" for i in range(0, min(T, 64), 8)])
        print("   per-col-block err:", [f"{float(d[..., i:i + 32].nan_to_num().mean()):.2e}" for i in range(0, c_out, 32)][:16])
        print("   got[0,0,:8]", got[0, 0, :8].tolist(), "\n   ref[0,0,:8]", ref[0, 0, :8].tolist())
    return err


@case
def gemm_min():
    _gemm_case("tf32", 1, 128, 32, 64, 1)          # one k-block, one tile, BN=64


@case
def gemm_k():
    _gemm_case("tf32", 1, 128, 256, 128, 1)        # 8 k-blocks (ring wraps for BN=128: 6 stages)


@case
def gemm_n256():
    _gemm_case("tf32", 2, 128, 512, 512, 1)


@case
def conv_k5():
    _gemm_case("tf32", 2, 128, 64, 128, 5)
    _gemm_case("tf32", 3, 64, 336, 512, 5, act="relu")
    _gemm_case("tf32", 2, 176, 80, 512, 5, act="tanh")
    _gemm_case("tf32", 5, 96, 512, 80, 5)


@case
def gemm_bf16():
    _gemm_case("bf16", 1, 128, 64, 64, 1)
    _gemm_case("bf16", 2, 128, 512, 512, 5, act="relu")


@case
def gemm_fp32():
    _gemm_case("fp32", 1, 128, 64, 64, 1)
    _gemm_case("fp32", 3, 64, 336, 512, 5, act="relu")
    _gemm_case("fp32", 2, 176, 80, 512, 5, act="tanh")
    _gemm_case("fp32", 2, 128, 512, 80, 5)


def _lstm_case(precision, B, T, I, H, persistent=False, seed=0):
    import torch
    from autoformer_b200 import layers, packing
    from oracle import rel_l2
    from oracle.layers import lstm_explicit
    torch.manual_seed(seed)
    k = 1.0 / H ** 0.5
    w_ih, w_hh = (torch.rand(4 * H, I) * 2 - 1) * k * 3, (torch.rand(4 * H, H) * 2 - 1) * k * 3
    b_ih, b_hh = (torch.rand(4 * H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k
    x = torch.randn(B, T, I)
    ref = lstm_explicit(x.double(), w_ih.double(), w_hh.double(), b_ih.double(), b_hh.double())
    layer = layers.LstmLayer(w_ih.cuda(), w_hh.cuda(), b_ih.cuda(), b_hh.cuda(), precision)
    f32 = torch.full((B, T, H), float("nan"), device="cuda")
    last = torch.full((B, H), float("nan"), device="cuda")
    t0 = time.time()
    layer(packing.to_act(x, precision).cuda(), B, T, hseq_f32=f32, h_last=last, persistent=persistent)
    torch.cuda.synchronize()
    dt = time.time() - t0
    got = f32.cpu()
    errs = [rel_l2(got[:, t], ref[:, t]) for t in (0, 1, 2, T // 2, T - 1)]
    print(f"  lstm {precision} B={B} T={T} I={I} H={H} persistent={persistent}: rel_l2={rel_l2(got, ref):.3e} "
          f"nan={int(torch.isnan(got).sum())} per-step(0,1,2,mid,last)={['%.2e' % e for e in errs]} "
          f"h_last={rel_l2(last.cpu(), ref[:, -1]):.2e} wall={dt * 1e3:.1f}ms")


@case
def lstm_step():
    _lstm_case("tf32", 4, 8, 64, 128)
    _lstm_case("tf32", 130, 16, 320, 512)
    _lstm_case("tf32", 64, 32, 512, 1024)


@case
def lstm_bf16():
    _lstm_case("bf16", 64, 32, 512, 1024)


@case
def lstm_fp32():
    _lstm_case("fp32", 4, 8, 64, 128)
    _lstm_case("fp32", 130, 16, 320, 512)
    _lstm_case("fp32", 64, 32, 512, 1024)
    _lstm_case("fp32", 256, 64, 512, 1024, persistent=True)


@case
def lstm_persistent():
    _lstm_case("tf32", 4, 8, 64, 128, persistent=True)
    _lstm_case("tf32", 256, 64, 512, 1024, persistent=True)


@case
def bilstm():
    import torch
    from autoformer_b200 import layers, packing
    from oracle import rel_l2
    from oracle.layers import lstm_stack
    from oracle.seeded import seeded_state_dict
    for H, freq, B, T in ((32, 32, 5, 64), (44, 22, 3, 88)):
        tmpl = {}
        from oracle.templates import _lstm
        _lstm(tmpl, "lstm", 512, H, 2, bidirectional=True)
        sd = seeded_state_dict(tmpl, 3)
        x = torch.randn(B, T, 512)
        ref = lstm_stack({k: v.double() for k, v in sd.items()}, "lstm", x.double(), 2, bidirectional=True, impl="explicit")
        ref_codes = torch.cat((ref[:, freq - 1::freq, :H], ref[:, ::freq, H:]), dim=-1)
        m = layers.BiLstmSmall({k: v.cuda() for k, v in sd.items()}, "lstm", 2, "tf32")
        out, codes = m(packing.to_operand(x, "tf32").cuda(), B, T, freq=freq, want_out=True)
        torch.cuda.synchronize()
        print(f"  bilstm H={H} B={B} T={T}: out rel_l2={rel_l2(out, ref):.3e} codes rel_l2={rel_l2(codes, ref_codes):.3e}")


@case
def autovc_small():
    import __graft_entry__
    __graft_entry__.smoke()


@case
def autovc_taps():
    import torch
    from autoformer_b200.factory.AutoVC import AutoVC
    from oracle import rel_l2, templates, centred_rel_l2
    from oracle.autovc import autovc_forward
    from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker
    cfgs = (((32, 256, 512, 32), 2, 128, "fp32"), ((44, 256, 512, 22), 2, 176, "fp32"),
            ((32, 256, 512, 32), 2, 128, "tf32"), ((32, 256, 512, 32), 2, 128, "bf16"))
    for args, B, T, prec in cfgs:
        sd = seeded_state_dict(templates.autovc_template(*args), 0)
        x, c_org, c_trg = synthetic_mel(B, T, 1234), synthetic_speaker(B, 1234, "org"), synthetic_speaker(B, 1234, "trg")
        rt = {}
        autovc_forward(sd, x, c_org, c_trg, args[0], args[3], taps=rt, dtype=torch.float64)
        m = AutoVC(*args)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        m.precision = prec
        m.collect_taps = True
        m(x.cuda(), c_org.cuda(), c_trg.cuda())
        torch.cuda.synchronize()
        print(f"  autovc {args} B={B} T={T} {prec}:")
        for k in rt:
            if k in m.taps:
                print(f"    {k:12s} rel_l2={rel_l2(m.taps[k], rt[k]):.3e} centred={centred_rel_l2(m.taps[k], rt[k]):.3e}")


@case
def melgan():
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    from autoformer_b200.melgan.modules import Generator
    from oracle import rel_l2, templates
    from oracle.melgan import melgan_forward
    from oracle.seeded import seeded_state_dict, synthetic_mel
    sd = seeded_state_dict(templates.melgan_template(), 4)
    for prec in ("fp32", "tf32", "bf16"):
        for B, T in ((1, 40), (2, 17)):
            mel = synthetic_mel(B, T, 6).transpose(1, 2).contiguous()
            rt = {}
            ref = melgan_forward(sd, mel, taps=rt, dtype=torch.float64)
            g = Generator(80, 32, 3)
            g.load_state_dict(sd)
            g = g.cuda().eval()
            g.precision = prec
            g.collect_taps = True
            wav = g(mel.cuda())
            torch.cuda.synchronize()
            print(f"  melgan {prec} B={B} T={T}: wav rel_l2={rel_l2(wav, ref):.3e}", {k: f"{rel_l2(g.taps[k], rt[k].transpose(1, 2)):.1e}" for k in g.taps})


@case
def lstmdv():
    import torch
    from autoformer_b200.factory.LstmDV import LstmDV
    from oracle import rel_l2, templates
    from oracle.lstmdv import lstmdv_forward
    from oracle.seeded import seeded_state_dict, synthetic_mel
    sd = seeded_state_dict(templates.lstmdv_template(), 3, lstm_gain=1.5)
    x = synthetic_mel(2, 100, 5)
    ref = lstmdv_forward(sd, x, dtype=torch.float64)
    for prec in ("fp32", "tf32", "bf16"):
        m = LstmDV()
        m.load_state_dict(sd)
        m = m.cuda().eval()
        m.precision = prec
        print(f"  lstmdv {prec}: rel_l2={rel_l2(m(x.cuda()), ref):.3e}")


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get("AVC_PROBE_CHILD"):
        CASES[names[0]]()
        return 0
    env = dict(os.environ, AVC_PROBE_CHILD="1")
    for n in names:
        print(f"== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, timeout=300,
                               capture_output=True, text=True)
            print(r.stdout[-6000:], end="")
            if r.returncode != 0:
                print(f"  !! exit {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired:
            print("  !! timeout")
        sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
