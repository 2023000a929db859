"""CPU tests of the host-side weight folding / packing against torch's own operators."""
import re
import os
import ctypes

import pytest
import torch
import torch.nn.functional as F

from autoformer_b200 import packing
from oracle.layers import lstm_explicit
from tests.emulate import emulate_conv_gemm, emulate_lstm_seq

torch.manual_seed(0)


def _cl(x):  # (B,C,T) -> channels-last (B,T,C)
    return x.transpose(1, 2).contiguous()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("c_in,c_out,k", [(336, 512, 5), (80, 512, 5), (512, 80, 5), (88, 352, 1), (80, 512, 7)])
def test_pack_conv_matches_conv1d(precision, c_in, c_out, k):
    B, T = 2, 19
    w = torch.randn(c_out, c_in, k)
    b = torch.randn(c_out)
    x = torch.randn(B, c_in, T)
    wp, bp, meta = packing.pack_conv(w, b, precision)
    assert wp.shape == (meta["n_pad"], meta["k_pad"]) and meta["n_pad"] % meta["block_n"] == 0
    assert meta["k_pad"] % packing.KC[precision] == 0
    ours = emulate_conv_gemm(wp, bp, meta, [_cl(x)], B, T, [-(k // 2)], [1])
    ref = _cl(F.conv1d(x.double(), w.double(), b.double(), padding=k // 2))
    tol = 1e-9 if precision == "tf32" else 1e-9
    # compare with the operand-rounded weights (packing rounds; the rounding itself is tested below)
    ref_r = _cl(F.conv1d(x.double(), packing.to_operand(w, precision).double(), b.double(), padding=k // 2))
    assert (ours - ref_r).abs().max() < tol
    assert (ours - ref).norm() / ref.norm() < (2e-3 if precision == "tf32" else 1e-2)


@pytest.mark.parametrize("c_in,c_out,k", [(336, 512, 5), (80, 96, 5), (512, 80, 1)])
def test_split_bf16_packing_is_fp32_grade(c_in, c_out, k):
    """"fp32" mode: [a_hi|a_lo] x [w_hi|w_hi] + a_hi x w_lo reproduces the fp32 product to ~2^-16."""
    B, T = 2, 13
    w = torch.randn(c_out, c_in, k)
    b = torch.randn(c_out)
    x = torch.randn(B, c_in, T)
    wp, bp, meta = packing.pack_conv(w, b, "fp32")
    assert meta["split"] and meta["channels"] == [2 * c_in, c_in] and wp.dtype == torch.bfloat16
    buf = packing.to_act(_cl(x), "fp32")
    assert buf.shape == (B, T, 2 * c_in)
    assert (packing.act_to_float(buf, "fp32") - _cl(x)).abs().max() < 2 ** -15 * x.abs().max()
    ours = emulate_conv_gemm(wp, bp, meta, [buf, buf[..., :c_in]], B, T, [-(k // 2)] * 2, [1, 1])
    ref = _cl(F.conv1d(x.double(), w.double(), b.double(), padding=k // 2))
    assert (ours - ref).norm() / ref.norm() < 3e-5


@pytest.mark.parametrize("c_in,c_out,k", [(336, 512, 5), (80, 96, 5), (512, 80, 1)])
def test_fp16x2_packing_rounds_only_the_activations(c_in, c_out, k):
    """"fp16x2" mode: a x w_hi + a x w_lo over the SAME fp16 buffer -- the weights are exact to ~2^-22, so with
    activations that are exactly representable in fp16 the product matches fp64 to fp32 rounding level, and with
    general activations the error is the 2^-12 of their rounding."""
    B, T = 2, 13
    w = torch.randn(c_out, c_in, k)
    b = torch.randn(c_out)
    x = torch.randn(B, c_in, T)
    wp, bp, meta = packing.pack_conv(w, b, "fp16x2")
    assert meta["dup"] and not meta.get("split") and meta["channels"] == [c_in, c_in] and wp.dtype == torch.float16
    assert meta["precision"] == "fp16x2" and meta["logical_channels"] == [c_in] and meta["logical_taps"] == [k]
    hi, lo = packing.split_f16(w)
    assert (hi.double() + lo.double() - w.double()).abs().max() < 2 ** -21 * w.abs().max()
    buf = packing.to_act(_cl(x), "fp16x2")
    assert buf.shape == (B, T, c_in) and buf.dtype == torch.float16 and packing.act_channels(c_in, "fp16x2") == c_in
    ours = emulate_conv_gemm(wp, bp, meta, [buf, buf], B, T, [-(k // 2)] * 2, [1, 1])
    ref = _cl(F.conv1d(x.double(), w.double(), b.double(), padding=k // 2))
    ref_q = _cl(F.conv1d(x.half().double(), w.double(), b.double(), padding=k // 2))     # activations rounded, weights exact
    assert (ours - ref_q).norm() / ref_q.norm() < 2e-6
    assert 1e-5 < (ours - ref).norm() / ref.norm() < 5e-4


def test_fp16x2_lstm_weight_layout():
    H, G = 128, 32
    w = torch.randn(4 * H, H) * 0.1
    p = packing.pack_lstm_hh(w, "fp16x2", G)
    assert p.shape == (4 * H, 2 * H) and p.dtype == torch.float16
    perm = packing.gate_permutation(H, G)
    assert ((p[:, :H].double() + p[:, H:].double()) - w[perm].double()).abs().max() < 2 ** -21 * w.abs().max()
    wi, bias = packing.pack_lstm_ih_fused(torch.randn(4 * H, 80), torch.randn(4 * H), torch.randn(4 * H), "fp16x2", G)
    assert wi.shape == (4 * H, 2 * 128) and wi.dtype == torch.float16 and bias.dtype == torch.float32
    assert float(wi[:, 80:128].abs().max()) == 0.0 and float(wi[:, 128 + 80:].abs().max()) == 0.0   # K padding


def test_split_lstm_hh_layout():
    H, G = 128, 32
    w = torch.randn(4 * H, H)
    p = packing.pack_lstm_hh(w, "fp32", G)
    assert p.shape == (4 * H, 2 * H) and p.dtype == torch.bfloat16
    perm = packing.gate_permutation(H, G)
    assert ((p[:, :H].float() + p[:, H:].float()) - w[perm]).abs().max() < 2 ** -15 * w.abs().max()


def test_round_tf32_is_rna():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.14159274, 1e-30, -7.5e8])
    r = packing.round_tf32(x)
    assert (r.view(torch.int32) & 0x1FFF).eq(0).all()
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[3] == -1.0 - 2 ** -10   # ties away from zero
    assert ((r - x).abs() <= x.abs() * 2 ** -11).all()


def test_fold_bn_matches_batchnorm_eval():
    c_in, c_out, B, T = 16, 24, 2, 11
    conv = torch.nn.Conv1d(c_in, c_out, 5, padding=2)
    bn = torch.nn.BatchNorm1d(c_out).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.5, 2)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    x = torch.randn(B, c_in, T)
    w, b = packing.fold_bn(conv.weight.detach(), conv.bias.detach(), bn.weight.detach(), bn.bias.detach(),
                           bn.running_mean, bn.running_var)
    with torch.no_grad():
        ref = bn(conv(x))
        ours = F.conv1d(x, w, b, padding=2)
    assert torch.allclose(ours, ref, atol=2e-5)


@pytest.mark.parametrize("G", [16, 32, 64])
def test_lstm_packing_matches_explicit_lstm(G):
    B, T, I, H = 3, 7, 40, 128
    w_ih, w_hh = torch.randn(4 * H, I) * 0.2, torch.randn(4 * H, H) * 0.2
    b_ih, b_hh = torch.randn(4 * H) * 0.1, torch.randn(4 * H) * 0.1
    x = torch.randn(B, T, I)
    perm = packing.gate_permutation(H, G)
    assert sorted(perm.tolist()) == list(range(4 * H))
    wp, bp, meta = packing.pack_lstm_ih(w_ih, b_ih, b_hh, "tf32", G)
    xp = emulate_conv_gemm(wp, bp, meta, [x], B, T, [0], [1])
    hh = packing.pack_lstm_hh(w_hh, "tf32", G)
    ours = emulate_lstm_seq(xp, hh, B, T, H, G)
    ref = lstm_explicit(x.double(), packing.round_tf32(w_ih).double(), packing.round_tf32(w_hh).double(),
                        b_ih.double(), b_hh.double())
    assert (ours - ref).abs().max() < 1e-6   # bias sum is formed in fp32


def test_bilstm_projection_layout():
    I, H = 24, 12
    ws = [torch.randn(4 * H, I) for _ in range(2)]
    bs = [torch.randn(4 * H) for _ in range(4)]
    wp, bp, meta = packing.pack_bilstm_ih(ws[0], bs[0], bs[1], ws[1], bs[2], bs[3], "tf32")
    x = torch.randn(1, 5, I)
    xp = emulate_conv_gemm(wp, bp, meta, [x], 1, 5, [0], [1])
    ref_f = x.double() @ packing.round_tf32(ws[0]).double().t() + (bs[0] + bs[1]).double()
    ref_r = x.double() @ packing.round_tf32(ws[1]).double().t() + (bs[2] + bs[3]).double()
    assert (xp[..., :4 * H] - ref_f).abs().max() < 1e-9 and (xp[..., 4 * H:] - ref_r).abs().max() < 1e-9


@pytest.mark.parametrize("r,c_in,c_out", [(8, 32, 16), (2, 16, 8)])
def test_conv_transpose_as_three_tap_conv(r, c_in, c_out):
    B, L = 2, 9
    p = r // 2 + r % 2
    w = torch.randn(c_in, c_out, 2 * r)
    b = torch.randn(c_out)
    x = torch.randn(B, c_in, L)
    ref = F.conv_transpose1d(x.double(), w.double(), b.double(), stride=r, padding=p, output_padding=r % 2)
    w3 = packing.conv_transpose_as_conv(w, r, p)
    wp, bp, meta = packing.pack_conv(w3, b.repeat(r), "tf32")
    ours = emulate_conv_gemm(wp, bp, meta, [_cl(x)], B, L, [-1], [1])       # [B][L][r*c_out]
    ours = ours.reshape(B, L * r, c_out)
    ref_r = F.conv_transpose1d(x.double(), packing.round_tf32(w).double(), b.double(), stride=r, padding=p)
    assert (ours - _cl(ref_r)).abs().max() < 1e-9
    assert ref.shape[-1] == L * r


def test_weight_norm_fold():
    v = torch.randn(6, 4, 3)
    g = torch.rand(6, 1, 1) + 0.5
    m = torch.nn.utils.weight_norm(torch.nn.Conv1d(4, 6, 3))
    with torch.no_grad():
        m.weight_v.copy_(v)
        m.weight_g.copy_(g)
    x = torch.randn(1, 4, 8)
    with torch.no_grad():
        ref = m(x)
    ours = F.conv1d(x, packing.fold_weight_norm(g, v), m.bias.detach())
    assert torch.allclose(ours, ref, atol=1e-6)


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every function include/avc_b200.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "avc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(avc_[a-z0-9_]+)\s*\(", header))
    assert {"avc_conv_gemm", "avc_lstm_seq", "avc_bilstm_small", "avc_concat_bcast"} <= declared
    from autoformer_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().avc_version() >= 100
