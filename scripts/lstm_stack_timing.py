"""Wavefront LSTM stack (avc_lstm_stack_ws) against the layer-by-layer weight-stationary kernels: CUDA-event times of
the LstmDV recurrence and a per-tick clock64 breakdown (profiling aid).
Stamps per (tick, CTA): 0 barrier arrival issued (own h stored, membar done), 1 grid barrier passed, 2 first operand group
landed (MMA thread), 3 accumulator ready (cell warps), 4 partial sums pushed, 5 both partial sums of the owned rows landed,
6 cell update done, 7 h stored, 8 second operand group landed, 9 next frame's W_ih part issued, 10 the frame's MMAs issued, 11-12 operand groups landed (polled by the producer thread).  clock64 is per SM: only differences within one CTA are meaningful."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import layers, ops, packing


def build(H, L, I, wavefront):
    torch.manual_seed(0)
    k = 1.0 / H ** 0.5
    ls = []
    for l in range(L):
        w_ih = (torch.rand(4 * H, I if l == 0 else H) * 2 - 1) * k
        w_hh, b_ih, b_hh = (torch.rand(4 * H, H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k, (torch.rand(4 * H) * 2 - 1) * k
        ls.append(layers.LstmLayer(w_ih.cuda(), w_hh.cuda(), b_ih.cuda(), b_hh.cuda(), "fp16x2"))
    return layers.LstmStack(ls, "fp16x2", wavefront)


def timed(fn, n=3):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(B=64, T=1000, H=768, L=3, I=80):
    x = packing.to_act(torch.randn(B, T, I), "fp16x2").cuda()
    out = {}
    for name, wf in (("wavefront", True), ("layer by layer", False)):
        st = build(H, L, I, wf)
        last = torch.empty(B, H, device="cuda")
        out[name] = timed(lambda: st.last_hidden(x, B, T, last, persistent=True))
        out[name + "_h"] = last.clone()
    st = build(H, L, I, True)
    ih, _ = st.layers[0].packs(packing.WS_GROUP)
    xp = torch.empty(B * T, 4 * H, dtype=torch.float32, device="cuda")
    ih(x, B, T, out2=xp)
    rec = timed(lambda: ops.lstm_stack_ws(xp, st.packs(), B, T, H))
    grid = L * (4 * H // 128) * 2
    nt = T + 2 * (L - 1)
    dbg = torch.zeros(nt * grid * 16, dtype=torch.int64, device="cuda")
    ops.lstm_stack_ws(xp, st.packs(), B, T, H, debug_clk=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu().view(nt, grid, 16).double()
    diff = float((out["wavefront_h"] - out["layer by layer_h"]).norm() / out["layer by layer_h"].norm())
    print(f"stack B={B} T={T} H={H} L={L}: wavefront {out['wavefront']:.3f} ms (recurrence alone {rec:.3f} ms = "
          f"{rec * 1e3 / nt:.2f} us/tick = {rec * 1.965e6 / nt:.0f} cycles @1965 MHz), layer by layer {out['layer by layer']:.3f} ms, "
          f"h_last rel diff {diff:.2e}")
    R2 = (4 * H // 128) * 2
    for l in range(L):
        cur, nxt = d[10 + 2 * L:T - 1, l * R2:(l + 1) * R2], d[11 + 2 * L:T, l * R2:(l + 1) * R2]
        f = lambda a: f"{a.mean():.0f}"
        for nm, sl in (("even CTAs (K half 0)", slice(0, None, 2)), ("odd CTAs (K half 1)", slice(1, None, 2))):
            c, n = cur[:, sl], nxt[:, sl]
            print(f"  layer {l} {nm}: h stored -> barrier passed {f(n[..., 1] - c[..., 7])} (after the arrival was issued "
                  f"{f(n[..., 1] - n[..., 0])}), passed -> operand groups landed {f(c[..., 11] - c[..., 1])} / {f(c[..., 12] - c[..., 1])} "
                  f"(seen by the MMA thread {f(c[..., 2] - c[..., 1])} / {f(c[..., 8] - c[..., 1])}), next frame's W_ih part issued {f(c[..., 9] - c[..., 2])} after the first group, "
                  f"first group -> all MMAs issued {f(c[..., 10] - c[..., 2])}, -> accumulator ready "
                  f"{f(c[..., 3] - c[..., 2])}, tmem ld + push {f(c[..., 4] - c[..., 3])}, wait for the pair's sums "
                  f"{f(c[..., 5] - c[..., 4])}, cell {f(c[..., 6] - c[..., 5])}, h store {f(c[..., 7] - c[..., 6])}, tick "
                  f"{f(n[..., 1] - c[..., 1])}")


if __name__ == "__main__":
    for B, T in ((64, 1000), (32, 1000), (2, 1000)):
        run(B=B, T=T)
