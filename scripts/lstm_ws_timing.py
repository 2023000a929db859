"""Per-frame clock64 breakdown of the weight-stationary small-batch LSTM recurrence (avc_lstm_seq_ws; profiling aid).
Stamps per (frame, CTA): 0 barrier arrival issued (own h stored, membar done), 1 grid barrier passed, 2 h slice landed
(MMA thread), 3 accumulator ready (cell warps), 4 partial sums pushed, 5 all partial sums of the owned
rows landed, 6 cell update done, 7 h stored.  clock64 is per SM: only differences within one CTA are meaningful."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing


def run(B=32, T=256, H=1024, precision="fp32"):
    torch.manual_seed(0)
    w_hh = torch.randn(4 * H, H) * 0.03
    hh = packing.pack_lstm_hh(w_hh, precision, packing.WS_GROUP).cuda()
    xp = (torch.randn(B * T, 4 * H) * 0.5).cuda()
    for _ in range(2):
        assert ops.lstm_seq_ws(xp, hh, B, T, H, precision=precision) is not None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.lstm_seq_ws(xp, hh, B, T, H, precision=precision); e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) * 1e3 / T
    grid_max = 4 * H // 128 * 8
    dbg = torch.zeros(T * grid_max * 8, dtype=torch.int64, device="cuda")
    ops.lstm_seq_ws(xp, hh, B, T, H, debug_clk=dbg, precision=precision)
    torch.cuda.synchronize()
    flat = dbg.cpu()
    # the grid is R x S with S = 8 or 4: find it from the stamps written for frame 1
    for S in (8, 4):
        grid = 4 * H // 128 * S
        d = flat[:T * grid * 8].view(T, grid, 8).double()
        if (d[1:, :, 1] > 0).all() and (flat[T * grid * 8:] == 0).all():
            break
    cur, nxt = d[10:T - 1], d[11:T]
    f = lambda a: f"{a.mean():.0f}"
    print(f"ws {precision} B={B} H={H} S={S} grid={grid}: {plain:.2f} us/frame = {plain * 1.965e3:.0f} cycles @1965 MHz; cycles: "
          f"h stored -> barrier passed {f(nxt[..., 1] - cur[..., 7])} (of which after the arrival was issued {f(nxt[..., 1] - nxt[..., 0])}), "
          f"passed -> h slice landed {f(cur[..., 2] - cur[..., 1])}, "
          f"MMAs -> accumulator ready {f(cur[..., 3] - cur[..., 2])}, "
          f"tmem ld + push {f(cur[..., 4] - cur[..., 3])}, "
          f"wait for the cluster's sums {f(cur[..., 5] - cur[..., 4])}, "
          f"cell {f(cur[..., 6] - cur[..., 5])}, "
          f"h store {f(cur[..., 7] - cur[..., 6])}, "
          f"frame {f(nxt[..., 1] - cur[..., 1])}")


if __name__ == "__main__":
    for prec in ("fp32", "fp16x2"):
        for B, H in ((32, 1024), (32, 768), (32, 512), (1, 1024), (64, 768)):
            run(B=B, H=H, precision=prec)
