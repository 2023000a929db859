"""Per-frame clock64 breakdown of the persistent LSTM kernel (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoformer_b200 import ops, packing

def run(precision, B=512, T=128, H=1024):
    torch.manual_seed(0)
    G = ops.choose_gate_group(B, H, True)
    w_hh = torch.randn(4 * H, H) * 0.03
    hh = packing.pack_lstm_hh(w_hh, precision, G).cuda()
    xp = torch.randn(B * T, 4 * H, device="cuda")
    m_tiles = (B + 127) // 128
    m_tiles = (m_tiles + 1) // 2 * 2 if m_tiles >= 2 else m_tiles
    grid = m_tiles * (H // G)
    for _ in range(2):
        ops.lstm_seq(xp, hh, B, T, H, precision, G, persistent=True)
    dbg = torch.zeros(T * grid * 6, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.lstm_seq(xp, hh, B, T, H, precision, G, persistent=True, debug_clk=dbg); e1.record()
    torch.cuda.synchronize()
    d = dbg.view(T, grid, 6).double().cpu()
    step = 2 if m_tiles >= 2 else 1
    lead = d[10:T - 1, 0::step]                  # leader CTAs (MMA stamps), frames 10..T-2
    nxt = d[11:T, 0::step, 0]
    f = lambda a: f"{a.mean():.0f}"
    print(f"{precision} B={B} H={H} G={G} grid={grid}: {e0.elapsed_time(e1) * 1e3 / T:.2f} us/frame; cycles: "
          f"start->first stage {f(lead[..., 1] - lead[..., 0])}, mainloop issue {f(lead[..., 2] - lead[..., 1])}, "
          f"issue end->acc ready {f(lead[..., 4] - lead[..., 2])}, cell {f(lead[..., 5] - lead[..., 4])}, "
          f"cell end->next frame start (barrier) {f(nxt - lead[..., 5])}, frame {f(nxt - lead[..., 0])}; "
          f"preload issue {f(lead[..., 3] - lead[..., 0])}")

import sys
if len(sys.argv) > 1 and sys.argv[1] == "small":
    run("fp32", B=32, T=256, H=768)
    run("fp32", B=64, T=256, H=768)
    run("fp32", B=32, T=256, H=1024)
    run("fp32", B=32, T=256, H=512)
    run("bf16", B=32, T=256, H=768)
else:
    for prec in ("fp32", "bf16", "tf32"):
        run(prec)
    run("fp32", H=512)
