"""Do two under-filled AutoVC batches overlap when they run on two streams with the persistent LSTM grids capped at half
the SMs?  (profiling aid for pipeline.convert_pairs)"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from autoformer_b200 import ops
from autoformer_b200.factory.AutoVC import AutoVC
from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker

args = (32, 256, 512, 32)
m = AutoVC(*args)
m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 0))
m = m.cuda().eval()
m.precision = "fp16x2"


def inputs(B, T, seed):
    return synthetic_mel(B, T, seed).cuda(), synthetic_speaker(B, seed, "org").cuda(), synthetic_speaker(B, seed, "trg").cuda()


def timed(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (B1, T1), (B2, T2) in (((212, 1024), (180, 992)), ((212, 512), (230, 480)), ((100, 1024), (90, 992))):
    a, b = inputs(B1, T1, 1), inputs(B2, T2, 2)
    m(*a); m(*b)
    m.freeze_weights()
    ops.LSTM_CTA_BUDGET = None
    seq = timed(lambda: (m(*a), m(*b)))
    ref = m(*a)[1].clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            o1 = m(*a)
        with torch.cuda.stream(s2):
            o2 = m(*b)
        cur.wait_stream(s1); cur.wait_stream(s2)
        return o1, o2

    for budget in (None, 74):
        ops.LSTM_CTA_BUDGET = budget
        par = timed(both)
        same = torch.equal(both()[0][1], ref)
        print(f"({B1} x {T1}) + ({B2} x {T2}): one after the other {seq:.2f} ms; two streams, LSTM CTA budget {budget}: {par:.2f} ms; "
              f"outputs bit-equal to the single-stream run: {same}")
    ops.LSTM_CTA_BUDGET = None
