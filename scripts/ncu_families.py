"""One forward of a named workload inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_melgan \
        python scripts/ncu_families.py melgan

Workloads (sizes of the bench's sub-records where ncu's ~40 replays per launch stay affordable):
  autovc   AutoVC(32,256,512,32) 512 x 128 frames, fp16x2 (BASELINE configs[1]; persistent LSTM kernels launched without
           the cooperative attribute, which ncu cannot replay -- AVC_LSTM_NO_COOP=1, same kernels)
  melgan   Generator(80,32,3) on B = 32 x 1000 frames (the vocoder of configs[3])
  lstmdv   LstmDV on B = 64 x 256 frames, fp16x2 (the weight-stationary small-batch recurrence)
  meta     MetaPool(44,256,512,22) on B = 128 x 176 frames, fp16x2 (every Meta / mixer kernel family)
"""
import os
import sys
import warnings

os.environ.setdefault("AVC_LSTM_NO_COOP", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch

from oracle import templates
from oracle.seeded import seeded_state_dict, synthetic_mel, synthetic_speaker


def build(what):
    if what == "autovc":
        from autoformer_b200.factory.AutoVC import AutoVC
        args = (32, 256, 512, 32)
        m = AutoVC(*args)
        m.load_state_dict(seeded_state_dict(templates.autovc_template(*args), 0))
        m = m.cuda().eval()
        m.precision = os.environ.get("AVC_NCU_PRECISION", "fp16x2")
        B, T = 512, 128
        x, co, ct = synthetic_mel(B, T, 1).cuda(), synthetic_speaker(B, 1, "org").cuda(), synthetic_speaker(B, 1, "trg").cuda()
        return lambda: m(x, co, ct)
    if what == "melgan":
        from autoformer_b200.melgan.modules import Generator
        g = Generator(80, 32, 3)
        g.load_state_dict(seeded_state_dict(templates.melgan_template(), 4))
        g = g.cuda().eval()
        g.precision = os.environ.get("AVC_NCU_PRECISION", "fp32")
        B, T = int(os.environ.get("AVC_NCU_B", 32)), 1000
        mel = synthetic_mel(B, T, 2).transpose(1, 2).contiguous().cuda()
        return lambda: g(mel)
    if what == "lstmdv":
        from autoformer_b200.factory.LstmDV import LstmDV
        m = LstmDV()
        m.load_state_dict(seeded_state_dict(templates.lstmdv_template(), 3, lstm_gain=1.5))
        m = m.cuda().eval()
        m.precision = os.environ.get("AVC_NCU_PRECISION", "fp16x2")
        x = synthetic_mel(64, 256, 3).cuda()
        return lambda: m(x)
    if what == "meta":
        from autoformer_b200.factory.MetaPool import MetaPool
        args = (44, 256, 512, 22)
        m = MetaPool(*args)
        m.load_state_dict(seeded_state_dict(templates.meta_template("pool", *args), 6))
        m = m.cuda().eval()
        m.precision = os.environ.get("AVC_NCU_PRECISION", "fp16x2")
        B = int(os.environ.get("AVC_NCU_B", 128))
        x, co, ct = synthetic_mel(B, 176, 4).cuda(), synthetic_speaker(B, 4, "org").cuda(), synthetic_speaker(B, 4, "trg").cuda()
        return lambda: m(x, co, ct)
    raise SystemExit(f"unknown workload {what}")


if __name__ == "__main__":
    fn = build(sys.argv[1])
    fn()
    fn()
    for obj in list(fn.__closure__ or []):          # serving mode: no per-forward weight digest inside the profiled range
        m = obj.cell_contents
        if hasattr(m, "freeze_weights"):
            m.freeze_weights()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", sys.argv[1])
