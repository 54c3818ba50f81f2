"""Cumulative phase times of the CUDA-graph forward: the launch sequence is cut after each phase (Engine.stop_after), captured
and replayed.   usage: python tools/gpu_phase_times.py [bf16|tf32] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
m = arch.GShiftNet().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict("full", 0))
x = make_clip(1, batch, 180, 320).to(dev)
prev = 0.0
for stop in ("mgaa_pair", "mgaa", "mffr", "scnet", None):
    eng = Engine(m, mode=dtype)
    eng.stop_after = stop
    with torch.no_grad():
        eng.forward(x)
        eng.use_graph = True
        for _ in range(3):
            eng.forward(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"through {stop or 'tail (whole forward)':22s}: {ms:7.3f} ms   (+{ms - prev:6.3f})")
    prev = ms
    del eng
