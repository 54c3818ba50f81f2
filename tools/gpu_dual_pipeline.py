"""Experiment: N independent forward pipelines (own engine, workspace, CUDA graph) replayed concurrently on N streams, so that
one pipeline's HBM-bound elementwise kernels overlap another's tensor-core convolutions.  usage: gpu_dual_pipeline.py [B] [N]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
sd = arch.seeded_state_dict("full", 0)
models, xs, streams = [], [], []
with torch.no_grad():
    for i in range(N):
        m = arch.GShiftNet().to(dev).eval()
        m.load_state_dict(sd)
        m.compute_dtype = "bf16"
        x = make_clip(1234 + i, B, 180, 320).to(dev)
        m(x)
        m._engine.use_graph = True
        m(x)
        models.append(m)
        xs.append(x)
        streams.append(torch.cuda.Stream(dev))
    torch.cuda.synchronize()

    def step():
        for m, x, s in zip(models, xs, streams):
            with torch.cuda.stream(s):
                m(x)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
print(f"{N} pipelines x {B} windows: {dt * 1e3:.2f} ms per round, {N * B / dt:.1f} frames/s")
