"""Phase timeline of one eager forward (multi-stream on): CUDA events around feat / MGAA x3 / MFFR / SCNet / tail."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda:0")
m = arch.GShiftNet().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict("full", 0))
m.compute_dtype = a.dtype
x = make_clip(1, a.batch, 180, 320).to(dev)
marks = []


def wrap(name):
    orig = getattr(Engine, name)

    def f(self, *args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        r = orig(self, *args, **kw)
        e1.record(torch.cuda.current_stream())
        marks.append((name + kw.get("sfx", ""), e0, e1))
        return r
    setattr(Engine, name, f)


for n in ("_mgaa", "_mffr", "_scnet", "_tail"):
    wrap(n)
with torch.no_grad():
    for _ in range(3):
        m(x)
    marks.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    m(x)
    t1.record()
    torch.cuda.synchronize()
print(f"eager forward {t0.elapsed_time(t1):.2f} ms (batch {a.batch}, {a.dtype})")
for name, e0, e1 in marks:
    print(f"  {name:10s} start {t0.elapsed_time(e0):7.2f} ms  dur {e0.elapsed_time(e1):7.2f} ms")
