"""Per-kernel time breakdown of one forward (single stream, eager, CUDA events around every launch)."""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="tf32")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--variant", default="full")
ap.add_argument("--hw", type=int, nargs=2, default=[180, 320])
a = ap.parse_args()
dev = torch.device("cuda:0")
m = (arch.GShiftNet if a.variant == "full" else arch.GShiftNet_S)().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict(a.variant, 0))
m.compute_dtype = a.dtype
x = make_clip(1, a.batch, a.hw[0], a.hw[1]).to(dev)
with torch.no_grad():
    m(x)
    m(x)
    eng = m._engine
    eng.profile = []
    m(x)
    torch.cuda.synchronize()
    prof, eng.profile = eng.profile, None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (k, fl, by, e0, e1) in prof:
    r = agg[k]
    r[0] += 1
    r[1] += e0.elapsed_time(e1) * 1e3
    r[2] += fl
tot = sum(r[1] for r in agg.values())
print(f"{a.variant} {a.hw} batch {a.batch} dtype {a.dtype}: {tot / 1e3:.2f} ms per step in kernels ({tot / 1e3 / a.batch:.2f} ms / frame)")
for k, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f"  {fl / us / 1e6:7.1f} TFLOP/s" if fl else ""
    print(f"{k:34s} n={n:4d} total={us:9.1f} us  avg={us / n:8.1f} us  share={us / tot:.3f}{extra}")
