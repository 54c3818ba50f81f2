"""A/B of an Engine attribute on the same box: two engines over the same model, graph-replayed alternately.
usage: python tools/gpu_ab_flag.py FLAG [bf16|tf32] [batch] [stop_after]      e.g.  gpu_ab_flag.py x16 bf16 6 scnet"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

flag = sys.argv[1]
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
stop = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
m = arch.GShiftNet().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict("full", 0))
x = make_clip(1, batch, 180, 320).to(dev)
engs = {}
with torch.no_grad():
    for val in (False, True):
        eng = Engine(m, mode=dtype)
        assert hasattr(eng, flag), flag
        setattr(eng, flag, val)
        eng.stop_after = stop
        eng.forward(x)
        eng.use_graph = True
        for _ in range(3):
            eng.forward(x)
        engs[val] = eng
    tot = {False: [], True: []}
    for _ in range(6):
        for val in (False, True):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                engs[val].forward(x)
            e1.record()
            torch.cuda.synchronize()
            tot[val].append(e0.elapsed_time(e1) / 10)
for val in (False, True):
    t = sorted(tot[val])
    print(f"{flag}={val!s:5s} {dtype} B{batch} through {stop or 'end'}: median {t[len(t) // 2]:.3f} ms  min {t[0]:.3f}  max {t[-1]:.3f}")
