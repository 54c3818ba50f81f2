"""Time the non-GEMM SCNet kernels per pyramid level (B=4, 180x320 base) and report achieved HBM GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for (H, W) in ((180, 320), (90, 160), (46, 80)):
    P = H * W
    f = lambda *s, dt=torch.float32: torch.randn(*s, device=dev).to(dt)
    res, r0, rr, xin, xout = f(B, P, 64), f(B, P, 64), f(B, P, 64), f(B, P, 64), f(B, P, 64)
    rrh = f(B, P, 64, dt=torch.bfloat16)
    tr = f(B, P, 64, dt=torch.bfloat16)
    td = f(B, 4 * P, 64)
    tu = f(B, P // 4, 64)
    add = f(B, 64)
    wm, w1, w2 = f(64), f(64, 64), f(64, 64)
    part = torch.empty(B * ((P + 127) // 128) * 66, device=dev)
    cnt = torch.zeros(B, device=dev, dtype=torch.int32)
    MB = B * P * 64 * 4 / 1e6
    t = timeit(lambda: C.call("fcvsr_context_block", res.data_ptr(), 64, wm.data_ptr(), w1.data_ptr(), w2.data_ptr(), part.data_ptr(), add.data_ptr(), cnt.data_ptr(), B, P, 0, st))
    line = f"B{B} {H}x{W}: context {t:6.1f} us {MB / t * 1e3:6.0f} GB/s"
    t = timeit(lambda: C.call("fcvsr_rcb_finish", res.data_ptr(), add.data_ptr(), r0.data_ptr(), rr.data_ptr(), B, P, rrh.data_ptr(), 1, 0, 0, 0, 0, 0, st))
    line += f" | rcb_finish {t:6.1f} us {3.5 * MB / t * 1e3:6.0f} GB/s"
    t = timeit(lambda: C.call("fcvsr_level_mix", xin.data_ptr(), 64, xout.data_ptr(), 64, rr.data_ptr(), 1.0, td.data_ptr(), tu.data_ptr(), B, H, W, tr.data_ptr(), 64, 0, 1, 0, st))
    line += f" | level_mix(td+tu) {t:6.1f} us {(3.5 + 4 + 0.25) * MB / t * 1e3:6.0f} GB/s"
    t = timeit(lambda: C.call("fcvsr_level_mix", xin.data_ptr(), 64, xout.data_ptr(), 64, rr.data_ptr(), 2.0, 0, tu.data_ptr(), B, H, W, tr.data_ptr(), 64, 0, 1, 0, st))
    line += f" | level_mix(tu) {t:6.1f} us {(3.5 + 0.25) * MB / t * 1e3:6.0f} GB/s"
    print(line)
