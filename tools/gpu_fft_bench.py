"""Time the FFT passes on the MGAA / MFFR shapes and report achieved HBM GB/s (read + write of the tensor)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C, bands  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (B, H, W, Cc) in ((1, 180, 320, 192), (4, 180, 320, 192), (4, 180, 320, 64), (4, 180, 320, 24), (1, 272, 480, 192)):
    Wf = W // 2 + 1
    x = torch.randn(B, H, W, Cc, device=dev)
    spec = torch.empty(B, H, Wf, Cc, 2, device=dev)
    y = torch.empty(B, H, W, Cc, device=dev)
    tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
    t1 = timeit(lambda: C.call("fcvsr_fft_r2c_w", x.data_ptr(), Cc, spec.data_ptr(), tw_w.data_ptr(), B, H, W, Cc, st))
    t2 = timeit(lambda: C.call("fcvsr_fft_c2c_h", spec.data_ptr(), spec.data_ptr(), tw_h.data_ptr(), 0, B, H, Wf, Cc, 0, 1.0, 0, 1, 0, st))
    t3 = timeit(lambda: C.call("fcvsr_fft_c2r_w", spec.data_ptr(), y.data_ptr(), Cc, tw_w.data_ptr(), B, H, W, Cc, 1.0, st))
    b_real, b_spec = x.numel() * 4, spec.numel() * 4
    print(f"B{B} {H}x{W} C={Cc}: r2c_w {t1:7.1f} us {(b_real + b_spec) / t1 / 1e3:7.0f} GB/s | c2c_h {t2:7.1f} us {2 * b_spec / t2 / 1e3:7.0f} GB/s | "
          f"c2r_w {t3:7.1f} us {(b_real + b_spec) / t3 / 1e3:7.0f} GB/s")
