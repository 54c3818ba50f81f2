"""Time the fused modulated-DCN backward (csrc/dcn_bwd.cu) on the SURVEY 8(a10) benchmark shape, kernel by kernel
(data = grad_input + grad_offset + grad_mask; weight = grad_weight; bias)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timed(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (B, H, W) in ((1, 180, 320), (4, 180, 320)):
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(B, 64, H, W, generator=g).to(dev)
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
    off = (2.0 * torch.randn(B, 288, H, W, generator=g)).to(dev)
    msk = torch.rand(B, 144, H, W, generator=g).to(dev)
    gy = torch.randn(B, 64, H, W, generator=g).to(dev)
    gx, goff, gm, gw, gb = (torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(msk), torch.zeros_like(w),
                            torch.zeros(64, device=dev))

    def run(gx_, gw_, gb_, goff_, gm_):
        C.call("fcvsr_modulated_deform_conv_backward", x.data_ptr(), w.data_ptr(), off.data_ptr(), msk.data_ptr(),
               gy.data_ptr(), gx_, gw_, gb_, goff_, gm_, B, 64, H, W, 64, 3, 3, 1, 1, 1, 1, 1, 1, 1, 16, SCR, st)

    fl = 2.0 * B * H * W * 64 * 64 * 9
    scratch = torch.empty(2 * x.numel(), device=dev)
    for SCR, tag in ((0, "NCHW scalar atomics"), (scratch.data_ptr(), "NHWC + red.v4")):
      t_data = timed(lambda: run(gx.data_ptr(), 0, 0, goff.data_ptr(), gm.data_ptr()))
      t_w = timed(lambda: run(0, gw.data_ptr(), 0, 0, 0))
      t_b = timed(lambda: run(0, 0, gb.data_ptr(), 0, 0))
      print(f"DCN backward [{tag}] B{B} 64->64 3x3 dg16 {H}x{W}: data {t_data:8.1f} us ({fl / t_data / 1e6:5.2f} TFLOP/s)  "
            f"weight {t_w:8.1f} us ({fl / t_w / 1e6:5.2f} TFLOP/s)  bias {t_b:6.1f} us")
