"""Bring-up helper: time conv_tc on the trunk shapes (CUDA events, 20 reps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C  # noqa: E402
from fcvsr_b200.engine import _ConvPack  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
shapes = [(1, 64, 64, 180, 320, 3), (1, 64, 128, 180, 320, 3), (1, 128, 64, 180, 320, 3), (1, 64, 64, 90, 160, 3),
          (1, 64, 64, 45, 80, 3), (1, 64, 64, 180, 320, 1), (1, 64, 1152, 180, 320, 1), (1, 64, 256, 360, 640, 3),
          (4, 64, 64, 180, 320, 3), (4, 64, 128, 180, 320, 3), (4, 128, 64, 180, 320, 3), (4, 64, 256, 360, 640, 3),
          (4, 64, 64, 90, 160, 3), (8, 128, 128, 180, 161, 1), (8, 128, 128, 180, 160, 1), (4, 256, 128, 180, 161, 1),
          (4, 64, 64, 180, 320, 1), (4, 64, 64, 90, 160, 1)]
print("FCVSR_TC_DBG =", os.environ.get("FCVSR_TC_DBG", "0"))
for (B, ci, co, H, W, k) in shapes:
    x = torch.randn(B, H, W, ci, device=dev)
    w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
    bvec = torch.randn(co, device=dev)
    bp = bvec.data_ptr() if os.environ.get("BIAS") else 0
    act = 2 if os.environ.get("BIAS") else 0
    pk = _ConvPack(w, None)
    y = torch.empty(B, H, W, co, device=dev)
    args = (x.data_ptr(), ci, pk.w_tc.data_ptr(), bp, 0, 0, 0, 0, y.data_ptr(), co, B, H, W, ci, co, k, act, 0.1, 0, 0, 0, 0, 0, 0, 0, st)
    for _ in range(3):
        C.call("fcvsr_conv2d_tc", *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        C.call("fcvsr_conv2d_tc", *args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * B * H * W * ci * co * k * k
    line = f"B{B} {ci}->{co} {H}x{W} k{k}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s"
    # bf16 operands
    pk16 = _ConvPack(w, None, op16=True)
    x16 = x.to(torch.bfloat16)
    y16 = torch.empty(B, H, W, co, device=dev, dtype=torch.bfloat16)
    a3 = (x16.data_ptr(), ci, pk16.w_tc.data_ptr(), bp, 0, 0, 0, 0, y16.data_ptr(), co, B, H, W, ci, co, k, act, 0.1, 0, 0, 0, 0, 1, 0, 1, st)
    if ci % 64 == 0:
        for _ in range(3):
            C.call("fcvsr_conv2d_tc", *a3)
        e0.record()
        for _ in range(20):
            C.call("fcvsr_conv2d_tc", *a3)
        e1.record()
        torch.cuda.synchronize()
        us3 = e0.elapsed_time(e1) / 20 * 1e3
        line += f"   | bf16: {us3:8.1f} us  {fl / us3 / 1e6:7.1f} TFLOP/s"
    print(line)
