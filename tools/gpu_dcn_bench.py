"""Time the fused modulated-DCN forward on the SURVEY 8(a10) benchmark shape (x [B,64,H,W], dg = 16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200.ops.dcn import modulated_deform_conv  # noqa: E402

dev = torch.device("cuda:0")
for (B, H, W) in ((1, 180, 320), (4, 180, 320)):
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(B, 64, H, W, generator=g).to(dev)
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    off = (2.0 * torch.randn(B, 288, H, W, generator=g)).to(dev)
    msk = torch.rand(B, 144, H, W, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(3):
            modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16)
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    fl = 2.0 * B * H * W * 64 * 64 * 9
    by = 4.0 * B * H * W * (64 + 64 + 288 + 144)
    print(f"modulated DCN B{B} 64->64 3x3 dg16 {H}x{W}: {us:8.1f} us  {fl / us / 1e6:6.2f} TFLOP/s  "
          f"{by / us / 1e3:7.1f} GB/s of algorithmic traffic")
