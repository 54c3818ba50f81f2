"""Time the fused modulated-DCN forward on the SURVEY 8(a10) benchmark shape (x [B,64,H,W], dg = 16): this library's operator
(tcgen05 and exact fp32 kernels) and ModulatedDeformConvPack next to torchvision.ops.deform_conv2d -- the operator the reference's
own alignment module calls (arch/SIDECVSR_J_L_fast_3x3_our.py:1783) -- on the same box, and the backward of both."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fcvsr_b200.ops.dcn as D  # noqa: E402
from fcvsr_b200.ops.dcn import modulated_deform_conv  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


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

    # ---- same-box anchors -------------------------------------------------------------------------------------------------
    try:
        import torchvision.ops as tvo
        with torch.no_grad():
            us_tv = timeit(lambda: tvo.deform_conv2d(x, off, w, b, padding=1, mask=msk))
            yt = tvo.deform_conv2d(x, off, w, b, padding=1, mask=msk)
            yo = modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16)
        print(f"   torchvision.ops.deform_conv2d (im2col + cuBLAS): {us_tv:8.1f} us  -> ratio {us_tv / us:5.2f}x; "
              f"max |ours - torchvision| = {float((yo - yt).abs().max()):.2e} (TF32 operands here, fp32 there)")
    except Exception as e:          # torchvision CUDA ops missing on this image
        print("   torchvision.ops.deform_conv2d unavailable:", type(e).__name__, e)
    D.PRECISION = "fp32"
    with torch.no_grad():
        us32 = timeit(lambda: modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 16))
    D.PRECISION = "tf32"
    print(f"   exact fp32 kernel (dcn.cu): {us32:8.1f} us")
    pack = D.ModulatedDeformConvPack(64, 64, 3, stride=1, padding=1, deformable_groups=16).to(dev)
    with torch.no_grad():
        pack.conv_offset_mask.weight.normal_(0, 0.02)
        pack.conv_offset_mask.bias.normal_(0, 0.5)
        us_pack = timeit(lambda: pack(x))
        D.PRECISION = "fp32"
        us_pack32 = timeit(lambda: pack(x), n=3, warm=1)
        D.PRECISION = "tf32"
    print(f"   ModulatedDeformConvPack (conv_offset_mask 64->432 + sigmoid + DCN): tensor-core path {us_pack:8.1f} us, "
          f"CUDA-core path {us_pack32:8.1f} us")
    # backward
    xg, og, mg, wg, bg = (t.clone().requires_grad_(True) for t in (x, off, msk, w, b))
    gy = torch.randn(B, 64, H, W, device=dev)

    def bwd_ours():
        y = modulated_deform_conv(xg, og, mg, wg, bg, 1, 1, 1, 1, 16)
        y.backward(gy)

    us_fb = timeit(bwd_ours, n=5)
    try:
        import torchvision.ops as tvo

        def bwd_tv():
            y = tvo.deform_conv2d(xg, og, wg, bg, padding=1, mask=mg)
            y.backward(gy)

        us_fb_tv = timeit(bwd_tv, n=5)
        print(f"   forward + backward (all five gradients): ours {us_fb:8.1f} us, torchvision {us_fb_tv:8.1f} us -> {us_fb_tv / us_fb:5.2f}x")
    except Exception as e:
        print(f"   forward + backward ours {us_fb:8.1f} us; torchvision unavailable: {type(e).__name__}")
