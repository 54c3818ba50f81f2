"""Shared helpers of the test-suite (clip synthesis, golden loading, NHWC conversion)."""
import os

import torch

from oracle.make_golden import make_clip, make_clip_rgb  # noqa: F401  (same seeded clips as the golden generator)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def nhwc(t):      # [B,C,H,W] -> [B,H,W,C] contiguous
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):      # [B,H,W,C] -> [B,C,H,W]
    return t.permute(0, 3, 1, 2).contiguous()


def psnr(a, b):
    """metric/psnr_ssim.py:314-316 style PSNR on [0,1]-clamped images, peak 1."""
    mse = torch.mean((a.clamp(0, 1) - b.clamp(0, 1)) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()
