"""On-GPU evaluation metrics adjacent to the output: Y-channel PSNR / SSIM with a cropped border.

Restates what the reference's evaluation driver does on the host after writing PNG files
(CVSR_train/metric/psnr_ssim.py: `calculate_psnr(res, gt, 4, test_y_channel=True)` / `calculate_ssim(...)` :447-478 on
single-channel uint8 frames): the frames stay on the device, `fcvsr_psnr_ssim_u8` (csrc/metrics.cu) reduces them to one
(PSNR, SSIM) pair per frame.  `fcvsr_b200.sequence.quantize_u8` produces the uint8 frames exactly as the driver's
clamp -> *255 -> astype(uint8) does.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _capi as C


def psnr_ssim_u8(sr: torch.Tensor, gt: torch.Tensor, crop_border: int = 4) -> Tuple[torch.Tensor, torch.Tensor]:
    """sr, gt: uint8 CUDA tensors [..., H, W] of equal shape (each leading index is one single-channel frame).
    Returns (psnr, ssim) float32 tensors of the leading shape; PSNR is +inf for identical frames."""
    if sr.shape != gt.shape:
        raise ValueError(f"Image shapes are different: {tuple(sr.shape)}, {tuple(gt.shape)}.")     # psnr_ssim.py:296
    if not (sr.is_cuda and gt.is_cuda):
        raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
    if sr.dtype != torch.uint8 or gt.dtype != torch.uint8:
        raise TypeError("psnr_ssim_u8 expects uint8 frames (see fcvsr_b200.sequence.quantize_u8)")
    h, w = sr.shape[-2:]
    lead = sr.shape[:-2]
    n = 1
    for d in lead:
        n *= d
    sr, gt = sr.contiguous(), gt.contiguous()
    hc, wc = h - 2 * crop_border, w - 2 * crop_border
    scratch = torch.empty(n * ((hc + 15) // 16) * ((wc + 15) // 16) * 2, device=sr.device, dtype=torch.float64)
    out = torch.empty(n, 2, device=sr.device, dtype=torch.float32)
    with torch.cuda.device(sr.device):
        C.call("fcvsr_psnr_ssim_u8", sr.data_ptr(), gt.data_ptr(), n, h, w, crop_border, scratch.data_ptr(), out.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
    return out[:, 0].reshape(lead), out[:, 1].reshape(lead)
