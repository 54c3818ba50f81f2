"""Host-side constants of the frequency kernels: band masks and FFT twiddles.

Band masks restate Split_freq.generate_freq_mask / forward (CVSR_train/arch/CVSR_freq.py:2016-2051,
:2078): Q gaussian difference masks on a 1024x1024 grid, resized to (H, W) with torchvision's bicubic
``Resize`` (the reference calls it with torchvision defaults, so the antialias behaviour is whatever
the installed torchvision does -- the same call is made here on purpose).  The device kernels apply
them to the half spectrum, which needs the Hermitian-symmetrised mask
    Msym(k) = (M(k) + M(-k)) / 2,   M = ifftshift(resized mask)
because Re ifft2(F * M) == irfft2(rfft2(x) * Msym) for real x.
Computed once per (Q, H, W) and cached on the device.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

_MASK1024: Dict[int, torch.Tensor] = {}
_MASKS: Dict[Tuple[int, int, int, str, str], torch.Tensor] = {}
_TW: Dict[Tuple[int, str], torch.Tensor] = {}


def _gaussian_masks_1024(q: int) -> torch.Tensor:
    if q not in _MASK1024:
        n = 1024
        step = math.sqrt(2.0 * (n / 2) ** 2) / q
        ax = (np.arange(n) - n // 2).astype(np.float64) ** 2
        r2 = ax[:, None] + ax[None, :]
        out, acc = [], None
        for i in range(q):
            g = torch.from_numpy(np.exp(-np.power(np.sqrt(r2), 2) / (2.0 * (step * (i + 1)) ** 2))).float()
            for prev in out:
                g = g - prev
            out.append(g)
        _MASK1024[q] = torch.stack(out, 0)
    return _MASK1024[q]


_IDEAL1024: Dict[int, torch.Tensor] = {}


def _ideal_masks_1024(q: int) -> torch.Tensor:
    """'ideal' mode of the RGB family (CVSR_freq_RGB.py:1493-1506): filled discs drawn with cv2.circle -- the same call is made
    here on purpose (its rasterisation is part of the reference's result), so this mode needs OpenCV on the host."""
    if q not in _IDEAL1024:
        try:
            import cv2
        except ImportError as e:           # pragma: no cover
            raise RuntimeError("the 'ideal' band masks of FCVSR / FCVSR_S are rasterised with cv2.circle as in the reference; "
                               "OpenCV (cv2) is required for these two classes") from e
        n = 1024
        step = math.sqrt(2.0 * (n / 2) ** 2) / q
        out = []
        for i in range(q):
            pf = np.zeros((n, n))
            cv2.circle(pf, (n // 2, n // 2), math.ceil((i + 1) * step), (1), -1)
            g = torch.from_numpy(pf).float()
            for prev in out:
                g = g - prev
            out.append(g)
        _IDEAL1024[q] = torch.stack(out, 0)
    return _IDEAL1024[q]


def symmetric_half_masks(q: int, h: int, w: int, device, mode: str = "gaussian") -> torch.Tensor:
    """[Q, H, W/2+1] float32 on `device`."""
    key = (q, h, w, str(device), mode)
    if key not in _MASKS:
        from torchvision.transforms import Resize, functional as TF
        base = _gaussian_masks_1024(q) if mode == "gaussian" else _ideal_masks_1024(q)
        m = Resize([h, w], interpolation=TF.InterpolationMode.BICUBIC)(base)
        m = torch.fft.ifftshift(m, dim=(1, 2))
        neg = torch.roll(torch.flip(m, dims=(1, 2)), shifts=(1, 1), dims=(1, 2))   # M(-k)
        msym = 0.5 * (m + neg)
        _MASKS[key] = msym[:, :, : w // 2 + 1].contiguous().to(device)
    return _MASKS[key]


def twiddles(n: int, device) -> torch.Tensor:
    """float2[n] = exp(-2 pi i k / n), computed in float64."""
    key = (n, str(device))
    if key not in _TW:
        k = np.arange(n, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / n), -np.sin(2 * np.pi * k / n)], axis=1).astype(np.float32)
        _TW[key] = torch.from_numpy(tw).contiguous().to(device)
    return _TW[key]
