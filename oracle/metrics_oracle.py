"""CPU restatement of the reference's evaluation metrics (TEST INFRASTRUCTURE ONLY; see oracle/fcvsr_oracle.py for the rules).

CVSR_train/metric/psnr_ssim.py: calculate_psnr :278-316, _ssim :318-350, calculate_ssim :353-399, as the evaluation driver
calls them (:447-478): single-channel uint8 frames as float64 [H,W,1], crop_border = 4, test_y_channel = True (to_y_channel
:201-214 is `float32(img) / 255 * 255` for one channel).  cv2.filter2D with the 11 x 11 Gaussian window is restated as a
separable float64 correlation over the valid region (the reference only keeps [5:-5, 5:-5], so the border mode is irrelevant).
`tests/test_oracle.py` pins this file to the live reference functions when /root/reference is present.
"""
from __future__ import annotations

import numpy as np


def _to_y(img: np.ndarray) -> np.ndarray:
    return (img.astype(np.float32) / 255.0) * 255.0           # psnr_ssim.py:210-214, single channel


def gaussian_kernel_11() -> np.ndarray:
    """cv2.getGaussianKernel(11, 1.5)"""
    k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2.0 * 1.5 ** 2))
    return k / k.sum()


def calculate_psnr(img1: np.ndarray, img2: np.ndarray, crop_border: int = 4) -> float:
    if crop_border:
        img1 = img1[crop_border:-crop_border, crop_border:-crop_border]
        img2 = img2[crop_border:-crop_border, crop_border:-crop_border]
    a, b = _to_y(img1).astype(np.float64), _to_y(img2).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else float(20.0 * np.log10(255.0 / np.sqrt(mse)))


def _filt(x: np.ndarray, g: np.ndarray) -> np.ndarray:
    """valid-region separable correlation: out[i, j] = sum_{u,v} g[u] g[v] x[i+u, j+v]"""
    h, w = x.shape
    t = sum(g[k] * x[:, k:w - 10 + k] for k in range(11))
    return sum(g[k] * t[k:h - 10 + k, :] for k in range(11))


def calculate_ssim(img1: np.ndarray, img2: np.ndarray, crop_border: int = 4) -> float:
    if crop_border:
        img1 = img1[crop_border:-crop_border, crop_border:-crop_border]
        img2 = img2[crop_border:-crop_border, crop_border:-crop_border]
    a, b = _to_y(img1).astype(np.float64), _to_y(img2).astype(np.float64)
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    g = gaussian_kernel_11()
    mu1, mu2 = _filt(a, g), _filt(b, g)
    s1, s2, s12 = _filt(a * a, g) - mu1 ** 2, _filt(b * b, g) - mu2 ** 2, _filt(a * b, g) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))
    return float(m.mean())
