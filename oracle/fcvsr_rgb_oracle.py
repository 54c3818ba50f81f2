"""CPU oracle for the RGB FCVSR family: `FCVSR` / `FCVSR_S` of CVSR_train/arch/CVSR_freq_RGB.py:2135-2202 / :2059-2128.

TEST INFRASTRUCTURE ONLY (same rules as oracle/fcvsr_oracle.py).  Plain PyTorch restatement, every function citing the
reference lines (file = CVSR_train/arch/CVSR_freq_RGB.py); pinned to the unmodified reference by tests/golden/fcvsr_rgb_*.pt
(made by oracle/make_golden_rgb.py) in tests/test_oracle.py.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from .fcvsr_oracle import _ca, _conv, _pack_spec, sac, warp_bilinear

SD = Dict[str, torch.Tensor]
_IDEAL: Dict[int, torch.Tensor] = {}


def ideal_masks_1024(q: int) -> torch.Tensor:
    """Split_freq.generate_freq_mask(1024, 1024), mode 'ideal' :1493-1506: filled discs (cv2.circle) of radius
    ceil((i+1) * l / q), each minus all previous masks."""
    if q not in _IDEAL:
        import cv2
        n = 1024
        step = math.sqrt((n / 2) ** 2 + (n / 2) ** 2) / q
        prev: List[torch.Tensor] = []
        for i in range(q):
            pf = np.zeros((n, n))
            cv2.circle(pf, (n // 2, n // 2), math.ceil((i + 1) * step), (1), -1)
            g = torch.from_numpy(pf).float()
            for pm in prev:
                g = g - pm
            prev.append(g)
        _IDEAL[q] = torch.stack(prev, 0)
    return _IDEAL[q]


def split_freq_ideal(x: torch.Tensor, q: int) -> List[torch.Tensor]:
    """Split_freq.forward :1555-1572 with the ideal masks (bicubic torchvision Resize of the 1024^2 masks)."""
    from torchvision.transforms import Resize, functional as TF
    h, w = x.shape[-2:]
    m = Resize([h, w], interpolation=TF.InterpolationMode.BICUBIC)(ideal_masks_1024(q)).to(x.dtype)
    f = torch.fft.fftshift(torch.fft.fftn(x, dim=(2, 3)), dim=(2, 3))
    return [torch.fft.ifftn(torch.fft.ifftshift(f * m[j], dim=(2, 3)), dim=(2, 3)).real for j in range(q)]


def div_enh(sd: SD, name: str, x: torch.Tensor, before: List[torch.Tensor]) -> torch.Tensor:
    """DivEnh.forward :1585-1612.  The reference builds BOTH x_before_sum and ex_before_sum from x_before (:1597-1603), so the
    previously enhanced bands never enter."""
    a, b = sd[name + ".a"], sd[name + ".b"]
    gate = lambda t: torch.sigmoid(_conv(sd, name + ".Conv", t))  # noqa: E731
    if not before:
        o = x - x.mean(dim=(2, 3), keepdim=True)
        return _ca(sd, name + ".ca", 0.2 * a * gate(o) * x + b * x)
    s = _ca(sd, name + ".ca", torch.stack(before, 0).sum(0))
    o = x - s + 0.2 * s
    return _ca(sd, name + ".ca", 0.2 * a * gate(o) * x + b * x) + _ca(sd, name + ".ca", 0.2 * a * gate(s) * x + b * x)


def mffr(sd: SD, x: torch.Tensor, q: int, p: str = "MFFRblock") -> torch.Tensor:
    """MultiFreq_Refinment.forward :1637-1655 (no skip connection in this family)."""
    bands = split_freq_ideal(x, q)[::-1]
    outs = [div_enh(sd, f"{p}.DivEnh_block.{i}", bands[i], bands[:i]) for i in range(q)]
    return _ca(sd, p + ".ca", torch.stack(outs, 0).sum(0))


def conv_blk(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ConvBlk :298-310: CA(conv2(PReLU(conv1(x)))) * res_scale (1.0)."""
    t = F.prelu(_conv(sd, name + ".conv1", x), sd[name + ".relu.weight"])
    return _ca(sd, name + ".CA", _conv(sd, name + ".conv2", t))


def iac(feat_in: torch.Tensor, pred_k: torch.Tensor, offsets: List[torch.Tensor], n_iter: int) -> torch.Tensor:
    """IAC :1009-1023: SAC(flow_warp(feat)) + F_bs[i], LeakyReLU 0.1 after every iteration (is_act_last = True)."""
    c = feat_in.shape[1]
    feat = feat_in
    for i in range(n_iter):
        taps = pred_k[:, i * 6 * c: i * 6 * c + 3 * c]           # F1; F2 is unused by SAC (:1041)
        bias = pred_k[:, n_iter * 6 * c + i * c: n_iter * 6 * c + (i + 1) * c]
        feat = F.leaky_relu(sac(warp_bilinear(feat, offsets[i]), taps) + bias, 0.1)
    return feat


def mgaa(sd: SD, x: torch.Tensor, n_iter: int, p: str = "MGAA") -> torch.Tensor:
    """MGAA.forward :1101-1180."""
    d = x.shape[1] // 3
    h, w = x.shape[-2:]
    x1, x2, x3 = x[:, :d], x[:, d:2 * d], x[:, 2 * d:]
    f1, f2, f3 = (_pack_spec(torch.fft.rfft2(t, norm="backward")) for t in (x1, x2, x3))
    fuse = lambda t: F.conv2d(F.relu(F.conv2d(t, sd[p + ".convfuse.0.weight"])), sd[p + ".convfuse.2.weight"])  # noqa: E731
    off_f = (f1 - f2) + fuse(torch.cat([f1, f2], 1))
    off_b = (f3 - f2) + fuse(torch.cat([f3, f2], 1))
    sim = F.conv2d(F.relu(F.conv2d(f2, sd[p + ".convcrt.0.weight"])), sd[p + ".convcrt.2.weight"])
    offs_f, offs_b = [], []
    for i in range(n_iter):
        for src, dst in ((off_f, offs_f), (off_b, offs_b)):
            o = conv_blk(sd, f"{p}.MConvB.{i}", src) * sim
            dst.append(torch.fft.irfft2(torch.complex(o[:, 0:2], o[:, 2:4]), s=(h, w), norm="backward"))
    pred_k = _conv(sd, p + ".F.1", _conv(sd, p + ".F.0", _conv(sd, p + ".conv_KP", x2)))
    al_f = iac(x1, pred_k, offs_f, n_iter)
    al_b = iac(x3, pred_k, offs_b, n_iter)
    return _conv(sd, p + ".conv3", torch.cat([al_f, al_b], 1))


def block(sd: SD, name: str, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """Block.forward :648-657."""
    res = [_conv(sd, name + ".body.2", F.leaky_relu(_conv(sd, name + ".body.0", x), 0.1)) for x in xs]
    down = [res[0]] + [F.interpolate(_conv(sd, name + ".down.0", r), scale_factor=0.5, mode="bilinear", align_corners=False)
                       for r in res[:-1]]
    up = [F.interpolate(_conv(sd, name + ".up.0", r), scale_factor=2.0, mode="bilinear", align_corners=False) for r in res[1:]] + [res[-1]]
    return [x + r + d + u for x, r, d, u in zip(xs, res, down, up)]


def scnet(sd: SD, xs: List[torch.Tensor], n_groups: int, p: str = "recorb1") -> List[torch.Tensor]:
    """SCNet :685-700 / SCGroup :659-683."""
    cur = xs
    for g in range(n_groups):
        t = cur
        for k in range(3):
            t = block(sd, f"{p}.body.{g}.body.{k}", t)
        cur = [x + _conv(sd, f"{p}.body.{g}.conv", r) for x, r in zip(cur, t)]
    return [x + r for x, r in zip(xs, cur)]


def forward(sd: SD, x: torch.Tensor, return_taps: bool = False):
    """FCVSR.forward :2170-2202 / FCVSR_S.forward :2096-2128."""
    n_iter = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("MGAA.MConvB."))
    q = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("MFFRblock.DivEnh_block."))
    n_groups = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("recorb1.body."))
    b, t, c, h, w = x.shape
    n = sd["recorb0.weight"].shape[0]
    taps = {}
    feats = _conv(sd, "feat_extract.0", x.reshape(b, t * c, h, w))
    f1, f2, f3 = feats[:, :3 * n], feats[:, 3 * n:4 * n], feats[:, 4 * n:]
    o1 = mgaa(sd, f1, n_iter)
    o3 = mgaa(sd, f3, n_iter)
    o2 = mgaa(sd, torch.cat([o1, f2, o3], 1), n_iter)
    taps["mgaa1"], taps["mgaa2"] = o1, o2
    l1 = mffr(sd, o2, q)
    taps["mffr"] = l1
    l2 = _conv(sd, "rconcat1", l1, stride=2)
    l3 = _conv(sd, "rconcat2", l2, stride=2)
    o_l1, o_l2, o_l3 = scnet(sd, [l1, l2, l3], n_groups)
    taps["sc_l1"], taps["sc_l3"] = o_l1, o_l3
    pw = sd["lrelu.weight"]
    u3_1 = F.pixel_shuffle(F.prelu(_conv(sd, "upconv1_L3", o_l3), pw), 2)
    u3_2 = F.pixel_shuffle(u3_1, 2)
    u2 = F.prelu(_conv(sd, "upconv1_L2", o_l2), pw)
    u2 = F.pixel_shuffle(u2 + _conv(sd, "upconv1_L2_2", torch.cat([u2, u3_1], 1)), 2)
    fuse = _conv(sd, "recorb0", _conv(sd, "upconv_fuse", torch.cat([o_l1, u2, u3_2], 1)))
    taps["fuse"] = fuse
    y = F.prelu(F.pixel_shuffle(_conv(sd, "upconv1", fuse), 2), pw)
    y = F.prelu(F.pixel_shuffle(_conv(sd, "upconv2", y), 2), pw)
    y = _conv(sd, "conv_last0", y) + F.interpolate(x[:, t // 2], scale_factor=4, mode="bilinear")
    return (y, taps) if return_taps else y
