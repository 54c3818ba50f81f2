"""Forward of the RGB FCVSR family (`FCVSR` / `FCVSR_S`, CVSR_train/arch/CVSR_freq_RGB.py:2059-2202) on the kernel library.

Built from the same autograd-capable operators as `fcvsr_b200.train_forward` -- convolutions (tcgen05 where the shape fits,
CUDA cores otherwise: the dense 128 -> 64 -> 4 offset blocks with kernels up to 11 x 11 run there), the 2-D real FFTs,
flow_warp and SAC are this repository's kernels with backward kernels; activations, residual sums, channel attention and
resampling are PyTorch elementwise glue -- so one code path serves inference (under torch.no_grad()) and training.  It is not
CUDA-graph'ed or fused like the Y-channel engine (fcvsr_b200.engine): this family is SURVEY 8(f1), built for API and result
parity first.

Line numbers refer to CVSR_train/arch/CVSR_freq_RGB.py.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from . import autograd as A
from . import bands
from .train_forward import _ca, _cl, _Ctx, _index_tensors


def _mgaa(cx: _Ctx, x: torch.Tensor) -> torch.Tensor:
    """MGAA.forward :1101-1180 on x [B,192,H,W]."""
    mg, n, acn = cx.m.MGAA, cx.m.n_feats, cx.m.ACNum
    B, _, H, W = x.shape
    inv, rows = _index_tensors(x.device, n, acn, with_bias_rows=True)
    x1, x2, x3 = x[:, :n], x[:, n:2 * n], x[:, 2 * n:]
    spec = A.rfft2(x)                                              # interleaved, group g at channels [128g, 128g+128)
    s1, s2, s3 = spec[:, :2 * n], spec[:, 2 * n:4 * n], spec[:, 4 * n:]
    w0 = mg.convfuse[0].weight                                     # cat([imag, real]) packing (:1118-1127) = weight permutation
    w0 = torch.cat([w0[:, :2 * n][:, inv], w0[:, 2 * n:][:, inv]], 1)
    w2 = mg.convfuse[2].weight[inv]

    def fuse(sa):
        return cx.conv(F.relu(cx.conv(_cl(torch.cat([sa, s2], 1)), w0)), w2) + (sa - s2)      # :1132-1133

    off_f, off_b = fuse(s1), fuse(s3)
    sim = cx.conv(F.relu(cx.conv(_cl(s2), mg.convcrt[0].weight[:, inv])), mg.convcrt[2].weight)   # :1134
    zs = []
    for i in range(acn):                                           # ConvBlk_i * x2_f_sim (:1137-1150), dense k = 2i+1 kernels
        blk = mg.MConvB[i]
        w1 = blk.conv1.weight[:, inv]
        for o in (off_f, off_b):
            t = cx.conv(F.prelu(cx.conv(_cl(o), w1), blk.relu.weight), blk.conv2.weight)
            v = _ca(blk.CA, t) * sim
            zs.append(torch.stack([v[:, 0], v[:, 2], v[:, 1], v[:, 3]], 1))      # complex(v[0:2], v[2:4]) interleaved
    offs = A.irfft2(_cl(torch.cat(zs, 1)), W)                      # channel (i*2+dir)*2 + (dx, dy)
    # kernel predictor (:1152-1153): live tap rows i*384 + c*3 + t re-ordered to [i][t][c], bias rows A*384 + i*64 + c
    kp = cx.conv(cx.conv(_cl(x2), mg.conv_KP), mg.F[0])
    pred = cx.conv(kp, mg.F[1].weight[rows], mg.F[1].bias[rows])   # [B, A*192 + A*64, H, W]
    taps, fbs = pred[:, :acn * 3 * n], pred[:, acn * 3 * n:]
    aligned = []
    for d, xin in enumerate((x1, x3)):                             # IAC (:1009-1023)
        feat = xin
        for i in range(acn):
            ch = (i * 2 + d) * 2
            wp = A.flow_warp(feat, offs[:, ch:ch + 2])
            feat = F.leaky_relu(A.sac(wp, taps[:, i * 3 * n:(i + 1) * 3 * n]) + fbs[:, i * n:(i + 1) * n], 0.1)
        aligned.append(feat)
    return cx.conv(_cl(torch.cat(aligned, 1)), mg.conv3)           # :1179 (no skip in this family)


def _mffr(cx: _Ctx, x: torch.Tensor) -> torch.Tensor:
    """MultiFreq_Refinment.forward :1637-1655 with Split_freq 'ideal' (:1555-1572) and DivEnh (:1585-1612)."""
    mf, q = cx.m.MFFRblock, cx.m.mffr_bands
    B, c, H, W = x.shape
    masks = bands.symmetric_half_masks(q, H, W, x.device, mode="ideal")
    spec = A.rfft2(x)
    stacked = torch.cat([spec * masks[j][None, None] for j in range(q)], 0)
    bnds = list(A.irfft2(_cl(stacked), W).split(B, 0))[::-1]
    outs: List[torch.Tensor] = []
    sb = None
    for i in range(q):
        de, xb = mf.DivEnh_block[i], bnds[i]
        a, b = de.a[None], de.b[None]
        gate = lambda t: torch.sigmoid(cx.conv(_cl(t), de.Conv))  # noqa: E731
        if i == 0:
            out = _ca(de.ca, 0.2 * a * gate(xb - xb.mean(dim=(2, 3), keepdim=True)) * xb + b * xb)
            sb = xb
        else:
            s = _ca(de.ca, sb)               # the reference forms both sums from x_before (:1597-1603)
            o = xb - s + 0.2 * s
            out = _ca(de.ca, 0.2 * a * gate(o) * xb + b * xb) + _ca(de.ca, 0.2 * a * gate(s) * xb + b * xb)
            sb = sb + xb
        outs.append(out)
    return _ca(mf.ca, torch.stack(outs, 0).sum(0))


def _block(cx: _Ctx, blk, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """Block.forward :648-657."""
    res = [cx.conv(F.leaky_relu(cx.conv(_cl(x), blk.body[0]), 0.1), blk.body[2]) for x in xs]
    down = [res[0]] + [cx.conv(_cl(F.avg_pool2d(r, 2)), blk.down[0]) for r in res[:-1]]     # 2x2 mean commutes with the 1x1 conv
    up = [F.interpolate(cx.conv(r, blk.up[0]), scale_factor=2.0, mode="bilinear", align_corners=False) for r in res[1:]] + [res[-1]]
    return [x + r + d + u for x, r, d, u in zip(xs, res, down, up)]


def forward_rgb(model, x: torch.Tensor, mode: str = "tf32") -> torch.Tensor:
    """x [B,7,3,H,W] -> [B,3,4H,4W] (FCVSR.forward :2170-2202 / FCVSR_S.forward :2096-2128)."""
    cx = _Ctx(model, mode)
    m, n = model, model.n_feats
    b, t, c, h, w = x.shape
    feats = cx.conv(_cl(x.reshape(b, t * c, h, w)), m.feat_extract[0])
    f1, f2, f3 = feats[:, :3 * n], feats[:, 3 * n:4 * n], feats[:, 4 * n:]
    o1 = _mgaa(cx, _cl(f1))
    o3 = _mgaa(cx, _cl(f3))
    o2 = _mgaa(cx, _cl(torch.cat([o1, f2, o3], 1)))
    l1 = _mffr(cx, o2)
    l2 = cx.conv(_cl(l1), m.rconcat1)
    l3 = cx.conv(l2, m.rconcat2)
    xs = [l1, l2, l3]
    cur = xs
    for grp in m.recorb1.body:                                      # SCNet :685-700 / SCGroup :659-683
        tt = cur
        for blk in grp.body:
            tt = _block(cx, blk, tt)
        cur = [xx + cx.conv(_cl(r), grp.conv) for xx, r in zip(cur, tt)]
    o_l1, o_l2, o_l3 = [xx + r for xx, r in zip(xs, cur)]
    pw = m.lrelu.weight
    u3_1 = F.pixel_shuffle(F.prelu(cx.conv(_cl(o_l3), m.upconv1_L3), pw), 2)
    u3_2 = F.pixel_shuffle(u3_1, 2)
    u2 = F.prelu(cx.conv(_cl(o_l2), m.upconv1_L2), pw)
    u2 = F.pixel_shuffle(u2 + cx.conv(_cl(torch.cat([u2, u3_1], 1)), m.upconv1_L2_2), 2)
    fuse = cx.conv(cx.conv(_cl(torch.cat([o_l1, u2, u3_2], 1)), m.upconv_fuse), m.recorb0)
    y = F.prelu(F.pixel_shuffle(cx.conv(fuse, m.upconv1), 2), pw)
    y = F.prelu(F.pixel_shuffle(cx.conv(_cl(y), m.upconv2), 2), pw)
    y = cx.conv(_cl(y), m.conv_last0)
    return y + F.interpolate(x[:, t // 2], scale_factor=4, mode="bilinear", align_corners=False)
