"""Differentiable FCVSR forward: GShiftNet.forward (CVSR_train/arch/CVSR_freq.py:2688-2756) for the training step.

Same arithmetic as `engine.Engine` (which it is parity-tested against), organised for autograd instead of for a CUDA graph:
the convolutions, FFTs, CorrBlock lookup, flow_warp and SAC are the kernel-library Functions of `fcvsr_b200.autograd` (forward
and backward kernels of this repository), the elementwise glue between them (activations, residual sums, channel attention,
soft-max pooling, pixel shuffles, bilinear resampling) is PyTorch, so autograd derives the rest.  Parameters are read from the
module itself, so `loss.backward()` fills `p.grad` exactly like the reference: the 2 x Freq_Inv `DivEnh.Conv` parameters get
no gradient and the dead half of `MGAA.F.1` gets exact zeros.

Line numbers refer to CVSR_train/arch/CVSR_freq.py.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from . import autograd as A
from . import bands


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous(memory_format=torch.channels_last)


def _mlp_vec(v: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, act) -> torch.Tensor:
    """Two 1x1 convolutions on a [B, C] vector (the pooled branch of CALayer / ContextBlock) as broadcast-multiply-sums:
    no library GEMM for a handful of 4..64-wide mat-vecs."""
    h = act((v.unsqueeze(1) * w1.flatten(1).unsqueeze(0)).sum(2))
    return (h.unsqueeze(1) * w2.flatten(1).unsqueeze(0)).sum(2)


def _ca(mod, x: torch.Tensor) -> torch.Tensor:
    """CALayer :1812-1828: x * sigmoid(W2 relu(W1 mean_hw(x)))."""
    g = torch.sigmoid(_mlp_vec(x.mean(dim=(2, 3)), mod.conv_du[0].weight, mod.conv_du[2].weight, F.relu))
    return x * g[:, :, None, None]


_IDX = {}


def _index_tensors(device, n: int, acn: int, with_bias_rows: bool = False):
    """Constant index tensors on the device, built once (a host-to-device copy per forward would also be illegal inside a CUDA-graph
    capture of the training step): `inv` = spectrum-packing permutation, `rows` = live rows of MGAA.F.1 in [iteration][tap][channel]
    order (plus, for the RGB family, the per-iteration bias rows)."""
    key = (str(device), n, acn, with_bias_rows)
    if key not in _IDX:
        j = torch.arange(2 * n)
        pi = torch.where(j < n, 2 * j + 1, 2 * (j - n))          # reference xk_f channel j -> interleaved float index
        inv = torch.argsort(pi)                                    # interleaved index i <- reference channel inv[i]
        ii, tt, cc = torch.meshgrid(torch.arange(acn), torch.arange(3), torch.arange(n), indexing="ij")
        rows = (ii * 6 * n + cc * 3 + tt).reshape(-1)
        if with_bias_rows:
            rows = torch.cat([rows, acn * 6 * n + torch.arange(acn * n)])
        _IDX[key] = (inv.to(device), rows.to(device))
    return _IDX[key]


class _Ctx:
    def __init__(self, model, mode: str):
        self.m, self.mode = model, mode

    def conv(self, x, mod_or_w, bias=None, stride=1):
        if isinstance(mod_or_w, torch.nn.Conv2d):
            return A.conv2d(x, mod_or_w.weight, mod_or_w.bias, mod_or_w.stride[0], self.mode)
        return A.conv2d(x, mod_or_w, bias, stride, self.mode)

    def conv_levels(self, xs, mod):
        """One convolution module applied to every pyramid level (BlockRCB / SCGroup): a single launch per pass in tf32 mode."""
        return A.conv2d_levels(xs, mod.weight, mod.bias, self.mode)


# ---------------------------------------------------------------------------------------------------------------------
def _mgaa(cx: _Ctx, x: torch.Tensor) -> torch.Tensor:
    """MGAAbk.forward :1442-1547 on x [B,192,H,W] (x1 | x2 | x3)."""
    mg, n, acn = cx.m.MGAA, cx.m.n_feats, cx.m.ACNum
    B, _, H, W = x.shape
    inv, rows = _index_tensors(x.device, n, acn)
    x1, x2, x3 = x[:, :n], x[:, n:2 * n], x[:, 2 * n:]
    spec = A.rfft2(x)                                              # [B,384,H,Wf]: group g at channels [128g, 128g+128)
    s1, s2, s3 = spec[:, :2 * n], spec[:, 2 * n:4 * n], spec[:, 4 * n:]
    # per-bin MLPs on the interleaved layout: the reference's cat([imag, real]) packing (:1456-1465) is a permutation of
    # the 1x1 weights' input / output channels
    w0 = mg.convfuse[0].weight
    w0 = torch.cat([w0[:, :2 * n][:, inv], w0[:, 2 * n:][:, inv]], 1)
    w4 = mg.convfuse[4].weight[inv]

    # the forward (x1) and backward (x3) branches share every weight: both run as ONE batch of 2B (at the training crop a
    # convolution launch is latency, not work)
    s22 = torch.cat([s2, s2], 0)
    sa = torch.cat([s1, s3], 0)
    h = F.relu(cx.conv(_cl(torch.cat([sa, s22], 1)), w0))
    h = F.relu(cx.conv(h, mg.convfuse[2].weight))
    ofb = cx.conv(h, w4) + (sa - s22)                              # :1472-1473, [of; ob]
    sim = cx.conv(F.relu(cx.conv(_cl(s2), mg.convcrt[0].weight[:, inv])), mg.convcrt[2].weight)   # :1474
    corr = A.corr_lookup(spec, 0, 2 * n)                           # corr_f feeds both branches (:1488)
    wc = mg.convcorr[0].weight[:, :2 * n + 81]                     # the two flow channels are zeros (:1484-1485)
    wc = torch.cat([wc[:, :2 * n][:, inv], wc[:, 2 * n:], wc.new_zeros(wc.shape[0], 15, 1, 1)], 1)   # K padded to 224
    zpad = spec.new_zeros(B, 15, H, spec.shape[3])

    h = F.relu(cx.conv(_cl(torch.cat([ofb, torch.cat([corr, corr], 0), torch.cat([zpad, zpad], 0)], 1)), wc))
    h = F.relu(cx.conv(h, mg.convcorr[2].weight))
    off_fb = cx.conv(h, mg.convcorr[4].weight)                     # [2B,4,H,Wf]: [off_f; off_b]
    sim2 = torch.cat([sim, sim], 0)
    zs = []
    for i in range(acn):                                           # ConvBlk_i * x2_f_sim (:1494-1498), both directions per call
        blk = mg.MConvB[i]
        t = cx.conv(off_fb, blk.conv1.weight)
        t = F.prelu(t, blk.relu.weight)
        t = cx.conv(t, blk.conv2.weight)
        v = (_ca(blk.CA, t) + t) * sim2
        z = torch.stack([v[:, 0], v[:, 2], v[:, 1], v[:, 3]], 1)   # complex(v[0:2], v[2:4]) interleaved
        zs.append(z[:B])
        zs.append(z[B:])
    offs = A.irfft2(_cl(torch.cat(zs, 1)), W)                      # [B, 4*ACNum, H, W]: channel (i*2+dir)*2 + (dx, dy)
    # kernel predictor (:1522-1523), live rows of F.1 only, re-ordered to [iteration][tap][channel]
    kp = cx.conv(cx.conv(_cl(x2), mg.conv_KP), mg.F[0])
    taps = cx.conv(kp, mg.F[1].weight[rows], mg.F[1].bias[rows])   # [B, ACNum*192, H, W]
    aligned = []
    for d, xin in enumerate((x1, x3)):                             # IAC (:1230-1250)
        feat = xin
        for i in range(acn):
            ch = (i * 2 + d) * 2
            wp = A.flow_warp(feat, offs[:, ch:ch + 2])
            feat = F.leaky_relu(A.sac(wp, taps[:, i * 3 * n:(i + 1) * 3 * n]) + xin, 0.1)
        aligned.append(feat)
    return cx.conv(_cl(torch.cat(aligned, 1)), mg.conv3) + x2       # :1529


def _mffr(cx: _Ctx, x: torch.Tensor) -> torch.Tensor:
    """MultiFreq_Refinment.forward :2201-2254 (Split_freq :2075-2101, DivEnh :2104-2133)."""
    mf, q = cx.m.MFFRblock, cx.m.Freq_Inv
    B, c, H, W = x.shape
    masks = bands.symmetric_half_masks(q, H, W, x.device)          # [Q,H,Wf], constants
    spec = A.rfft2(x)
    stacked = torch.cat([spec * masks[j][None, None] for j in range(q)], 0)
    bnds = list(A.irfft2(_cl(stacked), W).split(B, 0))[::-1]       # freq[::-1] (:2204-2205)
    outs: List[torch.Tensor] = []
    sb = se = None
    for i in range(q):
        de = mf.DivEnh_block[i]
        xb = bnds[i]
        a, b = de.a[None], de.b[None]
        if i == 0:
            o = xb - xb.mean(dim=(2, 3), keepdim=True)
            out = _ca(de.ca, 0.2 * a * o * xb + b * xb)
            sb, se = xb, out
        else:
            o = xb - sb + 0.2 * se
            out = _ca(de.ca, 0.2 * a * o * xb + b * xb) + _ca(de.ca, 0.2 * a * se * xb + b * xb)
            sb, se = sb + xb, se + out
        outs.append(out)
    return _ca(mf.ca, se) + x                                      # :2229-2230


def _context(gc, x: torch.Tensor) -> torch.Tensor:
    """ContextBlock :657-701."""
    B, c, H, W = x.shape
    logits = (x * gc.conv_mask.weight.view(1, c, 1, 1)).sum(1).view(B, H * W)
    prob = torch.softmax(logits, dim=1).view(B, 1, H, W)
    ctxv = (x * prob).sum(dim=(2, 3))                               # [B,C]
    t = _mlp_vec(ctxv, gc.channel_add_conv[0].weight, gc.channel_add_conv[2].weight, lambda v: F.leaky_relu(v, 0.2))
    return x + t[:, :, None, None]


def _block_rcb(cx: _Ctx, blk, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """BlockRCB.forward :766-777 with RCB :705-725."""
    a = [F.leaky_relu(t, 0.1) for t in cx.conv_levels(xs, blk.body[0])]
    r0s = cx.conv_levels(a, blk.body[2])
    c1 = [F.leaky_relu(t, 0.2) for t in cx.conv_levels(r0s, blk.RCB.body[0])]
    rs = cx.conv_levels(c1, blk.RCB.body[2])
    # ContextBlock (:657-701): soft-max pooling of the three levels in one launch, the 64 -> 64 -> 64 MLP on the 3B pooled vectors
    gc = blk.RCB.gcnet
    pooled = A.context_pool(rs, gc.conv_mask.weight)                               # [levels, B, 64]
    tadd = _mlp_vec(pooled.reshape(-1, pooled.shape[-1]), gc.channel_add_conv[0].weight, gc.channel_add_conv[2].weight,
                    lambda v: F.leaky_relu(v, 0.2)).view_as(pooled)
    res = A.rcb_tail(rs, tadd, r0s)                                                # lrelu_0.2(r + t) + r0 (:720-724)
    # Interpolate(0.5) of an even-sized map is the 2x2 mean, which commutes with the 1x1 `down` convolution (:753-757)
    tds = cx.conv_levels([_cl(F.avg_pool2d(r, 2)) for r in res[:-1]], blk.down[0])
    tus = cx.conv_levels(res[1:], blk.up[0])
    if len(xs) == 3 and xs[0].shape[1] == 64 and all(x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 for x in xs[:2]):
        return A.level_mix(xs, res, tds, tus)                                       # x + r + d + u (:771-776), one launch
    down = [res[0]] + tds
    up = [F.interpolate(u, scale_factor=2.0, mode="bilinear", align_corners=False) for u in tus] + [res[-1]]
    return [x + r + d + u for x, r, d, u in zip(xs, res, down, up)]


def _scnet(cx: _Ctx, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """SCNetbk :807-822 / SCGroupbk :781-803."""
    cur = xs
    for grp in cx.m.recorb1.body:
        t = cur
        for blk in grp.body:
            t = _block_rcb(cx, blk, t)
        cur = [x + c for x, c in zip(cur, cx.conv_levels(t, grp.conv))]
    return [x + r for x, r in zip(xs, cur)]


def forward_train(model, x: torch.Tensor, mode: str = "tf32") -> torch.Tensor:
    """x [B,7,1,H,W] -> [B,1,4H,4W]; differentiable with respect to every live parameter (and to x if it requires grad)."""
    if mode not in ("fp32", "tf32"):
        raise ValueError("the training forward runs in 'fp32' (CUDA-core convolutions) or 'tf32' (tcgen05) mode")
    cx = _Ctx(model, mode)
    m, n = model, model.n_feats
    b, t, c, h, w = x.shape
    feats = cx.conv(_cl(x.reshape(b, t * c, h, w)), m.feat_extract[0])
    f1, f2, f3 = feats[:, :3 * n], feats[:, 3 * n:4 * n], feats[:, 4 * n:]
    o1 = _mgaa(cx, _cl(f1))
    o3 = _mgaa(cx, _cl(f3))
    o2 = _mgaa(cx, _cl(torch.cat([o1, f2, o3], 1)))
    l1 = _mffr(cx, o2)
    l2 = cx.conv(l1, m.rconcat1)
    l3 = cx.conv(l2, m.rconcat2)
    o_l1, o_l2, o_l3 = _scnet(cx, [l1, l2, l3])
    pw = m.lrelu.weight
    u3_1 = F.pixel_shuffle(F.prelu(cx.conv(o_l3, m.upconv1_L3), pw), 2)
    u3_2 = F.pixel_shuffle(u3_1, 2)
    u2 = F.prelu(cx.conv(o_l2, m.upconv1_L2), pw)
    u2 = F.pixel_shuffle(u2 + cx.conv(_cl(torch.cat([u2, u3_1], 1)), m.upconv1_L2_2), 2)
    fuse = cx.conv(cx.conv(_cl(torch.cat([o_l1, u2, u3_2], 1)), m.upconv_fuse), m.recorb0)
    y = F.prelu(F.pixel_shuffle(cx.conv(fuse, m.upconv1), 2), pw)
    y = F.prelu(F.pixel_shuffle(cx.conv(_cl(y), m.upconv2), 2), pw)
    y = cx.conv(_cl(y), m.conv_last0)
    return y + F.interpolate(x[:, t // 2], scale_factor=4, mode="bilinear", align_corners=False)
