"""CPU oracle for the FCVSR / FCVSR-S per-clip x4 super-resolution forward.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch (CPU, fp32) restatement of the
reference algorithm in ``CVSR_train/arch/CVSR_freq.py``; it exists so the CUDA path can be
checked on a box where ``/root/reference`` is absent.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it, and only as the checker / the timed CPU arm.  The product package
(``fcvsr_b200``) never imports it.

Parity pin: ``tests/golden/*.pt`` were produced by the *unmodified* reference
(``oracle/make_golden.py`` imports ``/root/reference/CVSR_train/arch/CVSR_freq.py`` in the
build container) and ``tests/test_oracle.py`` checks this restatement against them
(max-abs <= 2e-5), so the oracle is pinned to the reference's own outputs.

Every function cites the reference lines it restates (file = CVSR_train/arch/CVSR_freq.py
unless another file is named).  The host-side visualisation side effects of the reference
(numpy colour wheels, featuremap_visual, out.cpu()) do not influence the returned
tensor and are not restated.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """nn.Conv2d with 'same' padding k//2 (all convs on the path use it)."""
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    return F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2)


def _ca(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """CALayer :1812-1828 -- x * sigmoid(W2 relu(W1 mean_hw(x))), no bias."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = F.relu(F.conv2d(y, sd[name + ".conv_du.0.weight"]))
    y = torch.sigmoid(F.conv2d(y, sd[name + ".conv_du.2.weight"]))
    return x * y


def _pack_spec(z: torch.Tensor) -> torch.Tensor:
    """:1456-1465 -- cat([imag, real]) of an rfft2 spectrum along channels."""
    return torch.cat([z.imag, z.real], dim=1)


# ----------------------------------------------------------------------------------------------
# MGAAbk pieces
# ----------------------------------------------------------------------------------------------
def warp_bilinear(x: torch.Tensor, flow_xy: torch.Tensor) -> torch.Tensor:
    """flow_warp :1188-1227.  flow_xy [B,2,H,W], channel 0 = dx, 1 = dy (pixels).
    out[c,y,x] = bilinear(x[c], y+dy, x+dx), zeros outside, align_corners=True."""
    b, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    px = gx.to(x) + flow_xy[:, 0]
    py = gy.to(x) + flow_xy[:, 1]
    nx = 2.0 * px / max(w - 1, 1) - 1.0
    ny = 2.0 * py / max(h - 1, 1) - 1.0
    grid = torch.stack((nx, ny), dim=3)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def sac(feat: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """SAC :1253-1276 (ksize 3).  taps [B, C*3, H, W] with channel c*3+t.  The reference applies
    kernel1 in BOTH passes (:1265 and :1273); kernel2 is dead.  Vertical pass then horizontal
    pass, replicate padding, the horizontal pass re-using the taps of the output pixel."""
    b, c, h, w = feat.shape
    k = taps.view(b, c, 3, h, w)
    fp = F.pad(feat, (0, 0, 1, 1), mode="replicate")
    v = sum(fp[:, :, t:t + h, :] * k[:, :, t] for t in range(3))
    vp = F.pad(v, (1, 1, 0, 0), mode="replicate")
    return sum(vp[:, :, :, t:t + w] * k[:, :, t] for t in range(3))


def iac(feat_in: torch.Tensor, pred_k: torch.Tensor, offsets: List[torch.Tensor], n_iter: int) -> torch.Tensor:
    """IAC :1230-1250.  Per iteration i only channels [i*384, i*384+192) of Pred_K are live."""
    c = feat_in.shape[1]
    feat = feat_in
    for i in range(n_iter):
        taps = pred_k[:, i * 6 * c: i * 6 * c + 3 * c]
        feat = sac(warp_bilinear(feat, offsets[i]), taps) + feat_in
        feat = F.leaky_relu(feat, 0.1)
    return feat


def corr_lookup(a_f: torch.Tensor, b_f: torch.Tensor) -> torch.Tensor:
    """CorrBlock :1279-1337 + bilinear_sampler :1340-1354 at integer coordinates.

    prod = a_f*b_f/sqrt(C) ([B,C=128,H,Wf]) is *memory-reinterpreted* as, per position
    p = y0*Wf+x0, a 64x2 image made of the 128 consecutive floats prod.flat[p*128:(p+1)*128]
    (:1334).  Output channel i*9+j samples that image at (col = x0+i-4, row = y0+j-4)
    (delta puts dy on the x coordinate, :1303-1309), zero outside."""
    bsz, c, h, wf = a_f.shape
    prod = (a_f * b_f / math.sqrt(float(c))).reshape(bsz, -1)
    npos = h * wf
    img = prod.view(bsz, npos, c // 2, 2)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(wf), indexing="ij")
    ys = ys.reshape(-1)
    xs = xs.reshape(-1)
    out = torch.zeros(bsz, 81, npos, dtype=a_f.dtype)
    pos = torch.arange(npos)
    for i in range(9):
        col = xs + i - 4
        for j in range(9):
            row = ys + j - 4
            ok = (col >= 0) & (col < 2) & (row >= 0) & (row < c // 2)
            if not bool(ok.any()):
                continue
            vals = img[:, pos[ok], row[ok], col[ok]]
            out[:, i * 9 + j, pos[ok]] = vals
    return out.view(bsz, 81, h, wf)


def conv_blk(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ConvBlk :344-357 -- conv k -> PReLU -> conv k -> CALayer(4, r=1)*1 + out."""
    t = F.conv2d(x, sd[name + ".conv1.weight"], padding=sd[name + ".conv1.weight"].shape[-1] // 2)
    t = F.prelu(t, sd[name + ".relu.weight"])
    t = F.conv2d(t, sd[name + ".conv2.weight"], padding=sd[name + ".conv2.weight"].shape[-1] // 2)
    return _ca(sd, name + ".CA", t) + t


def _mlp1x1(sd: SD, name: str, x: torch.Tensor, idx: Tuple[int, ...]) -> torch.Tensor:
    for n, i in enumerate(idx):
        x = F.conv2d(x, sd[f"{name}.{i}.weight"])
        if n + 1 < len(idx):
            x = F.relu(x)
    return x


def mgaa_offsets(sd: SD, x: torch.Tensor, n_iter: int, p: str = "MGAA") -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Frequency-domain offset estimation, MGAAbk.forward :1452-1505."""
    d = x.shape[1] // 3
    h, w = x.shape[-2:]
    x1, x2, x3 = x[:, :d], x[:, d:2 * d], x[:, 2 * d:]
    f1 = _pack_spec(torch.fft.rfft2(x1, norm="backward"))
    f2 = _pack_spec(torch.fft.rfft2(x2, norm="backward"))
    f3 = _pack_spec(torch.fft.rfft2(x3, norm="backward"))
    off_f = (f1 - f2) + _mlp1x1(sd, p + ".convfuse", torch.cat([f1, f2], 1), (0, 2, 4))
    off_b = (f3 - f2) + _mlp1x1(sd, p + ".convfuse", torch.cat([f3, f2], 1), (0, 2, 4))
    sim = _mlp1x1(sd, p + ".convcrt", f2, (0, 2))
    corr_f = corr_lookup(f1, f2)          # the backward branch re-uses corr_f (:1488)
    zero_flow = torch.zeros_like(f1[:, :2])   # coords1 - coords0 == 0 (:1484-1485)
    off_f = _mlp1x1(sd, p + ".convcorr", torch.cat([off_f, corr_f, zero_flow], 1), (0, 2, 4))
    off_b = _mlp1x1(sd, p + ".convcorr", torch.cat([off_b, corr_f, zero_flow], 1), (0, 2, 4))
    outs_f, outs_b = [], []
    for i in range(n_iter):
        for src, dst in ((off_f, outs_f), (off_b, outs_b)):
            o = conv_blk(sd, f"{p}.MConvB.{i}", src) * sim
            z = torch.complex(o[:, 0:2], o[:, 2:4])
            dst.append(torch.fft.irfft2(z, s=(h, w), norm="backward"))
    return outs_f, outs_b


def mgaa(sd: SD, x: torch.Tensor, n_iter: int, p: str = "MGAA") -> torch.Tensor:
    """MGAAbk.forward :1442-1547 (returns only `out`)."""
    d = x.shape[1] // 3
    x1, x2, x3 = x[:, :d], x[:, d:2 * d], x[:, 2 * d:]
    offs_f, offs_b = mgaa_offsets(sd, x, n_iter, p)
    pred_k = _conv(sd, p + ".F.1", _conv(sd, p + ".F.0", _conv(sd, p + ".conv_KP", x2)))
    al_f = iac(x1, pred_k, offs_f, n_iter)
    al_b = iac(x3, pred_k, offs_b, n_iter)
    return _conv(sd, p + ".conv3", torch.cat([al_f, al_b], 1)) + x2


# ----------------------------------------------------------------------------------------------
# MultiFreq_Refinment pieces
# ----------------------------------------------------------------------------------------------
_MASK_CACHE: Dict[Tuple[int, str], torch.Tensor] = {}


def band_masks_1024(q: int) -> torch.Tensor:
    """Split_freq.generate_freq_mask(1024, 1024), gaussian mode :2016-2051.
    M_n = G_n - sum(previous M) with G_n = exp(-r^2 / (2 (n+1)^2 l^2)), l = sqrt(512^2+512^2)/q."""
    key = (q, "gaussian")
    if key not in _MASK_CACHE:
        n = 1024
        length = math.sqrt((n / 2) ** 2 + (n / 2) ** 2) / q
        ax = np.arange(-(n // 2), n - n // 2, 1) ** 2
        r2 = ax[:, None] + ax[None, :]
        r = np.sqrt(r2.astype(np.float64))
        prev: List[torch.Tensor] = []
        for i in range(q):
            g = torch.from_numpy(np.exp(-np.power(r, 2) / (2 * (length * (i + 1)) ** 2))).float()
            for pm in prev:
                g = g - pm
            prev.append(g)
        _MASK_CACHE[key] = torch.stack(prev, 0)
    return _MASK_CACHE[key]


def band_masks(q: int, h: int, w: int) -> torch.Tensor:
    """:2078 -- torchvision bicubic Resize of the 1024^2 masks (torchvision default antialias)."""
    from torchvision.transforms import Resize, functional as TF
    return Resize([h, w], interpolation=TF.InterpolationMode.BICUBIC)(band_masks_1024(q))


def split_freq(x: torch.Tensor, q: int) -> List[torch.Tensor]:
    """Split_freq.forward :2075-2101: band_j = Re ifft2(ifftshift(fftshift(fft2 x) * M_j))."""
    h, w = x.shape[-2:]
    m = band_masks(q, h, w).to(x.dtype)
    f = torch.fft.fftshift(torch.fft.fftn(x, dim=(2, 3)), dim=(2, 3))
    bands = []
    for j in range(q):
        bands.append(torch.fft.ifftn(torch.fft.ifftshift(f * m[j], dim=(2, 3)), dim=(2, 3)).real)
    return bands


def div_enh(sd: SD, name: str, x: torch.Tensor, before: List[torch.Tensor], enh_before: List[torch.Tensor]) -> torch.Tensor:
    """DivEnh.forward :2114-2133."""
    a, b = sd[name + ".a"], sd[name + ".b"]
    if not before:
        o = x - x.mean(dim=(2, 3), keepdim=True)
        return _ca(sd, name + ".ca", 0.2 * a * o * x + b * x)
    sb = torch.stack(before, 0).sum(0)
    se = torch.stack(enh_before, 0).sum(0)
    o = x - sb + 0.2 * se
    t1 = _ca(sd, name + ".ca", 0.2 * a * o * x + b * x)
    t2 = _ca(sd, name + ".ca", 0.2 * a * se * x + b * x)
    return t1 + t2


def mffr(sd: SD, x: torch.Tensor, q: int, p: str = "MFFRblock") -> torch.Tensor:
    """MultiFreq_Refinment.forward :2201-2254."""
    bands = split_freq(x, q)[::-1]
    outs: List[torch.Tensor] = []
    for i in range(q):
        outs.append(div_enh(sd, f"{p}.DivEnh_block.{i}", bands[i], bands[:i], outs[:i]))
    return _ca(sd, p + ".ca", torch.stack(outs, 0).sum(0)) + x


# ----------------------------------------------------------------------------------------------
# SCNetbk pieces
# ----------------------------------------------------------------------------------------------
def context_block(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ContextBlock :657-701."""
    b, c, h, w = x.shape
    logits = F.conv2d(x, sd[name + ".conv_mask.weight"]).view(b, 1, h * w)
    prob = torch.softmax(logits, dim=2)
    ctx = torch.matmul(x.view(b, c, h * w), prob.transpose(1, 2)).view(b, c, 1, 1)
    t = F.conv2d(ctx, sd[name + ".channel_add_conv.0.weight"])
    t = F.conv2d(F.leaky_relu(t, 0.2), sd[name + ".channel_add_conv.2.weight"])
    return x + t


def rcb(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """RCB :705-725."""
    r = _conv(sd, name + ".body.2", F.leaky_relu(_conv(sd, name + ".body.0", x), 0.2))
    return F.leaky_relu(context_block(sd, name + ".gcnet", r), 0.2) + x


def block_rcb(sd: SD, name: str, xs: List[torch.Tensor]) -> List[torch.Tensor]:
    """BlockRCB.forward :766-777."""
    res = []
    for x in xs:
        r = _conv(sd, name + ".body.2", F.leaky_relu(_conv(sd, name + ".body.0", x), 0.1))
        res.append(rcb(sd, name + ".body.3", r))
    down = [res[0]] + [F.interpolate(_conv(sd, name + ".down.0", r), scale_factor=0.5, mode="bilinear",
                                     align_corners=False) for r in res[:-1]]
    up = [F.interpolate(_conv(sd, name + ".up.0", r), scale_factor=2.0, mode="bilinear",
                        align_corners=False) for r in res[1:]] + [res[-1]]
    return [x + r + d + u for x, r, d, u in zip(xs, res, down, up)]


def scnet(sd: SD, xs: List[torch.Tensor], n_groups: int, p: str = "recorb1") -> List[torch.Tensor]:
    """SCNetbk :807-822 / SCGroupbk :781-803."""
    cur = xs
    for g in range(n_groups):
        t = cur
        for k in range(3):
            t = block_rcb(sd, f"{p}.body.{g}.body.{k}", t)
        cur = [x + _conv(sd, f"{p}.body.{g}.conv", r) for x, r in zip(cur, t)]
    return [x + r for x, r in zip(xs, cur)]


# ----------------------------------------------------------------------------------------------
# whole forward
# ----------------------------------------------------------------------------------------------
def infer_config(sd: SD) -> Tuple[int, int, int]:
    """(ACNum, Freq_Inv, SCGroupN) from the state-dict keys."""
    a = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("MGAA.MConvB."))
    q = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("MFFRblock.DivEnh_block."))
    g = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("recorb1.body."))
    return a, q, g


def forward(sd: SD, x: torch.Tensor, return_taps: bool = False):
    """GShiftNet.forward :2688-2756 / GShiftNet_S.forward :2611-2646 (they differ only in
    hyper-parameters and 1x1 vs 3x3 up-convs, both carried by the state dict)."""
    n_iter, q, n_groups = infer_config(sd)
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
    u3 = F.prelu(_conv(sd, "upconv1_L3", o_l3), pw)
    u3_1 = F.pixel_shuffle(u3, 2)
    u3_2 = F.pixel_shuffle(u3_1, 2)
    u2 = F.prelu(_conv(sd, "upconv1_L2", o_l2), pw)
    u2 = F.pixel_shuffle(u2 + _conv(sd, "upconv1_L2_2", torch.cat([u2, u3_1], 1)), 2)
    fuse = _conv(sd, "recorb0", _conv(sd, "upconv_fuse", torch.cat([o_l1, u2, u3_2], 1)))
    taps["fuse"] = fuse
    y = F.prelu(F.pixel_shuffle(_conv(sd, "upconv1", fuse), 2), pw)
    y = F.prelu(F.pixel_shuffle(_conv(sd, "upconv2", y), 2), pw)
    y = _conv(sd, "conv_last0", y)
    y = y + F.interpolate(x[:, t // 2], scale_factor=4, mode="bilinear")
    return (y, taps) if return_taps else y


# ----------------------------------------------------------------------------------------------
# DCN operator oracle (ops/dcn/deform_conv.py:114-187; kernels ops/dcn/src/deform_conv_cuda_kernel.cu)
# ----------------------------------------------------------------------------------------------
def modulated_deform_conv(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                          groups=1, deformable_groups=1):
    """Plain-loop restatement of modulated_deformable_im2col (.cu:570-632) + the per-group GEMM
    (deform_conv_cuda.cpp:545-563).  mask=None gives DCNv1 (.cu:190-242).  Offset channel
    dg*2*kk + 2*(i*kW+j) is dh, +1 is dw; samples outside (-1,H)x(-1,W) contribute 0 and taps on
    the border ring use zero for out-of-image corners (dmcn_im2col_bilinear .cu:84-114)."""
    b, cin, h, w = x.shape
    cout, cin_g, kh, kw = weight.shape
    ho = (h + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    wo = (w + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    cpg = cin // deformable_groups
    ys, xs = torch.meshgrid(torch.arange(ho), torch.arange(wo), indexing="ij")
    cols = torch.zeros(b, cin, kh * kw, ho, wo, dtype=x.dtype)
    for g in range(deformable_groups):
        xg = x[:, g * cpg:(g + 1) * cpg]
        for i in range(kh):
            for j in range(kw):
                kidx = i * kw + j
                dh = offset[:, g * 2 * kh * kw + 2 * kidx]
                dw = offset[:, g * 2 * kh * kw + 2 * kidx + 1]
                py = (ys * stride - padding + i * dilation).to(x) + dh
                px = (xs * stride - padding + j * dilation).to(x) + dw
                valid = (py > -1) & (px > -1) & (py < h) & (px < w)
                y0 = torch.floor(py)
                x0 = torch.floor(px)
                ly, lx = py - y0, px - x0
                val = torch.zeros(b, cpg, ho, wo, dtype=x.dtype)
                for (yy, xx, wt) in ((y0, x0, (1 - ly) * (1 - lx)), (y0, x0 + 1, (1 - ly) * lx),
                                     (y0 + 1, x0, ly * (1 - lx)), (y0 + 1, x0 + 1, ly * lx)):
                    inb = (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1) & valid
                    yi = yy.clamp(0, h - 1).long()
                    xi = xx.clamp(0, w - 1).long()
                    idx = (yi * w + xi).view(b, 1, -1).expand(b, cpg, -1)
                    got = xg.reshape(b, cpg, -1).gather(2, idx).view(b, cpg, ho, wo)
                    val = val + got * (wt * inb.to(x)).unsqueeze(1)
                if mask is not None:
                    val = val * mask[:, g * kh * kw + kidx].unsqueeze(1)
                cols[:, g * cpg:(g + 1) * cpg, kidx] = val
    cols = cols.view(b, groups, cin_g * kh * kw, ho * wo)
    wg = weight.view(groups, cout // groups, cin_g * kh * kw)
    out = torch.einsum("gok,bgkp->bgop", wg, cols).reshape(b, cout, ho, wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def charbonnier_sum(sr: torch.Tensor, hr: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """CharbonnierLoss, CVSR_train/opt/loss.py:20-31 (sum reduction)."""
    d = sr - hr
    return torch.sqrt(d * d + eps).sum()
