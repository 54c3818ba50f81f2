"""Autograd-capable operators of the FCVSR training step, bound to the sm_100a kernel library.

The reference trains with `sr = model(frames); loss.backward()` (CVSR_train/train_LD_freqCVSR_22.py:243-251): every gradient
comes from ATen's autograd kernels (cuDNN dgrad / wgrad, cuFFT, grid_sampler_2d_backward).  Here the arithmetic-heavy
operators of the forward -- convolutions, the 2-D real FFTs, the CorrBlock lookup, flow_warp and SAC -- are
`torch.autograd.Function`s whose forward AND backward are calls into libfcvsr_b200.so (include/fcvsr_b200.h, "adjoints"
section); `fcvsr_b200.train_forward` strings them together with PyTorch's elementwise glue, so that
`fcvsr_b200.arch.GShiftNet(x)` is differentiable like the reference module.

Tensor convention: every activation is an NHWC-contiguous buffer exposed to PyTorch as the NCHW-logical tensor
`buf.permute(0, 3, 1, 2)` (= torch.channels_last), which is what the kernels read and what ATen's elementwise / interpolate
kernels keep.  Spectra are complex-interleaved along the channel axis: logical channel 2c = Re, 2c + 1 = Im of complex
channel c (the layout of fcvsr_fft_*).

Compute modes (`mode` arguments): "fp32" runs every convolution on the CUDA-core kernel (exact fp32; used by the golden-
gradient test), "tf32" runs forward and data-gradient convolutions whose shape fits on the tcgen05 kernel with operands
rounded to nearest TF32 (the contract's fp32 mode) and their weight gradients on the tcgen05 wgrad kernel (bf16 copies of
the activations / upstream gradients, fp32 accumulation; csrc/wgrad_tc.cu); every other weight gradient is fp32 FFMA.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _capi as C
from . import bands

F32 = torch.float32
WGRAD_TC = True        # "tf32" mode: weight gradients of 64-multiple-channel stride-1 convolutions on tcgen05 (bf16 operands)


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """NCHW-logical tensor -> contiguous [B,H,W,C] buffer (no copy when it already is channels_last)."""
    if t.dtype != F32:
        raise TypeError("fcvsr_b200 training operators are fp32")
    v = t.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _logical(buf: torch.Tensor) -> torch.Tensor:
    return buf.permute(0, 3, 1, 2)


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32)


def _tc_ok(cin: int, cout: int, k: int, stride: int) -> bool:
    """Shape envelope of fcvsr_conv2d_tc with TF32 operands (conv_tc.cu)."""
    return stride == 1 and k in (1, 3) and cin % 32 == 0 and (cout % 16 == 0 or cout < 16) and cout <= 2048


def _is_conv4(ci: int, co: int, k: int, stride: int, bias) -> bool:
    """The 4 -> 4 channel ConvBlk convolutions (CVSR_freq.py:344-357: k = 1 .. 11, no bias): dedicated kernels in both modes (exact
    fp32 FFMA) -- the generic kernels tile 64 pixels x 64 channels and spend 1/16 .. 1/256 of their work on this shape."""
    return ci == 4 and co == 4 and stride == 1 and (k & 1) and k <= 11 and bias is None


def _conv_launch(xh: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], stride: int, mode: str,
                 transposed: bool = False, x16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = conv(x, w) + bias on an NHWC buffer; w in the reference's [Cout, Cin, k, k] layout.  transposed: the data gradient --
    x is dy and the convolution runs with w's taps flipped and its channel axes swapped (stride 1 only).  x16: optional bf16
    buffer of x's shape that receives the bf16 copy of x in the same pass as the TF32-rounded one (tensor-core path only: the
    operand of the tcgen05 weight gradient)."""
    B, H, W, ci = xh.shape
    if transposed:
        ci_w, co, k, _ = w.shape
    else:
        co, ci_w, k, _ = w.shape
    assert ci_w == ci, (ci_w, ci)
    pad = k // 2
    ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(B, ho, wo, co, device=xh.device, dtype=F32)
    bp = bias.contiguous().data_ptr() if bias is not None else 0
    if _is_conv4(ci, co, k, stride, bias):
        # transposed (data gradient): w'[tap][co][ci] = w[k*k - 1 - tap][ci][co]
        wd = (w.flip(2, 3).permute(2, 3, 0, 1) if transposed else w.permute(2, 3, 1, 0)).contiguous()
        C.call("fcvsr_conv4x4", xh.data_ptr(), wd.data_ptr(), y.data_ptr(), B, H, W, k, _st())
        return y
    if mode == "tf32" and _tc_ok(ci, co, k, stride):
        xr = torch.empty_like(xh)                            # tcgen05 truncates TF32 operands: round to nearest first
        C.call("fcvsr_round_copy_dual", xh.data_ptr(), xr.data_ptr(), x16.data_ptr() if x16 is not None else 0, xh.numel(), _st())
        wt = torch.empty(max(co, 16), k * k * ci, device=xh.device, dtype=F32)
        wc = w.contiguous()
        C.call("fcvsr_pack_conv_weight", wc.data_ptr(), wt.data_ptr(), wc.shape[0], wc.shape[1], k, int(transposed), 16, _st())
        C.call("fcvsr_conv2d_tc", xr.data_ptr(), ci, wt.data_ptr(), bp, 0, 0, 0, 0, y.data_ptr(), co, B, H, W, ci, co, k,
               C.ACT_NONE, 0.0, 0, 0, 0, 0, 0, 0, 0, _st())
    else:
        assert not transposed
        wd = w.permute(2, 3, 1, 0).contiguous()              # [k*k][Cin][Cout]
        C.call("fcvsr_conv2d_direct", xh.data_ptr(), ci, 0, wd.data_ptr(), bp, 0, 0, 0, 0, y.data_ptr(), co, B, H, W, ci, co, k,
               stride, C.ACT_NONE, 0.0, 0, 0, 0, 0, 0, 0, 0, _st())
    return y


class _Conv2d(torch.autograd.Function):
    """nn.Conv2d(k, stride, padding k // 2) of the reference (every convolution of CVSR_freq.py uses that padding)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, mode):
        xh = _nhwc(x)
        co, ci, k, _ = w.shape
        # tcgen05 weight gradient ahead: its bf16 copy of the activation is made in the same pass as the forward's TF32-rounded
        # copy and is what backward keeps (half the bytes of the fp32 activation, one launch less per convolution)
        wg_tc = (mode == "tf32" and WGRAD_TC and stride == 1 and k in (1, 3) and ci % 64 == 0 and co % 64 == 0
                 and ctx.needs_input_grad[1] and _tc_ok(ci, co, k, stride))
        xb = torch.empty(xh.shape, device=xh.device, dtype=torch.bfloat16) if wg_tc else None
        with torch.cuda.device(x.device):
            y = _conv_launch(xh, w.detach(), None if bias is None else bias.detach(), stride, mode, x16=xb)
        ctx.save_for_backward(xb if wg_tc else xh, w)
        ctx.stride, ctx.mode, ctx.has_bias, ctx.wg_tc = stride, mode, bias is not None, wg_tc
        return _logical(y)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        xh, w = ctx.saved_tensors
        w = w.detach()
        stride, mode = ctx.stride, ctx.mode
        g = _nhwc(gy)
        B, H, W, ci = xh.shape
        co, _, k, _ = w.shape
        gx = gw = gb = None
        wg_tc = ctx.wg_tc and ctx.needs_input_grad[1]
        gb16 = torch.empty(g.shape, device=g.device, dtype=torch.bfloat16) if wg_tc else None
        g16_done = False
        with torch.cuda.device(xh.device):
            if ctx.needs_input_grad[0]:
                if _is_conv4(ci, co, k, stride, None if not ctx.has_bias else 1) or (mode == "tf32" and _tc_ok(co, ci, k, stride)):
                    # dx = conv(dy, w') with w'[ci][co][ky][kx] = w[co][ci][k-1-ky][k-1-kx]
                    tc_dgrad = not _is_conv4(ci, co, k, stride, None if not ctx.has_bias else 1)
                    dx = _conv_launch(g, w, None, 1, mode, transposed=True, x16=gb16 if tc_dgrad else None)
                    g16_done = tc_dgrad and gb16 is not None
                else:
                    dx = torch.empty(B, H, W, ci, device=xh.device, dtype=F32)
                    wt = w.permute(2, 3, 0, 1).contiguous()                      # [k*k][Cout][Cin]
                    C.call("fcvsr_conv2d_dgrad_direct", g.data_ptr(), co, wt.data_ptr(), dx.data_ptr(), ci, B, H, W, ci, co, k,
                           stride, _st())
                gx = _logical(dx)
            if ctx.needs_input_grad[1]:
                dw = torch.zeros(k * k, ci, co, device=xh.device, dtype=F32)
                if _is_conv4(ci, co, k, stride, None if not ctx.has_bias else 1):
                    C.call("fcvsr_conv4x4_wgrad", xh.data_ptr(), g.data_ptr(), dw.data_ptr(), B, H, W, k, _st())
                elif wg_tc:
                    # tensor-core weight gradient: bf16 copies of the activation (made by the forward pass) and of the upstream
                    # gradient (made together with the data gradient's TF32 copy when there is one), fp32 accumulation
                    if not g16_done:
                        C.call("fcvsr_round_copy_dual", g.data_ptr(), 0, gb16.data_ptr(), g.numel(), _st())
                    C.call("fcvsr_conv2d_wgrad_tc", xh.data_ptr(), ci, gb16.data_ptr(), co, dw.data_ptr(), B, H, W, ci, co, k, _st())
                else:
                    C.call("fcvsr_conv2d_wgrad", xh.data_ptr(), ci, g.data_ptr(), co, dw.data_ptr(), B, H, W, ci, co, k, stride, _st())
                gw = dw.view(k, k, ci, co).permute(3, 2, 0, 1)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                npix = g.shape[0] * g.shape[1] * g.shape[2]
                scratch = torch.empty(((npix + 63) // 64) * co, device=xh.device, dtype=F32)
                gb = torch.empty(co, device=xh.device, dtype=F32)
                C.call("fcvsr_colsum", g.data_ptr(), co, co, npix, scratch.data_ptr(), gb.data_ptr(), 0, _st())
        return gx, gw, gb, None, None


def conv2d(x, w, bias=None, stride: int = 1, mode: str = "tf32"):
    return _Conv2d.apply(x, w, bias, stride, mode)


def _conv_multi_launch(xs, wt: torch.Tensor, bias: Optional[torch.Tensor], co: int, k: int):
    """fcvsr_conv2d_tc_multi on NHWC fp32 buffers of different spatial size (same batch, same channels): TF32 operands, fp32 out."""
    import ctypes
    n = len(xs)
    B, ci = xs[0].shape[0], xs[0].shape[3]
    ys = [torch.empty(x.shape[0], x.shape[1], x.shape[2], co, device=x.device, dtype=F32) for x in xs]
    vp = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
    ia = lambda v: (ctypes.c_int * n)(*v)  # noqa: E731
    C.call("fcvsr_conv2d_tc_multi", n, vp(xs), ci, wt.data_ptr(), bias.contiguous().data_ptr() if bias is not None else 0, None, 0,
           vp(ys), co, ia([x.shape[1] for x in xs]), ia([x.shape[2] for x in xs]), B, ci, co, k, C.ACT_NONE, 0.0, 0, None, 0, 0, 0, _st())
    return ys


def _round_dual_multi(xs, y32, y16):
    """TF32-rounded and / or bf16 copies of up to three dense fp32 tensors in one launch."""
    import ctypes
    n = len(xs)
    vp = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]) if ts is not None else None  # noqa: E731
    C.call("fcvsr_round_copy_dual_multi", n, vp(xs), vp(y32), vp(y16), (ctypes.c_longlong * n)(*[t.numel() for t in xs]), _st())


class _Conv2dLevels(torch.autograd.Function):
    """The same stride-1 convolution (one weight, one bias) on several tensors of different spatial size -- the pyramid levels
    of a BlockRCB / SCGroup convolution (CVSR_freq.py:766-770, :797-803) -- in ONE tcgen05 launch for the forward and one for
    the data gradients (fcvsr_conv2d_tc_multi); the weight gradients of the levels accumulate into one buffer.  At the training
    crop (64 x 64 and below) a convolution launch is latency, not work, so three levels per launch is three times fewer of them."""

    @staticmethod
    def forward(ctx, w, bias, *xs):
        co, ci, k, _ = w.shape
        xh = [_nhwc(x) for x in xs]
        wg_tc = WGRAD_TC and ci % 64 == 0 and co % 64 == 0 and ctx.needs_input_grad[0]
        with torch.cuda.device(w.device):
            wt = torch.empty(max(co, 16), k * k * ci, device=w.device, dtype=F32)
            wc = w.detach().contiguous()
            C.call("fcvsr_pack_conv_weight", wc.data_ptr(), wt.data_ptr(), co, ci, k, 0, 16, _st())
            xr = [torch.empty_like(x) for x in xh]
            xb = [torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if wg_tc else None for x in xh]
            _round_dual_multi(xh, xr, xb if wg_tc else None)
            ys = _conv_multi_launch(xr, wt, None if bias is None else bias.detach(), co, k)
        ctx.save_for_backward(w, *(xb if wg_tc else xh))
        ctx.wg_tc, ctx.has_bias, ctx.n = wg_tc, bias is not None, len(xs)
        return tuple(_logical(y) for y in ys)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gys):
        w, *xsaved = ctx.saved_tensors
        w = w.detach()
        co, ci, k, _ = w.shape
        g = [_nhwc(t) for t in gys]
        need_x = any(ctx.needs_input_grad[2:])
        gw = gb = None
        gxs = [None] * ctx.n
        with torch.cuda.device(w.device):
            want16 = ctx.wg_tc and ctx.needs_input_grad[0]
            gr = [torch.empty_like(t) if need_x else None for t in g]
            g16 = [torch.empty(t.shape, device=t.device, dtype=torch.bfloat16) if want16 else None for t in g]
            if need_x or want16:
                _round_dual_multi(g, gr if need_x else None, g16 if want16 else None)
            if need_x:
                wt = torch.empty(max(ci, 16), k * k * co, device=w.device, dtype=F32)
                wc = w.contiguous()
                C.call("fcvsr_pack_conv_weight", wc.data_ptr(), wt.data_ptr(), co, ci, k, 1, 16, _st())
                dxs = _conv_multi_launch(gr, wt, None, ci, k)
                gxs = [_logical(d) for d in dxs]
            if ctx.needs_input_grad[0]:
                dw = torch.zeros(k * k, ci, co, device=w.device, dtype=F32)
                if ctx.wg_tc:            # the levels' weight gradients add up: one launch walks the tiles of all of them
                    import ctypes
                    n = len(xsaved)
                    C.call("fcvsr_conv2d_wgrad_tc_multi", n, (ctypes.c_void_p * n)(*[t.data_ptr() for t in xsaved]), ci,
                           (ctypes.c_void_p * n)(*[t.data_ptr() for t in g16]), co, dw.data_ptr(), xsaved[0].shape[0],
                           (ctypes.c_int * n)(*[t.shape[1] for t in xsaved]), (ctypes.c_int * n)(*[t.shape[2] for t in xsaved]), ci, co, k, _st())
                else:
                    for xs_l, g_l in zip(xsaved, g):
                        B, H, W, _ = xs_l.shape
                        C.call("fcvsr_conv2d_wgrad", xs_l.data_ptr(), ci, g_l.data_ptr(), co, dw.data_ptr(), B, H, W, ci, co, k, 1, _st())
                gw = dw.view(k, k, ci, co).permute(3, 2, 0, 1)
            if ctx.has_bias and ctx.needs_input_grad[1]:
                gb = torch.zeros(co, device=w.device, dtype=F32)
                for g_l in g:
                    npix = g_l.shape[0] * g_l.shape[1] * g_l.shape[2]
                    scratch = torch.empty(((npix + 63) // 64) * co, device=w.device, dtype=F32)
                    part = torch.empty(co, device=w.device, dtype=F32)
                    C.call("fcvsr_colsum", g_l.data_ptr(), co, co, npix, scratch.data_ptr(), part.data_ptr(), 0, _st())
                    gb += part
        return (gw, gb) + tuple(gxs)


class _ContextPool(torch.autograd.Function):
    """Soft-max attention pooling of the ContextBlock (CVSR_freq.py:657-690) for the pyramid levels of a BlockRCB in one launch:
    ctx[l, b, :] = sum_p softmax_p(w . x_l[b, p, :]) x_l[b, p, :].  Forward = the inference path's online-softmax kernel without
    its MLP (fcvsr_context_pool_multi), backward = one pass over x (fcvsr_context_pool_backward_multi).  As PyTorch glue this
    was ~10 forward and ~20 backward launches per level."""

    @staticmethod
    def forward(ctx, wmask, *xs):
        import ctypes
        n = len(xs)
        xh = [_nhwc(x) for x in xs]
        B, c = xh[0].shape[0], xh[0].shape[3]
        P = [x.shape[1] * x.shape[2] for x in xh]
        dev = xh[0].device
        wm = wmask.detach().reshape(-1).contiguous()
        pool = torch.empty(n, B, 66, device=dev, dtype=F32)
        partial = torch.empty(sum(B * ((p + 127) // 128) * 66 for p in P), device=dev, dtype=F32)
        counters = torch.zeros(n * B, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            C.call("fcvsr_context_pool_multi", n, (ctypes.c_void_p * n)(*[x.data_ptr() for x in xh]), c, wm.data_ptr(),
                   partial.data_ptr(), pool.data_ptr(), counters.data_ptr(), B, (ctypes.c_int * n)(*P), 0, _st())
        ctx.save_for_backward(wm, pool, *xh)
        ctx.wshape = wmask.shape
        return pool[:, :, :64].contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        import ctypes
        wm, pool, *xh = ctx.saved_tensors
        n = len(xh)
        B, c = xh[0].shape[0], xh[0].shape[3]
        P = [x.shape[1] * x.shape[2] for x in xh]
        dev = xh[0].device
        g = g.contiguous()
        dxs = [torch.empty_like(x) for x in xh]
        Parr = (ctypes.c_int * n)(*P)
        nblk = C.lib().fcvsr_context_pool_backward_blocks(n, Parr)
        dwpart = torch.empty(B * nblk, 64, device=dev, dtype=F32)
        with torch.cuda.device(dev):
            C.call("fcvsr_context_pool_backward_multi", n, (ctypes.c_void_p * n)(*[x.data_ptr() for x in xh]), c, wm.data_ptr(),
                   pool.data_ptr(), g.data_ptr(), (ctypes.c_void_p * n)(*[d.data_ptr() for d in dxs]), dwpart.data_ptr(), B, Parr,
                   _st())
        gw = dwpart.sum(0).view(ctx.wshape) if ctx.needs_input_grad[0] else None
        return (gw,) + tuple(_logical(d) for d in dxs)


class _RcbTail(torch.autograd.Function):
    """RCB tail r_l = lrelu_0.2(res_l + add[l, b]) + r0_l (CVSR_freq.py:720-724) for the pyramid levels of a BlockRCB: forward =
    the inference kernel (fcvsr_rcb_finish_multi, plain fp32 output), backward = one kernel for the three levels (the gradient of
    r0 is the incoming gradient itself).  As PyTorch glue: a broadcast add, an activation and an add per level, and their adjoints."""

    @staticmethod
    def forward(ctx, add, *ts):
        import ctypes
        n = len(ts) // 2
        res = [_nhwc(t) for t in ts[:n]]
        r0 = [_nhwc(t) for t in ts[n:]]
        addc = add.detach().contiguous()                                  # [n, B, 64]
        B = res[0].shape[0]
        outs = [torch.empty_like(t) for t in res]
        vp = lambda ptrs: (ctypes.c_void_p * n)(*ptrs)  # noqa: E731
        ia = lambda v: (ctypes.c_int * n)(*v)  # noqa: E731
        with torch.cuda.device(res[0].device):
            C.call("fcvsr_rcb_finish_multi", n, vp([t.data_ptr() for t in res]), vp([addc[i].data_ptr() for i in range(n)]),
                   vp([t.data_ptr() for t in r0]), vp([t.data_ptr() for t in outs]), None, None, ia([t.shape[1] for t in res]),
                   ia([t.shape[2] for t in res]), B, 0, 1, 0, _st())
        ctx.save_for_backward(addc, *res)
        ctx.n = n
        return tuple(_logical(t) for t in outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        import ctypes
        addc, *res = ctx.saved_tensors
        n = ctx.n
        g = [_nhwc(t) for t in gs]
        B = res[0].shape[0]
        gres = [torch.empty_like(t) for t in res]
        gadd = torch.zeros_like(addc)
        vp = lambda ptrs: (ctypes.c_void_p * n)(*ptrs)  # noqa: E731
        with torch.cuda.device(res[0].device):
            C.call("fcvsr_rcb_finish_backward_multi", n, vp([t.data_ptr() for t in res]), vp([addc[i].data_ptr() for i in range(n)]),
                   vp([t.data_ptr() for t in g]), vp([t.data_ptr() for t in gres]), vp([gadd[i].data_ptr() for i in range(n)]),
                   (ctypes.c_int * n)(*[t.shape[1] * t.shape[2] for t in res]), B, _st())
        return (gadd,) + tuple(_logical(t) for t in gres) + tuple(gs)


class _LevelMix(torch.autograd.Function):
    """BlockRCB cross-level sum (CVSR_freq.py:766-777) of a three-level pyramid in one launch (fcvsr_level_mix_multi):
        out_0 = x_0 + 2 r_0 + up2(tu_0),   out_1 = x_1 + r_1 + td_1 + up2(tu_1),   out_2 = x_2 + 2 r_2 + td_2
    (td_l: the `down` convolution of the pooled finer level, already at level l's size; tu_l: the `up` convolution of the coarser
    level, bilinearly up-sampled in the kernel; level 0 has d = r, level 2 has u = r).  Backward: identities, two scalings and
    ATen's adjoint of the bilinear up-sampling."""

    @staticmethod
    def forward(ctx, x0, x1, x2, r0, r1, r2, td1, td2, tu0, tu1):
        import ctypes
        xs = [_nhwc(t) for t in (x0, x1, x2)]
        rs = [_nhwc(t) for t in (r0, r1, r2)]
        tds = [None, _nhwc(td1), _nhwc(td2)]
        tus = [_nhwc(tu0), _nhwc(tu1), None]
        outs = [torch.empty_like(t) for t in xs]
        B = xs[0].shape[0]
        ptr = lambda t: t.data_ptr() if t is not None else 0  # noqa: E731
        vp = lambda ts: (ctypes.c_void_p * 3)(*[ptr(t) for t in ts])  # noqa: E731
        with torch.cuda.device(xs[0].device):
            C.call("fcvsr_level_mix_multi", 3, vp(xs), 64, vp(outs), 64, vp(rs), (ctypes.c_float * 3)(2.0, 1.0, 2.0), vp(tds), vp(tus), B,
                   (ctypes.c_int * 3)(*[t.shape[1] for t in xs]), (ctypes.c_int * 3)(*[t.shape[2] for t in xs]), None, 0, 0, 0, 1, _st())
        ctx.shapes = [tuple(t.shape) for t in (tu0, tu1)]
        return tuple(_logical(t) for t in outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g0, g1, g2):
        up = torch.ops.aten.upsample_bilinear2d_backward
        gtu0 = up(g0, [g0.shape[2], g0.shape[3]], list(ctx.shapes[0]), False, 2.0, 2.0)
        gtu1 = up(g1, [g1.shape[2], g1.shape[3]], list(ctx.shapes[1]), False, 2.0, 2.0)
        return g0, g1, g2, g0 * 2.0, g1, g2 * 2.0, g1, g2, gtu0, gtu1


def level_mix(xs, rs, tds, tus):
    """xs, rs: three-level lists [B,64,H_l,W_l]; tds = [td_1, td_2] (at the sizes of levels 1, 2); tus = [tu_0, tu_1] (at the sizes
    of levels 1, 2, i.e. half of the level they are added to)."""
    if len(xs) != 3 or xs[0].shape[1] != 64:
        raise ValueError("level_mix handles a three-level 64-channel pyramid")
    return list(_LevelMix.apply(*xs, *rs, *tds, *tus))


def rcb_tail(res, add, r0):
    """[lrelu_0.2(res_l + add[l][:, :, None, None]) + r0_l for l] -- res, r0: lists of up to three [B,64,H_l,W_l] tensors, add [len, B, 64]."""
    if len(res) > 3 or res[0].shape[1] != 64:
        raise ValueError("rcb_tail handles up to three 64-channel tensors")
    return list(_RcbTail.apply(add, *res, *r0))


def context_pool(xs, wmask):
    """xs: up to three [B,64,H_l,W_l] tensors, wmask: the ContextBlock's conv_mask weight [1,64,1,1] -> pooled context [len(xs), B, 64]."""
    if xs[0].shape[1] != 64 or len(xs) > 3:
        raise ValueError("context_pool handles up to three 64-channel tensors")
    return _ContextPool.apply(wmask, *xs)


def conv2d_levels(xs, w, bias=None, mode: str = "tf32"):
    """[conv2d(x, w, bias) for x in xs] for stride-1 convolutions; one launch per pass in "tf32" mode when the shape fits tcgen05."""
    co, ci, k, _ = w.shape
    if mode == "tf32" and 1 < len(xs) <= 3 and _tc_ok(ci, co, k, 1) and _tc_ok(co, ci, k, 1) and co >= 16:
        return list(_Conv2dLevels.apply(w, bias, *xs))
    return [conv2d(x, w, bias, 1, mode) for x in xs]


# ----------------------------------------------------------------------------------------------------------------------
# 2-D real FFTs (torch.fft.rfft2 / irfft2 with norm="backward", CVSR_freq.py:1452-1454, :1499-1504, :2082-2088)
# ----------------------------------------------------------------------------------------------------------------------
_COLW: Dict[Tuple[int, int, float, str], torch.Tensor] = {}


def _col_weights(H: int, W: int, interior: float, device) -> torch.Tensor:
    """[H, W/2+1] real mask: 1 on the DC and Nyquist columns, `interior` elsewhere (the r2c / c2r pair counts interior
    columns once / twice, so their adjoints carry 1/2 / 2 there)."""
    key = (H, W, interior, str(device))
    if key not in _COLW:
        wf = W // 2 + 1
        m = torch.full((H, wf), interior, dtype=F32)
        m[:, 0] = 1.0
        m[:, wf - 1] = 1.0
        _COLW[key] = m.contiguous().to(device)
    return _COLW[key]


def _rfft2_launch(xh: torch.Tensor) -> torch.Tensor:
    B, H, W, c = xh.shape
    wf = W // 2 + 1
    dev = xh.device
    spec = torch.empty(B, H, wf, 2 * c, device=dev, dtype=F32)
    tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
    C.call("fcvsr_fft_r2c_w", xh.data_ptr(), c, spec.data_ptr(), tw_w.data_ptr(), B, H, W, c, _st())
    C.call("fcvsr_fft_c2c_h", spec.data_ptr(), spec.data_ptr(), tw_h.data_ptr(), 0, B, H, wf, c, 0, 1.0, 0, 1, 0, _st())
    return spec


def _c2r_launch(zh: torch.Tensor, W: int, mask: Optional[torch.Tensor], scale: float) -> torch.Tensor:
    """inverse H pass (optional real column mask at load) + torch-semantics c2r W pass, result * scale"""
    B, H, wf, c2 = zh.shape
    c = c2 // 2
    dev = zh.device
    tmp = torch.empty_like(zh)
    y = torch.empty(B, H, W, c, device=dev, dtype=F32)
    tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
    C.call("fcvsr_fft_c2c_h", zh.data_ptr(), tmp.data_ptr(), tw_h.data_ptr(), mask.data_ptr() if mask is not None else 0, B, H, wf,
           c, 1, 1.0, 0, 1, 0, _st())
    C.call("fcvsr_fft_c2r_w", tmp.data_ptr(), y.data_ptr(), c, tw_w.data_ptr(), B, H, W, c, scale, _st())
    return y


class _Rfft2(torch.autograd.Function):
    """[B,C,H,W] real -> [B,2C,H,W/2+1] interleaved spectrum (unnormalised forward transform)."""

    @staticmethod
    def forward(ctx, x):
        xh = _nhwc(x)
        ctx.W = xh.shape[2]
        with torch.cuda.device(x.device):
            return _logical(_rfft2_launch(xh))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gz):
        # X[k] = sum_n x[n] e^{-i theta k n}  =>  dx[n] = Re sum_{k=0}^{W/2} G[k] e^{+i theta k n}: the c2r pass with the interior
        # columns halved (c2r counts them twice), after the conjugate-transpose (= unnormalised inverse) H pass
        g = _nhwc(gz)
        with torch.cuda.device(g.device):
            dx = _c2r_launch(g, ctx.W, _col_weights(g.shape[1], ctx.W, 0.5, g.device), 1.0)
        return _logical(dx)


class _Irfft2(torch.autograd.Function):
    """[B,2C,H,W/2+1] interleaved (any, not necessarily Hermitian) -> [B,C,H,W] = torch.fft.irfft2(z, s=(H, W))."""

    @staticmethod
    def forward(ctx, z, W):
        zh = _nhwc(z)
        ctx.W = W
        with torch.cuda.device(z.device):
            return _logical(_c2r_launch(zh, W, None, 1.0 / (zh.shape[1] * W)))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gx):
        # x[n] = s [Re T_0 + (-1)^n Re T_{W/2} + 2 sum_{0<k<W/2} Re(T_k e^{i theta k n})], T = inverse H pass of z
        #   =>  dT[k] = s c_k r2c(dx)[k] (c_k = 1 on the DC / Nyquist columns, whose imaginary gradient vanishes by itself, 2
        #       elsewhere) and dz = forward H pass of dT
        g = _nhwc(gx)
        B, H, W, c = g.shape
        wf = W // 2 + 1
        dev = g.device
        with torch.cuda.device(dev):
            spec = torch.empty(B, H, wf, 2 * c, device=dev, dtype=F32)
            tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
            C.call("fcvsr_fft_r2c_w", g.data_ptr(), c, spec.data_ptr(), tw_w.data_ptr(), B, H, W, c, _st())
            C.call("fcvsr_fft_c2c_h", spec.data_ptr(), spec.data_ptr(), tw_h.data_ptr(), _col_weights(H, W, 2.0, dev).data_ptr(), B, H,
                   wf, c, 0, 1.0 / (H * W), 0, 1, 0, _st())
        return _logical(spec), None


def rfft2(x):
    return _Rfft2.apply(x)


def irfft2(z, W: int):
    return _Irfft2.apply(z, W)


# ----------------------------------------------------------------------------------------------------------------------
# CorrBlock lookup (CVSR_freq.py:1279-1337)
# ----------------------------------------------------------------------------------------------------------------------
class _Corr(torch.autograd.Function):
    """spec [B, >=256, H, Wf] interleaved with the two 64-channel spectra at channel offsets a_off / b_off -> [B,81,H,Wf]."""

    @staticmethod
    def forward(ctx, spec, a_off, b_off):
        sh = _nhwc(spec)
        B, H, wf, ld = sh.shape
        out = torch.empty(B, H, wf, 81, device=sh.device, dtype=F32)
        with torch.cuda.device(sh.device):
            C.call("fcvsr_corr_gather", sh.data_ptr(), ld, a_off, b_off, out.data_ptr(), 81, B, H, wf, 128, 0, _st())
        ctx.save_for_backward(sh)
        ctx.offs = (a_off, b_off)
        return _logical(out)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (sh,) = ctx.saved_tensors
        g = _nhwc(gout)
        B, H, wf, ld = sh.shape
        ds = torch.zeros_like(sh)
        with torch.cuda.device(sh.device):
            C.call("fcvsr_corr_gather_backward", sh.data_ptr(), ld, ctx.offs[0], ctx.offs[1], g.data_ptr(), 81, ds.data_ptr(), ld, B,
                   H, wf, 128, _st())
        return _logical(ds), None, None


def corr_lookup(spec, a_off: int = 0, b_off: int = 128):
    return _Corr.apply(spec, a_off, b_off)


# ----------------------------------------------------------------------------------------------------------------------
# flow_warp (CVSR_freq.py:1188-1227) and SAC (:1253-1276)
# ----------------------------------------------------------------------------------------------------------------------
class _FlowWarp(torch.autograd.Function):
    """x [B,C,H,W], off [B,2,H,W] (channel 0 = dx, 1 = dy in pixels) -> bilinear samples, zeros outside, align_corners=True."""

    @staticmethod
    def forward(ctx, x, off):
        xh, oh = _nhwc(x), _nhwc(off)
        B, H, W, c = xh.shape
        y = torch.empty_like(xh)
        with torch.cuda.device(xh.device):
            C.call("fcvsr_flow_warp", xh.data_ptr(), c, oh.data_ptr(), 2, y.data_ptr(), c, B, H, W, c, _st())
        ctx.save_for_backward(xh, oh)
        return _logical(y)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        xh, oh = ctx.saved_tensors
        g = _nhwc(gy)
        B, H, W, c = xh.shape
        dx = torch.zeros_like(xh) if ctx.needs_input_grad[0] else None
        do = torch.empty_like(oh) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(xh.device):
            C.call("fcvsr_flow_warp_backward", xh.data_ptr(), c, oh.data_ptr(), 2, g.data_ptr(), c,
                   dx.data_ptr() if dx is not None else 0, c, do.data_ptr() if do is not None else 0, B, H, W, c, _st())
        return (_logical(dx) if dx is not None else None), (_logical(do) if do is not None else None)


class _Sac(torch.autograd.Function):
    """wp [B,C,H,W], taps [B,3C,H,W] with channel t*C + c -> vertical then horizontal 3-tap pass with the same taps."""

    @staticmethod
    def forward(ctx, wp, taps):
        wh, kh = _nhwc(wp), _nhwc(taps)
        B, H, W, c = wh.shape
        y = torch.empty_like(wh)
        with torch.cuda.device(wh.device):
            C.call("fcvsr_sac", wh.data_ptr(), c, kh.data_ptr(), 3 * c, y.data_ptr(), c, B, H, W, c, _st())
        ctx.save_for_backward(wh, kh)
        return _logical(y)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        wh, kh = ctx.saved_tensors
        g = _nhwc(gy)
        B, H, W, c = wh.shape
        dw = torch.empty_like(wh) if ctx.needs_input_grad[0] else None
        dk = torch.empty_like(kh) if ctx.needs_input_grad[1] else None
        scratch = torch.empty_like(wh)
        with torch.cuda.device(wh.device):
            C.call("fcvsr_sac_backward", wh.data_ptr(), c, kh.data_ptr(), 3 * c, g.data_ptr(), c, scratch.data_ptr(),
                   dk.data_ptr() if dk is not None else 0, 3 * c, dw.data_ptr() if dw is not None else 0, c, B, H, W, c, _st())
        return (_logical(dw) if dw is not None else None), (_logical(dk) if dk is not None else None)


def flow_warp(x, off):
    return _FlowWarp.apply(x, off)


def sac(wp, taps):
    return _Sac.apply(wp, taps)
