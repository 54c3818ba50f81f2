"""Deformable-convolution operator with the signature of the reference's CVSR_train/ops/dcn/deform_conv.py.

    deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64)                                                     (deform_conv.py:17-26,186)
    modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                          groups=1, deformable_groups=1)                            (deform_conv.py:117-127,187)
    DeformConv / DeformConvPack / ModulatedDeformConv / ModulatedDeformConvPack     (deform_conv.py:190-337)

Same tensor contract (NCHW contiguous CUDA tensors; fp32, and -- as the reference dispatches
AT_DISPATCH_FLOATING_TYPES_AND_HALF -- fp16 / fp64 tensors, which are computed in fp32 and returned in their own dtype; offset [B, dg*2*kh*kw, Ho, Wo] with (dh, dw)
interleaved per tap; mask [B, dg*kh*kw, Ho, Wo]; NotImplementedError on CPU, :46-47,:136-137).  The native
entry it binds is ``fcvsr_modulated_deform_conv_forward`` (include/fcvsr_b200.h), a fused gather + GEMM
kernel that replaces deform_conv_forward_cuda / modulated_deform_conv_cuda_forward
(ops/dcn/src/deform_conv_cuda.cpp:151,486) without the HBM column buffer.  ``im2col_step`` is accepted for
API compatibility (v1 still validates that it divides the batch, :49-51) but there is no column buffer
to chunk.

Autograd: ``deform_conv`` / ``modulated_deform_conv`` (and the DeformConv / ModulatedDeformConv modules) are
``torch.autograd.Function``s like the reference's (deform_conv.py:14-98,:114-183); their backward binds
``fcvsr_modulated_deform_conv_backward`` (csrc/dcn_bwd.cu), which replaces modulated_deform_conv_cuda_backward /
deform_conv_backward_input_cuda / deform_conv_backward_parameters_cuda (deform_conv_cuda.cpp:566,260,373).  The *Pack
modules train as well: their conv_offset / conv_offset_mask layer is an autograd op whose backward is the same entry with
zero offsets (a deformable convolution with zero offsets is the convolution).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .. import _capi as C


def _check(*tensors, allow_grad=False):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NotImplementedError("deformable convolution is CUDA-only (as the reference, deform_conv.py:46-47)")
        if t.dtype not in (torch.float32, torch.float16, torch.float64):
            raise TypeError("deformable convolution expects floating-point tensors (AT_DISPATCH_FLOATING_TYPES_AND_HALF, "
                            "deform_conv_cuda_kernel.cu:360)")
        if not allow_grad and torch.is_grad_enabled() and t.requires_grad:
            raise NotImplementedError("fcvsr_b200: this path has no backward kernel; call under torch.no_grad()")


def _require_fp32(*tensors):
    """The *Pack modules own fp32 parameters and hand raw pointers to the kernels: no dtype conversion on that path."""
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise TypeError("the DeformConvPack / ModulatedDeformConvPack modules of fcvsr_b200 are fp32")


def _out_hw(h, w, kh, kw, stride, padding, dilation):
    ho = (h + 2 * padding[0] - (dilation[0] * (kh - 1) + 1)) // stride[0] + 1
    wo = (w + 2 * padding[1] - (dilation[1] * (kw - 1) + 1)) // stride[1] + 1
    if ho <= 0 or wo <= 0:
        raise ValueError("convolution input is too small")
    return ho, wo


# "tf32": shapes in the tensor-core kernel's class (groups == 1, Cin % 32 == 0, Cout % 16 == 0, Cout <= 256) run on
# tcgen05 with TF32-rounded operands and fp32 accumulation (csrc/dcn_tc.cu, |error| <~ 1e-3 of the output scale, the same
# contract as the model's "tf32" mode); everything else, and everything under "fp32", runs the exact CUDA-core kernel.
PRECISION = "tf32"
BACKWARD_NHWC = True      # backward: NHWC scratch copies + red.global.add.v4.f32 when the channel counts allow (dcn_bwd.cu)


def _launch(x, offset, mask, weight, bias, stride, padding, dilation, groups, dg, off_bs=0, mask_bs=0, sigmoid=0):
    b, cin, h, w = x.shape
    cout, cin_g, kh, kw = weight.shape
    if cin_g * groups != cin:
        raise ValueError("weight shape does not match input channels / groups")
    ho, wo = _out_hw(h, w, kh, kw, stride, padding, dilation)
    y = x.new_empty(b, cout, ho, wo)
    st = torch.cuda.current_stream().cuda_stream
    args = (x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else 0, offset.data_ptr(),
            mask.data_ptr() if mask is not None else 0, y.data_ptr(), b, cin, h, w, cout, kh, kw, stride[0], stride[1],
            padding[0], padding[1], dilation[0], dilation[1], groups, dg, off_bs, mask_bs, sigmoid, st)
    if PRECISION not in ("tf32", "fp32"):
        raise ValueError(f"unknown DCN precision {PRECISION!r}")
    with torch.cuda.device(x.device):
        if PRECISION == "tf32" and groups == 1 and cin % 32 == 0 and (cin // dg) % 4 == 0 and cout % 16 == 0 and 16 <= cout <= 256:
            scratch = x.new_empty(x.numel())             # NHWC copy of the input made by the kernel
            rc = C.try_call("fcvsr_modulated_deform_conv_forward_tc", *args[:-1], scratch.data_ptr(), 0, st)
            if rc == 0:
                return y
            if rc != C.ERR_UNSUPPORTED:
                raise RuntimeError(f"fcvsr_modulated_deform_conv_forward_tc failed with status {rc}")
        C.call("fcvsr_modulated_deform_conv_forward", *args)
    return y


def _backward(x, offset, mask, weight, grad_out, stride, padding, dilation, groups, dg, need, with_bias):
    """One call of fcvsr_modulated_deform_conv_backward; need = (input, offset, mask, weight) flags.  Gradients are
    zero-filled here and accumulated by the kernels (deform_conv.py:155-159)."""
    b, cin, h, w = x.shape
    cout, _, kh, kw = weight.shape
    grad_out = grad_out.contiguous()
    gx = torch.zeros_like(x) if need[0] else None
    goff = torch.zeros_like(offset) if (offset is not None and need[1]) else None
    gmask = torch.zeros_like(mask) if (mask is not None and need[2]) else None
    gw = torch.zeros_like(weight) if need[3] else None
    gb = x.new_zeros(cout) if with_bias else None
    ptr = lambda t: t.data_ptr() if t is not None else 0  # noqa: E731
    fast = BACKWARD_NHWC and (cin // groups) % 4 == 0 and (cin // dg) % 4 == 0
    scratch = x.new_empty(2 * x.numel()) if fast else None      # NHWC copies of input / grad_input (vector reductions)
    with torch.cuda.device(x.device):
        C.call("fcvsr_modulated_deform_conv_backward", x.data_ptr(), weight.data_ptr(), ptr(offset), ptr(mask),
               grad_out.data_ptr(), ptr(gx), ptr(gw), ptr(gb), ptr(goff), ptr(gmask), b, cin, h, w, cout, kh, kw,
               stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], groups, dg, ptr(scratch),
               torch.cuda.current_stream().cuda_stream)
    return gx, goff, gmask, gw, gb


class DeformConvFunction(torch.autograd.Function):
    """deform_conv.py:14-98."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        _check(input, offset, weight, allow_grad=True)
        cur = min(im2col_step, input.shape[0])
        assert input.shape[0] % cur == 0, "im2col step must divide batchsize"
        stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
        kh, kw = weight.shape[2:]
        ho, wo = _out_hw(input.shape[2], input.shape[3], kh, kw, stride, padding, dilation)
        if tuple(offset.shape) != (input.shape[0], deformable_groups * 2 * kh * kw, ho, wo):
            raise ValueError(f"invalid offset shape {tuple(offset.shape)}")
        ctx.dtypes = (input.dtype, offset.dtype, weight.dtype)
        input, offset, weight = input.float().contiguous(), offset.float().contiguous(), weight.float().contiguous()
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        ctx.save_for_backward(input, offset, weight)
        return _launch(input, offset, None, weight, None, stride, padding, dilation, groups, deformable_groups).to(ctx.dtypes[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, weight = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.cfg
        n = ctx.needs_input_grad
        need_data = n[0] or n[1]                      # the reference computes both together (:76-82)
        gx, goff, _, gw, _ = _backward(input, offset, None, weight, grad_output.float(), stride, padding, dilation, groups, dg,
                                       (need_data, need_data, False, n[2]), False)
        cast = lambda t, d: t.to(d) if t is not None else None  # noqa: E731
        return cast(gx, ctx.dtypes[0]), cast(goff, ctx.dtypes[1]), cast(gw, ctx.dtypes[2]), None, None, None, None, None, None


class ModulatedDeformConvFunction(torch.autograd.Function):
    """deform_conv.py:114-183."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        _check(input, offset, mask, weight, bias, allow_grad=True)
        stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)   # scalars in the reference (:179-182)
        kh, kw = weight.shape[2:]
        ho, wo = _out_hw(input.shape[2], input.shape[3], kh, kw, stride, padding, dilation)
        if tuple(offset.shape) != (input.shape[0], deformable_groups * 2 * kh * kw, ho, wo):
            raise ValueError(f"invalid offset shape {tuple(offset.shape)}")
        if tuple(mask.shape) != (input.shape[0], deformable_groups * kh * kw, ho, wo):
            raise ValueError(f"invalid mask shape {tuple(mask.shape)}")
        ctx.dtypes = (input.dtype, offset.dtype, mask.dtype, weight.dtype, bias.dtype if bias is not None else None)
        input, offset, mask, weight = (t.float().contiguous() for t in (input, offset, mask, weight))
        bias = bias.float().contiguous() if bias is not None else None
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups, bias is not None)
        ctx.save_for_backward(input, offset, mask, weight)
        return _launch(input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups).to(ctx.dtypes[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight = ctx.saved_tensors
        stride, padding, dilation, groups, dg, with_bias = ctx.cfg
        n = ctx.needs_input_grad
        gx, goff, gmask, gw, gb = _backward(input, offset, mask, weight, grad_output.float(), stride, padding, dilation, groups,
                                            dg, (n[0], n[1], n[2], n[3]), with_bias and n[4])
        cast = lambda t, d: t.to(d) if (t is not None and d is not None) else t  # noqa: E731
        d = ctx.dtypes
        return cast(gx, d[0]), cast(goff, d[1]), cast(gmask, d[2]), cast(gw, d[3]), cast(gb, d[4]), None, None, None, None, None


def deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64):
    """deform_conv.py:186 (``DeformConvFunction.apply``; keywords accepted as well)."""
    return DeformConvFunction.apply(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step)


def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1):
    """deform_conv.py:187 (``ModulatedDeformConvFunction.apply``; keywords accepted as well)."""
    return ModulatedDeformConvFunction.apply(input, offset, mask, weight, bias, stride, padding, dilation, groups,
                                             deformable_groups)


def _offset_conv_raw(x, w, bias, s):
    cout, cin, kh, kw = w.shape
    b, _, h, wd = x.shape
    ho, wo = (h + 2 * (kh // 2) - kh) // s + 1, (wd + 2 * (kw // 2) - kw) // s + 1
    y = x.new_empty(b, cout, ho, wo)
    wp = w.detach().permute(2, 3, 1, 0).contiguous()
    with torch.cuda.device(x.device):
        C.call("fcvsr_conv2d_direct", x.data_ptr(), 0, 1, wp.data_ptr(), bias.data_ptr() if bias is not None else 0,
               0, 0, 0, 0, y.data_ptr(), 0, b, h, wd, cin, cout, kh, s, C.ACT_NONE, 0.0, 0, 0, 1, 0, 0, 0, 0,
               torch.cuda.current_stream().cuda_stream)
    return y


class _OffsetConvFunction(torch.autograd.Function):
    """conv_offset / conv_offset_mask as an autograd op: forward on fcvsr_conv2d_direct; backward on the DCN backward kernels
    with zero offsets and no mask -- a deformable convolution with zero offsets IS the convolution, so grad_input,
    grad_weight and grad_bias come from fcvsr_modulated_deform_conv_backward(offset = NULL)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        x, weight = x.contiguous(), weight.contiguous()
        ctx.stride, ctx.with_bias = stride, bias is not None
        ctx.save_for_backward(x, weight)
        return _offset_conv_raw(x, weight, bias, stride)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        x, weight = ctx.saved_tensors
        k, s = weight.shape[2], ctx.stride
        n = ctx.needs_input_grad
        gx, _, _, gw, gb = _backward(x, None, None, weight, grad_output, (s, s), (k // 2, k // 2), (1, 1), 1, 1,
                                     (n[0], False, False, n[1]), ctx.with_bias and n[2])
        return gx, gw, gb, None


def _offset_conv(x, conv: nn.Conv2d):
    """conv_offset / conv_offset_mask of the *Pack modules (deform_conv.py:243-250,:315-323) on our own
    convolution kernel (NCHW in, NCHW out), autograd-capable."""
    _check(x, conv.weight, conv.bias, allow_grad=True)
    _require_fp32(x, conv.weight, conv.bias)
    kh, kw = conv.weight.shape[2:]
    if kh != kw or conv.stride[0] != conv.stride[1] or conv.padding[0] != kh // 2 or conv.padding[1] != kw // 2:
        raise NotImplementedError("conv_offset: only square kernels with 'same' padding k//2 are supported")
    return _OffsetConvFunction.apply(x, conv.weight, conv.bias, conv.stride[0])


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        assert in_channels % groups == 0, "in_channels {} cannot be divisible by groups {}".format(in_channels, groups)
        assert out_channels % groups == 0, "out_channels {} cannot be divisible by groups {}".format(out_channels, groups)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class DeformConvPack(DeformConv):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding),
                                     bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        offset = _offset_conv(x, self.conv_offset)
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups = groups, deformable_groups
        self.with_bias = bias
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                                     self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset_mask = nn.Conv2d(self.in_channels,
                                          self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                          kernel_size=self.kernel_size, stride=_pair(self.stride),
                                          padding=_pair(self.padding), bias=True)
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        """chunk -> cat(o1, o2) -> sigmoid(mask) (:331-334) are fused away: offset is the first 2/3 of the
        conv_offset_mask output, the mask the last third (sigmoid applied inside the DCN kernel)."""
        _check(x, self.weight, self.bias, allow_grad=True)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: the reference's own sequence (:331-336) on autograd-capable ops
            out = _offset_conv(x, self.conv_offset_mask)
            o1, o2, mask = torch.chunk(out, 3, dim=1)
            return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride,
                                         self.padding, self.dilation, self.groups, self.deformable_groups)
        kh, kw = self.kernel_size
        dg = self.deformable_groups
        stride, padding, dilation = _pair(self.stride), _pair(self.padding), _pair(self.dilation)
        b, cin, h, w = x.shape
        c3 = dg * 3 * kh * kw
        tc = (PRECISION == "tf32" and self.groups == 1 and kh == kw and kh in (1, 3) and stride == (1, 1) and dilation == (1, 1)
              and padding == (kh // 2, kh // 2) and cin % 32 == 0 and (cin // dg) % 4 == 0 and self.out_channels % 16 == 0
              and 16 <= self.out_channels <= 256 and c3 % 16 == 0)
        if tc:
            # tensor-core path: ONE pixel-major TF32-rounded copy of x feeds both the conv_offset_mask convolution (tcgen05
            # implicit GEMM, NHWC output) and the fused gather + GEMM DCN kernel, which reads offsets and mask straight from
            # that NHWC output (sigmoid applied on the fly)
            y = self._forward_tc(x.contiguous(), b, cin, h, w, c3, kh, dg)
            if y is not None:
                return y
        out = _offset_conv(x, self.conv_offset_mask)
        b, c3, ho, wo = out.shape
        kk = self.kernel_size[0] * self.kernel_size[1]
        n_off = self.deformable_groups * 2 * kk
        mask = out.view(b, -1)[:, n_off * ho * wo:]
        stride, padding, dilation = _pair(self.stride), _pair(self.padding), _pair(self.dilation)
        y = x.new_empty(b, self.out_channels, ho, wo)
        with torch.cuda.device(x.device):
            C.call("fcvsr_modulated_deform_conv_forward", x.contiguous().data_ptr(), self.weight.data_ptr(),
                   self.bias.data_ptr() if self.bias is not None else 0, out.data_ptr(), mask.data_ptr(), y.data_ptr(), b,
                   self.in_channels, x.shape[2], x.shape[3], self.out_channels, self.kernel_size[0], self.kernel_size[1],
                   stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], self.groups,
                   self.deformable_groups, c3 * ho * wo, c3 * ho * wo, 1, torch.cuda.current_stream().cuda_stream)
        return y

    def _forward_tc(self, x, b, cin, h, w, c3, k, dg):
        st = torch.cuda.current_stream().cuda_stream
        key = (self.conv_offset_mask.weight.data_ptr(), self.conv_offset_mask.weight._version)
        if getattr(self, "_tc_pack_key", None) != key:          # [Cout][k*k*Cin] K-major, TF32-rounded (conv_tc.cu layout)
            wt = self.conv_offset_mask.weight.detach().permute(0, 2, 3, 1).reshape(c3, k * k * cin).contiguous()
            bits = wt.view(torch.int32)
            self._tc_pack = ((bits + 0x1000) & -8192).view(torch.float32)
            self._tc_pack_key = key
        xt = x.new_empty(b, h, w, cin)
        off = x.new_empty(b, h, w, c3)
        y = x.new_empty(b, self.out_channels, h, w)
        kk = k * k
        with torch.cuda.device(x.device):
            C.call("fcvsr_nchw_to_nhwc", x.data_ptr(), xt.data_ptr(), b, cin, h, w, 1, st)
            rc = C.try_call("fcvsr_conv2d_tc", xt.data_ptr(), cin, self._tc_pack.data_ptr(), self.conv_offset_mask.bias.data_ptr(), 0, 0,
                            0, 0, off.data_ptr(), c3, b, h, w, cin, c3, k, C.ACT_NONE, 0.0, 0, 0, 0, 0, 0, 0, 0, st)
            if rc == C.ERR_UNSUPPORTED:
                return None
            if rc != 0:
                raise RuntimeError(f"fcvsr_conv2d_tc failed with status {rc}")
            rc = C.try_call("fcvsr_modulated_deform_conv_forward_tc", xt.data_ptr(), self.weight.data_ptr(),
                            self.bias.data_ptr() if self.bias is not None else 0, off.data_ptr(),
                            off.data_ptr() + dg * 2 * kk * 4, y.data_ptr(), b, cin, h, w, self.out_channels, k, k, 1, 1, k // 2,
                            k // 2, 1, 1, 1, dg, 0, 0, 1, xt.data_ptr(), c3, st)
            if rc == C.ERR_UNSUPPORTED:
                return None
            if rc != 0:
                raise RuntimeError(f"fcvsr_modulated_deform_conv_forward_tc failed with status {rc}")
        return y
