"""Drop-in for the reference's compiled extension module `deform_conv_cuda` (CVSR_train/ops/dcn/src/deform_conv_cuda.cpp:681-695):
the same five entry points with the same argument order, so that the reference's own `deform_conv.py` runs on this library
by replacing `from . import deform_conv_cuda` with `from fcvsr_b200.ops import deform_conv_cuda`.

    deform_conv_forward_cuda               (.cpp:151;  call site deform_conv.py:52-57)
    deform_conv_backward_input_cuda        (.cpp:260;  deform_conv.py:76-82)
    deform_conv_backward_parameters_cuda   (.cpp:373;  deform_conv.py:86-92)
    modulated_deform_conv_cuda_forward     (.cpp:486;  deform_conv.py:144-148)
    modulated_deform_conv_cuda_backward    (.cpp:566;  deform_conv.py:161-166)

As in the reference the caller pre-allocates every output (and zero-fills the gradients, deform_conv.py:71-72,155-159); results
are written in place.  The scratch tensors of the reference ABI (`columns`, `ones`) are accepted and ignored: the kernels
(csrc/dcn.cu, dcn_tc.cu, dcn_bwd.cu) have no column buffer.  Tensors are NCHW contiguous CUDA tensors; fp16 / fp64 are
computed in fp32 (AT_DISPATCH_FLOATING_TYPES_AND_HALF in the reference).  Failures raise (the reference throws c10::Error
through pybind); the v1 entries return 1 like the reference's `int` functions.
"""
from __future__ import annotations

import torch

from . import dcn as _D


def _chk(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("input tensor has to be on GPU")            # TORCH_CHECK(input.is_cuda()) in the reference
        if not t.is_contiguous():
            raise RuntimeError("input tensor has to be contiguous")        # deform_conv_cuda.cpp:493-494


def _f(t):
    return t if t.dtype == torch.float32 else t.float()


def deform_conv_forward_cuda(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH, group,
                             deformable_group, im2col_step):
    _chk(input, weight, offset, output)
    if tuple(weight.shape[2:]) != (kH, kW):
        raise RuntimeError(f"kernel size should be consistent with weight, but got kH: {kH} kW: {kW} weight.size(2): "
                           f"{weight.shape[2]}, weight.size(3): {weight.shape[3]}")
    y = _D._launch(_f(input), _f(offset), None, _f(weight), None, (dH, dW), (padH, padW), (dilationH, dilationW), group,
                   deformable_group)
    if tuple(output.shape) != tuple(y.shape):
        output.resize_(y.shape)                                            # the reference views / resizes its output (.cpp:196-199)
    output.copy_(y)
    return 1


def deform_conv_backward_input_cuda(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH, padW, padH,
                                    dilationW, dilationH, group, deformable_group, im2col_step):
    _chk(input, offset, gradOutput, gradInput, gradOffset, weight)
    gx, goff, _, _, _ = _D._backward(_f(input), _f(offset), None, _f(weight), _f(gradOutput), (dH, dW), (padH, padW),
                                     (dilationH, dilationW), group, deformable_group, (True, True, False, False), False)
    gradInput.copy_(gx)                                                    # the reference overwrites both (.cpp:331-349)
    gradOffset.copy_(goff)
    return 1


def deform_conv_backward_parameters_cuda(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH, dilationW,
                                         dilationH, group, deformable_group, scale, im2col_step):
    _chk(input, offset, gradOutput, gradWeight)
    w_like = torch.zeros(gradWeight.shape, device=gradWeight.device, dtype=torch.float32)
    _, _, _, gw, _ = _D._backward(_f(input), _f(offset), None, w_like, _f(gradOutput), (dH, dW), (padH, padW), (dilationH, dilationW),
                                  group, deformable_group, (False, False, False, True), False)
    gradWeight.add_(gw.to(gradWeight.dtype), alpha=float(scale))           # gradWeight += scale * ... (.cpp:458-462)
    return 1


def modulated_deform_conv_cuda_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h,
                                       stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    _chk(input, weight, offset, mask, output)
    if tuple(weight.shape[2:]) != (kernel_h, kernel_w):
        raise RuntimeError(f"Input shape and kernel shape wont match: ({kernel_h} x {kernel_w} vs {weight.shape[2]} x {weight.shape[3]}).")
    y = _D._launch(_f(input), _f(offset), _f(mask), _f(weight), _f(bias) if with_bias else None, (stride_h, stride_w), (pad_h, pad_w),
                   (dilation_h, dilation_w), group, deformable_group)
    if tuple(output.shape) != tuple(y.shape):
        output.resize_(y.shape)
    output.copy_(y)


def modulated_deform_conv_cuda_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias,
                                        grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                                        dilation_h, dilation_w, group, deformable_group, with_bias):
    _chk(input, weight, offset, mask, grad_input, grad_weight, grad_offset, grad_mask, grad_output)
    gx, goff, gmask, gw, gb = _D._backward(_f(input), _f(offset), _f(mask), _f(weight), _f(grad_output), (stride_h, stride_w),
                                           (pad_h, pad_w), (dilation_h, dilation_w), group, deformable_group,
                                           (True, True, True, True), bool(with_bias))
    grad_input.copy_(gx)                                                   # per-sample results are assigned (.cpp:640-653) ...
    grad_offset.copy_(goff)
    grad_mask.copy_(gmask)
    grad_weight.add_(gw.to(grad_weight.dtype))                             # ... the parameter gradients accumulate (.cpp:667-676)
    if with_bias:
        grad_bias.add_(gb.to(grad_bias.dtype))
