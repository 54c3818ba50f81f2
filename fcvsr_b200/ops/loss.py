"""Training loss of the reference (CVSR_train/opt/loss.py:20-31) on the GPU.

    CharbonnierLoss(x, y, mean_res=False) -> 0-dim tensor:  sum(sqrt((x - y)^2 + 1e-4))

Same name, arguments and reduction (sum) as the reference function; CUDA fp32 only.  It is an autograd Function: the
backward binds ``fcvsr_charbonnier_loss_backward`` (what autograd derives for the reference's expression)."""
from __future__ import annotations

import torch

from .. import _capi as C

_EPS = 1e-4


class _Charbonnier(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, mean_res):
        x, y = x.contiguous(), y.contiguous()
        # the forward kernel reads 16-byte lanes: a contiguous view with a storage offset (sr[1:], narrow()) can be
        # misaligned, which the reference accepts -- realign it with a copy instead of failing
        if x.data_ptr() % 16:
            x = x.clone()
        if y.data_ptr() % 16:
            y = y.clone()
        scratch = torch.empty(max(592, x.shape[0]), device=x.device, dtype=torch.float64)
        out = torch.empty((), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            C.call("fcvsr_charbonnier_loss", x.data_ptr(), y.data_ptr(), x.numel(), x.shape[0], int(mean_res), _EPS,
                   scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.mean_res = bool(mean_res)
        ctx.save_for_backward(x, y, scratch)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, y, scratch = ctx.saved_tensors
        grad_out = grad_out.to(torch.float32).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            C.call("fcvsr_charbonnier_loss_backward", x.data_ptr(), y.data_ptr(), x.numel(), x.shape[0], int(ctx.mean_res),
                   _EPS, grad_out.data_ptr(), scratch.data_ptr(), gx.data_ptr() if gx is not None else 0,
                   gy.data_ptr() if gy is not None else 0, torch.cuda.current_stream().cuda_stream)
        return gx, gy, None


def CharbonnierLoss(x: torch.Tensor, y: torch.Tensor, mean_res: bool = False) -> torch.Tensor:
    if x.shape != y.shape:
        raise ValueError(f"shape mismatch {tuple(x.shape)} vs {tuple(y.shape)}")     # the reference prints "!!!" and fails in x - y
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise TypeError("fcvsr_b200 loss kernels are fp32")
    return _Charbonnier.apply(x, y, mean_res)
