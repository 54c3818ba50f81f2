"""Training-loss forward of the reference (CVSR_train/opt/loss.py:20-31) on the GPU.

    CharbonnierLoss(x, y, mean_res=False) -> 0-dim tensor:  sum(sqrt((x - y)^2 + 1e-4))

Same name, arguments and reduction (sum) as the reference function; CUDA fp32 only, forward only in this round (the
backward kernels of the model do not exist yet, so a call that needs autograd raises)."""
from __future__ import annotations

import torch

from .. import _capi as C


def CharbonnierLoss(x: torch.Tensor, y: torch.Tensor, mean_res: bool = False) -> torch.Tensor:
    if x.shape != y.shape:
        raise ValueError(f"shape mismatch {tuple(x.shape)} vs {tuple(y.shape)}")     # the reference prints "!!!" and fails in x - y
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise TypeError("fcvsr_b200 loss kernels are fp32")
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        raise NotImplementedError("fcvsr_b200: backward kernels are not implemented in this round; call under torch.no_grad()")
    x, y = x.contiguous(), y.contiguous()
    scratch = torch.empty(max(592, x.shape[0]), device=x.device, dtype=torch.float64)
    out = torch.empty((), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        C.call("fcvsr_charbonnier_loss", x.data_ptr(), y.data_ptr(), x.numel(), x.shape[0], int(mean_res), 1e-4,
               scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return out
