"""One data-parallel training step with the reference's semantics (CVSR_train/train_LD_freqCVSR_22.py:243-251):

    optimizer.zero_grad(); sr = model(frames); loss = CharbonnierLoss(sr, hr); loss.backward(); optimizer.step()

plus the gradient all-reduce that the mmedit configuration wraps around it (one replica per GPU,
fcvsr_redsLD_QP22.py:144).  Host logic only: `model` is any autograd-capable module -- the FCVSR forward of this package is not
yet, so today this runs the DCN modules / stand-ins and is covered by the 2-rank gloo test; the loss, the
optimizer and the reducer it is meant to be used with are `fcvsr_b200.ops.loss.CharbonnierLoss`, `fcvsr_b200.ops.optim.Adam`
and `fcvsr_b200.gradsync.GradAllReducer`.

Charbonnier is sum-reduced (opt/loss.py:30), so averaging the gradients over G ranks makes one step equal to a single-GPU
step on the concatenated batch with the learning rate divided by G -- the reference behaves the same way; it is documented,
not "fixed" (SURVEY 8e).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def train_step(model: torch.nn.Module, optimizer: torch.optim.Optimizer, frames: torch.Tensor, hr: torch.Tensor,
               loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], reducer=None) -> torch.Tensor:
    """Returns this rank's (detached) loss.  `reducer`: a GradAllReducer over model.parameters(), or None for one replica."""
    optimizer.zero_grad(set_to_none=True)
    if reducer is not None:
        reducer.start_step()
    sr = model(frames)
    loss = loss_fn(sr, hr)
    loss.backward()                      # the reducer's hooks launch each bucket's all-reduce as its gradients complete
    if reducer is not None:
        reducer.finish()
    optimizer.step()
    return loss.detach()


def replicas_in_sync(model: torch.nn.Module, group=None, atol: float = 0.0) -> bool:
    """Debug check: every rank holds the same parameters (max over ranks of |p - p_rank0| <= atol)."""
    import torch.distributed as dist
    ok = True
    for p in model.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0, group=group)
        ok = ok and float((p.detach() - ref).abs().max()) <= atol
    flag = torch.tensor([1.0 if ok else 0.0], device=next(model.parameters()).device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item() > 0)
