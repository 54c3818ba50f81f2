"""Restorer-level training step of the mmedit route, host logic around the kernel library.

The reference's REDS / Vimeo training goes `IterBasedRunner -> BasicVSR.train_step -> BasicRestorer.forward_train ->
generator(lq) -> pixel_loss` (mmedit_train/mmedit/models/restorers/basicvsr.py:85-117, basic_restorer.py:76-95), with
`MSELoss(mean)` and `fix_iter = 100` in the FCVSR configuration, Adam(lr 5e-6, betas (0.9, 0.99)) and the CosineRestart
schedule (configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:7-9,:112-128), checkpoints with optimizer state (:130).  This module
mirrors those pieces without mmcv (absent from the image):

  * `MSELoss` / `CharbonnierLoss` / `L1Loss`  -- mmedit's pixel losses (pixelwise_loss.py:54-190: loss_weight, reduction
    'mean' | 'sum'; element-wise `weight` masks are not supported) on `fcvsr_pixel_loss[_backward]`;
  * `BasicVSRRestorer.train_step(data_batch, optimizer)` -- same protocol and return value (`log_vars`, `num_samples`,
    `results`), the `step_counter` buffer and the `fix_iter` freeze of 'spynet' / 'edvr' parameters (FCVSR has none, so the
    freeze is a no-op exactly as in the reference);
  * `CosineRestartLR` -- mmcv's CosineRestartLrUpdaterHook as a per-iteration torch scheduler;
  * `save_checkpoint` / `load_checkpoint` -- the mmcv layout {'meta', 'state_dict', 'optimizer'}.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import _capi as C

_KIND = {"charbonnier": 0, "mse": 1, "l1": 2}


class _PixelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, kind, eps, scale):
        x, y = x.contiguous(), y.contiguous()
        scratch = torch.empty(592, device=x.device, dtype=torch.float64)
        out = torch.empty((), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            C.call("fcvsr_pixel_loss", x.data_ptr(), y.data_ptr(), x.numel(), kind, eps, scale, scratch.data_ptr(), out.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
        ctx.cfg = (kind, eps, scale)
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, y = ctx.saved_tensors
        kind, eps, scale = ctx.cfg
        grad_out = grad_out.to(torch.float32).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            C.call("fcvsr_pixel_loss_backward", x.data_ptr(), y.data_ptr(), x.numel(), kind, eps, scale, grad_out.data_ptr(),
                   gx.data_ptr() if gx is not None else 0, gy.data_ptr() if gy is not None else 0,
                   torch.cuda.current_stream().cuda_stream)
        return gx, gy, None, None, None


class _MMEditLoss(nn.Module):
    _kind = "mse"

    def __init__(self, loss_weight=1.0, reduction="mean", sample_wise=False, eps=1e-12):
        super().__init__()
        if reduction not in ("mean", "sum"):          # 'none' would return the element-wise map: not a training-step loss
            raise ValueError(f"Unsupported reduction mode: {reduction}. Supported ones are: ['mean', 'sum']")
        self.loss_weight, self.reduction, self.sample_wise, self.eps = loss_weight, reduction, sample_wise, eps

    def forward(self, pred, target, weight=None, **kwargs):
        if weight is not None:
            raise NotImplementedError("element-wise loss weights (masked_loss) are not built; the FCVSR configurations pass none")
        if pred.shape != target.shape:
            raise ValueError(f"shape mismatch {tuple(pred.shape)} vs {tuple(target.shape)}")
        if not (pred.is_cuda and target.is_cuda):
            raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
        if pred.dtype != torch.float32 or target.dtype != torch.float32:
            raise TypeError("fcvsr_b200 loss kernels are fp32")
        scale = self.loss_weight / pred.numel() if self.reduction == "mean" else self.loss_weight
        return _PixelLoss.apply(pred, target, _KIND[self._kind], float(self.eps), float(scale))


class MSELoss(_MMEditLoss):                     # pixelwise_loss.py:95-132
    _kind = "mse"

    def __init__(self, loss_weight=1.0, reduction="mean", sample_wise=False):
        super().__init__(loss_weight, reduction, sample_wise)


class L1Loss(_MMEditLoss):                      # pixelwise_loss.py:54-92
    _kind = "l1"

    def __init__(self, loss_weight=1.0, reduction="mean", sample_wise=False):
        super().__init__(loss_weight, reduction, sample_wise)


class CharbonnierLoss(_MMEditLoss):             # pixelwise_loss.py:135-190 (eps 1e-12, mean -- not opt/loss.py's sum / 1e-4)
    _kind = "charbonnier"


class BasicVSRRestorer(nn.Module):
    """BasicVSR restorer as FCVSR's configurations use it (restorers/basicvsr.py:36-117 + basic_restorer.py:76-95)."""

    def __init__(self, generator: nn.Module, pixel_loss: nn.Module, train_cfg: Optional[dict] = None, test_cfg: Optional[dict] = None):
        super().__init__()
        self.generator, self.pixel_loss = generator, pixel_loss
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.fix_iter = train_cfg.get("fix_iter", 0) if train_cfg else 0      # basicvsr.py:46
        self.is_weight_fixed = False
        self.register_buffer("step_counter", torch.zeros(1))                  # basicvsr.py:50
        self._steps = 0            # host mirror of step_counter: the reference compares the device buffer (a sync per step)

    def forward_train(self, lq: torch.Tensor, gt: torch.Tensor) -> dict:
        """basic_restorer.py:76-95: the target is the centre frame of the GT clip; `results` are host copies."""
        output = self.generator(lq)
        gt = gt[:, gt.shape[1] // 2]
        losses = dict(loss_pix=self.pixel_loss(output, gt.contiguous()))
        return dict(losses=losses, num_samples=len(gt.data), results=dict(lq=lq.cpu(), gt=gt.cpu(), output=output.detach().cpu()))

    def forward(self, lq, gt=None, test_mode=False, **kwargs):
        if test_mode:
            with torch.no_grad():
                return dict(lq=lq.cpu(), output=self.generator(lq).cpu())
        return self.forward_train(lq, gt)

    @staticmethod
    def parse_losses(losses: Dict[str, torch.Tensor]):
        """mmedit/models/base.py:78-105 (the .item() per logged value is the reference's own host synchronisation)."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        return loss, OrderedDict((k, v.item()) for k, v in log_vars.items())

    def train_step(self, data_batch: dict, optimizer: Dict[str, torch.optim.Optimizer]) -> dict:
        """basicvsr.py:85-117."""
        if self._steps < self.fix_iter:
            if not self.is_weight_fixed:
                self.is_weight_fixed = True
                for k, v in self.generator.named_parameters():
                    if "spynet" in k or "edvr" in k:
                        v.requires_grad_(False)
        elif self._steps == self.fix_iter:
            self.generator.requires_grad_(True)
        outputs = self(**data_batch, test_mode=False)
        loss, log_vars = self.parse_losses(outputs.pop("losses"))
        optimizer["generator"].zero_grad()
        loss.backward()
        optimizer["generator"].step()
        self._steps += 1
        self.step_counter += 1
        outputs.update({"log_vars": log_vars})
        return outputs

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._steps = int(self.step_counter.item())
        return out


class CosineRestartLR(torch.optim.lr_scheduler.LRScheduler):
    """mmcv CosineRestartLrUpdaterHook, by_epoch=False (fcvsr_redsLD_QP22.py:118-127): inside period i (weight w_i, start s_i,
    length T_i) lr = min_lr + 0.5 * w_i * (base_lr - min_lr) * (1 + cos(pi * (t - s_i) / T_i)); call step() once per iteration."""

    def __init__(self, optimizer, periods: Sequence[int], restart_weights: Sequence[float] = (1,), min_lr: Optional[float] = None,
                 min_lr_ratio: Optional[float] = None, last_epoch: int = -1):
        assert (min_lr is None) ^ (min_lr_ratio is None)
        assert len(periods) == len(restart_weights), "periods and restart_weights should have the same length."
        self.periods, self.restart_weights = list(periods), list(restart_weights)
        self.min_lr, self.min_lr_ratio = min_lr, min_lr_ratio
        self.cumulative = [sum(self.periods[:i + 1]) for i in range(len(self.periods))]
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        t = self.last_epoch
        idx = next((i for i, c in enumerate(self.cumulative) if t < c), None)
        if idx is None:
            raise ValueError(f"Current iteration {t} exceeds cumulative_periods {self.cumulative}")
        start = 0 if idx == 0 else self.cumulative[idx - 1]
        alpha = min((t - start) / self.periods[idx], 1.0)
        out = []
        for base in self.base_lrs:
            target = base * self.min_lr_ratio if self.min_lr_ratio is not None else self.min_lr
            out.append(target + 0.5 * self.restart_weights[idx] * (base - target) * (math.cos(math.pi * alpha) + 1.0))
        return out


def save_checkpoint(model: nn.Module, path: str, optimizer: Optional[Dict[str, torch.optim.Optimizer]] = None,
                    meta: Optional[dict] = None) -> None:
    """mmcv.runner.save_checkpoint layout (checkpoint_config save_optimizer=True, fcvsr_redsLD_QP22.py:130)."""
    ckpt = {"meta": dict(meta or {}), "state_dict": OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items())}
    if optimizer is not None:
        ckpt["optimizer"] = {k: o.state_dict() for k, o in optimizer.items()}
    torch.save(ckpt, path)


def load_checkpoint(model: nn.Module, path: str, optimizer: Optional[Dict[str, torch.optim.Optimizer]] = None,
                    strict: bool = True) -> dict:
    """resume_from / load_from (fcvsr_redsLD_QP22.py:141-142): restores the model, and the optimizers when given."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(ckpt["state_dict"], strict=strict)
    if optimizer is not None and "optimizer" in ckpt:
        for k, o in optimizer.items():
            o.load_state_dict(ckpt["optimizer"][k])
    return ckpt.get("meta", {})
