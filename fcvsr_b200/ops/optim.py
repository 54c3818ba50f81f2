"""Optimizer of the reference's training loop on the GPU: ``Adam(params, lr, betas, eps, weight_decay)`` with the semantics of
``torch.optim.Adam`` as the reference uses it (CVSR_train/train_LD_freqCVSR_22.py:204: lr = 5e-6, weight_decay = 1e-5, L2
decay folded into the gradient, no amsgrad).  It subclasses ``torch.optim.Optimizer``, so ``param_groups`` / ``state_dict`` and
the reference's ``MultiStepLR`` scheduler (:205) work unchanged; ``step()`` is one multi-tensor kernel launch per 64 tensors
(``fcvsr_adam_step``, csrc/optim.cu) instead of four elementwise kernels per parameter."""
from __future__ import annotations

import ctypes

import torch

from .. import _capi as C


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam state dict: torch stores `step` as a tensor -- normalised to int here (a tensor key would
        hash by identity and force one launch and one host sync per parameter); amsgrad / maximize groups are refused."""
        for g in state_dict.get("param_groups", []):
            if g.get("amsgrad") or g.get("maximize"):
                raise ValueError("fcvsr_b200.ops.optim.Adam does not implement amsgrad / maximize (the reference uses neither)")
        super().load_state_dict(state_dict)
        for st in self.state.values():
            if "step" in st:
                st["step"] = int(st["step"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("fcvsr_b200.ops.optim.Adam updates fp32 CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                if not p.is_contiguous():
                    raise RuntimeError("Adam expects contiguous parameters")
                by_step.setdefault((st["step"], p.device), []).append((p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"]))
            for (step, dev), items in by_step.items():
                n = len(items)
                arr = lambda k: (ctypes.c_void_p * n)(*[it[k].data_ptr() for it in items])  # noqa: E731
                numels = (ctypes.c_longlong * n)(*[it[0].numel() for it in items])
                with torch.cuda.device(dev):
                    C.call("fcvsr_adam_step", arr(0), arr(1), arr(2), arr(3), numels, n, float(group["lr"]),
                           float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                           float(group["weight_decay"]), int(step), torch.cuda.current_stream().cuda_stream)
        return loss
