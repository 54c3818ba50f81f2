"""Generate tests/golden/*_grads.pt: gradients of the UNMODIFIED reference (build container only).

    python oracle/make_golden_grads.py

Pins the backward of the training step (BASELINE config 4) before any backward kernel exists: the seeded
weights are loaded strictly into the reference ``GShiftNet`` / ``GShiftNet_S`` (CVSR_train/arch/CVSR_freq.py:2653,2577), one
forward + backward of the Charbonnier-sum loss (opt/loss.py:20-31) runs on a seeded clip / target, and for every parameter the
file keeps: whether it received a gradient at all (the ``DivEnh.Conv`` parameters do not), the gradient's
L2 norm and sum, and 16 strided samples.  ``tests/test_oracle.py`` checks that autograd through the oracle restatement
reproduces them; the CUDA backward kernels will be checked against the same file.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fcvsr_b200.arch import seeded_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import GOLD, make_clip  # noqa: E402

CASES = [
    dict(name="fcvsr_s_32_grads", variant="S", seed=0, clip_seed=5, target_seed=7, b=2, h=32, w=32),
    dict(name="fcvsr_full_32_grads", variant="full", seed=0, clip_seed=6, target_seed=8, b=1, h=32, w=32),
]


def target(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    return torch.rand(b, 1, 4 * h, 4 * w, generator=torch.Generator().manual_seed(seed))


def charbonnier_sum(sr: torch.Tensor, hr: torch.Tensor) -> torch.Tensor:
    d = sr - hr                                    # opt/loss.py:24-29 with mean_res = False
    return torch.sum(torch.sqrt(d * d + 1e-4))


def strided(t: torch.Tensor, n: int = 16) -> torch.Tensor:
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].clone()


def main() -> None:
    ref = ref_loader.load()
    for case in CASES:
        sd = seeded_state_dict(case["variant"], case["seed"])
        model = (ref.GShiftNet_S if case["variant"] == "S" else ref.GShiftNet)()
        model.load_state_dict(sd, strict=True)
        x = make_clip(case["clip_seed"], case["b"], case["h"], case["w"]).requires_grad_()
        hr = target(case["target_seed"], case["b"], case["h"], case["w"])
        loss = charbonnier_sum(model(x), hr)
        loss.backward()
        grads, none = {}, []
        for k, p in model.named_parameters():      # de-duplicated: the aliased RCB == body.3 modules appear once
            if p.grad is None:
                none.append(k)
            else:
                grads[k] = dict(norm=float(p.grad.norm()), sum=float(p.grad.double().sum()), amax=float(p.grad.abs().max()),
                                samples=strided(p.grad))
        out = dict(case=case, loss=float(loss), grads=grads, no_grad=none, dx=strided(x.grad, 64), dx_norm=float(x.grad.norm()))
        torch.save(out, os.path.join(GOLD, case["name"] + ".pt"))
        print(case["name"], "loss", float(loss), "params with grad", len(grads), "without", len(none))


if __name__ == "__main__":
    main()
