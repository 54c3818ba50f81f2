"""One data-parallel training step with the reference's semantics (CVSR_train/train_LD_freqCVSR_22.py:243-251):

    optimizer.zero_grad(); sr = model(frames); loss = CharbonnierLoss(sr, hr); loss.backward(); optimizer.step()

plus the gradient all-reduce that the mmedit configuration wraps around it (one replica per GPU,
fcvsr_redsLD_QP22.py:144).  `model` is any autograd-capable module; with `fcvsr_b200.arch.GShiftNet[_S]` the forward and
backward run on this repository's kernels (fcvsr_b200.train_forward / fcvsr_b200.autograd), the loss is
`fcvsr_b200.ops.loss.CharbonnierLoss`, the optimizer `fcvsr_b200.ops.optim.Adam` and the reducer
`fcvsr_b200.gradsync.GradAllReducer` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Charbonnier is sum-reduced (opt/loss.py:30), so averaging the gradients over G ranks makes one step equal to a single-GPU
step on the concatenated batch with the learning rate divided by G -- the reference behaves the same way; it is documented,
not "fixed" (SURVEY 8e).
"""
from __future__ import annotations

from typing import Callable

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


class GraphedTrainStep:
    """The training step with forward + backward replayed from ONE CUDA graph.

    The eager step is host-bound (about 18 k kernel launches per step for FCVSR at batch 8: 250 ms of Python / launch time for
    100 ms of device work), so the launch sequence of `loss_fn(model(frames), hr).backward()` is captured once on static
    input buffers and replayed; the gradient all-reduce (its buckets are issued back to back, NCCL over NVLink) and the
    optimizer step stay outside the graph -- Adam's bias corrections depend on the step count, which a captured launch would
    freeze.  Gradients live in the graph's memory pool and are rewritten by every replay (the whole-network capture pattern of
    the PyTorch CUDA-graphs notes), so `zero_grad` must not be called between steps.
    """

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn, frames: torch.Tensor, hr: torch.Tensor,
                 reducer=None, warmup: int = 3):
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.frames, self.hr = frames.clone(), hr.clone()
        dev = frames.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                   # eager warm-up: lazy initialisation (caches, function attributes), allocator
            for _ in range(warmup):
                optimizer.zero_grad(set_to_none=True)
                loss_fn(model(self.frames), self.hr).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn(model(self.frames), self.hr)
            self.loss.backward()

    def __call__(self, frames: torch.Tensor, hr: torch.Tensor) -> torch.Tensor:
        self.frames.copy_(frames, non_blocking=True)
        self.hr.copy_(hr, non_blocking=True)
        if self.reducer is not None:
            self.reducer.start_step()
        self.graph.replay()
        if self.reducer is not None:
            self.reducer.finish()            # hooks do not fire on a replay: every bucket goes out here, in order
        self.optimizer.step()
        return self.loss.detach()


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


def bench_train_step(dev, rank: int, world: int, steps: int = 5, warmup: int = 2, variant: str = "full", batch: int = 8,
                     size: int = 64, graph: bool = True) -> dict:
    """BASELINE config 4 for bench.py: FCVSR training step (forward + backward + Adam, Charbonnier-sum loss) on a per-GPU batch
    of `batch` synthetic 7 x size x size crops, data parallel with the NCCL gradient all-reduce when world > 1.  Device time
    (CUDA events), max over ranks; the all-reduce is timed separately in a second pass (`time_collectives`: every collective
    bracketed by events, which serialises it with the backward) so that the overlapped step time is not perturbed."""
    import torch.distributed as dist
    from . import arch
    from .gradsync import GradAllReducer
    from .ops.loss import CharbonnierLoss
    from .ops.optim import Adam
    cls = arch.GShiftNet if variant == "full" else arch.GShiftNet_S
    model = cls().to(dev).train()
    model.load_state_dict(arch.seeded_state_dict(variant, 0))
    model.compute_dtype = "tf32"
    opt = Adam(model.parameters(), lr=5e-6, weight_decay=1e-5)              # train_LD_freqCVSR_22.py:35,42,204
    red = GradAllReducer(model.parameters(), overlap=not graph) if world > 1 else None
    g = torch.Generator().manual_seed(99 + rank)
    frames = (torch.round(255 * torch.rand(batch, 7, 1, size, size, generator=g)) / 255).to(dev)
    hr = torch.rand(batch, 1, 4 * size, 4 * size, generator=g).to(dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stepper = GraphedTrainStep(model, opt, CharbonnierLoss, frames, hr, red) if graph else None

    def one():
        return stepper(frames, hr) if graph else train_step(model, opt, frames, hr, CharbonnierLoss, red)

    def run(n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = None
        for _ in range(n):
            loss = one()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync()
        return float(ms) / n, float(loss)

    run(warmup)
    ms, loss = run(steps)
    ar_ms = None
    if red is not None:
        red.time_collectives = True
        one()
        t = torch.tensor([red.allreduce_ms()], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t)
        red.time_collectives = False
        red.remove_hooks()
    nparam = sum(p.numel() for p in model.parameters())
    return {"metric": "training steps/sec (fwd + bwd + Adam, per-GPU batch %d of 7x%dx%d crops)" % (batch, size, size),
            "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "clips_per_s": world * batch * 1e3 / ms, "n_gpus": world,
            "steps": steps, "warmup": warmup, "loss": loss, "dtype": "tf32 (fp32 storage; TF32 tensor-core operands in forward and "
            "data-gradient convolutions; bf16 tcgen05 weight gradients with fp32 accumulation)", "variant": variant,
            "cuda_graph": "forward + backward replayed from one CUDA graph; all-reduce and Adam outside" if graph else False,
            "allreduce": None if red is None else {"backend": "nccl", "buckets": len(red.buckets), "bytes": 4 * nparam,
                                                   "ms_serialised": ar_ms,
                                                   "note": "sum of the bucket collectives' device time in a pass where each is "
                                                           "bracketed by events; in the timed steps they are issued back to back "
                                                           "after the backward graph"},
            "optimizer": "fcvsr_adam_step (multi-tensor), lr 5e-6, weight_decay 1e-5", "loss_fn": "Charbonnier sum (fcvsr_charbonnier_loss)"}
