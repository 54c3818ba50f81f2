"""Gradient all-reduce of the data-parallel training step (BASELINE config 4).

The reference trains with one replica per GPU and `find_unused_parameters=True`
(mmedit_train/configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:144): every step the 8.81 M fp32 gradients (35 MB) are summed
over the ranks and divided by the world size; 16 parameters (`MFFRblock.DivEnh_block.*.Conv.*`) never receive a gradient.
This is the only collective of the whole system -- inference shards over windows with no communication.

`GradAllReducer` packs the gradients into a few flat fp32 buckets (reverse parameter order, i.e. the order in which the
backward produces them), launches one asynchronous `all_reduce` per bucket on the process group (NCCL over NVLink on the GPU
box, gloo in the CPU tests) and writes the averaged values back into `.grad` in `finish()`.  Buckets are sized for launch
latency and overlap, not link count: NVSwitch gives every GPU full bandwidth to every peer, so a 35 MB gradient set is four
~9 MB collectives in flight behind the rest of the backward.

Ordering contract: collectives are issued in STRICT bucket order on every rank -- bucket k goes out from a gradient hook only
once buckets 0..k-1 have gone out, everything else goes out from `finish()` in index order -- so ranks can never disagree on
the order even if their hooks fire differently.  Unused parameters: a parameter without a local gradient contributes zeros
and keeps `grad is None`; the set of gradient-less parameters must be the same on every rank (it is static for FCVSR: the
dead `DivEnh.Conv` layers).  The first step verifies that with one flag all-reduce (the only host synchronisation of this
module; later steps have none).  Replicas start identical: rank 0's parameters are broadcast at construction.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 9.0, group=None, average: bool = True,
                 overlap: bool = True, broadcast: bool = True):
        seen, plist = set(), []
        for p in params:                           # aliased parameters (recorb1...RCB == body.3) appear once
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                plist.append(p)
        if broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
            with torch.no_grad():
                for p in plist:                    # forward order, identical on every rank
                    dist.broadcast(p.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        plist.reverse()                            # the backward reaches the last layers first
        self.group, self.average, self.overlap = group, average, overlap
        self.buckets: List[List[torch.nn.Parameter]] = []
        cap = int(bucket_mb * (1 << 20)) // 4
        cur, n = [], 0
        for p in plist:
            if cur and (n + p.numel() > cap or p.device != cur[0].device):
                self.buckets.append(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            self.buckets.append(cur)
        nb = len(self.buckets)
        self._sizes = [[p.numel() for p in b] for b in self.buckets]
        self._flat: List[Optional[torch.Tensor]] = [None] * nb
        self._work: List[Optional[object]] = [None] * nb
        self._have: List[Optional[List[bool]]] = [None] * nb
        self._pending = [0] * nb
        self._ready = [False] * nb
        self._next = 0                             # first bucket whose collective has not been issued yet
        self._steps = 0
        self._flags = None
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._hooks = []
        self.allreduce_events = []                 # (start, end) CUDA events per bucket of the last step, when enabled
        self.time_collectives = False
        if overlap:
            for b in self.buckets:
                for p in b:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.start_step()

    # ---- per-step protocol: start_step() -> backward() -> finish() -------------------------------------------------------
    def start_step(self) -> None:
        nb = len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        self._ready = [False] * nb
        self._work = [None] * nb
        self._have = [None] * nb
        self._next = 0
        self.allreduce_events = []

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._ready[bi] = True
            self._drain(False)

    def _drain(self, force: bool) -> None:
        """Issue collectives in strict index order: as far as the ready prefix reaches, or all of them (`force`)."""
        nb = len(self.buckets)
        while self._next < nb and (force or self._ready[self._next]):
            self._launch(self._next)
            self._next += 1

    def _launch(self, bi: int) -> None:
        b, sizes = self.buckets[bi], self._sizes[bi]
        total = sum(sizes)
        flat = self._flat[bi]
        if flat is None or flat.device != b[0].device:
            flat = torch.zeros(total, dtype=torch.float32, device=b[0].device)
            self._flat[bi] = flat
        chunks = flat.split(sizes)
        have = [p.grad is not None for p in b]
        self._have[bi] = have
        dst = [c for c, h in zip(chunks, have) if h]
        if dst:                                                   # one multi-tensor copy instead of a launch per parameter
            torch._foreach_copy_(dst, [p.grad.reshape(-1) for p, h in zip(b, have) if h])
        missing = [c for c, h in zip(chunks, have) if not h]
        if missing:
            torch._foreach_zero_(missing)
        ev = None
        if self.time_collectives and flat.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if ev is not None:
            self._work[bi].wait()              # timing mode: the stream waits for the collective so that the end event brackets it
            ev[1].record()
            self.allreduce_events.append(ev)

    def _check_unused_consistent(self) -> None:
        """First step only: every rank must have the same set of gradient-less parameters."""
        dev = self.buckets[0][0].device
        mine = torch.tensor([1.0 if h else 0.0 for hv in self._have for h in hv], dtype=torch.float32, device=dev)
        tot = mine.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        world = dist.get_world_size(self.group)
        bad = ((tot != 0) & (tot != world)).nonzero().flatten().tolist()
        if bad:
            raise RuntimeError(f"GradAllReducer: {len(bad)} parameters received a gradient on some ranks only (flat indices "
                               f"{bad[:8]}...); the set of unused parameters must be identical on every rank")

    def finish(self) -> None:
        """Issue what the hooks did not (parameters that got no gradient never fire one), wait, average, write back."""
        world = dist.get_world_size(self.group)
        self._drain(True)
        if self._steps == 0 and world > 1:
            self._check_unused_consistent()
        for bi, b in enumerate(self.buckets):
            self._work[bi].wait()
            flat = self._flat[bi]
            if self.average:
                flat.div_(world)
            have = self._have[bi]
            chunks = flat.split(self._sizes[bi])
            src = [c for c, h in zip(chunks, have) if h]
            if src:
                torch._foreach_copy_([p.grad.reshape(-1) if p.grad.is_contiguous() else p.grad for p, h in zip(b, have) if h],
                                     [c if p.grad.is_contiguous() else c.view_as(p) for c, p, h in zip(chunks, b, have) if h])
        self._steps += 1
        self.start_step_keep_events()

    def start_step_keep_events(self) -> None:
        ev = self.allreduce_events
        self.start_step()
        self.allreduce_events = ev

    def allreduce_ms(self) -> float:
        """Sum of the collectives' device time in the last step (time_collectives mode; synchronises)."""
        if not self.allreduce_events:
            return 0.0
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.allreduce_events)

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
