"""Gradient all-reduce of the data-parallel training step (BASELINE config 4).

The reference trains with one replica per GPU and `find_unused_parameters=True`
(mmedit_train/configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:144): every step the 8.81 M fp32 gradients (35 MB) are summed
over the ranks and divided by the world size; 16 parameters (`MFFRblock.DivEnh_block.*.Conv.*`) never receive a gradient.
This is the only collective of the whole system -- inference shards over windows with no communication.

`GradAllReducer` packs the gradients into a few flat fp32 buckets (reverse parameter order, i.e. the order in which the
backward produces them), launches one asynchronous `all_reduce` per bucket on the process group (NCCL over NVLink on the GPU
box, gloo in the CPU tests) as soon as the bucket's last gradient has been accumulated, and writes the averaged values back
into `.grad` in `finish()`.  Buckets are sized for launch latency and overlap, not link count: NVSwitch gives every GPU full
bandwidth to every peer, so a 35 MB gradient set is four ~9 MB collectives in flight behind the rest of the backward.
Parameters without a gradient contribute zeros and keep `grad is None` only if no rank produced one.

The model's own backward kernels are not built yet; this module is exercised by the 2-rank gloo test in
tests/test_gradsync.py and is what the training step will call.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 9.0, group=None, average: bool = True,
                 overlap: bool = True):
        seen, plist = set(), []
        for p in params:                           # aliased parameters (recorb1...RCB == body.3) appear once
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                plist.append(p)
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
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._work: List[Optional[object]] = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._bucket_of = {id(p): (bi, pi) for bi, b in enumerate(self.buckets) for pi, p in enumerate(b)}
        self._hooks = []
        if overlap:
            for b in self.buckets:
                for p in b:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.start_step()

    # ---- per-step protocol: start_step() -> backward() -> finish() -------------------------------------------------------
    def start_step(self) -> None:
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi, _ = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self._work[bi] is None:     # one backward per step: later accumulations are not re-sent
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        b = self.buckets[bi]
        sizes = [p.numel() for p in b]
        total = sum(sizes)
        flat = self._flat[bi]
        if flat is None or flat.device != b[0].device:
            # + one flag per parameter: "some rank produced a gradient" (so unused parameters keep grad None everywhere)
            flat = torch.zeros(total + len(b), dtype=torch.float32, device=b[0].device)
            self._flat[bi] = flat
        chunks = flat[:total].split(sizes)
        have = [p.grad is not None for p in b]
        dst = [c for c, h in zip(chunks, have) if h]
        if dst:                                                   # one multi-tensor copy instead of a launch per parameter
            torch._foreach_copy_(dst, [p.grad.reshape(-1) for p, h in zip(b, have) if h])
        for c, h in zip(chunks, have):
            if not h:
                c.zero_()
        flat[total:].copy_(torch.tensor([1.0 if h else 0.0 for h in have], dtype=torch.float32), non_blocking=True)
        self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Launch what the hooks did not (parameters that got no gradient never fire one), wait, average, write back."""
        world = dist.get_world_size(self.group)
        for bi in range(len(self.buckets)):
            if self._work[bi] is None:
                self._launch(bi)
        for bi, b in enumerate(self.buckets):
            self._work[bi].wait()
            flat = self._flat[bi]
            total = flat.numel() - len(b)
            if self.average:
                flat[:total].div_(world)
            flags = flat[total:].tolist()
            o = 0
            for i, p in enumerate(b):
                n = p.numel()
                if flags[i] > 0:
                    g = flat[o:o + n].view_as(p)
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.copy_(g)
                o += n
        self.start_step()

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
