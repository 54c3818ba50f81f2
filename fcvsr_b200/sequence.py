"""Sliding-window sequence inference around the per-clip forward, and its sharding over GPUs.

Restates the conventions of the reference drivers (not their I/O):
  * window of output frame t = LR frames clip(t-3 .. t+3, 0, N-1)  -- replicate edges
    (CVSR_train/test_LD_freqCVSR_S_FPS.py:14-17,64); 'reflection' mirrors instead
    (mmedit_train/mmedit/apis/restoration_video_inference.py:16-25);
  * frames whose height/width is not a multiple of 4 are zero-padded at the bottom/right and the
    output is cropped back (test_LD_freqCVSR.py:25-27,85-88: 270 -> 272 rows, 1088 -> 1080);
  * every output frame is an independent 7-frame window, so a sequence shards over ranks by output
    frame range with a 3-frame LR halo on each side and no communication.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

RADIUS = 3


def window_indices(t: int, n: int, mode: str = "replicate") -> List[int]:
    """LR frame indices of the 7-frame window centred on output frame t of an n-frame sequence."""
    idx = []
    for i in range(t - RADIUS, t + RADIUS + 1):
        if mode == "replicate":
            idx.append(min(max(i, 0), n - 1))
        elif mode == "reflection":
            j = -i if i < 0 else (2 * (n - 1) - i if i > n - 1 else i)
            idx.append(min(max(j, 0), n - 1))
        else:
            raise ValueError(f"unknown padding mode {mode!r}")
    return idx


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous output-frame range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def halo_range(lo: int, hi: int, n: int) -> Tuple[int, int]:
    """LR frames a rank needs for output frames [lo, hi): the range widened by the 3-frame halo."""
    if hi <= lo:
        return lo, lo
    return max(lo - RADIUS, 0), min(hi + RADIUS, n)


def pad_to_multiple(frames: torch.Tensor, m: int = 4) -> Tuple[torch.Tensor, int, int]:
    """Zero-pad [..., H, W] at the bottom/right to multiples of m; returns (padded, H, W)."""
    h, w = frames.shape[-2:]
    ph, pw = (-h) % m, (-w) % m
    if ph or pw:
        frames = torch.nn.functional.pad(frames, (0, pw, 0, ph))
    return frames, h, w


@torch.no_grad()
def super_resolve_sequence(model, frames: torch.Tensor, batch: int = 4, mode: str = "replicate", rank: int = 0,
                           world: int = 1, scale: int = 4) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """frames [N,1,H,W] (host or device) -> HR frames [hi-lo,1,4H,4W] of this rank's output range.

    Only the LR halo range of the rank is moved to the model's device; windows are gathered on the
    device and run `batch` at a time through `model` (the drop-in forward)."""
    n = frames.shape[0]
    lo, hi = shard_range(n, rank, world)
    dev = next(model.parameters()).device
    h_lo, h_hi = halo_range(lo, hi, n)
    local, h, w = pad_to_multiple(frames[h_lo:h_hi].to(dev, non_blocking=True))
    outs = []
    for t0 in range(lo, hi, batch):
        ts = range(t0, min(t0 + batch, hi))
        idx = torch.tensor([[j - h_lo for j in window_indices(t, n, mode)] for t in ts], device=dev)
        clips = local[idx.reshape(-1)].view(len(ts), 2 * RADIUS + 1, *local.shape[1:])
        y = model(clips)
        outs.append(y[..., : scale * h, : scale * w])
    out = torch.cat(outs, 0) if outs else frames.new_zeros(0, 1, scale * h, scale * w)
    return out, (lo, hi)
