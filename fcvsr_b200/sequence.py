"""Sliding-window sequence inference around the per-clip forward, and its sharding over GPUs.

Restates the conventions of the reference drivers (not their file I/O):
  * window of output frame t = 7 LR frames around t with one of the reference's edge paddings
      'replicate'          clip(t-3 .. t+3, 0, N-1)      CVSR_train/test_LD_freqCVSR_S_FPS.py:14-17,64
      'reflection'         mirror about the edge frame   mmedit .../pipelines/augmentation.py:856-877
      'reflection_circle'  [6,5,4,0,1,2,3] at t = 0, [6,5,0,1,2,3,4] at t = 1: same lines; the mode of the FCVSR REDS
                           test pipeline (configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:31)
      'circle'             [4,5,6,0,1,2,3] at t = 0       same lines
      'pad_sequence'       the sequence padded ONCE with frames [6,5,4] in front and [n-5,n-6,n-7] behind, then a plain
                           sliding window ([5,4,0,1,2,3,4] at t = 1): mmedit/apis/restoration_video_inference.py:16-25
  * frames whose height/width is not a multiple of 4 are zero-padded at the bottom/right and the
    output is cropped back (test_LD_freqCVSR.py:25-27,85-88: 270 -> 272 rows, 1088 -> 1080);
  * the evaluation driver stores clamp(sr, 0, 1) * 255 truncated to uint8 (test_LD_freqCVSR.py:91-93): `to_uint8`;
  * every output frame is an independent 7-frame window, so a sequence shards over ranks by output
    frame range with an LR halo on each side and no communication.  The halo is 3 frames for
    'replicate' / 'reflection' and up to 6 frames at the sequence ends for the circle / pad_sequence modes.

Long sequences stream: LR frames are uploaded chunk by chunk from (pinned) host memory on a copy stream while the
previous chunk computes, and HR frames are downloaded into a pinned host tensor the same way, so the device only ever
holds two chunks (`stream_chunk`).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

RADIUS = 3
MODES = ("replicate", "reflection", "reflection_circle", "circle", "pad_sequence")


def window_indices(t: int, n: int, mode: str = "replicate") -> List[int]:
    """LR frame indices of the 7-frame window centred on output frame t of an n-frame sequence
    (GenerateFrameIndiceswithPadding.__call__, augmentation.py:856-877, with num_input_frames = 7)."""
    if mode not in MODES:
        raise ValueError(f"unknown padding mode {mode!r}")
    last, width = n - 1, 2 * RADIUS + 1
    idx = []
    for i in range(t - RADIUS, t + RADIUS + 1):
        if i < 0:
            j = {"replicate": 0, "reflection": -i, "reflection_circle": t + RADIUS - i, "circle": width + i,
                 "pad_sequence": RADIUS - i}[mode]
        elif i > last:
            j = {"replicate": last, "reflection": 2 * last - i, "reflection_circle": (t - RADIUS) - (i - last),
                 "circle": i - width, "pad_sequence": 2 * last - RADIUS - i}[mode]
        else:
            j = i
        idx.append(min(max(j, 0), last))       # sequences shorter than the padding reach: clamp (the reference would index out of range)
    return idx


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous output-frame range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def halo_range(lo: int, hi: int, n: int, mode: str = "replicate") -> Tuple[int, int]:
    """LR frames [h_lo, h_hi) needed for output frames [lo, hi): the hull of their windows under `mode`."""
    if hi <= lo:
        return lo, lo
    h_lo, h_hi = n, 0
    for t in range(lo, hi):
        w = window_indices(t, n, mode)
        h_lo, h_hi = min(h_lo, min(w)), max(h_hi, max(w) + 1)
    return h_lo, h_hi


def pad_to_multiple(frames: torch.Tensor, m: int = 4) -> Tuple[torch.Tensor, int, int]:
    """Zero-pad [..., H, W] at the bottom/right to multiples of m; returns (padded, H, W)."""
    h, w = frames.shape[-2:]
    ph, pw = (-h) % m, (-w) % m
    if ph or pw:
        frames = torch.nn.functional.pad(frames, (0, pw, 0, ph))
    return frames, h, w


def quantize_u8(y: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[B,1,Hp,Wp] fp32 on the GPU -> [B,1,h,w] uint8 = trunc(clamp(y[..., :h, :w], 0, 1) * 255) (test_LD_freqCVSR.py:85-93)."""
    from . import _capi as C
    if not y.is_cuda:
        raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
    y = y.contiguous()
    b, hp, wp = y.shape[0] * y.shape[1], y.shape[-2], y.shape[-1]
    out = torch.empty(y.shape[0], y.shape[1], h, w, device=y.device, dtype=torch.uint8)
    with torch.cuda.device(y.device):
        C.call("fcvsr_quantize_u8", y.data_ptr(), out.data_ptr(), b, hp, wp, h, w, torch.cuda.current_stream().cuda_stream)
    return out


def _run_range(model, local, h_lo, lo, hi, n, mode, batch, scale, h, w, to_uint8):
    dev = local.device
    outs = []
    # the window indices of the whole range travel to the device ONCE: a per-launch torch.tensor(..., device=...) is a blocking
    # pageable copy that makes the host wait for the previous forward before it can queue the next one
    idx_all = torch.tensor([[j - h_lo for j in window_indices(t, n, mode)] for t in range(lo, hi)], dtype=torch.long).to(dev)
    for t0 in range(lo, hi, batch):
        ts = range(t0, min(t0 + batch, hi))
        idx = idx_all[t0 - lo:t0 - lo + len(ts)]
        clips = local[idx.reshape(-1)].view(len(ts), 2 * RADIUS + 1, *local.shape[1:])
        y = model(clips)
        outs.append(quantize_u8(y, scale * h, scale * w) if to_uint8 else y[..., : scale * h, : scale * w])
    return outs


@torch.no_grad()
def super_resolve_sequence(model, frames: torch.Tensor, batch: int = 4, mode: str = "replicate", rank: int = 0,
                           world: int = 1, scale: int = 4, to_uint8: bool = False, stream_chunk: Optional[int] = None,
                           out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """frames [N,1,H,W] (host or device) -> HR frames [hi-lo,1,4H,4W] of this rank's output range.

    Without `stream_chunk` the rank's whole LR halo range is moved to the model's device at once and the result stays
    there.  With `stream_chunk` = K (frames on the host, ideally pinned) the range is processed K output frames at a time:
    the next chunk's LR frames are uploaded on a copy stream while the current chunk computes, and every chunk's HR frames
    are copied asynchronously into `out` (a host tensor [hi-lo,1,4H,4W], allocated pinned if not given), which is returned."""
    n = frames.shape[0]
    lo, hi = shard_range(n, rank, world)
    dev = next(model.parameters()).device
    hh, ww = frames.shape[-2:]
    odt = torch.uint8 if to_uint8 else torch.float32
    if stream_chunk is None or frames.is_cuda or hi <= lo:
        h_lo, h_hi = halo_range(lo, hi, n, mode)
        local, h, w = pad_to_multiple(frames[h_lo:h_hi].to(dev, non_blocking=True))
        outs = _run_range(model, local, h_lo, lo, hi, n, mode, batch, scale, h, w, to_uint8)
        res = torch.cat(outs, 0) if outs else torch.zeros(0, 1, scale * hh, scale * ww, dtype=odt, device=dev)
        return res, (lo, hi)
    # ---- streaming path ------------------------------------------------------------------------------------------------
    if out is None:
        out = torch.empty(hi - lo, 1, scale * hh, scale * ww, dtype=odt).pin_memory()
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    chunks = [(c, min(c + stream_chunk, hi)) for c in range(lo, hi, stream_chunk)]

    def upload(c):
        a, b = chunks[c]
        h_lo, h_hi = halo_range(a, b, n, mode)
        with torch.cuda.stream(copy):
            t = frames[h_lo:h_hi].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return t, h_lo, ev

    nxt = upload(0)
    pending = []                                   # (device result, event) kept alive until their download was issued
    for c, (a, b) in enumerate(chunks):
        lr, h_lo, ev = nxt
        if c + 1 < len(chunks):
            nxt = upload(c + 1)
        main.wait_event(ev)
        lr.record_stream(main)
        local, h, w = pad_to_multiple(lr)
        y = torch.cat(_run_range(model, local, h_lo, a, b, n, mode, batch, scale, h, w, to_uint8), 0)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(copy):
            copy.wait_event(done)
            out[a - lo:b - lo].copy_(y, non_blocking=True)
            y.record_stream(copy)
        pending.append(y)
    copy.synchronize()
    return out, (lo, hi)
