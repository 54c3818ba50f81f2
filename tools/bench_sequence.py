#!/usr/bin/env python
"""BASELINE config 3: FCVSR inference over a synthetic N-frame REDS4-shaped LR sequence (180x320 -> 720x1280), sharded by
output-frame range with 3-frame LR halos over the ranks (strong scaling: the sequence is fixed).

    python tools/bench_sequence.py [--frames 100] [--batch 4] [--dtype bf16|tf32] [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sequence.py ...

Every rank uploads only its halo range from pinned host memory, runs its windows `batch` at a time through the drop-in
forward (CUDA graph) and copies its HR frames back to the host; the time is device time (CUDA events), max over ranks,
and includes the H2D / D2H copies.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.sequence import halo_range, shard_range, super_resolve_sequence  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--height", type=int, default=180)
ap.add_argument("--width", type=int, default=320)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
m = arch.GShiftNet().to(dev).eval()
m.load_state_dict(arch.seeded_state_dict("full", 0))
m.compute_dtype = a.dtype
g = torch.Generator().manual_seed(1234)
seq = (torch.round(255 * torch.rand(a.frames, 1, a.height, a.width, generator=g)) / 255).pin_memory()
lo, hi = shard_range(a.frames, rank, world)
host_out = torch.empty(hi - lo, 1, 4 * a.height, 4 * a.width).pin_memory()


def run():
    y, _ = super_resolve_sequence(m, seq, batch=a.batch, rank=rank, world=world)
    host_out.copy_(y, non_blocking=True)


with torch.no_grad():
    m(seq[:7].unsqueeze(0).to(dev).expand(a.batch, -1, -1, -1, -1).contiguous())
    m._engine.use_graph = True
    run()
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
if rank == 0:
    ms = min(times)
    h_lo, h_hi = halo_range(lo, hi, a.frames)
    print(json.dumps({"metric": "x4 SR output frames/sec over a fixed sequence (180x320->720x1280)", "value": a.frames / ms * 1e3,
                      "unit": "frames/s", "n_gpus": world, "scaling": "strong", "ms_per_sequence": ms, "frames": a.frames,
                      "dtype": a.dtype, "windows_per_launch": a.batch,
                      "config": {"workload": f"FCVSR over a synthetic {a.frames}-frame {a.height}x{a.width} sequence, "
                                             "output-frame ranges sharded over ranks with 3-frame LR halos, H2D/D2H inside the timed region",
                                 "rank0_lr_frames_uploaded": h_hi - h_lo}}))
if world > 1:
    dist.destroy_process_group()
