"""CPU tests of the sequence driver logic (window indices, padding, sharding) incl. a 2-rank gloo run."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcvsr_b200 import sequence as S


def test_window_indices_replicate_matches_reference_convention():
    # test_LD_freqCVSR_S_FPS.py:14-17: np.clip(range(t-3, t+4), 0, N-1)
    assert S.window_indices(0, 100) == [0, 0, 0, 0, 1, 2, 3]
    assert S.window_indices(50, 100) == [47, 48, 49, 50, 51, 52, 53]
    assert S.window_indices(99, 100) == [96, 97, 98, 99, 99, 99, 99]
    assert S.window_indices(1, 3) == [0, 0, 0, 1, 2, 2, 2]


def test_window_indices_reflection():
    assert S.window_indices(0, 100, "reflection") == [3, 2, 1, 0, 1, 2, 3]
    assert S.window_indices(99, 100, "reflection") == [96, 97, 98, 99, 98, 97, 96]


def test_shards_cover_sequence_and_halos_suffice():
    for n in (1, 5, 100, 101):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = S.shard_range(n, r, world)
                covered += list(range(lo, hi))
                h_lo, h_hi = S.halo_range(lo, hi, n)
                for t in range(lo, hi):
                    assert all(h_lo <= j < h_hi for j in S.window_indices(t, n))
            assert covered == list(range(n))


def test_pad_to_multiple_and_crop():
    x = torch.rand(3, 1, 270, 480)
    p, h, w = S.pad_to_multiple(x)
    assert p.shape[-2:] == (272, 480) and (h, w) == (270, 480)
    assert torch.equal(p[..., :270, :], x) and float(p[..., 270:, :].abs().max()) == 0.0


class _Center(torch.nn.Module):
    """Stand-in for the SR model: x4 nearest up-sampling of the centre frame (checks the plumbing)."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return torch.nn.functional.interpolate(x[:, 3], scale_factor=4, mode="nearest")


def test_sequence_runner_single_rank():
    frames = torch.rand(9, 1, 6, 10)
    out, (lo, hi) = S.super_resolve_sequence(_Center(), frames, batch=4)
    assert (lo, hi) == (0, 9) and out.shape == (9, 1, 24, 40)
    ref = torch.nn.functional.interpolate(frames, scale_factor=4, mode="nearest")
    assert torch.equal(out, ref)


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    frames = torch.rand(n, 1, 6, 10, generator=g)          # same sequence on every rank
    out, (lo, hi) = S.super_resolve_sequence(_Center(), frames, batch=3, rank=rank, world=world)
    counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([hi - lo]))
    assert sum(int(c) for c in counts) == n
    ref = torch.nn.functional.interpolate(frames[lo:hi], scale_factor=4, mode="nearest")
    assert torch.equal(out, ref)
    dist.barrier()
    dist.destroy_process_group()


def test_sequence_sharding_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)
