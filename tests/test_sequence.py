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
    # GenerateFrameIndiceswithPadding 'reflection' (mmedit augmentation.py:860-861,869-870)
    assert S.window_indices(0, 100, "reflection") == [3, 2, 1, 0, 1, 2, 3]
    assert S.window_indices(99, 100, "reflection") == [96, 97, 98, 99, 98, 97, 96]


def _mmedit_indices(t, n, mode, width=7):
    """Transcription of the index arithmetic documented in GenerateFrameIndiceswithPadding's docstring examples
    (augmentation.py:816-824): checked against those examples below before it is used as the expectation."""
    pad, last = width // 2, n - 1
    out = []
    for i in range(t - pad, t + pad + 1):
        if i < 0:
            out.append({"replicate": 0, "reflection": -i, "reflection_circle": t + pad - i, "circle": width + i}[mode])
        elif i > last:
            out.append({"replicate": last, "reflection": 2 * last - i, "reflection_circle": (t - pad) - (i - last),
                        "circle": i - width}[mode])
        else:
            out.append(i)
    return out


def test_window_indices_all_mmedit_modes():
    # the reference's own docstring examples (current_idx = 0, num_input_frames = 5)
    assert _mmedit_indices(0, 100, "replicate", 5) == [0, 0, 0, 1, 2]
    assert _mmedit_indices(0, 100, "reflection", 5) == [2, 1, 0, 1, 2]
    assert _mmedit_indices(0, 100, "reflection_circle", 5) == [4, 3, 0, 1, 2]
    assert _mmedit_indices(0, 100, "circle", 5) == [3, 4, 0, 1, 2]
    # FCVSR REDS test pipeline (fcvsr_redsLD_QP22.py:31) and pad_sequence (restoration_video_inference.py:16-25)
    assert S.window_indices(0, 100, "reflection_circle") == [6, 5, 4, 0, 1, 2, 3]
    assert S.window_indices(1, 100, "reflection_circle") == [6, 5, 0, 1, 2, 3, 4]
    assert S.window_indices(99, 100, "reflection_circle") == [96, 97, 98, 99, 95, 94, 93]
    assert S.window_indices(0, 100, "circle") == [4, 5, 6, 0, 1, 2, 3]
    for mode in S.MODES[:4]:
        for t in range(100):
            assert S.window_indices(t, 100, mode) == _mmedit_indices(t, 100, mode)


def test_pad_sequence_mode():
    # mmedit restoration_video_inference.pad_sequence: cat([data[1+p:1+2p].flip, data, data[-1-2p:-1-p].flip]) then a sliding window
    n, p = 12, 3
    data = list(range(n))
    padded = data[1 + p:1 + 2 * p][::-1] + data + data[-1 - 2 * p:-1 - p][::-1]
    for t in range(n):
        assert padded[t:t + 2 * p + 1] == S.window_indices(t, n, "pad_sequence")
    assert S.window_indices(0, n, "pad_sequence") == S.window_indices(0, n, "reflection_circle") == [6, 5, 4, 0, 1, 2, 3]


def test_shards_cover_sequence_and_halos_suffice():
    for mode in S.MODES:
        for n in (1, 5, 13, 100, 101):
            for world in (1, 2, 4, 8):
                covered = []
                for r in range(world):
                    lo, hi = S.shard_range(n, r, world)
                    covered += list(range(lo, hi))
                    h_lo, h_hi = S.halo_range(lo, hi, n, mode)
                    assert 0 <= h_lo <= h_hi <= n
                    for t in range(lo, hi):
                        assert all(h_lo <= j < h_hi for j in S.window_indices(t, n, mode))
                assert covered == list(range(n))
    # interior shards only ever need the 3-frame halo; the circle modes reach 6 frames at the sequence ends
    assert S.halo_range(40, 60, 100, "reflection_circle") == (37, 63)
    assert S.halo_range(0, 10, 100, "reflection_circle") == (0, 13)
    assert S.halo_range(0, 2, 100, "reflection_circle") == (0, 7)


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
    for mode in S.MODES:
        out, (lo, hi) = S.super_resolve_sequence(_Center(), frames, batch=4, mode=mode)
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
