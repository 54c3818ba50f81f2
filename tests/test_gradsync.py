"""CPU tests (gloo, world size 2) of the data-parallel gradient all-reduce (BASELINE config 4)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcvsr_b200.gradsync import GradAllReducer


class _Net(torch.nn.Module):
    """Small stand-in with the two quirks of the reference model: an aliased sub-module (registered twice, as
    recorb1...RCB == body.3, CVSR_freq.py:736,751) and parameters that never receive a gradient (DivEnh.Conv, :2104-2133)."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.b = torch.nn.Conv2d(8, 8, 3, padding=1)
        self.alias = self.b
        self.unused = torch.nn.Conv2d(8, 8, 3, padding=1)
        self.c = torch.nn.Conv2d(8, 1, 1)

    def forward(self, x):
        return self.c(torch.relu(self.alias(torch.relu(self.a(x)))))


def _worker(rank, world, port, overlap, bucket_mb):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = _Net()                                               # same weights on every rank
    red = GradAllReducer(net.parameters(), bucket_mb=bucket_mb, overlap=overlap)
    assert sum(len(b) for b in red.buckets) == 8               # a, b, unused, c (weight + bias each); the alias counted once
    for step in range(2):
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(4, 3, 8, 8, generator=g)
        net.zero_grad(set_to_none=True)
        torch.sqrt(net(x) ** 2 + 1e-4).sum().backward()        # Charbonnier-sum on this rank's batch
        local = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        red.finish()
        for n, p in net.named_parameters():
            if n.startswith("unused"):
                assert p.grad is None                          # no rank produced one: stays None (find_unused_parameters)
                continue
            parts = [torch.zeros_like(local[n]) for _ in range(world)]
            dist.all_gather(parts, local[n])
            want = sum(parts) / world
            assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7), (step, n)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,bucket_mb", [(True, 9.0), (False, 9.0), (True, 0.001)])
def test_gradient_allreduce_two_ranks_gloo(overlap, bucket_mb):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, overlap, bucket_mb), nprocs=2, join=True)


def test_buckets_follow_reverse_parameter_order():
    net = _Net()
    red = GradAllReducer(net.parameters(), bucket_mb=0.001, overlap=False)
    order = [id(p) for b in red.buckets for p in b]
    want = []
    for p in net.parameters():
        if id(p) not in want:
            want.append(id(p))
    assert order == want[::-1] and len(red.buckets) > 1


def _charb(sr, hr):
    d = sr - hr
    return torch.sqrt(d * d + 1e-4).sum()                      # opt/loss.py:20-31


def _train_worker(rank, world, port):
    from fcvsr_b200.train import replicas_in_sync, train_step
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = _Net()
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, weight_decay=1e-5)
    red = GradAllReducer(net.parameters())
    # single-process twin: the concatenated batch with the learning rate divided by the world size (sum-reduced loss)
    torch.manual_seed(0)
    twin = _Net()
    opt_twin = torch.optim.Adam(twin.parameters(), lr=5e-4, weight_decay=1e-5)
    for step in range(3):
        xs = [torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(10 * step + r)) for r in range(world)]
        hs = [torch.rand(2, 1, 8, 8, generator=torch.Generator().manual_seed(77 * step + r)) for r in range(world)]
        train_step(net, opt, xs[rank], hs[rank], _charb, red)
        opt_twin.zero_grad(set_to_none=True)
        (_charb(twin(torch.cat(xs)), torch.cat(hs)) / world).backward()
        opt_twin.step()
    assert replicas_in_sync(net)
    for (n, p), q in zip(net.named_parameters(), twin.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), n
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_train_step_two_ranks_gloo():
    """Two replicas stepping through fcvsr_b200.train.train_step stay identical and equal a single process on the concatenated
    batch whose (sum-reduced) loss is divided by the world size -- the 1/G effective learning rate of SURVEY 8e."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_train_worker, args=(2, port), nprocs=2, join=True)


class _TwoBranch(torch.nn.Module):
    """Two independent branches whose backward order depends on the order they are summed in."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Linear(4, 4)
        self.q = torch.nn.Linear(4, 4)
        self.r = torch.nn.Linear(4, 4)

    def forward(self, x, flip, use_r):
        a, b = self.p(x), self.q(x)
        y = (b.sum() + a.sum()) if flip else (a.sum() + b.sum())
        return y + self.r(x).sum() if use_r else y


def _order_worker(rank, world, port, mismatch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                                    # different initial weights: the constructor broadcasts rank 0's
    net = _TwoBranch()
    red = GradAllReducer(net.parameters(), bucket_mb=1e-5)     # one bucket per parameter
    ref = [p.detach().clone() for p in net.parameters()]
    for r in ref:
        dist.broadcast(r, src=0)
    assert all(torch.equal(p.detach(), r) for p, r in zip(net.parameters(), ref))
    x = torch.randn(3, 4, generator=torch.Generator().manual_seed(5 + rank))
    # rank 1 sums the branches in the other order (its hooks fire in another order); with `mismatch` branch r is used on
    # rank 0 only, which the first-step consistency check must reject on every rank instead of hanging
    use_r = (rank == 0) if mismatch else False
    net(x, flip=bool(rank), use_r=use_r).backward()
    local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    if mismatch:
        with pytest.raises(RuntimeError, match="some ranks only"):
            red.finish()
    else:
        red.finish()
        for p, g in zip(net.parameters(), local):
            if g is None:
                assert p.grad is None
                continue
            parts = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            assert torch.allclose(p.grad, sum(parts) / world, rtol=1e-6, atol=1e-7)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mismatch", [False, True])
def test_collective_order_is_rank_independent(mismatch):
    """Hooks firing in different orders on different ranks must not reorder the collectives (strict bucket order), and a
    rank-dependent set of unused parameters is detected on the first step."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_order_worker, args=(2, port, mismatch), nprocs=2, join=True)
