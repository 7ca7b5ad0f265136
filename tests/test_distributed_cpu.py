"""CPU, world_size 2 over gloo: the host-side data-parallel logic (shard ranges, owner split of a global minibatch,
flat gradient all-reduce that counts the replicated KL once)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpzoo_b200.distributed import FlatGradReducer, owner_split, shard_range
    torch.manual_seed(0)
    N = 103
    # a "model": shared params and a per-spot param; loss = sum_n f(theta, v_n, x_n) - KL(theta)
    theta = torch.nn.Parameter(torch.randn(7, dtype=torch.float64))
    W = torch.nn.Parameter(torch.randn(3, 4, dtype=torch.float64))
    x = torch.randn(N, dtype=torch.float64)
    idx = torch.randperm(N)[:40]                                   # one global minibatch, as utilities.py:605
    lo, hi = shard_range(N, world, rank)
    mask, local = owner_split(idx, N, world, rank)
    v = torch.nn.Parameter(torch.ones(hi - lo, dtype=torch.float64))

    def ll(xs, vs):
        return (torch.sin(xs * theta.sum()) * vs * W.pow(2).sum()).sum()

    kl = lambda: (theta ** 2).sum() + W.abs().sum()
    red = FlatGradReducer([theta, W])
    loss = ll(x[lo:hi][local], v[local]) - kl() / world               # kl_weight = 1/world
    loss.backward()
    total = red.all_reduce(loss)
    # single-process reference on the same global index set
    th2, W2, v2 = (torch.nn.Parameter(t.detach().clone()) for t in (theta, W, torch.ones(N, dtype=torch.float64)))
    ref = (torch.sin(x[idx] * th2.sum()) * v2[idx] * W2.pow(2).sum()).sum() - ((th2 ** 2).sum() + W2.abs().sum())
    ref.backward()
    ok = (torch.allclose(total, ref.detach(), rtol=1e-12) and torch.allclose(theta.grad, th2.grad, rtol=1e-12)
          and torch.allclose(W.grad, W2.grad, rtol=1e-12)
          and torch.allclose(v.grad, v2.grad[lo:hi], rtol=1e-12))
    ret[rank] = bool(ok) and int(mask.sum()) == len(local)
    dist.destroy_process_group()


def test_shard_ranges_cover():
    from gpzoo_b200.distributed import shard_range
    for n, w in ((103, 2), (32768, 8), (5, 8), (1000000, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.timeout(120)
def test_flat_allreduce_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
