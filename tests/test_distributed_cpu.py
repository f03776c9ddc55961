"""world_size-2 gloo test of the one-process-per-GPU Lloyd driver (kmcuda_b200/distributed.py).

The shard steps are supplied by a numpy stand-in (this test is about the collective plumbing: the
all-reduce of sums / integer counts / reassignment counter and the stop rule); on GPUs the same driver
runs over kmcuda_b200.shard.Shard with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyBackend:
    """stand-in with the three shard steps of include/kmcuda_b200.h"""

    def assign(self, X, C, assign, prev, changed):
        d = ((X[:, None, :].double() - C[None].double()) ** 2).sum(-1)
        new = d.argmin(1).to(torch.int32)
        prev.copy_(assign)
        changed += int((new != assign).sum())
        assign.copy_(new)

    def partial_sums(self, X, assign, sums, counts):
        sums.zero_()
        counts.zero_()
        sums.index_add_(0, assign.long(), X)
        counts.copy_(torch.bincount(assign.long(), minlength=sums.shape[0]).to(torch.int32))

    def finish_update(self, sums, counts, C, ccounts):
        C.copy_(sums / counts[:, None].float())
        ccounts.copy_(counts)


class AllReduceExchange:
    """stand-in with the interface of kmcuda_b200.shard.PeerExchange (update = partial sums + sum over the ranks)"""

    def __init__(self):
        self.calls = 0

    def update(self, backend, X, assign, sums, counts):
        self.calls += 1
        backend.partial_sums(X, assign, sums, counts)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)


def _worker(rank, world, port, X, C0, out, use_exchange=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kmcuda_b200.distributed import sharded_lloyd
    n = X.shape[0]
    lo, hi = rank * n // world, (rank + 1) * n // world
    C = C0.clone()
    ex = AllReduceExchange() if use_exchange else None
    assign, iters = sharded_lloyd(NumpyBackend(), X[lo:hi].contiguous(), C, total_samples=n, tolerance=0.0, exchange=ex)
    assert ex is None or ex.calls == iters - 1        # one exchange per centroid update
    out[rank] = (assign.numpy().copy(), C.numpy().copy(), iters)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_exchange", [False, True])
def test_sharded_lloyd_two_ranks_equals_single_rank(use_exchange):
    """use_exchange: the update's sums go through the `exchange=` route of sharded_lloyd (on GPUs: PeerExchange)"""
    pytest.importorskip("kmcuda_b200")
    rng = np.random.default_rng(0)
    centers = rng.random((6, 8)) * 10
    X = torch.from_numpy((centers[rng.integers(0, 6, 3000)] + rng.standard_normal((3000, 8)) * 0.2).astype(np.float32))
    C0 = X[torch.from_numpy(rng.choice(3000, 6, replace=False))].clone()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29533 + int(use_exchange), X, C0, out, use_exchange), nprocs=2, join=True)
    a = np.concatenate([out[0][0], out[1][0]])
    # single "rank" reference of the same driver
    sys.path.insert(0, ROOT)
    from kmcuda_b200.distributed import sharded_lloyd
    C1 = C0.clone()
    a1, it1 = sharded_lloyd(NumpyBackend(), X, C1, total_samples=3000, tolerance=0.0)
    assert out[0][2] == out[1][2] == it1
    assert np.array_equal(a, a1.numpy())
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=0, atol=0)      # every rank ends with the same centroids
    np.testing.assert_allclose(out[0][1], C1.numpy(), rtol=1e-5)
