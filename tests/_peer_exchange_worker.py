#!/usr/bin/env python
"""One rank of the peer-memory exchange test (launched by torchrun from test_parity_r2_gpu.py, >= 2 GPUs):
sharded Lloyd over `kmcuda_b200.shard.PeerExchange` next to the same run over the NCCL all-reduce.  Prints
PEER_EXCHANGE_OK on rank 0 when
  * the totals of one exchange equal the fp64 sum of the ranks' partial sums to fp32 accuracy and are bit-identical on all
    ranks, the counts exact,
  * 12 iterations in a row stay in step (double-buffered partial sums, no host synchronisation in between),
  * whole runs over the two exchange routes end with the same assignments and centroids within 1e-5."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from kmcuda_b200.shard import Shard, PeerExchange
    from kmcuda_b200.distributed import sharded_lloyd
    n_total, D, K = 400000, 128, 256
    rng = np.random.default_rng(5)
    centers = rng.random((K, D), dtype=np.float32)
    Xall = (centers[rng.integers(0, K, n_total)] + 0.2 * rng.standard_normal((n_total, D), dtype=np.float32)).astype(np.float32)
    C0 = Xall[rng.choice(n_total, K, replace=False)].copy()
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    X = torch.from_numpy(Xall[lo:hi]).cuda()
    n = X.shape[0]
    sh = Shard(n, D, K)
    ex = PeerExchange(K, D)

    # ---- one exchange against an independent sum
    a = torch.from_numpy(rng.integers(0, K, n_total)[lo:hi].astype(np.int32)).cuda()
    part = torch.zeros((K, D), device="cuda")
    pcnt = torch.zeros(K, dtype=torch.int32, device="cuda")
    sh.partial_sums(X, a, part, pcnt)
    gathered = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    gc = [torch.empty_like(pcnt) for _ in range(world)]
    dist.all_gather(gc, pcnt)
    tot = torch.zeros((K, D), device="cuda")
    cnt = torch.zeros(K, dtype=torch.int32, device="cuda")
    ex.update(sh, X, a, tot, cnt)
    torch.cuda.synchronize()
    assert ex.error() == 0
    exp = gathered[0].clone()
    for g in gathered[1:]:
        exp += g                                   # rank order, fp32: what the kernel does
    assert torch.equal(tot, exp), float((tot - exp).abs().max())
    assert torch.equal(cnt, sum(gc[1:], gc[0].clone()))
    alltot = [torch.empty_like(tot) for _ in range(world)]
    dist.all_gather(alltot, tot)
    assert all(torch.equal(alltot[0], t) for t in alltot)

    # ---- many iterations back to back, different data each time, no host sync in between
    outs = []
    for it in range(12):
        ai = ((a.long() + it * 7) % K).to(torch.int32)
        t_i = torch.zeros((K, D), device="cuda")
        c_i = torch.zeros(K, dtype=torch.int32, device="cuda")
        ex.update(sh, X, ai, t_i, c_i)
        outs.append((ai, t_i, c_i))
    torch.cuda.synchronize()
    assert ex.error() == 0
    for ai, t_i, c_i in outs:
        p_i = torch.zeros((K, D), device="cuda")
        pc_i = torch.zeros(K, dtype=torch.int32, device="cuda")
        sh.partial_sums(X, ai, p_i, pc_i)
        g_i = [torch.empty_like(p_i) for _ in range(world)]
        dist.all_gather(g_i, p_i)
        e_i = g_i[0].clone()
        for g in g_i[1:]:
            e_i += g
        assert torch.equal(t_i, e_i)
        dist.all_reduce(pc_i)
        assert torch.equal(c_i, pc_i)

    # ---- whole runs over both routes
    C1 = torch.from_numpy(C0).cuda()
    a1, it1 = sharded_lloyd(sh, X, C1, n_total, tolerance=0.0005, exchange=ex)
    C2 = torch.from_numpy(C0).cuda()
    a2, it2 = sharded_lloyd(sh, X, C2, n_total, tolerance=0.0005)
    torch.cuda.synchronize()
    assert it1 == it2, (it1, it2)
    assert float((a1 != a2).float().mean()) < 1e-4
    live = torch.isfinite(C2).all(1)             # (a cluster that lost all members has no centroid in either run)
    assert torch.equal(live, torch.isfinite(C1).all(1)) and int(live.sum()) > K // 2
    rel = ((C1[live] - C2[live]).abs() / C2[live].abs().amax(1, keepdim=True)).max()
    assert float(rel) < 1e-5, float(rel)
    ex.close()
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK world=%d iterations=%d" % (world, it1), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
