#!/usr/bin/env python
"""Three full Lloyd iterations (assign pass + member sums + normalise) on N x 256 @ 1024 resident samples through the
shard-level API: the workload of the launch lists / ncu captures of the update kernels (profiles/README.md).
    python tools/iteration_probe.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kmcuda_b200.shard import Shard

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
D, K = 256, 1024
g = torch.Generator(device="cuda").manual_seed(777)
X = torch.rand((n, D), generator=g, device="cuda", dtype=torch.float32)
C = X[torch.randperm(n, generator=g, device="cuda")[:K]].contiguous()
sh = Shard(n, D, K)
a = torch.full((n,), -1, dtype=torch.int32, device="cuda")
prev = torch.full((n,), -1, dtype=torch.int32, device="cuda")
ch = torch.zeros(1, dtype=torch.int32, device="cuda")
sums = torch.zeros((K, D), dtype=torch.float32, device="cuda")
counts = torch.zeros(K, dtype=torch.int32, device="cuda")
cc = torch.zeros(K, dtype=torch.int32, device="cuda")
for it in range(3):
    ch.zero_()
    sh.assign(X, C, a, prev, ch)
    sh.partial_sums(X, a, sums, counts)
    sh.finish_update(sums, counts, C, cc)
    torch.cuda.synchronize()
    print("iteration %d: %d reassignments, pipeline error 0x%x" % (it + 1, int(ch.item()), sh.last_error()), flush=True)

# phase times of 10 more iterations (CUDA events)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
for e in ev:
    e[0].record()
    sh.assign(X, C, a, prev, ch)
    e[1].record()
    sh.partial_sums(X, a, sums, counts)
    e[2].record()
    sh.finish_update(sums, counts, C, cc)
    e[3].record()
torch.cuda.synchronize()
ph = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / len(ev) for j in range(3)]
print("ITER lib=%s n=%d assign %.4f ms, member sums %.4f ms, normalise %.4f ms, iteration %.4f ms" % (
    os.environ.get("KMCUDA_B200_LIB", "product").split("/")[-2:][0], n, ph[0], ph[1], ph[2],
    ev[0][0].elapsed_time(ev[-1][3]) / len(ev)), flush=True)
