#!/usr/bin/env python
"""Timing of the member-sum pass (kmcuda_b200_partial_sums) of the library in KMCUDA_B200_LIB on N x 256 @ 1024:
cluster sizes as one assignment pass against random rows leaves them (skewed) and uniformly random labels (balanced).
    KMCUDA_B200_LIB=variants/x/libKMCUDA.so python tools/sums_probe.py [n]"""
import json
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
sh.assign(X, C, a, prev, ch)
ab = torch.randint(0, K, (n,), generator=g, device="cuda", dtype=torch.int32)
sums = torch.zeros((K, D), dtype=torch.float32, device="cuda")
counts = torch.zeros(K, dtype=torch.int32, device="cuda")
out = {"lib": os.environ.get("KMCUDA_B200_LIB", "product"), "n": n}
for name, lab in (("skewed", a), ("balanced", ab)):
    sh.partial_sums(X, lab, sums, counts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sh.partial_sums(X, lab, sums, counts)
    e1.record()
    torch.cuda.synchronize()
    out[name + "_ms"] = round(e0.elapsed_time(e1) / 5, 4)
    out[name + "_max_cluster"] = int(counts.max().item())
    out[name + "_checksum"] = float(sums.double().sum().item())
print("SUMS " + json.dumps(out), flush=True)
