#!/usr/bin/env python
"""Phase table (KMCUDA_B200_TIMING=1) of the bench's end-to-end call: kmeans_cuda(init=import, tolerance=1, yinyang_t=0)
on 8M x 256 @ 1024 with pinned host buffers.  Run on the GPU box: python tools/e2e_probe.py [n] [pageable]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmcuda_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
pinned = not (len(sys.argv) > 2 and sys.argv[2] == "pageable")
D, K = 256, 1024
g = torch.Generator().manual_seed(1)
Xh = torch.empty((n, D), dtype=torch.float32, pin_memory=pinned)
Xh.copy_(torch.rand((n, D), generator=g))
print("n = %d, %s host buffers" % (n, "pinned" if pinned else "pageable"), flush=True)
Ch = Xh[:K].numpy().copy()
Ah = torch.empty(n, dtype=torch.int32, pin_memory=pinned)
m = ctypes.c_uint32(0)
for i in range(4):
    os.environ["KMCUDA_B200_TIMING"] = "1" if i == 3 else "0"
    t = time.perf_counter()
    rc = kmcuda_b200._lib.kmeans_cuda(3, ctypes.byref(m), ctypes.c_float(1.0), ctypes.c_float(0.0), 0, n, D, K, 0, 1, -1,
                                      0, 0, ctypes.c_void_p(Xh.data_ptr()), ctypes.c_void_p(Ch.ctypes.data),
                                      ctypes.c_void_p(Ah.data_ptr()), None)
    print("call %d rc=%d %.1f ms" % (i, rc, (time.perf_counter() - t) * 1e3), flush=True)
