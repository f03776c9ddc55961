"""multi-GPU bring-up (run with gpurun --gpus 2): single-process device mask + NCCL, and torchrun bench"""
import os, sys, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmcuda_b200 as km
print("devices", km.device_count(), flush=True)
rng = np.random.default_rng(0)
X = rng.random((200000, 64), dtype=np.float32)
C0 = X[rng.choice(len(X), 100, replace=False)].copy()
# two assignment passes with one update in between (tolerance=0.99): trajectories cannot diverge yet
t = time.time(); c1, a1 = km.kmeans_cuda(X, 100, init=C0, device=1, tolerance=0.99, yinyang_t=0, verbosity=1); t1 = time.time() - t
t = time.time(); c2, a2 = km.kmeans_cuda(X, 100, init=C0, device=3, tolerance=0.99, yinyang_t=0, verbosity=1); t2 = time.time() - t
print("1 gpu %.2fs, 2 gpus %.2fs, assignments equal %.6f, centroid max rel diff %.3g" % (t1, t2, (a1 == a2).mean(), np.abs(c1 - c2).max() / np.abs(c1).max()), flush=True)
# clustered data: whole runs must agree
centers = rng.random((20, 64), dtype=np.float32) * 10
Xc = (centers[rng.integers(0, 20, 200000)] + rng.standard_normal((200000, 64)).astype(np.float32) * 0.3).astype(np.float32)
Cc = Xc[rng.choice(len(Xc), 20, replace=False)].copy()
c1, a1 = km.kmeans_cuda(Xc, 20, init=Cc, device=1, tolerance=0.0001, yinyang_t=0)
c2, a2 = km.kmeans_cuda(Xc, 20, init=Cc, device=3, tolerance=0.0001, yinyang_t=0)
c3, a3 = km.kmeans_cuda(Xc, 20, init=Cc, device=0, tolerance=0.0001, yinyang_t=0.1)
print("clustered: 1 vs 2 gpus equal %.6f; all-gpu yinyang vs 1-gpu lloyd %.6f; centroid rel diff %.3g" % ((a1 == a2).mean(), (a1 == a3).mean(), np.abs(c1 - c2).max() / np.abs(c1).max()), flush=True)
nb1 = km.knn_cuda(5, X[:20000], c1, a1[:20000], device=1)
nb2 = km.knn_cuda(5, X[:20000], c1, a1[:20000], device=3)
print("knn 1 vs 2 gpus equal:", np.array_equal(nb1, nb2), flush=True)
# Yinyang iterations on two shards (bounds, tensor-core local step and NCCL update per shard)
y1, b1 = km.kmeans_cuda(X, 100, init=C0, device=1, tolerance=0.03, yinyang_t=0.1)
y2, b2 = km.kmeans_cuda(X, 100, init=C0, device=3, tolerance=0.03, yinyang_t=0.1)
print("yinyang 1 vs 2 gpus: assignments equal %.6f, centroid max rel diff %.3g" % ((b1 == b2).mean(), np.abs(y1 - y2).max() / np.abs(y1).max()), flush=True)
