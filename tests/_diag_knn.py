"""diagnostic (not a pytest module): knn tensor-core path statistics"""
import os, sys, ctypes
import numpy as np
os.environ["KMCUDA_B200_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
import kmcuda_b200
ours = O.load_c_api(kmcuda_b200.LIB_PATH)
def km(X, C0, tol):
    N, D = X.shape; K = C0.shape[0]; C = C0.copy(); A = np.zeros(N, np.uint32); m = ctypes.c_uint32(0)
    assert ours.kmeans_cuda(3, ctypes.byref(m), tol, 0.0, 0, N, D, K, 3, 1, -1, 0, 0, X.ctypes.data, C.ctypes.data, A.ctypes.data, None) == 0
    return C, A
def knn(k, X, C, A):
    out = np.zeros((len(X), k), np.uint32)
    assert ours.knn_cuda(k, 0, X.shape[0], X.shape[1], C.shape[0], 1, -1, 0, 0, X.ctypes.data, C.ctypes.data, A.ctypes.data, out.ctypes.data) == 0
    return out
rng = np.random.default_rng(1)
for (n, d, kc, k) in ((30000, 48, 200, 10), (60000, 64, 300, 10)):
    X = rng.random((n, d), dtype=np.float32)
    C0 = X[rng.choice(n, kc, replace=False)].copy()
    C, A = km(X, C0, 0.05)
    print("case", n, d, kc, k, flush=True)
    sys.stderr.flush()
    knn(k, X, C, A)
