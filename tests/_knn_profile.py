"""profiling helper (not a pytest module): one knn_cuda call of this library on C5-shaped blob data"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmcuda_b200
lib = kmcuda_b200._lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
K, k, D = 200, 10, 256
rng = np.random.default_rng(777)
centers = rng.random((K, D), dtype=np.float32)
A = rng.integers(0, K, n).astype(np.uint32)
X = centers[A] + 0.05 * rng.standard_normal((n, D), dtype=np.float32)
C = np.stack([X[A == c].mean(0) for c in range(K)]).astype(np.float32)
out = np.zeros((n, k), np.uint32)
lib.knn_cuda.restype = ctypes.c_int
rc = lib.knn_cuda(k, 0, n, D, K, 1, -1, 0, 0, ctypes.c_void_p(X.ctypes.data), ctypes.c_void_p(C.ctypes.data),
                  ctypes.c_void_p(A.ctypes.data), ctypes.c_void_p(out.ctypes.data))
print("rc", rc, out[:2])
