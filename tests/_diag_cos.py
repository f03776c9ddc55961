"""diagnostic: cosine runs, ours vs reference, per tolerance (not a pytest module)"""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
import kmcuda_b200
ours = O.load_c_api(kmcuda_b200.LIB_PATH); ref = O.reference_lib()
def run(lib, X, C0, tol, yy, metric, verbosity=0):
    N, D = X.shape; K = C0.shape[0]
    C = C0.copy(); A = np.zeros(N, np.uint32); m = ctypes.c_uint32(0)
    rc = lib.kmeans_cuda(3, ctypes.byref(m), tol, yy, metric, N, D, K, 3, 1, -1, 0, verbosity, X.ctypes.data, C.ctypes.data, A.ctypes.data, None)
    assert rc == 0
    return C, A
rng = np.random.default_rng(74)
X = rng.standard_normal((30000, 32)).astype(np.float32); X /= np.linalg.norm(X, axis=1, keepdims=True); X = X.astype(np.float32)
C0 = X[rng.choice(30000, 64, replace=False)].copy()
for tol, yy in ((0.99, 0), (0.5, 0), (0.2, 0), (0.12, 0), (0.04, 0), (0.04, 0.1), (0.01, 0), (0.01, 0.1)):
    Co, Ao = run(ours, X, C0, tol, yy, 1); Cr, Ar = run(ref, X, C0, tol, yy, 1)
    print("tol", tol, "yy", yy, "equal", (Ao == Ar).mean(), "cdiff", np.abs(Co - Cr).max(), "norms", np.linalg.norm(Co, axis=1)[:3], np.linalg.norm(Cr, axis=1)[:3], flush=True)
print("--- ours log"); sys.stdout.flush(); run(ours, X, C0, 0.04, 0.1, 1, 1); sys.stdout.flush()
print("--- ref log"); sys.stdout.flush(); run(ref, X, C0, 0.04, 0.1, 1, 1)
