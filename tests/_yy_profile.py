"""profiling helper (not a pytest module): one Yinyang run of this library on uniform data"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmcuda_b200
lib = kmcuda_b200._lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
rng = np.random.default_rng(777)
X = rng.random((n, 256), dtype=np.float32)
C = X[rng.choice(n, 1024, replace=False)].copy()
A = np.zeros(n, np.uint32); m = ctypes.c_uint32(0)
lib.kmeans_cuda.restype = ctypes.c_int
rc = lib.kmeans_cuda(3, ctypes.byref(m), ctypes.c_float(0.01), ctypes.c_float(0.1), 0, n, 256, 1024, 3, 1, -1, 0, 0,
                     ctypes.c_void_p(X.ctypes.data), ctypes.c_void_p(C.ctypes.data), ctypes.c_void_p(A.ctypes.data), None)
print("rc", rc)
