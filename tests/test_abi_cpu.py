"""CPU tests of the drop-in boundary: exported symbols, argument validation, Python-surface errors."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = src.split("#ifdef __cplusplus\n#include <string>")[0]
    return sorted(set(re.findall(r"\b(k(?:means|nn)_cuda|kmcuda_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import kmcuda_b200
    lib = ctypes.CDLL(kmcuda_b200.LIB_PATH)
    names = _declared_functions("kmcuda.h") + _declared_functions("kmcuda_b200.h")
    assert "kmeans_cuda" in names and "knn_cuda" in names and "kmcuda_b200_assign" in names
    for n in names:
        assert hasattr(lib, n), n


def test_enum_values_cross_the_abi_unchanged():
    import kmcuda_b200 as km
    assert (km.SUCCESS, km.INVALID_ARGUMENTS, km.NO_SUCH_DEVICE, km.MEMORY_ALLOCATION_FAILURE, km.RUNTIME_ERROR,
            km.MEMORY_COPY_ERROR) == (0, 1, 2, 3, 4, 5)
    assert (km.INIT_RANDOM, km.INIT_PLUSPLUS, km.INIT_AFKMC2, km.INIT_IMPORT) == (0, 1, 2, 3)
    hdr = open(os.path.join(ROOT, "include", "kmcuda.h")).read()
    order = [m for m in re.findall(r"\b(kmcuda[A-Z]\w+)\b", hdr.split("} KMCUDAResult;")[0])]
    assert order[:6] == ["kmcudaSuccess", "kmcudaInvalidArguments", "kmcudaNoSuchDevice",
                         "kmcudaMemoryAllocationFailure", "kmcudaRuntimeError", "kmcudaMemoryCopyError"]


def _call(lib, N, D, K, tol=0.01, yy=0.1, samples=True, centroids=True, assignments=True, device=1):
    X = np.zeros((max(N, 1), max(D, 1)), np.float32)
    C = np.zeros((min(max(K, 1), 1000), max(D, 1)), np.float32)
    A = np.zeros(max(N, 1), np.uint32)
    return lib.kmeans_cuda(1, None, ctypes.c_float(tol), ctypes.c_float(yy), 0, N, D, K, 0, device, -1, 0, 0,
                           X.ctypes.data if samples else None, C.ctypes.data if centroids else None,
                           A.ctypes.data if assignments else None, None)


def test_argument_validation_order_matches_reference():
    """reference check_kmeans_args (kmcuda.cc:19-61): cluster / feature / sample-count checks come first"""
    import kmcuda_b200 as km
    lib = km._lib
    assert _call(lib, 100, 4, 1) == km.INVALID_ARGUMENTS          # clusters < 2
    assert _call(lib, 100, 4, 0xFFFFFFFF) == km.INVALID_ARGUMENTS
    assert _call(lib, 100, 0, 5) == km.INVALID_ARGUMENTS          # no features
    assert _call(lib, 3, 4, 5) == km.INVALID_ARGUMENTS            # fewer samples than clusters
    import torch
    if not torch.cuda.is_available():
        assert _call(lib, 100, 4, 5) == km.NO_SUCH_DEVICE
        assert _call(lib, 100, 4, 5, device=0) == km.NO_SUCH_DEVICE
    rc = lib.knn_cuda(0, 0, 10, 2, 2, 1, -1, 0, 0, None, None, None, None)
    assert rc == km.INVALID_ARGUMENTS


def test_python_surface_errors_like_libKMCUDA():
    """reference src/test.py:189-205 (test_crap) and python.cc:88-157"""
    import kmcuda_b200 as km
    arr = np.random.rand(100, 2).astype(np.float32)
    with pytest.raises(TypeError):
        km.kmeans_cuda(arr, 5, metric=3)
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, metric="manhattan")
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, init="bogus")
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 1)
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr[0], 5)
    with pytest.raises(ValueError):
        km.kmeans_cuda(np.random.rand(10, 3).astype(np.float16), 5)     # odd feature count in fp16 mode
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, init=np.zeros((4, 2), np.float32))        # wrong centroid count
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, init=np.zeros((5, 3), np.float32))        # wrong feature count
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, tolerance=100)                            # rejected by the C layer
    with pytest.raises(ValueError):
        km.kmeans_cuda(arr, 5, yinyang_t=10)
    with pytest.raises(ValueError):
        km.knn_cuda(0, arr, np.zeros((5, 2), np.float32), np.zeros(100, np.uint32))
    with pytest.raises(ValueError):
        km.knn_cuda(3, arr, np.zeros((5, 2), np.float32), np.zeros(99, np.uint32))
    assert km.supports_fp16 is True


def test_no_cpu_fallback_in_product_package():
    """the product never imports the oracle"""
    for root, _, files in os.walk(os.path.join(ROOT, "kmcuda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(root, f)


def test_libKMCUDA_module_imports_from_the_same_shared_object():
    """reference src/python.cc:33-54: the .so is itself the Python module `libKMCUDA`"""
    import importlib.util
    import kmcuda_b200 as km
    spec = importlib.util.spec_from_file_location("libKMCUDA", km.LIB_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.supports_fp16 is True
    arr = np.random.rand(100, 2).astype(np.float32)
    with pytest.raises(TypeError):
        mod.kmeans_cuda(arr, 5, metric=3)
    with pytest.raises(ValueError):
        mod.kmeans_cuda(arr, 1)
    with pytest.raises(ValueError):
        mod.kmeans_cuda(arr, 5, init="bogus")
    with pytest.raises(ValueError):
        mod.kmeans_cuda(arr, 5, init=np.zeros((4, 2), np.float32))
    with pytest.raises(ValueError):
        mod.knn_cuda(0, arr, np.zeros((5, 2), np.float32), np.zeros(100, np.uint32))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="No such CUDA device"):
            mod.kmeans_cuda(arr, 5)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line (library chatter and NCCL banners go to stderr).  Without a
    GPU the reference arm times the CPU oracle port on a bounded sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kmeans_assign_points_per_sec" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
