"""GPU bring-up / diagnostics script (not a pytest): run under gpurun, prints a staged report.

    python tests/gpu_bringup.py [stage ...]      # default: all stages

Every stage is wrapped so that a failure is reported and the next stage still runs.
"""
import ctypes
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402

IMPORT, L2 = 3, 0


def log(*a):
    print(*a, flush=True)


def c_kmeans(lib, X, C0, tol, yy, metric=0, verbosity=0, init=IMPORT, seed=3, fp16=False):
    X = np.ascontiguousarray(X)
    N, D = X.shape
    if fp16:
        D //= 2
    K = C0.shape[0] if hasattr(C0, "shape") else int(C0)
    C = np.array(C0, copy=True, order="C") if hasattr(C0, "shape") else np.zeros((K, X.shape[1]), X.dtype)
    A = np.zeros(N, np.uint32)
    m = ctypes.c_uint32(0)
    rc = lib.kmeans_cuda(init, ctypes.byref(m), tol, yy, metric, N, D, K, seed, 1, -1, int(fp16), verbosity,
                         X.ctypes.data, C.ctypes.data, A.ctypes.data, None)
    return rc, C, A


def c_knn(lib, k, X, C, A, metric=0, verbosity=0):
    N, D = X.shape
    out = np.zeros((N, k), np.uint32)
    rc = lib.knn_cuda(k, metric, N, D, C.shape[0], 1, -1, 0, verbosity, X.ctypes.data,
                      np.ascontiguousarray(C).ctypes.data, np.ascontiguousarray(A).ctypes.data, out.ctypes.data)
    return rc, out


def data(n, d, k, seed=777):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d), dtype=np.float32)
    C = X[rng.choice(n, k, replace=k > n)].copy()
    if k > n:
        C += rng.random((k, d), dtype=np.float32) * 0.05
    return X, C


def stage_env():
    import torch
    log("cpu_count", os.cpu_count(), "torch", torch.__version__, "cuda", torch.cuda.is_available(),
        torch.cuda.get_device_name(0) if torch.cuda.is_available() else None)
    os.system("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv")


def stage_reference():
    """pins the oracle: C restatement vs the unmodified reference rebuilt for sm_100"""
    ref = O.reference_lib()
    for (n, d, k) in [(3000, 2, 50), (3000, 64, 100), (2000, 256, 1024)]:
        X, C0 = data(n, d, k)
        t = time.time()
        rc, C, A = c_kmeans(ref, X, C0, 1.0, 0.0)
        tr = time.time() - t
        t = time.time()
        a, _, changed = O.assign_lloyd(X, C0)
        to = time.time() - t
        log("reference rc=%d n=%d d=%d k=%d  equal_to_oracle=%s mismatches=%d  t_ref=%.3fs t_oracle=%.3fs" %
            (rc, n, d, k, np.array_equal(a, A), int((a != A).sum()), tr, to))


def stage_exact():
    """our exact SIMT path (KMCUDA_B200_FORCE_EXACT=1) vs the reference"""
    import kmcuda_b200 as km
    ours = O.load_c_api(km.LIB_PATH)
    ref = O.reference_lib()
    os.environ["KMCUDA_B200_FORCE_EXACT"] = "1"
    try:
        for (n, d, k) in [(3000, 2, 50), (5000, 64, 100), (20000, 256, 1024), (1000, 7, 3), (4097, 100, 33)]:
            X, C0 = data(n, d, k)
            rc1, _, A1 = c_kmeans(ours, X, C0, 1.0, 0.0)
            rc2, _, A2 = c_kmeans(ref, X, C0, 1.0, 0.0)
            log("exact n=%d d=%d k=%d rc=%d/%d equal=%s mismatches=%d" %
                (n, d, k, rc1, rc2, np.array_equal(A1, A2), int((A1 != A2).sum())))
    finally:
        os.environ.pop("KMCUDA_B200_FORCE_EXACT", None)


def expected_scores(X, C, scale):
    """numpy model of what the tensor-core kernel accumulates: fp16(s*x).fp16(s*c) - s^2 csq/2"""
    xs = (X.astype(np.float32) * np.float32(scale)).astype(np.float16).astype(np.float64)
    cs = (C.astype(np.float32) * np.float32(scale)).astype(np.float16).astype(np.float64)
    csq = (C.astype(np.float64) ** 2).sum(1)
    return xs @ cs.T - 0.5 * scale * scale * csq[None, :]


def stage_tc_scores():
    """tcgen05 kernel numerics: dumped approximate scores vs a numpy model; both descriptor variants"""
    import torch
    from kmcuda_b200.shard import Shard
    os.environ["KMCUDA_B200_DUMP_SCORES"] = "1"
    try:
        for swap in ("0",):
            os.environ["KMCUDA_B200_PACK_SWAP"] = swap
            for (n, d, k) in [(128, 64, 256), (300, 64, 256), (512, 256, 1024), (700, 128, 300), (256, 72, 50)]:
                X, C0 = data(n, d, k, seed=5)
                Xt, Ct = torch.from_numpy(X).cuda(), torch.from_numpy(C0).cuda()
                sh = Shard(n, d, k)
                a = torch.full((n,), -1, dtype=torch.int32, device="cuda")
                prev = torch.full((n,), -1, dtype=torch.int32, device="cuda")
                ch = torch.zeros(1, dtype=torch.int32, device="cuda")
                t = time.time()
                sh.assign(Xt, Ct, a, prev, ch)
                torch.cuda.synchronize()
                dt = time.time() - t
                err = sh.last_error()
                info = sh.last_pass_info()
                st = sh.debug_stats()
                got = sh.debug_scores(n, k)
                exp = expected_scores(X, C0, st["scale"])
                diff = np.abs(got - exp)
                # the same without the bias term, to tell descriptor problems of the bias block apart
                csq = (C0.astype(np.float64) ** 2).sum(1)
                nobias = np.abs(got - (exp + 0.5 * st["scale"] ** 2 * csq[None, :]))
                ref_a, _, _ = O.assign_lloyd(X, C0)
                ours = a.cpu().numpy().astype(np.uint32)
                log("tc swap=%s n=%d d=%d k=%d err=0x%x tc=%s recheck=%d ovf=%d scale=%g cmax=%.3f dcmax=%.3g "
                    "max|got-exp|=%.4g (no-bias model %.4g) mean|exp|=%.3g assign_mismatch=%d t=%.3fs" %
                    (swap, n, d, k, err, info[0], info[1], info[2], st["scale"], st["cmax"], st["dcmax"],
                     diff.max(), nobias.max(), np.abs(exp).mean(), int((ours != ref_a).sum()), dt))
                if diff.max() > 0.5:
                    bad = np.argwhere(diff > 0.5)
                    log("   bad entries: %d of %d; rows %s cols %s" %
                        (len(bad), diff.size, np.unique(bad[:, 0])[:12], np.unique(bad[:, 1])[:12]))
                    log("   got[0,:6]", got[0, :6], "exp[0,:6]", exp[0, :6])
                    log("   got[1,:6]", got[1, :6], "exp[1,:6]", exp[1, :6])
                sh.close()
    finally:
        os.environ.pop("KMCUDA_B200_DUMP_SCORES", None)
        os.environ.pop("KMCUDA_B200_PACK_SWAP", None)


def _time_assign(sh, Xt, Ct, iters=5):
    import torch
    n = Xt.shape[0]
    a = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    prev = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    ch = torch.zeros(1, dtype=torch.int32, device="cuda")
    sh.assign(Xt, Ct, a, prev, ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        a.fill_(-1)
        e0.record()
        sh.assign(Xt, Ct, a, prev, ch)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return a, min(ts), sum(ts) / len(ts)


def stage_tc_large():
    """tensor-core path at scale: parity against the reference, first timings"""
    import torch
    import kmcuda_b200 as km
    from kmcuda_b200.shard import Shard
    ref = O.reference_lib()
    for (n, d, k) in [(100000, 256, 1024), (1000000, 256, 1024)]:
        X, C0 = data(n, d, k)
        Xt, Ct = torch.from_numpy(X).cuda(), torch.from_numpy(C0).cuda()
        sh = Shard(n, d, k)
        a, tmin, tavg = _time_assign(sh, Xt, Ct)
        err, info = sh.last_error(), sh.last_pass_info()
        t = time.time()
        rc, _, A = c_kmeans(ref, X, C0, 1.0, 0.0)
        tref = time.time() - t
        ours = a.cpu().numpy().astype(np.uint32)
        log("tc_large n=%d d=%d k=%d err=0x%x tc=%s recheck=%d (%.1f%%) ovf=%d  t_min=%.3fms t_avg=%.3fms -> %.3g pts/s; "
            "reference whole call %.3fs; mismatches vs reference=%d" %
            (n, d, k, err, info[0], info[1], 100.0 * info[1] / n, info[2], tmin, tavg, n / (tmin * 1e-3), tref,
             int((ours != A).sum())))
        os.environ["KMCUDA_B200_FORCE_EXACT"] = "1"
        try:
            she = Shard(n, d, k)
            ae, tmin, tavg = _time_assign(she, Xt, Ct, iters=2)
            log("   exact SIMT pass: t_min=%.3fms -> %.3g pts/s; equal to tc: %s" %
                (tmin, n / (tmin * 1e-3), bool((ae == a).all().item())))
            she.close()
        finally:
            os.environ.pop("KMCUDA_B200_FORCE_EXACT", None)
        sh.close()
    # headline size, generated on the device
    n, d, k = 8000000, 256, 1024
    g = torch.Generator(device="cuda").manual_seed(777)
    Xt = torch.rand((n, d), generator=g, device="cuda", dtype=torch.float32)
    Ct = Xt[torch.randperm(n, generator=g, device="cuda")[:k]].contiguous()
    sh = Shard(n, d, k)
    a, tmin, tavg = _time_assign(sh, Xt, Ct, iters=5)
    err, info = sh.last_error(), sh.last_pass_info()
    log("tc_8M err=0x%x tc=%s recheck=%d (%.1f%%) ovf=%d t_min=%.3fms t_avg=%.3fms -> %.4g pts/s, %.1f TFLOP/s algorithmic" %
        (err, info[0], info[1], 100.0 * info[1] / n, info[2], tmin, tavg, n / (tmin * 1e-3),
         2.0 * n * k * d / (tmin * 1e-3) / 1e12))
    sh.close()


def stage_update():
    import torch
    from kmcuda_b200.shard import Shard
    for (n, d, k) in [(5000, 32, 40), (20000, 256, 128)]:
        X, C0 = data(n, d, k)
        a, prev, _ = O.assign_lloyd(X, C0)
        Cexp, cnt_exp = O.adjust(X, C0, prev, a, np.zeros(k, np.uint32))
        Xt = torch.from_numpy(X).cuda()
        at = torch.from_numpy(a.astype(np.int32)).cuda()
        sums = torch.zeros((k, d), device="cuda")
        counts = torch.zeros(k, dtype=torch.int32, device="cuda")
        Ct = torch.zeros((k, d), device="cuda")
        cc = torch.zeros(k, dtype=torch.int32, device="cuda")
        sh = Shard(n, d, k)
        sh.partial_sums(Xt, at, sums, counts)
        sh.finish_update(sums, counts, Ct, cc)
        torch.cuda.synchronize()
        got = Ct.cpu().numpy()
        rel = np.abs(got - Cexp) / np.maximum(np.abs(Cexp), 1e-30)
        log("update n=%d d=%d k=%d counts_equal=%s max_rel_err=%.3g" %
            (n, d, k, np.array_equal(cc.cpu().numpy().astype(np.uint32), cnt_exp), np.nanmax(rel)))
        sh.close()


def blobs():
    rng = np.random.RandomState(0)
    arr = np.empty((13000, 2), dtype=np.float32)
    arr[:2000] = rng.rand(2000, 2) + [0, 2]
    arr[2000:4000] = rng.rand(2000, 2) - [0, 2]
    arr[4000:6000] = rng.rand(2000, 2) + [2, 0]
    arr[6000:8000] = rng.rand(2000, 2) - [2, 0]
    arr[8000:10000] = rng.rand(2000, 2) - [2, 2]
    arr[10000:] = rng.rand(3000, 2) + [2, 2]
    return arr


def stage_runs():
    """whole runs through the C ABI: ours vs the reference (Lloyd, Yinyang, init methods)"""
    import kmcuda_b200 as km
    ours = O.load_c_api(km.LIB_PATH)
    ref = O.reference_lib()
    X = blobs()
    rng = np.random.default_rng(1)
    C0 = X[rng.choice(len(X), 50, replace=False)].copy()
    for name, tol, yy, init in [("lloyd-import", 0.01, 0.0, IMPORT), ("yinyang-import", 0.01, 0.1, IMPORT),
                                ("lloyd-random", 0.01, 0.0, 0), ("lloyd-kmeans++", 0.01, 0.0, 1),
                                ("yinyang-kmeans++", 0.01, 0.1, 1)]:
        sys.stdout.flush()
        log("--- %s: ours (verbosity 1) ---" % name)
        rc1, C1, A1 = c_kmeans(ours, X, C0 if init == IMPORT else 50, tol, yy, verbosity=1, init=init)
        sys.stdout.flush()
        log("--- %s: reference (verbosity 1) ---" % name)
        rc2, C2, A2 = c_kmeans(ref, X, C0 if init == IMPORT else 50, tol, yy, verbosity=1, init=init)
        sys.stdout.flush()
        log("run %s rc=%d/%d assignments equal=%.4f centroid max rel diff=%.3g" %
            (name, rc1, rc2, (A1 == A2).mean(), np.nanmax(np.abs(C1 - C2) / np.maximum(np.abs(C2), 1e-6))))
    # k-NN
    rc, C, A = c_kmeans(ref, X, 50, 0.01, 0.1, init=1, seed=777)
    for k in (10, 50):
        rc1, n1 = c_knn(ours, k, X, C, A, verbosity=1)
        rc2, n2 = c_knn(ref, k, X, C, A, verbosity=1)
        log("knn k=%d rc=%d/%d differing entries=%d of %d" % (k, rc1, rc2, int((n1 != n2).sum()), n1.size))
    Xb, Cb = data(20000, 48, 200)
    rc, Cb2, Ab = c_kmeans(ref, Xb, Cb, 0.05, 0.0)
    rc1, n1 = c_knn(ours, 10, Xb, Cb2, Ab, verbosity=1)
    rc2, n2 = c_knn(ref, 10, Xb, Cb2, Ab, verbosity=1)
    log("knn 20000x48 k=10 rc=%d/%d differing entries=%d of %d" % (rc1, rc2, int((n1 != n2).sum()), n1.size))


STAGES = [("env", stage_env), ("reference", stage_reference), ("exact", stage_exact),
          ("tc_scores", stage_tc_scores), ("update", stage_update), ("runs", stage_runs),
          ("tc_large", stage_tc_large)]

if __name__ == "__main__":
    want = sys.argv[1:]
    for name, fn in STAGES:
        if want and name not in want:
            continue
        log("\n========== stage %s ==========" % name)
        t0 = time.time()
        try:
            fn()
        except Exception:
            log("STAGE %s FAILED:\n%s" % (name, traceback.format_exc()))
        log("---------- stage %s done in %.1fs" % (name, time.time() - t0))
