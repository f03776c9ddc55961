#!/usr/bin/env python
"""bench.py -- headline benchmark: k-means assignment step, points/sec, 8M x 256 fp32 @ 1024 clusters.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P]

A "step" is ONE assignment pass of the hot path (the reference's kmeans_assign_lloyd, src/kmeans.cu:
293-364; here: tcgen05 distance filter + exact fp32 re-check + bookkeeping) over the rank's shard of
synthetic samples that are already resident in HBM.  One process per GPU (torchrun for N > 1), shards
are independent (no data-path collective in the assignment step) -> "scaling": "weak".

Keys beyond the base contract: `e2e` (same metric through the reference-facing C ABI kmeans_cuda() with
HOST buffers: H2D of the samples and D2H of the assignments inside the timed region), `roofline` (the
tcgen05 kernel against the measured bf16 tensor peak of MEASURED_PEAKS.json), `cpu_baseline` (the C
oracle port on the host cores, bounded sample), `clocks`.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libKMCUDA.so, src-d/kmcuda rebuilt for
sm_100 -- the reference has no CPU implementation, it is a CUDA library) through the same C ABI call on
a bounded sample of the same workload; if that library cannot be loaded it times the CPU oracle port.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.pop("NCCL_DEBUG", None)   # some hosts export NCCL_DEBUG=VERSION; NCCL prints its banner on stdout
# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL banners, C-library progress
# messages) is sent to stderr, the JSON line goes to the saved descriptor
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


METRIC = "kmeans_assign_points_per_sec"
UNIT = "points/s"
N_POINTS, D, K = 8000000, 256, 1024
WORKLOAD = "k-means assignment step, %dx%d fp32 samples (U[0,1)) @ %d clusters per GPU (BASELINE configs[1] shape)"
IMPORT = 3


def _rank_info():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs a few hundred ms to start, longer than the default timed region, so it is started early
    (`start`, then `wait_ready`), the caller brackets the timed region with `mark()` and only the samples whose
    timestamps fall inside the marks are used."""

    def __init__(self, index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.index = index
        self.marks = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=5.0):
        t = time.time()
        while self.proc and time.time() - t < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.02)
        return False

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons, all_sm = [], [], [], set(), []
        lo, hi = (self.marks[0] - 0.02, self.marks[1] + 0.02) if len(self.marks) >= 2 else (0.0, 1e18)
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    clk, cmax = float(f[1]), float(f[2])
                except ValueError:
                    continue
                all_sm.append(clk)
                if not (lo <= ts <= hi):
                    continue
                sm.append(clk)
                mx.append(cmax)
                try:
                    pw.append(float(f[3]))
                except ValueError:
                    pass
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["reasons"] = sorted(reasons)
            out["samples"] = len(sm)
            if pw:
                out["power_w_max"] = max(pw)
        elif all_sm:
            out["note"] = "no sample fell inside the timed region (%d outside it)" % len(all_sm)
        return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_baseline(sample_points=65536):
    """the C oracle port (oracle/kmcuda_oracle.c, OpenMP) on a bounded sample of the same workload"""
    from oracle import oracle as O
    cores = O.set_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
    rng = np.random.default_rng(777)
    X = rng.random((sample_points, D), dtype=np.float32)
    C = rng.random((K, D), dtype=np.float32)
    O.assign_lloyd(X[:1024], C)  # warm the library / OpenMP pool
    t = time.perf_counter()
    O.assign_lloyd(X, C)
    dt = time.perf_counter() - t
    out = {"value": sample_points / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d points of the same %d-feature x %d-cluster workload, one pass, %.1f s" %
                     (sample_points, D, K, dt)}
    try:  # north_star's named CPU reference: sklearn KMeans labelling on the host cores
        from sklearn.cluster import KMeans
        km = KMeans(n_clusters=K, init=C, n_init=1, max_iter=1, algorithm="lloyd", tol=0).fit(X[:4096])
        km.cluster_centers_ = C.astype(km.cluster_centers_.dtype)
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=os.cpu_count()):
            km.predict(X[:4096])
            t = time.perf_counter()
            km.predict(X)
            dt2 = time.perf_counter() - t
        out["sklearn_predict"] = {"value": sample_points / dt2, "unit": UNIT, "cores": os.cpu_count(),
                                  "sample": "%d points, KMeans.predict" % sample_points}
    except Exception as e:  # pragma: no cover
        out["sklearn_predict"] = {"unavailable": repr(e)[:100]}
    return out


def c_api(path):
    from oracle import oracle as O
    return O.load_c_api(path)


def time_c_abi_host(lib, X_host, C_host, device_mask, steps, warmup):
    """kmeans_cuda(init=import, tolerance=1.0, yinyang_t=0): exactly one assignment pass, host buffers"""
    n = X_host.shape[0]
    A = np.empty(n, np.uint32)
    Cw = np.array(C_host, copy=True)
    m = ctypes.c_uint32(0)

    def call():
        rc = lib.kmeans_cuda(IMPORT, ctypes.byref(m), 1.0, 0.0, 0, n, D, K, 0, device_mask, -1, 0, 0,
                             X_host.ctypes.data, Cw.ctypes.data, A.ctypes.data, None)
        if rc != 0:
            raise RuntimeError("kmeans_cuda returned %d" % rc)

    for _ in range(warmup):
        call()
    t = time.perf_counter()
    for _ in range(steps):
        call()
    return (time.perf_counter() - t) / steps, A


def run_reference(args):
    rank, local, world = _rank_info()
    if rank != 0:
        return
    from oracle import oracle as O
    sample = min(args.points, 2000000)
    rng = np.random.default_rng(777)
    X = rng.random((sample, D), dtype=np.float32)
    C = X[rng.choice(sample, K, replace=False)].copy()
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (args.points, D, K), "l2": "inputs larger than L2"}}
    kind, note = "reference", ""
    try:
        import torch
        if not (O.reference_available() and torch.cuda.is_available()):
            raise RuntimeError("oracle/_ref/libKMCUDA.so or GPU missing")
        ref = O.reference_lib()
        dt, _ = time_c_abi_host(ref, X, C, 1 << local, args.steps, args.warmup)
        note = ("unmodified src-d/kmcuda rebuilt for sm_100 (oracle/_ref), one GPU, kmeans_cuda(import, tolerance=1, "
                "yinyang_t=0) with host buffers = H2D + transpose + one assign pass + D2H; bounded sample of %d points"
                % sample)
        cores = 0
    except Exception as e:
        kind = "port"
        sample = 65536
        X = X[:sample]
        t = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            O.assign_lloyd(X, C)
        dt = (time.perf_counter() - t) / max(1, min(args.steps, 3))
        note = "CPU oracle port (reference library unavailable: %s); %d points" % (repr(e)[:80], sample)
        cores = os.cpu_count()
    v = sample / dt
    line.update({"value": v, "ms_per_step": dt * 1e3,
                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": note},
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local, world = _rank_info()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import kmcuda_b200
    from kmcuda_b200.shard import Shard

    n = args.points
    g = torch.Generator(device="cuda").manual_seed(777 + rank)
    X = torch.rand((n, D), generator=g, device="cuda", dtype=torch.float32)
    gc = torch.Generator(device="cuda").manual_seed(777)
    C = torch.rand((K, D), generator=gc, device="cuda", dtype=torch.float32)  # same centroids on every rank
    sh = Shard(n, D, K)
    a = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    prev = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    changed = torch.zeros(1, dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        a.fill_(-1)
        sh.assign(X, C, a, prev, changed)
    barrier()
    if sh.last_error():
        raise RuntimeError("tensor-core pipeline error 0x%x" % sh.last_error())
    sampler.wait_ready()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        sh.assign(X, C, a, prev, changed)
    e1.record()
    barrier()
    sampler.mark()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
    tc_used, rechecked, overflowed = sh.last_pass_info()
    if not tc_used or sh.last_error():
        raise RuntimeError("the tensor-core path did not run cleanly (tc=%s err=0x%x)" % (tc_used, sh.last_error()))
    kt = sh.kernel_times(min(args.steps, 64))
    kernel_ms = sum(kt) / len(kt)

    if args.skip_extras:
        if rank == 0:
            emit({"metric": METRIC, "value": world * n / (ms_per_step * 1e-3), "unit": UNIT,
                  "ms_per_step": ms_per_step, "kernel_ms": kernel_ms, "note": "profiling run, extras skipped"})
        return
    # ---- end to end through the reference-facing C ABI with host buffers (pinned), rank-local shard
    e2e_steps = max(1, min(args.steps, 3))
    Xh = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
    Xh.copy_(X)
    Ch = C.cpu().numpy()
    del X
    torch.cuda.empty_cache()
    barrier()
    dt, A = time_c_abi_host(kmcuda_b200._lib, Xh.numpy(), Ch, 1 << local, e2e_steps, 1)
    te = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_dt = float(te.item())
    same = bool((torch.from_numpy(A.astype(np.int32)).cuda() == a).all().item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        flops = 2.0 * n * K * D
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * n / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 tensor-core filter (f32 accumulate) + f32 exact re-check",
            "data": "synthetic",
            "config": {"workload": WORKLOAD % (n, D, K), "parallelism": "sample-sharded x%d, no collective in the step" % world,
                       "l2": "inputs (%.1f GB per GPU) larger than L2, no flush needed" % (n * D * 4 / 1e9),
                       "rows_rechecked_exactly": rechecked, "rows_full_exact_fallback": overflowed},
            "e2e": {"value": world * n / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": n * D * 4 + K * D * 4,
                    "d2h_bytes_per_step": n * 4 + K * D * 4, "steps": e2e_steps,
                    "call": "kmeans_cuda(init=import, tolerance=1.0, yinyang_t=0) with pinned host buffers",
                    "equal_to_resident_result": same},
            "gpu_launches": args.steps * 9,
            "roofline": {"bound": "tensor", "kernel": "tc_assign_kernel", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                         "peak_source": "%s bf16_tflops (burst) of MEASURED_PEAKS.json; fp16 and bf16 share the tcgen05 rate" % peak_kind,
                         "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops,
                         "algorithmic_hbm_bytes_per_launch": n * (D * 4 + 4)},
            "clocks": clocks,
        }
        line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS, help="samples per GPU (default: the headline 8M)")
    ap.add_argument("--skip-extras", action="store_true", help="profiling runs: no e2e / cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
