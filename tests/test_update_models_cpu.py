"""CPU models of the update path's device-side algorithms (no GPU needed).  Each model restates the CUDA code step by
step, so an indexing mistake in the kernel's logic shows up here.

1. `simt_kernels.cu::warp_lower_bound`: the 33-ary warp search (32 probes per step) behind `segment_offsets_kernel`
   returns np.searchsorted(keys, c, 'left') and terminates for every interval size.
2. `simt_kernels.cu::cluster_sums_kernel` / `combine_partials_kernel`: CTA b owns the sorted positions
   [b * L, (b + 1) * L); every (chunk, cluster) run goes to slot b + c.  The slots are unique, inside the workspace of
   `update_partial_rows`, and the combine kernel's range [beg / L, (end - 1) / L] visits exactly the written runs of a
   cluster, in order.
3. `exchange.cu`: partial sums double-buffered by iteration parity with a "partial sums complete" flag per (rank, peer)
   and NO "reads complete" flag.  Under arbitrary interleavings of the ranks no rank ever reads a buffer that a peer has
   already overwritten for a later iteration.
4. One launch site for the warp-per-cluster kernel (source lint).
5. `assign_tc.cu::prep_grid_barrier`: the sense-reversing grid barrier of the one-launch preparation separates the phases,
   never deadlocks and leaves its two words reusable by the next launch / graph replay.
"""
import numpy as np
import pytest


# ---------------------------------------------------------------------------------------------- 1. warp_lower_bound
def warp_lower_bound(keys, c):
    n = len(keys)
    lo, hi = 0, n
    steps = 0
    while hi - lo > 32:
        s = hi - lo
        q = [lo + ((lane + 1) * s) // 33 for lane in range(32)]
        assert all(lo < x < hi for x in q) and all(q[i] < q[i + 1] for i in range(31))
        pred = [keys[x] >= c for x in q]
        t = pred.index(True) if any(pred) else 32
        q_prev = q[t - 1] if t > 0 else q[0]
        q_t = q[t] if t < 32 else q[31]
        new_hi = q_t if t < 32 else hi
        new_lo = q_prev + 1 if t > 0 else lo
        assert new_hi - new_lo < hi - lo          # progress
        lo, hi = new_lo, new_hi
        steps += 1
    for lane in range(32):
        p = lo + lane
        if p < hi and keys[p] >= c:
            return p, steps
    return hi, steps


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 34, 64, 65, 66, 1000, 1089, 35937, 100003])
def test_warp_lower_bound_equals_searchsorted(n):
    rng = np.random.default_rng(n)
    K = max(2, min(n, 300))
    keys = np.sort(rng.integers(0, K + 1, n))           # K = the "unassigned" key
    for c in list(range(0, K + 2)) if K < 50 else list(rng.integers(0, K + 2, 40)) + [0, K, K + 1]:
        got, steps = warp_lower_bound(keys.tolist(), int(c))
        assert got == int(np.searchsorted(keys, c, "left")), (n, c)
        assert steps <= 5                               # 33^4 > 10^6


def test_warp_lower_bound_degenerate_key_patterns():
    for keys in ([5] * 1000, [0] * 500 + [9] * 500, list(range(1000)), [0] * 999 + [7]):
        arr = np.array(keys)
        for c in (0, 1, 5, 6, 7, 9, 10, 999, 1000):
            assert warp_lower_bound(keys, c)[0] == int(np.searchsorted(arr, c, "left"))


# ------------------------------------------------------------------------------------- 2. chunked member sums layout
def chunk_runs(keys, offsets, K, L):
    """cluster_sums_kernel: the (slot, cluster, lo, hi) runs every CTA writes"""
    total = offsets[K]
    n = len(keys)
    runs = []
    for b in range((n + L - 1) // L):
        lo = b * L
        hi = min(total, lo + L)
        while lo < hi:
            c = keys[lo]
            e = min(hi, offsets[c + 1])
            assert e > lo
            runs.append((b + c, c, lo, e))
            lo = e
    return runs


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("L", [4, 16, 512])
def test_chunked_member_sums_slots_are_unique_and_combine_sees_every_run(seed, L):
    rng = np.random.default_rng(seed)
    K = int(rng.integers(2, 40))
    n = int(rng.integers(1, 3000))
    p = rng.random(K + 1) ** 3                      # skewed sizes, some clusters (nearly) empty
    p[rng.integers(0, K)] = 0.0
    keys = np.sort(rng.choice(K + 1, n, p=p / p.sum()))   # key K = unassigned
    offsets = np.searchsorted(keys, np.arange(K + 2), "left")   # offsets[K] = first unassigned position
    runs = chunk_runs(keys.tolist(), offsets.tolist(), K, L)
    slots = [r[0] for r in runs]
    assert len(set(slots)) == len(slots)
    assert all(0 <= s < (n + L - 1) // L + K for s in slots)    # update_partial_rows(n, K)
    covered = np.zeros(n, bool)
    for _, c, lo, hi in runs:
        assert np.all(keys[lo:hi] == c) and not covered[lo:hi].any()
        covered[lo:hi] = True
    assert np.array_equal(covered, keys < K)
    by_slot = {s: (c, lo, hi) for s, c, lo, hi in runs}
    for c in range(K):                                   # combine_partials_kernel
        beg, end = int(offsets[c]), int(offsets[c + 1])
        if end <= beg:
            continue
        pos = beg
        for b in range(beg // L, (end - 1) // L + 1):
            cc, lo, hi = by_slot[b + c]
            assert cc == c and lo == pos
            pos = hi
        assert pos == end


# ------------------------------------------------------------------------------------ 3. exchange without "reads done"
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("seed", range(5))
def test_double_buffered_exchange_never_reads_an_overwritten_buffer(world, seed):
    """Each rank runs, per iteration it = 1..T (stream order): write own partial buffer[it & 1] := it; store flag
    (it) into every peer's flag row; spin until the own flag row shows >= it for every peer; read every peer's
    buffer[it & 1].  A scheduler picks which rank advances by one micro-step; the value read must always be `it`."""
    rng = np.random.default_rng(seed * 10 + world)
    T = 12
    buf = [[0, 0] for _ in range(world)]                 # buf[rank][parity] = iteration whose sums it holds
    flags = [[0] * world for _ in range(world)]          # flags[owner][from]
    # micro-steps of one iteration: 0 = write, 1..world = signal peer (1 + p), world + 1 = wait, then world reads
    state = [(1, 0) for _ in range(world)]               # (iteration, micro-step)
    steps_per_it = 1 + world + 1 + world
    done = 0
    guard = 0
    while done < world:
        guard += 1
        assert guard < 10 ** 6
        r = int(rng.integers(0, world))
        it, ms = state[r]
        if it > T:
            continue
        if ms == 0:
            buf[r][it & 1] = it
        elif ms <= world:
            flags[ms - 1][r] = it
        elif ms == world + 1:
            if not all(flags[r][p] >= it for p in range(world)):
                continue                                   # still spinning
        else:
            p = ms - (world + 2)
            assert buf[p][it & 1] == it, (r, p, it, buf[p])
        ms += 1
        if ms == steps_per_it:
            it, ms = it + 1, 0
            if it > T:
                done += 1
        state[r] = (it, ms)


# ------------------------------------------------------------------------------------ 4. one launch site per grid contract
def test_warp_per_cluster_kernel_is_launched_through_its_helper_only():
    """`segment_offsets_kernel` needs one WARP per cluster.  A launch site that kept the old one-thread-per-cluster grid
    left most offsets unwritten and hung the strict update (round 2); every caller must go through
    `launch_segment_offsets`."""
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmcuda_b200", "csrc")
    sites = []
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            text = open(os.path.join(csrc, name)).read()
            sites += [(name, m.start()) for m in re.finditer(r"segment_offsets_kernel\s*<<<", text)]
    assert len(sites) == 1 and sites[0][0] == "simt_kernels.cu", sites
    text = open(os.path.join(csrc, "simt_kernels.cu")).read()
    helper = text.index("static void launch_segment_offsets(")
    assert helper < sites[0][1] < text.index("}", sites[0][1]) and text.count("launch_segment_offsets(") >= 4
    launch = text[sites[0][1]:text.index(";", sites[0][1])]
    assert "* 32" in launch.replace("*32", "* 32")           # (K + 1) warps


# ------------------------------------------------------------------------------------ 5. the preparation's grid barrier
@pytest.mark.parametrize("nblocks", [1, 2, 16, 148])
@pytest.mark.parametrize("seed", range(3))
def test_sense_reversing_grid_barrier_separates_phases_and_returns_to_zero(nblocks, seed):
    """assign_tc.cu::prep_grid_barrier, thread 0 of every CTA: g = gen; if atomicAdd(count, 1) == nblocks - 1 then
    count = 0, gen += 1 else spin until gen != g.  Three barriers per launch, several launches in a row (a CUDA-graph
    replay reuses the two words without any host-side reset): under arbitrary interleavings no CTA enters phase p + 1
    before every CTA has finished phase p, nobody deadlocks, and the words are back to (0, launches * 3)."""
    rng = np.random.default_rng(seed * 1000 + nblocks)
    count, gen = 0, 0
    launches, barriers = 3, 3
    total_phases = launches * (barriers + 1)
    # per CTA: phase index, micro-state: 0 = working in phase, 1 = read gen, 2 = arrived (spinning), 3 = past barrier
    phase = [0] * nblocks
    st = [0] * nblocks
    seen_gen = [0] * nblocks
    finished_phase = [0] * nblocks             # number of phases whose work is complete
    guard = 0
    while min(phase) < total_phases - 1 or any(s != 0 for s in st) or min(finished_phase) < total_phases:
        guard += 1
        assert guard < 4 * 10 ** 6, "deadlock"
        b = int(rng.integers(0, nblocks))
        if phase[b] == total_phases - 1 and st[b] == 0 and finished_phase[b] == total_phases:
            continue
        if st[b] == 0:                          # do the phase's work
            if finished_phase[b] == phase[b]:
                # entering / working in phase[b]: every CTA must have finished all earlier phases of this launch
                launch_first = (phase[b] // (barriers + 1)) * (barriers + 1)
                assert all(fp >= phase[b] or phase[b] == launch_first for fp in finished_phase), (phase, finished_phase)
                finished_phase[b] += 1
            if (phase[b] + 1) % (barriers + 1) == 0:
                # last phase of a launch: no barrier, the next launch starts when ALL CTAs are done (stream order)
                if phase[b] < total_phases - 1 and min(finished_phase) >= phase[b] + 1:
                    phase[b] += 1
                continue
            st[b] = 1
        elif st[b] == 1:
            seen_gen[b] = gen
            st[b] = 2
        elif st[b] == 2:
            count += 1
            if count == nblocks:
                count = 0
                gen += 1
                st[b] = 3
            else:
                st[b] = 4
        elif st[b] == 4:
            if gen != seen_gen[b]:
                st[b] = 3
        elif st[b] == 3:
            phase[b] += 1
            st[b] = 0
    assert count == 0 and gen == launches * barriers
