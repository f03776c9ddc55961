"""CPU models of two algorithmic claims the CUDA code relies on (no GPU needed).

1. Yinyang local step (kmcuda_b200/csrc/yinyang.cu): the reference scans the centroids in
   index order with a running (min, second) pair and a per-centroid skip test (kmeans.cu:620-668).  The library
   instead takes the two smallest elements of the multiset {ub} U {distances of all unpruned centroids} U {bounds
   of the pruned groups}.  With valid bounds both give the same (nearest, ub', bound of the nearest's group) --
   including ties.  Checked here on random instances built from consistent geometry (distances that moved by at
   most the drift), with integer distances to provoke ties.

2. k-NN threshold (assign_tc.cu MODE 2): the kk-th largest of per-bucket maxima over
   DISJOINT column buckets is a lower bound of the kk-th largest value, and it is reasonably tight when there are
   many more buckets than kk.
"""
import numpy as np
import pytest

FLT_MAX = np.float32(3.4028235e38)


def reference_scan(ub, a, lb, groups, drift, maxdrift, dist):
    """kmeans.cu:620-668 restated (also yinyang.cu::yy_local_scan_kernel)"""
    mn, sec, near = ub, FLT_MAX, a
    for c in range(len(groups)):
        if c == a:
            continue
        g = groups[c]
        b = lb[g]
        if b >= ub:
            if b < sec:
                sec = b
            continue
        b = b + maxdrift[g] - drift[c]
        if sec < b:
            continue
        d = dist[c]
        if d < mn:
            sec, mn, near = mn, d, c
        elif d < sec:
            sec = d
    return near, mn, sec


def two_smallest(ub, a, lb, groups, drift, maxdrift, dist):
    """yinyang.cu::yy_finish_kernel / yy_rows_cta_kernel: order independent"""
    G = len(lb)
    gsize = np.bincount(groups, minlength=G)
    p1 = FLT_MAX
    for g in range(G):
        if lb[g] >= ub and gsize[g] - (1 if g == groups[a] else 0) > 0:
            p1 = min(p1, lb[g])
    cand = [(dist[c], c) for c in range(len(groups)) if c != a and not (lb[groups[c]] >= ub)]
    if cand:
        dmin, cbest = min(cand)
    else:
        dmin, cbest = FLT_MAX, None
    moved = cbest is not None and dmin < ub
    near = cbest if moved else a
    mn = dmin if moved else ub
    rest = min([d for (d, c) in cand if not (moved and c == cbest)] + [p1])
    sec = min(ub, rest) if moved else rest
    return near, mn, sec


@pytest.mark.parametrize("seed", range(40))
def test_two_smallest_equals_reference_order_scan(seed):
    rng = np.random.default_rng(seed)
    K = int(rng.integers(3, 40))
    G = int(rng.integers(1, max(2, K // 3)))
    groups = rng.integers(0, G, K)
    levels = int(rng.integers(3, 30))                      # few distinct values -> many exact ties
    d_prev = rng.integers(1, levels + 1, K).astype(np.float32)
    drift = rng.integers(0, 3, K).astype(np.float32)
    move = np.array([rng.integers(-int(t), int(t) + 1) for t in drift], dtype=np.float32)
    dist = np.maximum(d_prev + move, 0).astype(np.float32)  # |dist - d_prev| <= drift
    a = int(rng.integers(0, K))
    ub = dist[a]                                            # tightened upper bound = exact distance to the own centroid
    maxdrift = np.array([drift[groups == g].max() if (groups == g).any() else 0 for g in range(G)], dtype=np.float32)
    lb = np.empty(G, np.float32)
    for g in range(G):
        members = [c for c in range(K) if groups[c] == g and c != a]
        true_min_prev = min([d_prev[c] for c in members]) if members else FLT_MAX
        slack = np.float32(rng.integers(0, 3))              # bounds are valid, not necessarily tight
        lb_old = np.float32(max(true_min_prev - slack, 0)) if members else np.float32(rng.integers(0, levels))
        lb[g] = lb_old - maxdrift[g]                        # the decayed bound the local step sees
    exp = reference_scan(ub, a, lb, groups, drift, maxdrift, dist)
    got = two_smallest(ub, a, lb, groups, drift, maxdrift, dist)
    assert got == exp, (got, exp)


@pytest.mark.parametrize("kk", [2, 4, 11, 16])
def test_bucket_maxima_give_a_valid_and_tight_knn_threshold(kk):
    rng = np.random.default_rng(kk)
    ranks = []
    for _ in range(200):
        ncols = int(rng.integers(3 * 128, 40 * 128))
        v = rng.standard_normal(ncols).astype(np.float32)
        # 64 disjoint buckets per row: (column half, block parity, 4-column group within the 64-column half)
        col = np.arange(ncols)
        blk, within = col // 128, col % 128
        bucket = (within // 64) * 32 + (blk % 2) * 16 + (within % 64) // 4
        bmax = np.full(64, -np.inf, np.float32)
        np.maximum.at(bmax, bucket, v)
        thr = np.sort(bmax)[::-1][kk - 1]
        kth = np.sort(v)[::-1][kk - 1]
        assert thr <= kth                                   # valid: kk distinct columns reach thr
        ranks.append(int((v >= thr).sum()))                 # how many columns the threshold lets through
    assert np.median(ranks) <= 2 * kk + 2                   # tight: about kk..2kk candidates before the margin


# ---------------------------------------------------------------------------------------------------------------
# 3. Yinyang bounds refresh on the tensor cores (assign_tc.cu MODE 3): the per-group maximum of the fp16 scores,
#    widened by the filter's error bound E, gives a VALID and TIGHT lower bound of the distance to the nearest
#    centroid of the group:  s^2 d^2 = |x^|^2 - 2 score  >=  |x^|^2 - 2 (max_g acc + E).
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["uniform", "offset", "blobs"])
def test_refresh_lower_bounds_from_group_maxima_are_valid_and_tight(kind):
    from test_margin_cpu import _filter_model
    rng = np.random.default_rng(99)
    n, d, k, G = 300, 96, 120, 12
    if kind == "uniform":
        X = rng.random((n, d))
    elif kind == "offset":
        X = 30.0 + rng.standard_normal((n, d))
    else:
        centers = rng.random((k, d))
        X = centers[rng.integers(0, k, n)] + 0.05 * rng.standard_normal((n, d))
    X = X.astype(np.float32)
    C = (X[rng.choice(n, k, replace=False)] + 0.01 * rng.standard_normal((k, d))).astype(np.float32)
    groups = rng.integers(0, G, k)
    acc, E, s, mu = _filter_model(X, C, centred=True)
    a = ((X - mu[None]).astype(np.float32) * np.float32(s)).astype(np.float32)
    xa2 = (a.astype(np.float32) ** 2).sum(1, dtype=np.float32)             # fp32 sum, as the converter's FFMA2 chain
    xa2lo = xa2 * np.float32(1.0 - 1.0e-4)
    Eb = (0.5 * (2.0 * E * 1.001 + 1e-30)).astype(np.float32)              # the kernel uses half of its margin (>= E)
    dist = np.sqrt(((X.astype(np.float64)[:, None, :] - C.astype(np.float64)[None]) ** 2).sum(-1))
    worst_slack = 0.0
    for g in range(G):
        cols = np.flatnonzero(groups == g)
        if len(cols) == 0:
            continue
        run = acc[:, cols].max(1)
        t = xa2lo - np.float32(2.0) * (run + Eb)
        lb = np.where(t > 0, np.sqrt(np.maximum(t, 0)) / np.float32(s) * np.float32(1.0 - 4.0e-6), 0.0)
        true_min = dist[:, cols].min(1)
        assert (lb <= true_min * (1 + 1e-7)).all(), float((lb - true_min).max())          # valid
        # slack in units of the typical centroid distance (the bound is absolute: next to a centroid -- the sample's
        # own one, which the kernel handles exactly -- it degenerates to 0)
        worst_slack = max(worst_slack, float((true_min - lb).max() / np.median(dist)))
    assert worst_slack < 0.05, worst_slack                                                # tight enough to prune with
