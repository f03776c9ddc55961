"""CPU model of the tensor-core filter's error bound (kmcuda_b200/csrc/assign_tc.cu).

The CUDA kernel only FILTERS with fp16 operands; it is correct iff the margin it derives per row really bounds
|approximate score - exact score|, so that the reference's fp32 winner is always among the candidates.  This test
re-derives the same quantities in numpy (fp16 rounding = round-to-nearest-even as `__float2half_rn`, products
accumulated in fp32) on adversarially scaled random data and checks the two properties the design rests on:

  1. |acc - s^2 (x.c - |c|^2 / 2)| <= E for every (row, centroid), with E exactly as the epilogue computes it;
  2. the float64 arg-min centroid of every row lies within `margin = 2 E (1.001)` of the row's best score.

It needs no GPU: it validates the arithmetic argument, not the kernel's plumbing (the GPU parity tests do that).
"""
import numpy as np
import pytest

KB = 64


def _scale_for(cmax):
    """tc_prep_scale_kernel: s = 2^k with s * cmax in [32, 64)"""
    if not (cmax > 0):
        return 1.0
    _, e = np.frexp(np.float32(cmax))
    return float(np.ldexp(1.0, 6 - int(e)))


def _filter_model(X, C, centred=False, residual="measured"):
    """returns (acc [n][K] fp32, E [n], s, mu) following the kernel's prep + converter + MMA + bias step.
    centred: both operands relative to mu = mean of the centroids (round 2, L2 metric): the table holds
    fp16(s fl32(c - mu)), the converter produces fl32(s x - s mu) in one rounding, the bias is ||fl32(c - mu)||^2."""
    X = X.astype(np.float32)
    C = C.astype(np.float32)
    n, D = X.shape
    nkb = (D + KB - 1) // KB
    mu = C.astype(np.float64).mean(0).astype(np.float32) if centred else np.zeros(D, np.float32)
    if not np.isfinite(mu).all():
        mu = np.zeros(D, np.float32)
    Cc = (C - mu[None]).astype(np.float32)                                # fl32(c - mu): the table operand before scaling
    csq = (Cc.astype(np.float64) ** 2).sum(1).astype(np.float32)          # ||c - mu||^2 (double accumulation -> fp32)
    cmax_raw = np.sqrt(csq.max())
    s = np.float32(_scale_for(cmax_raw))
    cmax = np.float32(cmax_raw * s * 1.001)
    mun = np.float32(np.sqrt(((mu.astype(np.float64) * float(s)) ** 2).sum()) * 1.001)
    cs = (Cc * s).astype(np.float32)
    ch = cs.astype(np.float16)                                            # fp16 centroid table
    dcmax = np.float32(np.sqrt(((cs - ch.astype(np.float32)).astype(np.float64) ** 2).sum(1)).max() * 1.0001)
    # bias: -(s^2 ||c - mu||^2 / 2) as three fp16 terms
    h = (np.float32(-0.5) * ((s * csq).astype(np.float32) * s)).astype(np.float32)   # same order as the prep kernel
    b0 = h.astype(np.float16)
    r1 = (h - b0.astype(np.float32)).astype(np.float32)
    b1 = r1.astype(np.float16)
    b2 = (r1 - b1.astype(np.float32)).astype(np.float32).astype(np.float16)
    # converter: a = fma(x, s, -mu s) (one rounding; s = 2^k, so = s * fl32(x - mu)), fp16 RN, norms of the rounded
    # vector and of the residual
    a = ((X - mu[None]).astype(np.float32) * s).astype(np.float32)
    ah = a.astype(np.float16)
    if residual == "measured":        # MODE 1 / 2: the converter sums |x~|^2 and |a - x~|^2
        nx = np.sqrt((ah.astype(np.float64) ** 2).sum(1)) * 1.0001
        nd = np.sqrt(((a - ah.astype(np.float32)).astype(np.float64) ** 2).sum(1)) * 1.0001
    else:                             # MODE 0 / 3: both are bounded from |a|^2 (fp32 sum) alone, as the epilogue does
        with np.errstate(over="ignore"):
            na = np.sqrt((a.astype(np.float32) ** 2).sum(1, dtype=np.float32).astype(np.float64)) * 1.0001
        nd = na * 4.8834e-4 + np.sqrt(float(nkb * KB)) * 2.99e-8
        nx = na + nd
    # MMA: fp16 x fp16 products are exact in fp32; accumulation order inside the tensor core is unspecified ->
    # model it with fp32 accumulation (np.matmul on float32 inputs), the bound has an explicit term for it
    acc = ah.astype(np.float32) @ ch.astype(np.float32).T
    acc = (acc + b0.astype(np.float32)[None] + b1.astype(np.float32)[None] + b2.astype(np.float32)[None]).astype(np.float32)
    xn = nx + nd
    E = nx * dcmax + nd * cmax + nd * dcmax
    E = E + (nkb * KB + 16) * 2.4e-7 * nx * cmax
    xu, cu = xn + mun, cmax + mun                                         # norms of the uncentred vectors (upper bounds)
    E = E + 6.0e-7 * (cu * cu + xu * cu)                                  # MODE 0's allowance for the reference's own rounding
    return acc, E.astype(np.float64), float(s), mu


def _cases():
    rng = np.random.default_rng(12345)
    out = []
    for (n, d, k, kind) in [(300, 256, 64, "uniform"), (200, 100, 33, "normal"), (200, 480, 50, "offset"),
                            (300, 32, 16, "wide"), (100, 8, 5, "tiny"), (200, 64, 40, "huge"), (150, 256, 30, "blobs")]:
        if kind == "uniform":
            X = rng.random((n, d))
        elif kind == "normal":
            X = rng.standard_normal((n, d))
        elif kind == "offset":            # far from the origin relative to the spread: the hard case for the filter
            X = 50.0 + rng.standard_normal((n, d))
        elif kind == "wide":
            X = rng.standard_normal((n, d)) * (10.0 ** rng.integers(-6, 6, size=d))
        elif kind == "tiny":
            X = rng.standard_normal((n, d)) * 1e-20
        elif kind == "huge":
            X = rng.standard_normal((n, d)) * 1e15
        else:
            centers = rng.random((k, d))
            X = centers[rng.integers(0, k, n)] + 0.01 * rng.standard_normal((n, d))
        X = X.astype(np.float32)
        C = X[rng.choice(n, k, replace=False)] + (0.01 * np.abs(X).mean() * rng.standard_normal((k, d))).astype(np.float32)
        out.append(pytest.param(X, C.astype(np.float32), id="%s_%dx%d_k%d" % (kind, n, d, k)))
    return out


@pytest.mark.parametrize("residual", ["measured", "analytic"])
@pytest.mark.parametrize("centred", [False, True])
@pytest.mark.parametrize("X,C", _cases())
def test_margin_bounds_the_fp16_filter_error(X, C, centred, residual):
    acc, E, s, mu = _filter_model(X, C, centred, residual)
    if residual == "analytic":
        # rows with an element beyond the fp16 range take the exact pass (the kernel's `na < 65000` guard)
        a_inf = ~np.isfinite(acc).all(1)
        if a_inf.any():
            keep = ~a_inf
            X, acc, E = X[keep], acc[keep], E[keep]
    # exact score of the (centred) operands in real arithmetic; it differs from the uncentred score by a per-row
    # constant only, so its arg-max is the true nearest centroid
    Xd, Cd = X.astype(np.float64) - mu.astype(np.float64), C.astype(np.float64) - mu.astype(np.float64)
    exact = (s * s) * (Xd @ Cd.T - 0.5 * (Cd ** 2).sum(1)[None])
    true_best = (((X.astype(np.float64)[:, None, :] - C.astype(np.float64)[None]) ** 2).sum(-1)).argmin(1)
    gap_ok = exact[np.arange(len(X)), true_best] >= exact.max(1) - 1e-9 * np.abs(exact).max()
    assert gap_ok.all()                      # centring is a per-row shift: same winner
    err = np.abs(acc.astype(np.float64) - exact)
    assert np.isfinite(acc).all()
    worst = (err / E[:, None]).max()
    assert worst <= 1.0, "error exceeds the bound: %.3f x E" % worst
    # the bound is meant to be tight enough to be useful, not just safe (except when every operand rounds exactly)
    assert worst > 1e-4 or err.max() == 0
    # containment: the true nearest centroid is within the margin of the row's best approximate score
    margin = 2.0 * E * 1.001 + 1e-30
    best_true = exact.argmax(1)
    rows = np.arange(len(X))
    assert (acc[rows, best_true] >= acc.max(1) - margin).all()


def test_candidate_counts_are_small_on_the_benchmark_distribution():
    """U[0,1)^256 @ 1024 (the reference's benchmark data): most rows have ONE candidate, the rest a handful --
    what makes the exact re-check cheap (18 % of rows, 2.3 candidates each)"""
    rng = np.random.default_rng(777)
    X = rng.random((2000, 256), dtype=np.float32)
    C = X[rng.choice(2000, 1024, replace=False)].copy()
    C += (0.01 * rng.standard_normal(C.shape)).astype(np.float32)
    acc, E, _, _ = _filter_model(X, C)
    margin = 2.0 * E * 1.001
    cand = (acc >= (acc.max(1) - margin)[:, None]).sum(1)
    assert (cand >= 1).all()
    assert 0.05 < (cand > 1).mean() < 0.35
    assert cand.max() <= 16
    # round 2: centred operands (|x - mu| |c - mu| is 4x smaller than |x| |c| on this data) -> 4x narrower margin
    acc_c, E_c, _, _ = _filter_model(X, C, centred=True)
    cand_c = (acc_c >= (acc_c.max(1) - 2.0 * E_c * 1.001)[:, None]).sum(1)
    assert (cand_c >= 1).all()
    assert (cand_c > 1).mean() < 0.5 * (cand > 1).mean()


def _knn_model(X, C, assign):
    """MODE 2: for every query x and candidate y (cluster B = assign[y]) the epilogue's translation-invariant score
    g = (x - c_B)~ . (y - c_B)~ + bias(y) - s^2 |x - c_B|^2 / 2  and the margin E of the segment (x, B)"""
    X = X.astype(np.float32)
    C = C.astype(np.float32)
    n, D = X.shape
    nkb = (D + KB - 1) // KB
    Yc = (X - C[assign]).astype(np.float32)                                  # table rows, centred on their own cluster
    ysq = (Yc.astype(np.float64) ** 2).sum(1).astype(np.float32)
    yabs_raw = np.sqrt(((np.abs(X) + np.abs(C[assign])).astype(np.float64) ** 2).sum(1)).max()
    cmax_raw = np.sqrt(ysq.max())
    s = np.float32(_scale_for(cmax_raw))
    cmax = np.float32(cmax_raw * s * 1.001)
    ys = (Yc * s).astype(np.float32)
    yh = ys.astype(np.float16)
    dcmax = np.float32(np.sqrt(((ys - yh.astype(np.float32)).astype(np.float64) ** 2).sum(1)).max() * 1.0001)
    yabs = np.float32(yabs_raw * s * 1.0001)
    h = (np.float32(-0.5) * ((s * ysq).astype(np.float32) * s)).astype(np.float32)
    b0 = h.astype(np.float16)
    r1 = (h - b0.astype(np.float32)).astype(np.float32)
    b1 = r1.astype(np.float16)
    b2 = (r1 - b1.astype(np.float32)).astype(np.float32).astype(np.float16)
    bias = b0.astype(np.float32) + b1.astype(np.float32) + b2.astype(np.float32)
    g = np.empty((n, n), np.float32)
    E = np.empty((n, n), np.float64)
    for B in np.unique(assign):
        cols = np.nonzero(assign == B)[0]
        a = ((X - C[B]).astype(np.float32) * s).astype(np.float32)          # the converter's re-centred query
        ah = a.astype(np.float16)
        nx = np.sqrt((ah.astype(np.float64) ** 2).sum(1)) * 1.0001
        nd = np.sqrt(((a - ah.astype(np.float32)).astype(np.float64) ** 2).sum(1)) * 1.0001
        a2 = (a.astype(np.float64) ** 2).sum(1).astype(np.float32)          # Kahan sum in the kernel
        nraw = np.sqrt(((np.abs(X) + np.abs(C[B])).astype(np.float64) ** 2).sum(1)) * s
        goff = (np.float32(0.5) * a2).astype(np.float32)
        acc = (ah.astype(np.float32) @ yh[cols].astype(np.float32).T + bias[cols][None]).astype(np.float32)
        g[:, cols] = (acc - goff[:, None]).astype(np.float32)
        xn = nx + nd
        e = nx * dcmax + nd * cmax + nd * dcmax + (nkb * KB + 16) * 2.4e-7 * nx * cmax
        e = e + 2.0e-6 * (cmax * cmax + xn * cmax) + 2.0e-6 * xn * xn
        e = e + 1.2e-7 * (nraw * cmax + yabs * xn) + 4.8e-7 * (goff + xn * cmax)
        E[:, cols] = e[:, None]
    return g, E, float(s)


@pytest.mark.parametrize("kind", ["blobs_far_from_origin", "uniform", "offset_uniform"])
def test_knn_centred_score_error_is_bounded_and_resolves_tight_clusters(kind):
    rng = np.random.default_rng(len(kind))
    n, d, k = 600, 128, 12
    if kind == "blobs_far_from_origin":       # the case that defeats uncentred fp16 operands
        centers = 20.0 + rng.random((k, d))
        assign = rng.integers(0, k, n)
        X = (centers[assign] + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    else:
        X = rng.random((n, d)).astype(np.float32) + (100.0 if kind == "offset_uniform" else 0.0)
        X = X.astype(np.float32)
        centers = X[rng.choice(n, k, replace=False)]
        assign = ((X[:, None, :] - centers[None]) ** 2).sum(-1).argmin(1)
    C = np.stack([X[assign == c].mean(0) if (assign == c).any() else centers[c] for c in range(k)]).astype(np.float32)
    g, E, s = _knn_model(X, C, assign)
    Xd = X.astype(np.float64)
    d2 = ((Xd[:, None, :] - Xd[None]) ** 2).sum(-1)
    g_true = -0.5 * s * s * d2
    err = np.abs(g.astype(np.float64) - g_true)
    assert (err <= E).all(), (err / E).max()
    # usefulness: with margin 2E the 11 nearest (self included) pull in only a few extra candidates per query
    kk = 11
    order = np.sort(g, axis=1)[:, ::-1]
    thr = order[:, kk - 1] - 2.0 * 1.001 * E.max(1)
    extra = (g >= thr[:, None]).sum(1) - kk
    assert np.median(extra) <= 6, np.median(extra)
