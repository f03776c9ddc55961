"""CPU tests of the oracle (oracle/kmcuda_oracle.c): arithmetic self-checks and the same statistical
pins the reference's own test-suite uses against scikit-learn (reference src/test.py:157-187,598-606)."""
import numpy as np
import pytest

from oracle import oracle as O


def blobs():
    """the reference's KmeansTests data (reference src/test.py:158-169)"""
    rng = np.random.RandomState(0)
    arr = np.empty((13000, 2), dtype=np.float32)
    arr[:2000] = rng.rand(2000, 2) + [0, 2]
    arr[2000:4000] = rng.rand(2000, 2) - [0, 2]
    arr[4000:6000] = rng.rand(2000, 2) + [2, 0]
    arr[6000:8000] = rng.rand(2000, 2) - [2, 0]
    arr[8000:10000] = rng.rand(2000, 2) - [2, 2]
    arr[10000:] = rng.rand(3000, 2) + [2, 2]
    return arr


def test_fma_rd_matches_fenv_rounding():
    L = O.lib()
    rng = np.random.default_rng(0)
    cases = []
    for i in range(20000):
        a, b, c = rng.standard_normal(3).astype(np.float32)
        if i % 7 == 0:
            c = np.float32(-(a * b))
        if i % 11 == 0:
            a = np.float32(a * 1e-20)
        if i % 13 == 0:
            c = np.float32(c * 1e20)
        if i % 17 == 0:
            b = np.float32(0)
        cases.append((a, b, c))
    cases += [(np.float32(3e38), np.float32(2), np.float32(0)), (np.float32(-3e38), np.float32(2), np.float32(0)),
              (np.float32(1e-45), np.float32(0.5), np.float32(0)), (np.float32(0), np.float32(0), np.float32(0)),
              (np.float32(1), np.float32(1), np.float32(-1)), (np.float32(np.inf), np.float32(1), np.float32(1))]
    for a, b, c in cases:
        r1 = np.float32(L.ko_fma_rd(a, b, c))
        r2 = np.float32(L.ko_fma_rd_fenv(a, b, c))
        assert r1.tobytes() == r2.tobytes() or (np.isnan(r1) and np.isnan(r2)), (a, b, c, r1, r2)


@pytest.mark.parametrize("n,d,k", [(500, 2, 7), (300, 33, 10), (400, 256, 64)])
def test_assign_matches_float64_truth_except_ties(n, d, k):
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((n, d)).astype(np.float32)
    C = X[rng.choice(n, k, replace=False)] + rng.standard_normal((k, d)).astype(np.float32) * 0.1
    a, prev, changed = O.assign_lloyd(X, C)
    arg, best, second = O.assign_truth(X, C)
    exempt = O.tie_exempt(best, second)
    assert changed == n
    assert (prev == 0xFFFFFFFF).all()
    assert ((a == arg) | exempt).all()


def test_assign_nan_semantics():
    """reference kmeans.cu:312,327-329,349-356: NaN first feature -> K; NaN centroid never wins;
    nothing wins -> assignment untouched"""
    rng = np.random.default_rng(3)
    X = rng.random((64, 8), dtype=np.float32)
    C = rng.random((5, 8), dtype=np.float32)
    X[3, 0] = np.nan
    X[9, 4] = np.nan          # NaN elsewhere: every score is NaN, nothing wins
    C[2, :] = np.nan
    start = np.full(64, 1234, np.uint32)
    a, prev, changed = O.assign_lloyd(X, C, assign=start)
    assert a[3] == 5
    assert a[9] == 1234 and prev[9] == 0xFFFFFFFF
    assert not (a == 2).any()
    ok = np.ones(64, bool)
    ok[[3, 9]] = False
    sc = ((X[ok, None, :] - C[None, [0, 1, 3, 4], :]) ** 2).sum(-1)
    assert np.array_equal(np.array([0, 1, 3, 4])[sc.argmin(1)], a[ok])


def test_ties_pick_lowest_index():
    X = np.zeros((4, 4), np.float32)
    C = np.ones((6, 4), np.float32)
    C[3] = 5
    a, _, _ = O.assign_lloyd(X, C)
    assert (a == 0).all()


def test_adjust_is_the_cluster_mean():
    rng = np.random.default_rng(5)
    X = rng.random((3000, 24), dtype=np.float32)
    C0 = X[:20].copy()
    a, prev, _ = O.assign_lloyd(X, C0)
    C1, cnt = O.adjust(X, C0, prev, a, np.zeros(20, np.uint32))
    for c in range(20):
        m = X[a == c]
        assert cnt[c] == len(m)
        np.testing.assert_allclose(C1[c], m.astype(np.float64).mean(0), rtol=2e-6)


def test_empty_cluster_becomes_nan():
    X = np.ones((10, 3), np.float32)
    C0 = np.array([[1, 1, 1], [50, 50, 50]], np.float32)
    a, prev, _ = O.assign_lloyd(X, C0)
    C1, cnt = O.adjust(X, C0, prev, a, np.zeros(2, np.uint32))
    assert cnt[1] == 0 and np.isnan(C1[1]).all()


def _sklearn_one_step_changed(X, centroids, assignments):
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=len(centroids), init=centroids, n_init=1, max_iter=1, algorithm="lloyd").fit(X)
    # labels of the first E-step from our centroids
    d = ((X[:, None, :].astype(np.float64) - centroids[None].astype(np.float64)) ** 2).sum(-1)
    return (d.argmin(1) != assignments).mean()


@pytest.mark.parametrize("yy", [0.0, 0.1])
def test_kmeans_run_validates_like_reference(yy):
    """reference src/test.py:176-183 `_validate`: one more Lloyd step from the returned centroids must
    change fewer than `tolerance` of the labels"""
    X = blobs()
    rng = np.random.default_rng(0)
    C0 = X[rng.choice(len(X), 50, replace=False)]
    C, a, lines = O.kmeans(X, C0, tolerance=0.01, yinyang_t=yy)
    assert lines >= 2
    assert not np.isnan(C).any()
    assert _sklearn_one_step_changed(X, C, a) < 0.01


def test_yinyang_equals_lloyd_up_to_ties():
    X = blobs()[::4]
    rng = np.random.default_rng(1)
    C0 = X[rng.choice(len(X), 40, replace=False)]
    Cl, al, _ = O.kmeans(X, C0, tolerance=0.0, yinyang_t=0.0, max_iter=60)
    Cy, ay, _ = O.kmeans(X, C0, tolerance=0.0, yinyang_t=0.25, max_iter=60)
    assert (al == ay).mean() > 0.995


def test_knn_matches_sklearn_exactly():
    """reference src/test.py:598-606: k=10 neighbours must equal sklearn's brute force"""
    from sklearn.neighbors import NearestNeighbors
    X = blobs()[::3]
    rng = np.random.default_rng(2)
    C0 = X[rng.choice(len(X), 30, replace=False)]
    C, a, _ = O.kmeans(X, C0, tolerance=0.01, yinyang_t=0.0)
    got, pairs = O.knn(10, X, C, a)
    nb = NearestNeighbors(n_neighbors=11).fit(X)
    dist, idx = nb.kneighbors(X)
    # drop self; compare as distance multisets to be robust to exact fp ties
    d_got = np.sqrt(((X[:, None, :].astype(np.float64) - X[got].astype(np.float64)) ** 2).sum(-1))
    assert np.allclose(d_got, dist[:, 1:], rtol=0, atol=1e-6)
    assert (np.diff(d_got, axis=1) >= -1e-7).all()
    assert 0 < pairs < len(X) ** 2
    assert (got == idx[:, 1:]).mean() > 0.995
