"""CPU tests: the NumPy oracle against INDEPENDENT published implementations.

The reference ships no code, tests or vectors (README.md:1-2) and FAISS is not installable here, so the
oracle's conventions (BASELINE.json north_star: fp32 NumPy / FAISS-CPU IndexFlat + Kmeans semantics) are pinned
against the exact brute-force k-NN and Lloyd step of scikit-learn and SciPy's cdist:

  * against the committed fixture tests/golden/sklearn_pin.npz (written by make_golden_sklearn.py, which does
    not import oracle/), and
  * live, on fresh seeded inputs, when scikit-learn / SciPy import (they ship in this image).

Tolerances: the third-party results are float64; the oracle is fp32 sgemm, so ids may differ only where two
scores are within 1e-5 (north_star's tie rule) and distances agree to 2e-6 absolute.
"""
import os

import numpy as np
import pytest

from oracle import flat_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
PIN = np.load(os.path.join(HERE, "golden", "sklearn_pin.npz"))


def ids_equal_up_to_ties(I, I_ref, D, D_ref, tol=1e-5):
    return O.check_topk(D, I, D_ref, I_ref, tie_tol=tol) == 0


# ------------------------------------------------------------------ committed fixture (no sklearn needed)
def test_l2_search_matches_sklearn_and_scipy_fixture():
    xb, xq, k = PIN["xb"], PIN["xq"], int(PIN["k"])
    D, I = O.search_ref(xb, xq, k, O.METRIC_L2, block_rows=700)
    for name in ("skl", "scipy"):
        D_ref, I_ref = PIN[f"{name}_l2_D"], PIN[f"{name}_l2_I"]
        assert ids_equal_up_to_ties(I, I_ref, D, D_ref)
        assert np.array_equal(I, I_ref)          # no ties in this fixture: identical, not merely tie-equivalent
        assert np.allclose(D, D_ref, atol=2e-6)  # squared distances, as FAISS IndexFlatL2 returns them


def test_ip_search_matches_scipy_and_sklearn_cosine_fixture():
    xb, xq, k = PIN["xb"], PIN["xq"], int(PIN["k"])
    D, I = O.search_ref(xb, xq, k, O.METRIC_IP)
    n_sc = PIN["scipy_ip_I"].shape[0]
    assert np.array_equal(I[:n_sc], PIN["scipy_ip_I"])
    assert np.allclose(D[:n_sc], PIN["scipy_ip_D"], atol=2e-6)
    # cosine ranking == inner-product ranking only for unit rows; bf16 rounding moves norms by < 2^-9, so
    # compare after normalising the same way: oracle IP on re-normalised fp64 rows
    xbn = xb / np.linalg.norm(xb.astype(np.float64), axis=1, keepdims=True)
    xqn = xq / np.linalg.norm(xq.astype(np.float64), axis=1, keepdims=True)
    Dn, In = O.search_ref(xbn, xqn, k, O.METRIC_IP, dtype=np.float64)
    assert ids_equal_up_to_ties(In, PIN["skl_cos_I"], Dn, PIN["skl_cos_D"])
    assert np.allclose(Dn, PIN["skl_cos_D"], atol=2e-6)


def test_tie_rule_same_set_as_sklearn_ordered_by_id():
    xb, xq = PIN["tie_xb"], PIN["tie_xq"]
    D, I = O.search_ref(xb, xq, 4, O.METRIC_L2)
    skl = PIN["tie_skl_I"]
    # query 0 == rows 7, 50, 333 (three exact duplicates), query 1 == rows 9, 120
    assert list(I[0, :3]) == [7, 50, 333] and set(skl[0, :3]) == {7, 50, 333}
    assert list(I[1, :2]) == [9, 120] and set(skl[1, :2]) == {9, 120}
    assert np.array_equal(np.sort(I, axis=1), np.sort(skl, axis=1))
    assert np.allclose(D, PIN["tie_skl_D"], atol=2e-6)


def test_exclusion_matches_sklearn_on_the_filtered_matrix():
    D, I = O.search_ref(PIN["xb"], PIN["ex_xq"], int(PIN["k"]), O.METRIC_L2, self_ids=PIN["ex_self_ids"],
                        group_db=PIN["ex_group_db"], group_q=PIN["ex_group_q"])
    assert ids_equal_up_to_ties(I, PIN["ex_skl_l2_I"], D, PIN["ex_skl_l2_D"])
    assert np.array_equal(I, PIN["ex_skl_l2_I"])
    assert np.allclose(D, PIN["ex_skl_l2_D"], atol=2e-6)


def test_kmeans_step_matches_sklearn_fixture():
    xb, cent = PIN["xb"], PIN["km_centroids"]
    a, dist = O.kmeans_assign_ref(xb, cent)
    assert np.array_equal(a, PIN["km_skl_assign"])
    assert np.allclose(dist, PIN["km_skl_dist"], atol=3e-6)
    newc, counts, _ = O.kmeans_update_ref(xb, a, cent)
    assert np.array_equal(counts, PIN["km_skl_counts"])
    assert np.allclose(newc, PIN["km_skl_new_centroids"], atol=1e-6)


# ------------------------------------------------------------------ live, on fresh inputs
sk_neighbors = pytest.importorskip("sklearn.neighbors")


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("n,d,nq,k,seed", [(2000, 64, 50, 10, 1), (777, 33, 20, 50, 2), (5000, 128, 30, 1, 3),
                                           (300, 16, 10, 300, 4)])
def test_live_l2_against_sklearn_brute(n, d, nq, k, seed):
    rng = np.random.default_rng(seed)
    xb, xq = _unit(rng, n, d), _unit(rng, nq, d)
    nn = sk_neighbors.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(xb.astype(np.float64))
    dist, ind = nn.kneighbors(xq.astype(np.float64))
    D, I = O.search_ref(xb, xq, k, O.METRIC_L2, block_rows=613)
    assert ids_equal_up_to_ties(I, ind, D, dist ** 2)
    assert np.allclose(D, dist ** 2, atol=3e-6)


@pytest.mark.parametrize("n,d,nq,k,seed", [(3000, 48, 40, 10, 5), (900, 20, 25, 64, 6)])
def test_live_ip_against_scipy_float64(n, d, nq, k, seed):
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(seed)
    xb, xq = _unit(rng, n, d) * rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32), _unit(rng, nq, d)
    # inner product from SciPy's squared distances: q.x = (|q|^2 + |x|^2 - |q-x|^2) / 2, all float64
    dd = cdist(xq.astype(np.float64), xb.astype(np.float64), "sqeuclidean")
    ip = 0.5 * ((xq.astype(np.float64) ** 2).sum(1)[:, None] + (xb.astype(np.float64) ** 2).sum(1)[None, :] - dd)
    o = np.argsort(-ip, axis=1, kind="stable")[:, :k]
    D, I = O.search_ref(xb, xq, k, O.METRIC_IP)
    assert ids_equal_up_to_ties(I, o, D, np.take_along_axis(ip, o, 1))
    assert np.allclose(D, np.take_along_axis(ip, o, 1), atol=3e-6)


def test_live_kmeans_step_against_sklearn():
    from sklearn.cluster import KMeans
    from sklearn.metrics import pairwise_distances_argmin_min
    rng = np.random.default_rng(7)
    x = _unit(rng, 4000, 40)
    cent = x[rng.choice(4000, 32, replace=False)].copy()
    a_ref, d_ref = pairwise_distances_argmin_min(x.astype(np.float64), cent.astype(np.float64), metric="sqeuclidean")
    a, dist = O.kmeans_assign_ref(x, cent)
    assert np.array_equal(a, a_ref)
    assert np.allclose(dist, d_ref, atol=3e-6)
    km = KMeans(n_clusters=32, init=cent.astype(np.float64), n_init=1, max_iter=1, algorithm="lloyd", tol=0.0)
    km.fit(x.astype(np.float64))
    newc, counts, _ = O.kmeans_update_ref(x, a, cent)
    assert counts.min() > 0
    assert np.allclose(newc, km.cluster_centers_, atol=1e-6)


def test_vectorised_select_equals_per_row_select():
    """The block-maximum bound + batched lexsort must give what a plain per-row stable sort gives,
    including ties that straddle the cut, -inf (excluded) scores and short rows."""
    rng = np.random.default_rng(11)
    for n, k in ((5000, 10), (5000, 50), (1500, 7), (300, 20), (4096, 1)):
        s = rng.standard_normal((37, n)).astype(np.float32)
        s[:, ::7] = np.float32(0.25)            # many exact ties
        s[3] = 1.0                               # a whole row of ties
        s[5, : n - 3] = -np.inf                  # fewer than k finite scores
        v, i = O._select_topk(s, 100, k)
        for r in range(s.shape[0]):
            o = np.lexsort((np.arange(n), -s[r]))[:k]
            assert np.array_equal(i[r], o + 100)
            assert np.array_equal(v[r], s[r][o])
