"""CPU tests: the NumPy oracle against the naive C double loop and the golden
fixtures.  (The reference has no tests or vectors - README.md only - so these
pin the oracle itself; DESIGN.md says "parity unpinned".)"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import flat_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "flat_small.npz"))


def naive_search(lib, xb, xq, k, metric, self_ids=None, group_db=None, group_q=None):
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    nq, d = xq.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    sp = gp = gq = None
    if self_ids is not None:
        s64 = np.ascontiguousarray(self_ids, np.int64)
        sp = s64.ctypes.data_as(C.c_void_p)
    if group_db is not None:
        gdb = np.ascontiguousarray(group_db, np.int32)
        gqq = np.ascontiguousarray(group_q, np.int32)
        gp, gq = gdb.ctypes.data_as(C.c_void_p), gqq.ctypes.data_as(C.c_void_p)
    lib.naive_search.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.naive_search(xb.ctypes.data, xb.shape[0], xq.ctypes.data, nq, d, k, metric, sp, gp, gq, D.ctypes.data,
                          I.ctypes.data)
    assert rc == 0
    return D, I


def rand_unit(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("metric", [O.METRIC_IP, O.METRIC_L2])
@pytest.mark.parametrize("n,d,nq,k", [(64, 8, 5, 3), (300, 17, 9, 10), (50, 4, 3, 50), (10, 6, 4, 16)])
def test_oracle_matches_naive_loop(naive_lib, metric, n, d, nq, k):
    rng = np.random.default_rng(n * 31 + d)
    xb, xq = rand_unit(rng, n, d), rand_unit(rng, nq, d)
    D, I = O.search_ref(xb, xq, k, metric, block_rows=37)
    Dn, In = naive_search(naive_lib, xb, xq, k, metric)
    assert O.check_topk(D, I, Dn, In, tie_tol=1e-5, metric=metric) == 0
    fin = np.isfinite(Dn)
    assert np.array_equal(np.isfinite(D), fin)
    assert np.allclose(D[fin], Dn[fin], atol=2e-6)
    assert np.array_equal(I < 0, In < 0)


def test_tie_rule_lower_id_first(naive_lib):
    xb = np.zeros((12, 4), np.float32)
    xb[:, 0] = 1.0                      # every row identical: all scores tie
    xq = np.array([[1, 0, 0, 0]], np.float32)
    for metric in (O.METRIC_IP, O.METRIC_L2):
        D, I = O.search_ref(xb, xq, 5, metric, block_rows=5)
        assert I.tolist() == [[0, 1, 2, 3, 4]]
        Dn, In = naive_search(naive_lib, xb, xq, 5, metric)
        assert In.tolist() == [[0, 1, 2, 3, 4]]


def test_padding_when_k_exceeds_n():
    rng = np.random.default_rng(1)
    xb, xq = rand_unit(rng, 3, 8), rand_unit(rng, 2, 8)
    D, I = O.search_ref(xb, xq, 6, O.METRIC_IP)
    assert (I[:, 3:] == -1).all() and np.isneginf(D[:, 3:]).all()
    D, I = O.search_ref(xb, xq, 6, O.METRIC_L2)
    assert (I[:, 3:] == -1).all() and np.isposinf(D[:, 3:]).all()
    D, I = O.search_ref(np.zeros((0, 8), np.float32), xq, 4, O.METRIC_IP)
    assert (I == -1).all()


def test_l2_is_squared_and_ascending():
    rng = np.random.default_rng(2)
    xb, xq = rng.standard_normal((200, 12)).astype(np.float32), rng.standard_normal((7, 12)).astype(np.float32)
    D, I = O.search_ref(xb, xq, 9, O.METRIC_L2)
    assert (np.diff(D, axis=1) >= 0).all()
    direct = ((xq[:, None, :] - xb[I]) ** 2).sum(-1)
    assert np.allclose(D, direct, rtol=1e-4, atol=1e-4)


def test_exclusion_matches_naive(naive_lib):
    rng = np.random.default_rng(3)
    n, d, nq, k = 400, 16, 40, 8
    xb = rand_unit(rng, n, d)
    self_ids = rng.integers(0, n, nq)
    xq = xb[self_ids] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
    gdb = (np.arange(n) // 3).astype(np.int32)
    gq = gdb[self_ids].copy()
    gq[::3] = -1
    for metric in (O.METRIC_IP, O.METRIC_L2):
        D, I = O.search_ref(xb, xq, k, metric, self_ids=self_ids, group_db=gdb, group_q=gq, block_rows=150)
        Dn, In = naive_search(naive_lib, xb, xq, k, metric, self_ids, gdb, gq)
        assert O.check_topk(D, I, Dn, In, tie_tol=1e-5, metric=metric) == 0
        for i in range(nq):
            assert self_ids[i] not in I[i]
            if gq[i] >= 0:
                assert not (gdb[I[i][I[i] >= 0]] == gq[i]).any()


def test_golden_fixture_is_reproduced(naive_lib):
    xb, xq, k = GOLD["xb"], GOLD["xq"], int(GOLD["k"])
    for name, metric in (("ip", O.METRIC_IP), ("l2", O.METRIC_L2)):
        D, I = O.search_ref(xb, xq, k, metric)
        assert np.array_equal(I, GOLD[f"I_{name}"]) and np.allclose(D, GOLD[f"D_{name}"], atol=1e-6)
        Dn, In = naive_search(naive_lib, xb, xq, k, metric)
        assert O.check_topk(GOLD[f"D_{name}"], GOLD[f"I_{name}"], Dn, In, tie_tol=1e-5, metric=metric) == 0
        D, I = O.search_ref(xb, xq, k, metric, self_ids=GOLD["self_ids"], group_db=GOLD["group_db"],
                            group_q=GOLD["group_q"])
        assert np.array_equal(I, GOLD[f"I_{name}_excl"])
    # duplicated rows 7 / 100 / 650 tie for query 3: lower id first
    assert GOLD["I_ip"][3][:3].tolist() == [7, 100, 650]


def test_merge_ref_equals_unsharded():
    rng = np.random.default_rng(4)
    xb, xq = rand_unit(rng, 1000, 24), rand_unit(rng, 30, 24)
    for metric in (O.METRIC_IP, O.METRIC_L2):
        D, I = O.search_ref(xb, xq, 10, metric)
        parts = []
        for lo, hi in ((0, 333), (333, 700), (700, 1000)):
            Dp, Ip = O.search_ref(xb[lo:hi], xq, 10, metric)
            parts.append((Dp, np.where(Ip >= 0, Ip + lo, -1)))
        Dm, Im = O.merge_ref([p[0] for p in parts], [p[1] for p in parts], 10, metric)
        assert np.array_equal(Im, I) and np.allclose(Dm, D, atol=1e-6)


def test_bf16_round_is_rne():
    x = np.array([1.0, 1.00390625, 1.0 + 2 ** -8 + 2 ** -9, -3.14159, 1e-30, 65504.0], np.float32)
    r = O.bf16_round(x)
    assert r[0] == 1.0
    assert r[1] == 1.0            # exactly half way between 1.0 and 1.0078125: ties to even
    assert r[2] == np.float32(1.0078125)
    assert np.all(np.abs(r - x) <= np.abs(x) * 2.0 ** -8)
    assert np.array_equal(O.bf16_bits_to_f32(O.bf16_bits(x)), r)


def test_kmeans_refs(naive_lib):
    xb, cent = GOLD["xb"], GOLD["km_centroids"]
    a, dist = O.kmeans_assign_ref(xb, cent, block_rows=97)
    assert np.array_equal(a, GOLD["km_assign"])
    n, d = xb.shape
    an = np.empty(n, np.int32)
    dn = np.empty(n, np.float32)
    naive_lib.naive_kmeans_assign.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    naive_lib.naive_kmeans_assign(np.ascontiguousarray(xb).ctypes.data, n, np.ascontiguousarray(cent).ctypes.data,
                                  cent.shape[0], d, an.ctypes.data, dn.ctypes.data)
    differ = a != an
    # a disagreement is only allowed where the two best distances tie
    assert np.all(np.abs(dist[differ] - dn[differ]) <= 1e-5)
    assert np.allclose(dist, dn, atol=1e-5)
    newc, counts, _ = O.kmeans_update_ref(xb, a, cent)
    assert counts.sum() == n and np.allclose(newc, GOLD["km_new_centroids"], atol=1e-6)
    j = int(np.argmax(counts))
    assert np.allclose(newc[j], xb[a == j].mean(0), atol=1e-6)


def test_synth_rows_chunk_invariance():
    a = O.synth_rows(1234, 0, 70000, 8)
    b = O.synth_rows(1234, 65000, 66000, 8)
    assert np.array_equal(a[65000:66000], b)
    assert np.allclose(np.linalg.norm(a[:100], axis=1), 1.0, atol=1e-5)


def test_ivf_oracle_consistency():
    rng = np.random.default_rng(6)
    n, d, nlist, nq, k = 800, 12, 9, 14, 6
    xb, xq = rand_unit(rng, n, d), rand_unit(rng, nq, d)
    cent = xb[:nlist].copy()
    for metric in (O.METRIC_IP, O.METRIC_L2):
        a = O.ivf_assign_ref(cent, xb, metric)
        assert a.min() >= 0 and a.max() < nlist
        # probing every list == brute force
        allp = np.tile(np.arange(nlist), (nq, 1))
        D, I = O.ivf_search_ref(xb, a, xq, k, allp, metric)
        D_ref, I_ref = O.search_ref(xb, xq, k, metric)
        assert np.array_equal(I, I_ref) and np.allclose(D, D_ref, atol=1e-6)
        # a single probed list only returns rows of that list, in oracle order
        probes = O.ivf_probe_ref(cent, xq, 2, metric)
        D, I = O.ivf_search_ref(xb, a, xq, k, probes, metric)
        for i in range(nq):
            got = I[i][I[i] >= 0]
            assert np.all(np.isin(a[got], probes[i]))
            member = np.nonzero(np.isin(a, probes[i]))[0]
            assert len(got) == min(k, len(member))
        # skipped probes (-1) and empty result padding
        D, I = O.ivf_search_ref(xb, a, xq[:2], k, np.full((2, 3), -1), metric)
        assert (I == -1).all()


def test_triplet_oracle_rules():
    I = np.array([[5, 6, 7, 8], [9, -1, -1, -1], [1, 2, 3, 4]], np.int64)
    D = np.array([[.9, .8, .7, .6], [.5, 0, 0, 0], [.4, .3, .2, .1]], np.float32)
    pos = np.array([6, 3, -1], np.int64)
    T = O.build_triplets_ref(D, I, pos, skip_top=0, per_anchor=2, metric=O.METRIC_IP, limit=0.85, anchor_base=10)
    assert T[0].tolist() == [[10, 6, 7], [10, 6, 8]]     # 5 is above the limit, 6 is the positive itself
    assert T[1].tolist() == [[11, 3, 9], [-1, -1, -1]]   # list ends at the first -1
    assert (T[2] == -1).all()                            # no positive -> no triplet


def test_kmeans_split_empty_ref_properties():
    """Empty clusters take half of the largest cluster; totals are conserved; order and tie rule are fixed."""
    rng = np.random.default_rng(5)
    K, d = 12, 6
    cent = rng.standard_normal((K, d)).astype(np.float32)
    counts = np.array([7, 0, 3, 9, 0, 9, 1, 0, 2, 2, 0, 5])
    c2, n2, n_split = O.kmeans_split_empty_ref(cent, counts, eps=1.0 / 1024)
    assert n_split == 4 and n2.sum() == counts.sum() and (n2 > 0).all()
    # cluster 1 takes half of cluster 3 (9 = first maximum): 4 / 5; cluster 4 then takes half of cluster 5 (9)
    assert n2[1] == 4 and n2[4] == 4 and n2[3] in (5, 2, 3) and n2[5] in (5, 2, 3)
    up, down = np.float32(1 + 1 / 1024), np.float32(1 - 1 / 1024)
    # cluster 7 takes half of cluster 0 (7 is the largest left): check the perturbation pattern on an untouched donor
    assert np.array_equal(c2[7][0::2], cent[0][0::2] * up) and np.array_equal(c2[7][1::2], cent[0][1::2] * down)
    assert np.array_equal(c2[0][0::2], cent[0][0::2] * down) and np.array_equal(c2[0][1::2], cent[0][1::2] * up)
    untouched = [2, 6, 8, 9, 11]
    assert np.array_equal(c2[untouched], cent[untouched]) and np.array_equal(n2[untouched], counts[untouched])
    # nothing to split from: all donors have < 2 points
    c3, n3, s3 = O.kmeans_split_empty_ref(cent, np.array([1, 0, 1, 0] + [1] * 8))
    assert s3 == 0 and np.array_equal(c3, cent)
    # no empties: identity
    c4, n4, s4 = O.kmeans_split_empty_ref(cent, np.arange(1, K + 1))
    assert s4 == 0 and np.array_equal(c4, cent) and np.array_equal(n4, np.arange(1, K + 1))
