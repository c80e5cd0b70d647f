"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI
(ctypes -> libcvdb_b200.so), against the NumPy oracle on the same seeded inputs,
against the committed golden fixture, and - at sizes the oracle cannot reach -
through size-independent properties.

Tolerances (BASELINE.json north_star):
  * exact storage: indices identical to the fp32 oracle except for ties within 1e-5
  * bf16 storage : recall@k >= 0.99 against the fp32 oracle on the unrounded
                   inputs, distances within 1e-2 relative; against the oracle run on
                   the bf16-rounded inputs (isolates the kernel) indices identical
                   up to ties within 2e-5
"""
import os

import numpy as np
import pytest
import torch

from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "flat_small.npz"))
M = {"ip": O.METRIC_IP, "l2": O.METRIC_L2}


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def make_index(d, metric, storage):
    from cloudvectordb_b200 import IndexFlat
    return IndexFlat(d, metric, storage, device=0)


def assert_parity(D, I, D_ref, I_ref, metric, tie_tol, dtol=1e-4):
    assert D.shape == D_ref.shape and I.shape == I_ref.shape
    assert I.dtype == np.int64 and D.dtype == np.float32
    assert np.array_equal(I < 0, I_ref < 0)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=tie_tol, metric=M[metric]) == 0
    fin = np.isfinite(D_ref)
    assert np.array_equal(np.isfinite(D), fin)
    assert np.allclose(D[fin], D_ref[fin], atol=dtol, rtol=1e-4)
    pad = ~fin
    if pad.any():
        assert np.all(np.isneginf(D[pad]) if metric == "ip" else np.isposinf(D[pad]))


# ---- kernel error isolated: inputs already exactly representable in bf16 ----------------
@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("n,d,nq,k", [
    (1000, 64, 7, 10),        # one partial tile
    (5000, 768, 129, 10),     # two query tiles, 12 K blocks
    (3000, 100, 64, 1),       # d not a multiple of 64, top-1 path
    (3000, 100, 64, 50),
    (3000, 100, 33, 128),
    (4097, 384, 300, 10),     # one row past a tile edge
    (257, 16, 1, 5),          # single query
    (5, 32, 3, 10),           # k > n: padding
    (3000, 64, 40, 300),      # large k
])
def test_bf16_kernel_matches_oracle_on_rounded_inputs(metric, n, d, nq, k):
    rng = np.random.default_rng(n + d + nq + k)
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    idx = make_index(d, metric, "bf16")
    idx.add(xb)
    assert idx.ntotal == n and idx.d == d
    D, I = idx.search(xq, k)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("metric,n,d,nq,k", [
    ("ip", 5000, 768, 300, 10),    # K = 768: TMEM + shared-memory tail (variant 2), all-TMEM N=64 (variant 4)
    ("l2", 5000, 600, 300, 10),    # 10 K blocks: a partial tail
    ("l2", 7000, 384, 513, 1),     # K <= 512: 128-column accumulators, top-1 path
    ("ip", 3000, 100, 64, 50),
    ("l2", 300, 20, 5, 7),         # tiny: the peer CTA of a pair sees only out-of-bounds rows
    ("ip", 40000, 256, 700, 10),   # several slices and query tiles: shared thresholds in play
])
def test_every_kernel_variant_matches_oracle(variant, metric, n, d, nq, k):
    rng = np.random.default_rng(variant * 1000 + n)
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    idx = make_index(d, metric, "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant)
    assert idx.last_work()["variant"] == variant
    idx.close()
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [768, 770, 829])
def test_l2_with_13_k_blocks(variant, d):
    """d = 768 plus the three L2 norm columns needs a 13th K block (one MMA K-step of it)."""
    rng = np.random.default_rng(d)
    n, nq, k = 6000, 400, 10
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_L2)
    idx = make_index(d, "l2", "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant)
    if variant:
        assert idx.last_work()["variant"] == variant
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "l2", tie_tol=2e-5)


@pytest.mark.parametrize("variant", [1, 3])
def test_exact_storage_on_both_streaming_variants(variant):
    rng = np.random.default_rng(77)
    n, d, nq, k = 12000, 200, 150, 10
    xb, xq = unit_rows(rng, n, d), unit_rows(rng, nq, d)
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    idx = make_index(d, "ip", "exact")
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=1e-5, dtol=1e-5)


def test_threshold_sharing_is_result_neutral():
    """debug flag 4 turns the cross-slice threshold sharing off: same answer either way."""
    rng = np.random.default_rng(78)
    n, d, nq, k = 60000, 64, 300, 10
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    idx = make_index(d, "ip", "bf16")
    idx.add(xb)
    D0, I0 = idx.search(xq, k, force_slices=16)
    D1, I1 = idx.search(xq, k, force_slices=16, debug_flags=4)
    idx.close()
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)


@pytest.mark.parametrize("slices", [2, 3, 7, 40])
def test_database_slices_do_not_change_results(slices):
    rng = np.random.default_rng(slices)
    n, d, nq, k = 20000, 128, 200, 20
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    idx = make_index(d, "ip", "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k, force_slices=slices)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=2e-5)


def test_golden_fixture_all_variants():
    xb, xq, k = GOLD["xb"], GOLD["xq"], int(GOLD["k"])
    for metric in ("ip", "l2"):
        for storage in ("bf16", "exact"):
            idx = make_index(xb.shape[1], metric, storage)
            idx.add(xb)
            D, I = idx.search(xq, k)
            assert np.array_equal(I, GOLD[f"I_{metric}"]), (metric, storage)
            assert np.allclose(D, GOLD[f"D_{metric}"], atol=2e-6)
            idx.set_groups(GOLD["group_db"])
            D, I = idx.search(xq, k, self_ids=GOLD["self_ids"], group_q=GOLD["group_q"])
            assert np.array_equal(I, GOLD[f"I_{metric}_excl"]), (metric, storage, "excl")
            assert np.allclose(D, GOLD[f"D_{metric}_excl"], atol=2e-6)
            idx.close()


def test_ties_resolve_to_lower_id():
    d = 32
    xb = np.zeros((600, d), np.float32)
    xb[:, 0] = 1.0                      # all rows identical -> every score ties
    xq = np.zeros((3, d), np.float32)
    xq[:, 0] = 1.0
    for storage in ("bf16", "exact"):
        for metric in ("ip", "l2"):
            idx = make_index(d, metric, storage)
            idx.add(xb)
            for slices in (0, 3):
                D, I = idx.search(xq, 12, force_slices=slices)
                assert np.array_equal(I, np.tile(np.arange(12), (3, 1))), (storage, metric, slices)
            idx.close()


def test_adversarial_ascending_scores():
    """Rows sorted by increasing similarity: every row is an insert for the filter."""
    rng = np.random.default_rng(5)
    n, d, nq, k = 6000, 64, 130, 10
    xq = O.bf16_round(unit_rows(rng, nq, d))
    xb = O.bf16_round(unit_rows(rng, n, d))
    order = np.argsort(xb @ xq[0])
    xb = np.ascontiguousarray(xb[order])
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    idx = make_index(d, "ip", "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=2e-5)
    assert I[0, 0] == n - 1


def test_zero_vectors_and_empty_inputs():
    d = 48
    idx = make_index(d, "ip", "bf16")
    D, I = idx.search(np.ones((4, d), np.float32), 5)          # empty index
    assert (I == -1).all() and np.isneginf(D).all()
    idx.add(np.zeros((0, d), np.float32))                       # empty add
    assert idx.ntotal == 0
    idx.add(np.zeros((300, d), np.float32))                     # zero vectors: all scores 0
    D, I = idx.search(np.zeros((2, d), np.float32), 4)
    assert np.array_equal(I, np.tile(np.arange(4), (2, 1))) and (D == 0).all()
    D, I = idx.search(np.zeros((0, d), np.float32), 4)          # no queries
    assert D.shape == (0, 4) and I.shape == (0, 4)
    idx.reset()
    assert idx.ntotal == 0
    idx.close()


def test_incremental_add_and_bf16_input_and_device_tensors():
    rng = np.random.default_rng(11)
    n, d, nq, k = 9000, 96, 70, 10
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_L2)
    idx = make_index(d, "l2", "bf16")
    idx.add(xb[:1000])                                            # numpy f32, host
    idx.add(torch.from_numpy(xb[1000:4000]).to(torch.bfloat16))   # torch bf16, host
    idx.add(torch.from_numpy(xb[4000:]).cuda())                   # torch f32, device
    assert idx.ntotal == n
    D, I = idx.search(torch.from_numpy(xq).cuda().to(torch.bfloat16), k)
    assert D.is_cuda and I.dtype == torch.int64
    assert_parity(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, "l2", tie_tol=2e-5)
    D, I = idx.search(xq, k, id_base=1_000_000)
    assert np.array_equal(I - 1_000_000, I_ref)
    idx.close()


# ---- exact storage: fp32 fidelity --------------------------------------------------------
@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("n,d,nq,k", [(20000, 384, 256, 10), (9000, 100, 130, 10), (3000, 768, 50, 64)])
def test_exact_storage_matches_fp32_oracle(metric, n, d, nq, k):
    rng = np.random.default_rng(n + d)
    xb, xq = unit_rows(rng, n, d), unit_rows(rng, nq, d)           # NOT rounded
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    idx = make_index(d, metric, "exact")
    idx.add(xb)
    D, I = idx.search(xq, k)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=1e-5, dtol=1e-5)


def test_config1_exact_100k_x_384():
    """BASELINE.json configs[0]: exact top-10 IP, 100k x 384 fp32, 10k queries."""
    xb = O.synth_rows(1234, 0, 100_000, 384)
    xq = O.synth_rows(5678, 0, 10_000, 384)
    D_ref, I_ref = O.search_ref(xb, xq, 10, O.METRIC_IP)
    idx = make_index(384, "ip", "exact")
    idx.add(xb)
    D, I = idx.search(xq, 10)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=1e-5, dtol=1e-5)


# ---- bf16 storage against the fp32 oracle on UNROUNDED inputs (north_star tolerances) ---
def test_bf16_recall_against_fp32_oracle():
    xb = O.synth_rows(1234, 0, 200_000, 768)
    xq = O.synth_rows(5678, 0, 1000, 768)
    D_ref, I_ref = O.search_ref(xb, xq, 10, O.METRIC_IP)
    idx = make_index(768, "ip", "bf16")
    idx.add(xb)
    D, I = idx.search(xq, 10)
    idx.close()
    assert O.recall_at_k(I, I_ref) >= 0.99
    assert np.all(np.abs(D - D_ref) <= 1e-2 * np.abs(D_ref) + 1e-6)


# ---- exclusion (hard-negative mining) -----------------------------------------------------
@pytest.mark.parametrize("nq", [150, 37, 5])          # two query tiles / one partial tile split over the four epilogue warps
@pytest.mark.parametrize("metric,k", [("ip", 10), ("l2", 50)])
def test_self_and_group_exclusion(metric, k, nq):
    rng = np.random.default_rng(21)
    n, d = 6000, 96
    xb = O.bf16_round(unit_rows(rng, n, d))
    self_ids = rng.integers(0, n, nq).astype(np.int32)
    xq = xb[self_ids].copy()
    gdb = (np.arange(n) // 4).astype(np.int32)
    gq = gdb[self_ids].copy()
    gq[::5] = -1
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric], self_ids=self_ids, group_db=gdb, group_q=gq)
    idx = make_index(d, metric, "bf16")
    idx.add(xb)
    idx.set_groups(gdb)
    D, I = idx.search(xq, k, self_ids=self_ids, group_q=gq)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5)
    for i in range(nq):
        assert self_ids[i] not in I[i]


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [1, 5, 20, 100, 300])
def test_group_exclusion_with_large_groups_of_near_duplicates(k, variant):
    """Every query has 40 near-duplicates in its own group: they outscore everything else, fill the candidate
    buffers (the group check runs when a buffer is sorted, not per row) and must all be gone from the result.
    Several database slices, every buffer size (k = 5 / 20 / 100 / 300), all kernel variants."""
    rng = np.random.default_rng(100 + k)
    d, n_groups, per = 64, 60, 40
    base = unit_rows(rng, n_groups, d)
    dup = base[:, None, :] + 0.02 * rng.standard_normal((n_groups, per, d), dtype=np.float32)
    filler = unit_rows(rng, 9000, d)
    xb = np.concatenate([dup.reshape(-1, d), filler]).astype(np.float32)
    gdb = np.concatenate([np.repeat(np.arange(n_groups), per), n_groups + np.arange(9000) // 3]).astype(np.int32)
    perm = rng.permutation(len(xb))
    xb, gdb = O.bf16_round(xb[perm]), gdb[perm]
    self_ids = np.flatnonzero(gdb < n_groups)[::7].astype(np.int32)        # queries = some of the duplicates
    xq, gq = xb[self_ids].copy(), gdb[self_ids].copy()
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP, self_ids=self_ids, group_db=gdb, group_q=gq)
    idx = make_index(d, "ip", "bf16")
    idx.add(xb)
    idx.set_groups(gdb)
    D, I = idx.search(xq, k, self_ids=self_ids, group_q=gq, force_variant=variant, force_slices=3)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=2e-5)
    assert not np.any(gdb[I] == gq[:, None])


def test_mine_hard_negatives_self_join():
    from cloudvectordb_b200 import mine_hard_negatives
    rng = np.random.default_rng(22)
    n, d, k = 5000, 64, 50
    emb = O.bf16_round(unit_rows(rng, n, d))
    groups = (np.arange(n) // 4).astype(np.int32)
    D, I = mine_hard_negatives(emb, k, groups, chunk=1536)
    D_ref, I_ref = O.search_ref(emb, emb, k, O.METRIC_IP, self_ids=np.arange(n), group_db=groups, group_q=groups)
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=2e-5)


# ---- shard / merge ---------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_merge_topk_equals_unsharded(metric):
    from cloudvectordb_b200 import merge_topk
    rng = np.random.default_rng(31)
    n, d, nq, k = 30000, 64, 257, 10
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    full = make_index(d, metric, "bf16")
    full.add(xb)
    D_full, I_full = full.search(xq, k)
    full.close()
    Ds, Is = [], []
    for lo, hi in ((0, 9000), (9000, 9001), (9001, 30000)):
        part = make_index(d, metric, "bf16")
        part.add(xb[lo:hi])
        Dp, Ip = part.search(xq, k, id_base=lo)
        part.close()
        Ds.append(Dp)
        Is.append(Ip)
    Dm, Im = merge_topk(np.stack(Ds), np.stack(Is), k, metric)          # host path
    assert np.array_equal(Im, I_full) and np.array_equal(Dm, D_full)    # merge is pure selection: bit-equal
    Dg, Ig = merge_topk(torch.from_numpy(np.stack(Ds)).cuda(), torch.from_numpy(np.stack(Is)).cuda(), k, metric)
    assert np.array_equal(Ig.cpu().numpy(), I_full) and np.array_equal(Dg.cpu().numpy(), D_full)
    Dr, Ir = O.merge_ref(Ds, Is, k, M[metric])
    assert np.array_equal(Im, Ir)


# ---- k-means ----------------------------------------------------------------------------------
def test_kmeans_step_matches_oracle_and_golden():
    from cloudvectordb_b200 import IndexFlat, Kmeans
    xb, cent = GOLD["xb"], GOLD["km_centroids"]
    km = Kmeans(xb.shape[1], cent.shape[0], niter=1, storage="exact", device=0)
    km.train(xb, init_centroids=cent)
    quantizer = IndexFlat(xb.shape[1], "l2", "exact", 0)       # the assignment on its own
    quantizer.add(cent)
    assign, dist = quantizer.assign(xb)
    quantizer.close()
    differ = assign != GOLD["km_assign"]
    assert np.all(np.abs(dist[differ] - GOLD["km_dist"][differ]) <= 1e-5)   # only genuine ties may differ
    assert np.allclose(dist, GOLD["km_dist"], atol=1e-5)
    assert np.array_equal(km.last_counts.cpu().numpy(), GOLD["km_counts"])
    assert np.allclose(km.centroids.cpu().numpy(), GOLD["km_new_centroids"], atol=1e-5)


def test_kmeans_bf16_assign_and_update_large():
    from cloudvectordb_b200 import Kmeans
    rng = np.random.default_rng(41)
    n, d, K = 50_000, 384, 1024
    x = O.bf16_round(unit_rows(rng, n, d))
    cent = O.bf16_round(x[rng.choice(n, K, replace=False)])
    km = Kmeans(d, K, niter=1, storage="bf16", device=0)
    xt = torch.from_numpy(x).cuda().to(torch.bfloat16)
    km.centroids = torch.from_numpy(cent).cuda()
    assign, obj = km.step(xt)
    a_ref, dist_ref = O.kmeans_assign_ref(x, cent)
    assign = assign.cpu().numpy()
    differ = assign != a_ref
    if differ.any():   # only genuine near-ties may differ
        d_alt = ((x[differ] - cent[assign[differ]]) ** 2).sum(1)
        assert np.all(np.abs(d_alt - dist_ref[differ]) <= 2e-5)
    assert differ.mean() < 1e-3
    newc_ref, counts_ref, _ = O.kmeans_update_ref(x, assign, cent)
    assert np.array_equal(km.last_counts.cpu().numpy(), counts_ref)
    assert np.allclose(km.centroids.cpu().numpy(), newc_ref, rtol=1e-3, atol=1e-5)
    assert abs(float(obj) - float(dist_ref.sum())) <= 1e-3 * float(dist_ref.sum())
    # objective never increases over Lloyd iterations
    km2 = Kmeans(d, 64, niter=4, storage="bf16", device=0)
    km2.train(xt)
    assert all(b <= a * (1 + 1e-5) for a, b in zip(km2.obj, km2.obj[1:]))


@pytest.mark.parametrize("K,d,n_empty", [(16, 8, 3), (1500, 100, 200), (70000, 32, 5000), (64, 384, 63), (5, 3, 0)])
def test_kmeans_split_empty_matches_oracle(K, d, n_empty):
    """cvdb_kmeans_split_empty against the oracle restatement: bit-exact centroids, counts and split count."""
    from cloudvectordb_b200 import _C
    rng = np.random.default_rng(K + d)
    cent = rng.standard_normal((K, d)).astype(np.float32)
    counts = rng.integers(1, 50, K).astype(np.int32)
    counts[rng.integers(0, K, 5)] = 49                       # ties for the largest cluster
    counts[rng.choice(K, n_empty, replace=False)] = 0
    if K == 64:
        counts[counts > 0] = 40                              # a single donor is split again and again, then runs dry
    c_ref, n_ref, s_ref = O.kmeans_split_empty_ref(cent, counts)
    ct, nt = torch.from_numpy(cent).cuda(), torch.from_numpy(counts).cuda()
    ns = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    _C.check(_C.lib().cvdb_kmeans_split_empty(ct.data_ptr(), nt.data_ptr(), K, d, 1.0 / 1024, ns.data_ptr(),
                                              int(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert int(ns.item()) == s_ref
    assert np.array_equal(nt.cpu().numpy().astype(np.int64), n_ref)
    assert np.array_equal(ct.cpu().numpy(), c_ref)


def test_kmeans_reseeds_empty_clusters():
    """Two identical initial centroids leave one cluster empty after the first assignment (ties go to the lower
    id); with split_empty it is re-seeded from the largest cluster, without it the old centroid stays."""
    from cloudvectordb_b200 import Kmeans
    rng = np.random.default_rng(8)
    n, d, K = 4000, 32, 8
    x = O.bf16_round(unit_rows(rng, n, d))
    init = x[:K].copy()
    init[5] = init[2]
    for split in (True, False):
        km = Kmeans(d, K, niter=1, storage="exact", device=0, split_empty=split)
        km.train(x, init_centroids=init)
        a_ref, _ = O.kmeans_assign_ref(x, init)
        newc, counts, _ = O.kmeans_update_ref(x, a_ref, init)
        assert counts[5] == 0
        if split:
            newc, counts, s = O.kmeans_split_empty_ref(newc, counts)
            assert s == 1 and int(km.last_nsplit.item()) == 1
        assert np.array_equal(km.last_counts.cpu().numpy().astype(np.int64), counts)
        assert np.allclose(km.centroids.cpu().numpy(), newc, rtol=1e-4, atol=1e-6)
    km = Kmeans(d, K, niter=6, storage="exact", device=0)
    km.train(x, init_centroids=init)
    assert int((km.last_counts > 0).sum().item()) == K          # no cluster stays empty


# ---- full-size properties (sizes the oracle cannot reach) ----------------------------------
def test_large_database_properties():
    """2M x 768 bf16: planted neighbours are found, results are sorted, a
    sharded run merges to the identical answer, and a torch fp32 matmul on a
    query subsample agrees."""
    from cloudvectordb_b200 import merge_topk
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    n, d, nq, k = 2_000_000, 768, 2048, 10
    xb = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for r0 in range(0, n, 1 << 19):
        r1 = min(n, r0 + (1 << 19))
        xb[r0:r1] = torch.nn.functional.normalize(torch.randn((r1 - r0, d), generator=g, device=dev), dim=1).bfloat16()
    planted = torch.randint(0, n, (nq,), generator=g, device=dev)
    xq = xb[planted].clone()                       # query i is exactly database row planted[i]
    idx = make_index(d, "ip", "bf16")
    idx.reserve(n)
    idx.add(xb)
    D, I = idx.search(xq, k)
    # (1) the planted row wins (its score is |x|^2 ~ 1, everything else ~ 0.1)
    assert torch.equal(I[:, 0], planted) or bool(((I[:, 0] == planted) | (D[:, 0] == D[:, 1])).all())
    # (2) sorted descending, ids valid and unique per query
    assert bool((D[:, 1:] <= D[:, :-1]).all()) and bool((I >= 0).all()) and bool((I < n).all())
    assert all(len(set(r.tolist())) == k for r in I[:64].cpu())
    # (3) torch fp32 reference on a query subsample
    qs = xq[:32].float()
    s = torch.cat([qs @ xb[r0:r0 + (1 << 19)].float().T for r0 in range(0, n, 1 << 19)], dim=1)
    v_ref, i_ref = torch.topk(s, k, dim=1)
    assert O.check_topk(D[:32].cpu().numpy(), I[:32].cpu().numpy(), v_ref.cpu().numpy(), i_ref.cpu().numpy(),
                        tie_tol=2e-5) == 0
    # (4) idempotence and shard/merge equality
    D2, I2 = idx.search(xq, k)
    assert torch.equal(I, I2) and torch.equal(D, D2)
    idx.close()
    Ds, Is = [], []
    for lo, hi in ((0, 700_000), (700_000, 2_000_000)):
        part = make_index(d, "ip", "bf16")
        part.add(xb[lo:hi])
        Dp, Ip = part.search(xq, k, id_base=lo)
        part.close()
        Ds.append(Dp)
        Is.append(Ip)
    Dm, Im = merge_topk(torch.stack(Ds), torch.stack(Is), k, "ip")
    assert torch.equal(Im, I) and torch.equal(Dm, D)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("k", [249, 504, 505, 1016, 1017, 2048])
def test_large_k_uses_the_in_memory_sort(k, variant):
    """k > 248: candidate buffers of 1024..4096 keys compacted in place (L2) instead of in registers."""
    rng = np.random.default_rng(k)
    n, d, nq = 6000, 48, 270
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    xb[4000] = xb[17]                                   # a tie somewhere in the list
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_L2)
    idx = make_index(d, "l2", "bf16")
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant, force_slices=3)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, "l2", tie_tol=2e-5)


def test_fp16_inputs():
    """float16 rows / queries (numpy or torch, host or device) are converted on the way in."""
    rng = np.random.default_rng(92)
    n, d, nq, k = 4000, 72, 90, 10
    xb16 = unit_rows(rng, n, d).astype(np.float16)
    xq16 = unit_rows(rng, nq, d).astype(np.float16)
    # what the engine stores: fp16 value rounded to bf16
    xb, xq = O.bf16_round(xb16.astype(np.float32)), O.bf16_round(xq16.astype(np.float32))
    for metric in ("ip", "l2"):
        D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
        idx = make_index(d, metric, "bf16")
        idx.add(xb16[:1500])                                        # numpy fp16, host
        idx.add(torch.from_numpy(xb16[1500:]).cuda())               # torch fp16, device
        D, I = idx.search(xq16, k)
        assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5)
        D, I = idx.search(torch.from_numpy(xq16).cuda(), k)
        assert_parity(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, metric, tie_tol=2e-5)
        idx.close()
    ex = make_index(d, "ip", "exact")                               # exact storage keeps the fp16 values exactly
    ex.add(xb16)
    D, I = ex.search(xq16, k)
    ex.close()
    D_ref, I_ref = O.search_ref(xb16.astype(np.float32), xq16.astype(np.float32), k, O.METRIC_IP)
    assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=1e-5, dtol=1e-5)


def test_query_batches_larger_than_one_launch_are_chunked():
    """More than 65 536 queries (262 144 for k = 1) go through several launches inside one call."""
    rng = np.random.default_rng(91)
    n, d = 3000, 32
    xb = O.bf16_round(unit_rows(rng, n, d))
    for nq, k in ((70_000, 5), (270_000, 1)):
        xq = O.bf16_round(unit_rows(rng, nq, d))
        D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
        idx = make_index(d, "ip", "bf16")
        idx.add(xb)
        D, I = idx.search(xq, k)
        if k == 1:
            A, dist = idx.assign(torch.from_numpy(xq).cuda())
            assert np.array_equal(A.cpu().numpy(), I[:, 0])
        idx.close()
        assert_parity(D, I, D_ref, I_ref, "ip", tie_tol=2e-5)


def test_errors_are_reported_not_swallowed():
    from cloudvectordb_b200 import _C
    idx = make_index(16, "ip", "bf16")
    idx.add(np.ones((10, 16), np.float32))
    with pytest.raises(_C.CvdbError):
        idx.search(np.ones((2, 16), np.float32), 0)
    with pytest.raises(_C.CvdbError):
        idx.search(np.ones((2, 16), np.float32), _C.MAX_K + 1)
    with pytest.raises(ValueError):
        idx.search(np.ones((2, 8), np.float32), 3)
    with pytest.raises(_C.CvdbError):
        idx.search(np.ones((2, 16), np.float32), 3, group_q=np.zeros(2, np.int32))   # no groups set
    idx.close()


@pytest.mark.parametrize("storage", ["bf16", "exact"])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_nonfinite_rows_are_counted_and_rejected_on_request(metric, storage):
    """SURVEY.md 4.2 adversarial row "NaN/Inf rejection": add()/search() with check_finite=True raise and leave the
    index as it was; without it the rows are counted (cvdb_index_nonfinite_rows) and nothing else changes."""
    rng = np.random.default_rng(77)
    d, k = 72, 5
    xb = O.bf16_round(unit_rows(rng, 700, d))
    xq = O.bf16_round(unit_rows(rng, 40, d))
    idx = make_index(d, metric, storage)
    idx.add(xb[:500], check_finite=True)
    assert idx.ntotal == 500 and idx.nonfinite_rows() == 0
    bad = xb[500:].copy()
    bad[3, 7] = np.nan
    bad[90, 0] = np.inf
    bad[91, d - 1] = -np.inf
    with pytest.raises(ValueError, match="3 of 200 rows"):
        idx.add(bad, check_finite=True)
    assert idx.ntotal == 500 and idx.nonfinite_rows() == 3
    with pytest.raises(ValueError, match="3 of 200 rows"):                    # same from device memory
        idx.add(torch.from_numpy(bad).cuda(), check_finite=True)
    assert idx.ntotal == 500 and idx.nonfinite_rows() == 6
    idx.add(xb[500:], check_finite=True)                                      # the clean batch goes in
    D, I = idx.search(xq, k, check_finite=True)
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5)
    qbad = xq.copy()
    qbad[11, 5] = np.nan
    with pytest.raises(ValueError, match="1 of 40 queries"):
        idx.search(qbad, k, check_finite=True)
    D2, I2 = idx.search(qbad, k)                                              # unchecked: other queries unaffected
    keep = np.arange(40) != 11
    assert np.array_equal(I2[keep], I[keep]) and np.array_equal(D2[keep], D[keep])
    assert idx.nonfinite_rows() == 8
    idx.truncate(600)
    assert idx.ntotal == 600
    D3, I3 = idx.search(xq, k)
    D_ref3, I_ref3 = O.search_ref(xb[:600], xq, k, M[metric])
    assert_parity(D3, I3, D_ref3, I_ref3, metric, tie_tol=2e-5)
    from cloudvectordb_b200 import _C
    with pytest.raises(_C.CvdbError):
        idx.truncate(601)
    idx.close()


# ---- seeded fuzz over shapes, k boundaries, metrics, storages and kernel variants ------------------------
def _fuzz_cases():
    rng = np.random.default_rng(20261018)
    edge_k = [1, 2, 12, 13, 28, 29, 60, 61, 124, 125, 200, 504]
    edge_n = [1, 127, 128, 129, 255, 256, 257, 511, 513, 1000, 4099]
    edge_q = [1, 2, 127, 128, 129, 255, 256, 257, 300, 513]
    edge_d = [1, 3, 8, 13, 61, 64, 65, 125, 128, 200, 384, 509, 512, 515, 765, 768, 829]
    cases = []
    for i in range(60):
        d = int(rng.choice(edge_d))
        n = int(rng.choice(edge_n)) if i % 3 else int(rng.integers(1, 9000))
        nq = int(rng.choice(edge_q)) if i % 2 else int(rng.integers(1, 700))
        k = int(rng.choice(edge_k))
        metric = "ip" if rng.random() < 0.5 else "l2"
        storage = "exact" if (i % 5 == 4 and d <= 256) else "bf16"
        variant = int(rng.choice([0, 1, 2, 3, 4])) if storage == "bf16" else int(rng.choice([0, 1, 3]))
        padded = d + (3 if metric == "l2" else 0)
        if variant == 4 and padded > 768:
            variant = 2
        if variant in (2, 4) and padded > 832:
            variant = 1
        cases.append((i, metric, storage, variant, n, d, nq, k))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}-v{c[3]}-n{c[4]}-d{c[5]}-q{c[6]}-k{c[7]}")
def test_fuzz_against_oracle(case):
    i, metric, storage, variant, n, d, nq, k = case
    rng = np.random.default_rng(1000 + i)
    xb, xq = unit_rows(rng, n, d), unit_rows(rng, nq, d)
    if storage == "bf16":
        xb, xq = O.bf16_round(xb), O.bf16_round(xq)
    if i % 4 == 0 and n > 8:            # duplicated rows: ties must resolve to the lower id
        xb[n // 2] = xb[1]
        xb[n - 1] = xb[1]
        xq[0] = xb[1]
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    idx = make_index(d, metric, storage)
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant, force_slices=(i % 7) if i % 2 else 0)
    idx.close()
    assert_parity(D, I, D_ref, I_ref, metric, tie_tol=2e-5 if storage == "bf16" else 1e-5)
