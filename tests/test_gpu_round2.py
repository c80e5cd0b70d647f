"""GPU tests (pytest -m gpu) of the round-2 host/ABI work: the key exchange format of the shard layer, the
state of an index whose rows were re-stored list-major, stream ordering on a shared handle, the zero-score
tie rule across slices, the pipelined host ingest, and the small edge cases the advisor listed."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu
M = {"ip": O.METRIC_IP, "l2": O.METRIC_L2}


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


# ---- shard exchange in keys --------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("k", [1, 10, 50])
def test_keys_exchange_equals_unsharded_search(metric, k):
    """search_keys on three shards (one of them EMPTY, one a single row) + merge_keys == one search of the whole
    matrix, bit for bit (the merge is pure selection), and == the oracle."""
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(31 + k)
    n, d, nq = 30000, 64, 257
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    xq_t = torch.from_numpy(xq).cuda()
    full = IndexFlat(d, metric, "bf16", 0)
    full.add(xb)
    D_full, I_full = full.search(xq_t, k)
    one = full.merge_keys(full.search_keys(xq_t, k)[None], k)          # a single list round-trips
    assert torch.equal(one[0], D_full) and torch.equal(one[1], I_full)
    keys = []
    shards = []
    for lo, hi in ((0, 9000), (9000, 9000), (9000, 9001), (9001, n)):
        part = IndexFlat(d, metric, "bf16", 0)
        if hi > lo:
            part.add(xb[lo:hi])
        keys.append(part.search_keys(xq_t, k, id_base=lo))
        shards.append(part)
    Dm, Im = shards[0].merge_keys(torch.stack(keys), k)
    assert torch.equal(Im, I_full) and torch.equal(Dm, D_full)
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    assert O.check_topk(Dm.cpu().numpy(), Im.cpu().numpy(), D_ref, I_ref, tie_tol=2e-5, metric=M[metric]) == 0
    # keys are sorted best-first as unsigned integers and carry the global id in the low word
    kk = keys[3].cpu().numpy().view(np.uint64)
    assert np.all(kk[:, :-1] >= kk[:, 1:])
    ids = (~kk & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert ids[kk != 0].min() >= 9001 and ids[kk != 0].max() < n
    for p in shards:
        p.close()
    full.close()


def test_keys_reject_ids_beyond_32_bits_and_host_pointers():
    from cloudvectordb_b200 import IndexFlat, _C
    idx = IndexFlat(16, "ip", "bf16", 0)
    idx.add(np.ones((4, 16), np.float32))
    q = torch.ones((2, 16), device="cuda")
    with pytest.raises(_C.CvdbError):
        idx.search_keys(q, 2, id_base=(1 << 32) - 3)
    with pytest.raises(ValueError):
        idx.search_keys(np.ones((2, 16), np.float32), 2)
    idx.close()


# ---- an index whose rows were re-stored list-major ------------------------------------------------------------
def test_flat_search_on_a_regrouped_index_returns_the_ids_rows_were_added_under():
    """After cvdb_index_group_by_list the rows are permuted in HBM.  A flat search / assign (before or after more
    rows are added) must still return insertion-order ids; what works on stored positions must refuse."""
    from cloudvectordb_b200 import IndexFlat, _C
    rng = np.random.default_rng(3)
    n, d, nlist, nq, k = 20000, 64, 50, 200, 10
    xb, xq = O.bf16_round(unit_rows(rng, n + 500, d)), O.bf16_round(unit_rows(rng, nq, d))
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(xb[:n])
    lists = torch.from_numpy(rng.integers(0, nlist, n).astype(np.int32)).cuda()
    st = int(torch.cuda.current_stream().cuda_stream)
    _C.check(_C.lib().cvdb_index_group_by_list(idx._h, lists.data_ptr(), nlist, st))
    D, I = idx.search(xq, k)                                   # grouped: flat search translates ids
    D_ref, I_ref = O.search_ref(xb[:n], xq, k, O.METRIC_IP)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    a, _ = idx.assign(xq)
    assert np.array_equal(a, I_ref[:, 0]) or O.check_topk(D[:, :1], a[:, None].astype(np.int64), D_ref[:, :1], I_ref[:, :1], 2e-5) == 0
    idx.add(xb[n:])                                            # un-grouped, still permuted
    D, I = idx.search(xq, k)
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    keys = idx.search_keys(torch.from_numpy(xq).cuda(), k, id_base=7)
    Dk, Ik = idx.merge_keys(keys[None], k)
    assert np.array_equal(Ik.cpu().numpy(), I + 7)
    with pytest.raises(_C.CvdbError, match="stored position"):
        idx.set_groups(np.zeros(n + 500, np.int32))
    with pytest.raises(_C.CvdbError, match="list-major"):
        idx.search(xq, k, self_ids=np.zeros(nq, np.int32))
    with pytest.raises(_C.CvdbError, match="list-major"):
        idx.save("/tmp/cvdb_should_not_exist.bin")
    with pytest.raises(_C.CvdbError):
        idx.truncate(n - 1)                                    # below the grouped rows
    idx.truncate(n)                                            # dropping the rows added since is fine
    D, I = idx.search(xq, k)
    D_ref, I_ref = O.search_ref(xb[:n], xq, k, O.METRIC_IP)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    idx.reset()                                                # back to a plain flat index
    idx.add(xb[:1000])
    idx.set_groups(np.arange(1000, dtype=np.int32) // 4)
    D, I = idx.search(xq, k, self_ids=np.zeros(nq, np.int32))
    assert not np.any(I == 0)
    idx.close()


# ---- one handle, two streams -------------------------------------------------------------------------------
def test_two_streams_on_one_handle_do_not_race_on_the_scratch():
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(5)
    n, d, k = 400_000, 128, 10
    xb = O.bf16_round(unit_rows(rng, n, d))
    q1, q2 = O.bf16_round(unit_rows(rng, 900, d)), O.bf16_round(unit_rows(rng, 300, d))
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(xb)
    t1, t2 = torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(4):                                       # back to back, alternating streams, no host sync
        with torch.cuda.stream(s1):
            outs.append(idx.search(t1, k))
        with torch.cuda.stream(s2):
            outs.append(idx.search(t2, k))
    torch.cuda.synchronize()
    D1, I1 = O.search_ref(xb, q1, k, O.METRIC_IP)
    D2, I2 = O.search_ref(xb, q2, k, O.METRIC_IP)
    for i, (D, I) in enumerate(outs):
        Dr, Ir = (D1, I1) if i % 2 == 0 else (D2, I2)
        assert O.check_topk(D.cpu().numpy(), I.cpu().numpy(), Dr, Ir, tie_tol=2e-5) == 0
    idx.close()


# ---- ties at exactly zero across slices (-0.0 / +0.0 are one score) ------------------------------------------
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("slices", [1, 3, 7])
def test_zero_scores_tie_by_lower_id_across_slices(variant, slices):
    """A zero query scores +0.0 against every row; a one-hot query scores exactly 0 against all rows that are
    orthogonal to it, and -0.0 where the row's entry is -0.0.  Lower id must win in every slice layout."""
    from cloudvectordb_b200 import IndexFlat
    n, d, k = 9000, 64, 12
    rng = np.random.default_rng(variant * 10 + slices)
    xb = O.bf16_round(unit_rows(rng, n, d))
    xb[:, 0] = 0.0
    xb[::3, 0] = -0.0                                           # 0 * 1 = -0.0 for these rows
    xb[5000:5006, 0] = 0.5                                      # six genuinely positive scores, late in the matrix
    xq = np.zeros((300, d), np.float32)
    xq[1:, 0] = 1.0                                             # query 0: all zero; the others: one-hot
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(xb)
    D, I = idx.search(xq, k, force_variant=variant, force_slices=slices)
    idx.close()
    assert np.array_equal(I[0], np.arange(k)) and np.all(D[0] == 0)
    want = np.concatenate([np.arange(5000, 5006), np.arange(k - 6)])
    assert np.array_equal(I[1:], np.broadcast_to(want, (299, k)))
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    assert np.array_equal(I, I_ref)


# ---- host ingest ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_pipelined_host_add_equals_device_add(dtype, tmp_path):
    """> 64 MB of host rows go through the pinned double-buffered pipeline (several chunks, ragged tail): the
    stored rows must equal those of a device-side add, from pageable memory, from pinned memory and from a file."""
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(8)
    n, d = 150_001, 384                                          # fp32: 230 MB = 3.4 chunks; bf16: 115 MB
    x = O.bf16_round(unit_rows(rng, n, d))
    xt = torch.from_numpy(x)
    if dtype == "bfloat16":
        xt = xt.to(torch.bfloat16)
    ref = IndexFlat(d, "l2", "bf16", 0)
    ref.add(xt.cuda())
    rb = ref.save(str(tmp_path / "ref.bin")) or open(tmp_path / "ref.bin", "rb").read()
    for how in ("pageable", "pinned", "file"):
        idx = IndexFlat(d, "l2", "bf16", 0)
        if how == "pageable":
            idx.add(xt)
        elif how == "pinned":
            idx.add(xt.pin_memory())
        else:
            raw = x if dtype == "float32" else O.bf16_bits(x)
            raw.tofile(tmp_path / "rows.bin")
            assert idx.add_from_file(str(tmp_path / "rows.bin"), dtype=dtype, chunk_rows=70_000) == n
        assert idx.ntotal == n
        idx.save(str(tmp_path / "got.bin"))
        assert open(tmp_path / "got.bin", "rb").read() == rb, how
        idx.close()
    ref.close()


def test_host_add_larger_than_4_GiB():
    """A single add() of 4.6 GB of fp32 host rows: sizes and byte offsets beyond 32 bits on the ingest path.
    The first 1.4M rows repeat one 4000-row tile (cheap to build); the last 100k rows -- the ones that lie past
    the 4 GiB mark -- are unique, so finding each of them as its own nearest neighbour proves they arrived."""
    from cloudvectordb_b200 import IndexFlat
    n_rep, n_uni, d = 1_400_000, 100_000, 768
    n = n_rep + n_uni
    rng = np.random.default_rng(12)
    tile = O.bf16_round(unit_rows(rng, 4000, d))
    x = np.empty((n, d), np.float32)
    for r0 in range(0, n_rep, 4000):
        x[r0:r0 + 4000] = tile
    x[n_rep:] = O.bf16_round(unit_rows(rng, n_uni, d))
    assert x.nbytes > (1 << 32) and (n_rep * d * 4) < (1 << 32) + (1 << 28)
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(x)
    assert idx.ntotal == n
    probe = np.array([n_rep, n_rep + 1, n_rep + 50_000, n - 2, n - 1, 1_398_102, 1_399_999])
    D, I = idx.search(x[probe], 1)
    want = np.where(probe >= n_rep, probe, probe % 4000)           # repeated rows: the lowest twin wins the tie
    assert np.array_equal(I[:, 0], want)
    assert np.allclose(D[:, 0], (x[probe] * x[probe]).sum(1), rtol=1e-5)
    idx.close()


# ---- small edges -------------------------------------------------------------------------------------------------
def test_ivf_search_on_an_empty_index_is_all_padding():
    from cloudvectordb_b200 import IndexIVFFlat
    ivf = IndexIVFFlat(16, 4, "l2", device=0)
    ivf.train(None, centroids=np.eye(4, 16, dtype=np.float32))
    D, I = ivf.search(np.ones((3, 16), np.float32), 5, nprobe=2)
    assert np.all(I == -1) and np.all(np.isposinf(D)) and D.shape == (3, 5)
    ivf.close()


def test_kmeans_with_fewer_points_than_clusters_raises():
    from cloudvectordb_b200 import Kmeans
    km = Kmeans(8, 50, niter=1, device=0)
    with pytest.raises(ValueError, match="at least k points"):
        km.train(np.ones((10, 8), np.float32))


def test_merge_topk_host_path_reuses_no_allocation_and_checks_copies():
    from cloudvectordb_b200 import merge_topk
    rng = np.random.default_rng(2)
    for _ in range(3):
        Dp = -np.sort(-rng.standard_normal((4, 50, 10)).astype(np.float32), axis=2)
        Ip = rng.permutation(4 * 50 * 10).reshape(4, 50, 10).astype(np.int64)
        D, I = merge_topk(Dp, Ip, 10, "ip")
        Dr, Ir = O.merge_ref(list(Dp), list(Ip), 10, O.METRIC_IP)
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr)


def test_handleless_calls_run_on_the_device_that_owns_the_data():
    """Kmeans / IVF / merge on cuda:1 while the current device is cuda:0 (advisor finding, round 1)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cloudvectordb_b200 import IndexIVFFlat, Kmeans, merge_topk
    assert torch.cuda.current_device() == 0
    rng = np.random.default_rng(4)
    x = O.bf16_round(unit_rows(rng, 20000, 32))
    km = Kmeans(32, 16, niter=2, device=1)
    km.train(x, init_centroids=x[:16])
    a_ref, _ = O.kmeans_assign_ref(x, x[:16])
    newc, counts, _ = O.kmeans_update_ref(x, a_ref, x[:16])
    km1 = Kmeans(32, 16, niter=1, device=1)
    km1.train(x, init_centroids=x[:16])
    assert np.array_equal(km1.last_counts.cpu().numpy(), counts)
    ivf = IndexIVFFlat(32, 16, "l2", device=1)
    ivf.train(None, centroids=km.centroids)
    ivf.add(x)
    D, I = ivf.search(x[:100], 5, nprobe=16)
    D_ref, I_ref = O.search_ref(x, x[:100], 5, O.METRIC_L2)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5, metric=O.METRIC_L2) == 0
    Dg = torch.from_numpy(D_ref[None]).to("cuda:1")
    Ig = torch.from_numpy(I_ref[None]).to("cuda:1")
    Dm, Im = merge_topk(Dg, Ig, 5, "l2")
    assert torch.equal(Im.cpu(), torch.from_numpy(I_ref))
    assert torch.cuda.current_device() == 0
    ivf.close()


# ---- slices that start from an earlier slice's result (inheritance) ---------------------------------------------
@pytest.mark.parametrize("variant,k", [(1, 10), (2, 50), (3, 100), (2, 200), (1, 20)])
def test_slice_inheritance_is_result_neutral(variant, k):
    """Many database slices (more than 32, so merge lanes own several lists), slices long enough to inherit: the
    lists of different slices overlap and the merge must drop the copies.  Same answer with inheritance on
    (debug flag 128) and off, and equal to the oracle; with exclusion as well."""
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(variant * 100 + k)
    n, d, nq = 400_000, 64, 300
    xb = O.bf16_round(unit_rows(rng, n, d))
    xb[1000:1040] = xb[7]                                   # a block of exact duplicates (ties across and inside slices)
    xb[390_000:390_020] = xb[7]
    xq = O.bf16_round(unit_rows(rng, nq, d))
    xq[0] = xb[7]
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(xb)
    slices = 40 if variant != 1 else 45
    D, I = idx.search(xq, k, force_variant=variant, force_slices=slices, debug_flags=128)   # inheritance on
    assert idx.last_work()["n_slices"] >= 33
    D0, I0 = idx.search(xq, k, force_variant=variant, force_slices=slices)
    assert np.array_equal(I, I0) and np.array_equal(D, D0)
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    assert all(len(set(r[r >= 0])) == (r >= 0).sum() for r in I)           # no id twice
    groups = (np.arange(n) // 4).astype(np.int32)
    idx.set_groups(groups)
    self_ids = rng.integers(0, n, nq).astype(np.int32)
    gq = groups[self_ids]
    D, I = idx.search(xb[self_ids], k, self_ids=self_ids, group_q=gq, force_variant=variant, force_slices=slices,
                      debug_flags=128)
    D_ref, I_ref = O.search_ref(xb, xb[self_ids], k, O.METRIC_IP, self_ids=self_ids, group_db=groups, group_q=gq)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    idx.close()


# ---- pooled thresholds across slices (order statistics of finished slices bound what later slices collect) ------
@pytest.mark.parametrize("variant,k,metric", [(1, 10, "ip"), (2, 50, "ip"), (3, 100, "ip"), (2, 200, "ip"), (1, 3, "l2"),
                                              (2, 17, "l2"), (3, 248, "ip")])
def test_pooled_thresholds_are_result_neutral(variant, k, metric):
    """Many slices, pooled thresholds forced on for short items (debug flag 512): identical to the run without them
    (flag 256) and equal to the oracle -- on iid rows, with blocks of exact duplicates spread over slices (ties at the
    k-th score must survive the pooled bound), on rows sorted by ASCENDING score (every slice beats all earlier
    ones) and DESCENDING score (the first slices already hold the answer), and with self + group exclusion."""
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(variant * 1000 + k)
    n, d, nq = 300_000, 64, 200
    code = O.METRIC_IP if metric == "ip" else O.METRIC_L2
    xb = O.bf16_round(unit_rows(rng, n, d))
    for r0 in (500, 60_000, 150_000, 299_000):                 # the same row in many slices: ties across slices
        xb[r0:r0 + 70] = xb[3]
    xq = O.bf16_round(unit_rows(rng, nq, d))
    xq[0] = xb[3]
    slices = 37 if variant != 1 else 41
    for order in ("iid", "ascending", "descending"):
        if order != "iid":
            s = xb @ xq[1]
            xb = xb[np.argsort(s if order == "ascending" else -s, kind="stable")]
        idx = IndexFlat(d, metric, "bf16", 0)
        idx.add(xb)
        D, I = idx.search(xq, k, force_variant=variant, force_slices=slices, debug_flags=512)
        assert idx.last_work()["n_slices"] >= 33
        D0, I0 = idx.search(xq, k, force_variant=variant, force_slices=slices, debug_flags=256)
        assert np.array_equal(I, I0) and np.array_equal(D, D0), order
        D_ref, I_ref = O.search_ref(xb, xq, k, code)
        assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5, metric=code) == 0, order
        if order == "iid" and metric == "ip":
            groups = (np.arange(n) // 4).astype(np.int32)
            idx.set_groups(groups)
            self_ids = rng.integers(0, n, nq).astype(np.int32)
            gq = groups[self_ids]
            D, I = idx.search(xb[self_ids], k, self_ids=self_ids, group_q=gq, force_variant=variant, force_slices=slices,
                              debug_flags=512)
            D_ref, I_ref = O.search_ref(xb, xb[self_ids], k, O.METRIC_IP, self_ids=self_ids, group_db=groups, group_q=gq)
            assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
        idx.close()
