"""GPU parity of the IVF-Flat widening: inverted lists from
the k=1 assignment, nprobe search = exact search restricted to the probed lists."""
import numpy as np
import pytest
import torch

from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu
M = {"ip": O.METRIC_IP, "l2": O.METRIC_L2}


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def build(metric, n, d, nlist, rng, centroids=None, chunks=1):
    from cloudvectordb_b200 import IndexIVFFlat
    xb = O.bf16_round(unit_rows(rng, n, d))
    ivf = IndexIVFFlat(d, nlist, metric, device=0)
    if centroids is None:
        centroids = O.bf16_round(xb[rng.choice(n, nlist, replace=False)])
    ivf.train(None, centroids=centroids)
    for part in np.array_split(xb, chunks):
        ivf.add(part)
    return ivf, xb, centroids


@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("n,d,nlist,nq,k,nprobe", [
    (20000, 64, 64, 300, 10, 4),
    (30000, 128, 257, 129, 10, 16),    # short lists (~117 rows): one partial tile each
    (5000, 96, 16, 50, 50, 3),         # long lists, several tiles, larger k
    (3000, 40, 100, 1, 1, 5),          # single query, top-1
])
def test_ivf_search_matches_oracle_on_probed_lists(metric, n, d, nlist, nq, k, nprobe):
    rng = np.random.default_rng(n + nlist)
    ivf, xb, cent = build(metric, n, d, nlist, rng, chunks=3)
    xq = O.bf16_round(unit_rows(rng, nq, d))
    assert ivf.ntotal == n
    # assignment: equal to the oracle's except where the two best centroids tie
    a_gpu = ivf.list_of_row().cpu().numpy()
    a_ref = O.ivf_assign_ref(cent, xb, M[metric])
    differ = a_gpu != a_ref
    if differ.any():
        s = xb[differ] @ cent.T if metric == "ip" else -((xb[differ][:, None, :] - cent[None]) ** 2).sum(-1)
        gap = np.abs(s[np.arange(differ.sum()), a_gpu[differ]] - s[np.arange(differ.sum()), a_ref[differ]])
        assert np.all(gap <= 2e-5)
    D, I = ivf.search(xq, k, nprobe=nprobe)
    probes = ivf.probe(xq, nprobe).cpu().numpy()
    p_ref = O.ivf_probe_ref(cent, xq, nprobe, M[metric])
    assert (probes == p_ref).mean() > 0.99
    # the search itself, given the engine's own lists and probes
    D_ref, I_ref = O.ivf_search_ref(xb, a_gpu, xq, k, probes, M[metric])
    assert np.array_equal(I < 0, I_ref < 0)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5, metric=M[metric]) == 0
    fin = np.isfinite(D_ref)
    assert np.allclose(D[fin], D_ref[fin], atol=1e-4)
    ivf.close()


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_ivf_with_all_lists_probed_is_exact_search(metric):
    """nprobe == nlist scans everything: must equal the brute-force oracle (no dependence on the lists)."""
    rng = np.random.default_rng(5)
    n, d, nlist, nq, k = 12000, 80, 37, 200, 10
    ivf, xb, _ = build(metric, n, d, nlist, rng, chunks=2)
    xq = O.bf16_round(unit_rows(rng, nq, d))
    D, I = ivf.search(torch.from_numpy(xq).cuda(), k, nprobe=nlist)
    D_ref, I_ref = O.search_ref(xb, xq, k, M[metric])
    assert O.check_topk(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, tie_tol=2e-5, metric=M[metric]) == 0
    ivf.close()


def test_ivf_train_with_kmeans_and_recall():
    rng = np.random.default_rng(9)
    n, d, nlist, nq, k = 100_000, 64, 256, 500, 10
    centres = rng.standard_normal((512, d)).astype(np.float32)
    xb = centres[rng.integers(0, 512, n)] + 0.3 * rng.standard_normal((n, d)).astype(np.float32)
    xb = O.bf16_round(xb / np.linalg.norm(xb, axis=1, keepdims=True))
    xq = xb[rng.choice(n, nq, replace=False)] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
    xq = O.bf16_round(xq / np.linalg.norm(xq, axis=1, keepdims=True))
    from cloudvectordb_b200 import IndexIVFFlat
    ivf = IndexIVFFlat(d, nlist, "l2", device=0)
    ivf.train(xb, niter=5)
    ivf.add(xb)
    _, I_exact = O.search_ref(xb, xq, k, O.METRIC_L2)
    recalls = []
    for nprobe in (1, 8, 64, 256):
        _, I = ivf.search(xq, k, nprobe=nprobe)
        recalls.append(O.recall_at_k(I, I_exact))
    assert all(b >= a - 1e-9 for a, b in zip(recalls, recalls[1:]))   # more probes never hurt
    assert recalls[-1] >= 0.999 and recalls[1] > 0.5
    ivf.close()


def test_ivf_empty_lists_and_padding():
    rng = np.random.default_rng(11)
    d, nlist = 32, 8
    cent = O.bf16_round(unit_rows(rng, nlist, d))
    from cloudvectordb_b200 import IndexIVFFlat
    ivf = IndexIVFFlat(d, nlist, "ip", device=0)
    ivf.train(None, centroids=cent)
    xb = O.bf16_round(cent[[0, 0, 3]] * 0.9)          # only lists 0 and 3 get rows
    ivf.add(xb)
    xq = cent.copy()
    D, I = ivf.search(xq, 4, nprobe=2)
    probes = ivf.probe(xq, 2).cpu().numpy()
    D_ref, I_ref = O.ivf_search_ref(xb, ivf.list_of_row().cpu().numpy(), xq, 4, probes, O.METRIC_IP)
    assert np.array_equal(I, I_ref)
    assert np.all(np.isneginf(D[I < 0]))
    ivf.close()


def test_ivf_add_after_search_regroups():
    """Rows added after a first search are picked up: the lists are rebuilt, ids stay insertion order."""
    from cloudvectordb_b200 import IndexIVFFlat
    rng = np.random.default_rng(13)
    n, d, nlist, nq, k, nprobe = 9000, 48, 32, 120, 10, 6
    xb = O.bf16_round(unit_rows(rng, n, d))
    xq = O.bf16_round(unit_rows(rng, nq, d))
    cent = O.bf16_round(xb[rng.choice(n, nlist, replace=False)])
    ivf = IndexIVFFlat(d, nlist, "ip", device=0)
    ivf.train(None, centroids=cent)
    for lo, hi in ((0, 4000), (4000, 4001), (4001, 9000)):
        ivf.add(xb[lo:hi])
        D, I = ivf.search(xq, k, nprobe=nprobe)          # groups (again) on demand
        a = ivf.list_of_row().cpu().numpy()
        assert a.shape[0] == hi and int(ivf.list_offsets[-1]) == hi
        probes = ivf.probe(xq, nprobe).cpu().numpy()
        D_ref, I_ref = O.ivf_search_ref(xb[:hi], a, xq, k, probes, O.METRIC_IP)
        assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
        assert I.max() < hi
    ivf.close()


def test_ivf_duplicate_rows_tie_by_row_id():
    """Equal scores inside and across lists come back in row-id order although rows are stored unordered."""
    from cloudvectordb_b200 import IndexIVFFlat
    rng = np.random.default_rng(17)
    d, nlist = 32, 4
    cent = O.bf16_round(unit_rows(rng, nlist, d))
    base = O.bf16_round(unit_rows(rng, 40, d))
    xb = np.concatenate([base, base, base])              # every row three times: ids i, i+40, i+80
    ivf = IndexIVFFlat(d, nlist, "ip", device=0)
    ivf.train(None, centroids=cent)
    ivf.add(xb)
    D, I = ivf.search(base, 3, nprobe=nlist)
    assert np.array_equal(I, np.stack([np.arange(40), np.arange(40) + 40, np.arange(40) + 80], 1))
    D1, I1 = ivf.search(base, 1, nprobe=nlist)           # top-1 path
    assert np.array_equal(I1[:, 0], np.arange(40))
    ivf.close()
