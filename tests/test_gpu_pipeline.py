"""GPU tests of the rows either side of the search path: triplet assembly from the
mined neighbours, and the on-disk format / streaming add."""
import os

import numpy as np
import pytest
import torch

from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("metric,limit", [("ip", None), ("ip", 0.35), ("l2", 1.2)])
def test_triplets_from_mined_negatives(metric, limit):
    from cloudvectordb_b200 import build_triplets, mine_hard_negatives
    rng = np.random.default_rng(3)
    n, d, k = 3000, 48, 20
    emb = O.bf16_round(unit_rows(rng, n, d))
    groups = (np.arange(n) // 3).astype(np.int32)
    D, I = mine_hard_negatives(emb, k, groups, metric=metric)
    positives = np.where(np.arange(n) % 3 == 2, np.arange(n) - 1, np.arange(n) + 1).astype(np.int64)
    positives[::7] = -1
    m = O.METRIC_IP if metric == "ip" else O.METRIC_L2
    for skip_top, per_anchor in ((0, 1), (2, 4), (18, 5)):
        T = build_triplets(D, I, positives, skip_top=skip_top, per_anchor=per_anchor, metric=metric, limit=limit,
                           anchor_base=100)
        T_ref = O.build_triplets_ref(D, I, positives, skip_top, per_anchor, m, limit, anchor_base=100)
        assert np.array_equal(T, T_ref)
        valid = T[:, :, 0] >= 0
        a, p, ng = T[valid][:, 0] - 100, T[valid][:, 1], T[valid][:, 2]
        assert np.all(groups[a] == groups[p]) and np.all(groups[a] != groups[ng]) and np.all(a != ng)
    Tg = build_triplets(torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda(), positives, per_anchor=2, metric=metric)
    assert Tg.is_cuda and np.array_equal(Tg.cpu().numpy(), O.build_triplets_ref(D, I, positives, 0, 2, m))


@pytest.mark.parametrize("metric,storage", [("ip", "bf16"), ("l2", "bf16"), ("l2", "exact")])
def test_save_load_roundtrip(tmp_path, metric, storage):
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(4)
    n, d, nq, k = 5000, 72, 60, 10
    xb, xq = unit_rows(rng, n, d), unit_rows(rng, nq, d)
    a = IndexFlat(d, metric, storage)
    a.add(xb)
    D0, I0 = a.search(xq, k)
    path = str(tmp_path / "flat.cvdb")
    a.save(path, chunk_rows=1024)
    a.close()
    b = IndexFlat.load(path, chunk_rows=999)
    assert b.ntotal == n and b.d == d and b.metric == metric and b.storage == storage
    D1, I1 = b.search(xq, k)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)      # same bytes in HBM -> bit-identical results
    b.add(xb[:10])                                                # a loaded index keeps growing
    assert b.ntotal == n + 10
    b.close()
    with open(path, "r+b") as f:                                  # truncated file is refused
        f.truncate(os.path.getsize(path) - 100)
    with pytest.raises(ValueError):
        IndexFlat.load(path)


def test_streaming_add_from_file(tmp_path):
    from cloudvectordb_b200 import IndexFlat
    rng = np.random.default_rng(5)
    n, d, nq, k = 7001, 40, 30, 5
    xb, xq = O.bf16_round(unit_rows(rng, n, d)), O.bf16_round(unit_rows(rng, nq, d))
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP)
    p32, p16 = str(tmp_path / "x.f32"), str(tmp_path / "x.bf16")
    xb.tofile(p32)
    O.bf16_bits(xb).tofile(p16)
    for path, dt in ((p32, "float32"), (p16, "bfloat16")):
        idx = IndexFlat(d, "ip", "bf16")
        assert idx.add_from_file(path, dt, chunk_rows=1000) == n
        D, I = idx.search(xq, k)
        idx.close()
        assert O.check_topk(D, I, D_ref, I_ref, tie_tol=2e-5) == 0
    bad = str(tmp_path / "bad.f32")
    xb.reshape(-1)[:-3].tofile(bad)
    idx = IndexFlat(d, "ip", "bf16")
    with pytest.raises(ValueError):
        idx.add_from_file(bad)
    idx.close()


def test_example_pipeline_runs(tmp_path, capsys):
    """examples/build_vectordb.py: mining -> triplets -> save/load -> IVF on a small synthetic corpus."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("build_vectordb", os.path.join(root, "examples", "build_vectordb.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["--rows", "20000", "--dim", "64", "--nlist", "64", "--out", str(tmp_path / "c.cvdb")]) == 0
    out = capsys.readouterr().out
    assert "PIPELINE_OK" in out and "triplets" in out
