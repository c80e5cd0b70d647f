"""GPU parity of the symmetric self-join (cvdb_selfjoin_*, mine_hard_negatives(symmetric=True)): every tile of
X.X^T is computed once and selected in both directions; the result must equal the plain self-join and the
oracle bit for bit up to ties, for every row, with self / group exclusion, ties across directions, several chunk
schedules, both kernel configurations, and when the column buffers overflow (adversarial row order)."""
import numpy as np
import pytest
import torch

from oracle import flat_oracle as O

pytestmark = pytest.mark.gpu


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def oracle_join(emb, k, groups):
    n = emb.shape[0]
    return O.search_ref(emb, emb, k, O.METRIC_IP, self_ids=np.arange(n), group_db=groups, group_q=groups)


def check(D, I, D_ref, I_ref, tol=2e-5):
    assert np.array_equal(I < 0, I_ref < 0)
    assert O.check_topk(D, I, D_ref, I_ref, tie_tol=tol) == 0
    fin = np.isfinite(D_ref)
    assert np.allclose(D[fin], D_ref[fin], atol=1e-4)


def test_schedule_covers_rows_in_order_and_at_most_doubles():
    from cloudvectordb_b200 import selfjoin_schedule
    for n in (1, 255, 256, 257, 5000, 200_000, 6_250_000):
        s = selfjoin_schedule(n, 65536, 256 if n < 100_000 else 8192)
        assert s[0][0] == 0 and sum(m for _, m in s) == n
        for (r0, m), (r1, _) in zip(s, s[1:]):
            assert r1 == r0 + m and r0 % 256 == 0 and m <= 65536 and (r0 == 0 or m <= r0)


@pytest.mark.parametrize("n,d,k,chunk,first", [
    (20_000, 64, 10, 2048, 256),     # many chunks, K <= 512 configuration
    (9_000, 768, 50, 4096, 512),     # K = 768 configuration (TMEM + shared-memory tail), k = 50
    (6_001, 96, 100, 65536, 256),    # ragged size, k = 100
    (3_000, 40, 2, 256, 256),        # smallest k, smallest chunks
    (700, 32, 20, 65536, 256),       # fewer rows than three chunks
    (30_000, 64, 10, 8192, 8192),    # a seed of 8192 rows
    (5_000, 64, 10, 65536, 8192),    # the seed covers every row: all plain
    (70_000, 64, 10, 65536, None),   # the default seed (65 536 rows) and one short chunk after it
    (4_000, 48, 12, 65536, None),    # default seed larger than the corpus
])
def test_symmetric_join_equals_plain_join_and_oracle(n, d, k, chunk, first):
    from cloudvectordb_b200 import mine_hard_negatives
    rng = np.random.default_rng(n + k)
    emb = O.bf16_round(unit_rows(rng, n, d))
    groups = (np.arange(n) // 4).astype(np.int32)
    groups[::9] = -1                                          # some rows without a group
    D_ref, I_ref = oracle_join(emb, k, groups)
    D, I = mine_hard_negatives(emb, k, groups, chunk=chunk, symmetric=True, first_chunk=first)
    check(D, I, D_ref, I_ref)
    assert not np.any(I == np.arange(n)[:, None])
    same = (groups[np.clip(I, 0, None)] == groups[:, None]) & (groups[:, None] >= 0) & (I >= 0)
    assert not same.any()
    D_p, I_p = mine_hard_negatives(emb, k, groups, chunk=max(chunk, 1024))
    assert O.check_topk(D, I, D_p, I_p, tie_tol=2e-5) == 0


def test_ties_between_the_two_directions_go_to_the_lower_id():
    """Exact duplicates spread over the matrix: a row's neighbours at equal score come partly from the row direction
    (later rows) and partly from the column direction (earlier anchors); the lower id must win everywhere."""
    from cloudvectordb_b200 import mine_hard_negatives
    rng = np.random.default_rng(5)
    n, d, k = 8_000, 64, 6
    emb = O.bf16_round(unit_rows(rng, n, d))
    dup = [10, 300, 1100, 2500, 4097, 6000, 7999]
    emb[dup] = emb[10]
    D, I = mine_hard_negatives(emb, k, None, chunk=1024, symmetric=True, first_chunk=256)
    D_ref, I_ref = oracle_join(emb, k, None)
    assert np.array_equal(I, I_ref)
    for r in dup:
        assert list(I[r, :6]) == [x for x in dup if x != r]


def test_adversarial_row_order_overflows_the_column_buffers_and_is_recomputed():
    """Rows converge to one direction, so every later anchor beats everything before it: whole chunks pass the column
    thresholds, the 256-key buffers overflow, the rows are flagged and recomputed exactly by the plain search."""
    from cloudvectordb_b200 import IndexFlat, mine_hard_negatives_symmetric
    rng = np.random.default_rng(9)
    n, d, k = 12_000, 64, 10
    u = unit_rows(rng, 1, d)
    noise = unit_rows(rng, n, d)
    scale = np.linspace(2.0, 0.02, n, dtype=np.float32)[:, None]      # later rows: closer to u, and to each other
    emb = u + scale * noise
    emb = O.bf16_round(emb / np.linalg.norm(emb, axis=1, keepdims=True))
    groups = (np.arange(n) // 4).astype(np.int32)
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(emb)
    idx.set_groups(groups)
    stats = {}
    D, I = mine_hard_negatives_symmetric(idx, k, emb=emb, groups=groups, chunk=4096, first_chunk=256, stats=stats)
    idx.close()
    assert stats["dirty_rows"] > 0
    D_ref, I_ref = oracle_join(emb, k, groups)
    check(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref)


def test_symmetric_join_at_mining_scale_matches_plain_join():
    """300k x 768, k = 50, groups of four, default chunking: every row against the plain (row-direction only) join."""
    from cloudvectordb_b200 import IndexFlat, mine_hard_negatives, mine_hard_negatives_symmetric
    n, d, k = 300_000, 768, 50
    g = torch.Generator(device="cuda").manual_seed(3)
    emb = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1).to(torch.bfloat16)
    groups = (torch.arange(n, device="cuda") // 4).to(torch.int32)
    idx = IndexFlat(d, "ip", "bf16", 0)
    idx.add(emb)
    idx.set_groups(groups)
    stats = {}
    D, I = mine_hard_negatives_symmetric(idx, k, emb=emb, groups=groups, stats=stats)
    assert stats["dirty_rows"] == 0
    D_p, I_p = mine_hard_negatives(emb, k, groups, index=idx)
    idx.close()
    assert O.check_topk(D.cpu().numpy(), I.cpu().numpy(), D_p.cpu().numpy(), I_p.cpu().numpy(), tie_tol=2e-5) == 0
    assert float((I == I_p).float().mean()) > 0.9999


def test_selfjoin_argument_checks():
    from cloudvectordb_b200 import IndexFlat, _C
    lib = _C.lib()
    idx = IndexFlat(32, "l2", "bf16", 0)
    idx.add(np.ones((300, 32), np.float32))
    assert lib.cvdb_selfjoin_begin(idx._h, 10, None) == _C.EINVAL          # L2 index
    idx.close()
    idx = IndexFlat(32, "ip", "bf16", 0)
    idx.add(np.ones((300, 32), np.float32))
    assert lib.cvdb_selfjoin_chunk(idx._h, 0, 256, 0, 1, None) == _C.EINVAL    # no join open
    assert lib.cvdb_selfjoin_begin(idx._h, 1, None) == _C.ELIMIT
    assert lib.cvdb_selfjoin_begin(idx._h, 10, None) == 0
    keys = torch.empty((300, 10), dtype=torch.int64, device="cuda")
    assert lib.cvdb_selfjoin_chunk(idx._h, 64, 100, 0, keys.data_ptr(), None) == _C.EINVAL   # unaligned start
    assert lib.cvdb_selfjoin_chunk(idx._h, 256, 100, 0, keys.data_ptr(), None) == _C.EINVAL  # past the end
    assert lib.cvdb_selfjoin_seed(idx._h, 100, 0, keys.data_ptr(), None) == _C.EINVAL         # seed not a multiple of 128
    assert lib.cvdb_selfjoin_end(idx._h) == 0
    idx.close()
