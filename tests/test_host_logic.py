"""CPU tests for the host side: the C-ABI library loads and exports every
symbol include/cvdb_b200.h declares, argument validation that needs no GPU,
the shard arithmetic, and the world_size-2 shard/merge path over gloo with
oracle-backed stand-ins for the per-rank search."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cloudvectordb_b200 import _C
from cloudvectordb_b200.sharded import ShardedIndex, shard_bounds
from oracle import flat_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cvdb_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cvdb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _C.lib()
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cvdb_b200.h but not exported"
    assert set(names) == set(_C.SIGNATURES), "ctypes signatures and header disagree"
    assert lib.cvdb_version() >= 100
    assert lib.cvdb_kernel_launches() >= 0


def test_header_constants_match_binding():
    txt = open(os.path.join(ROOT, "include", "cvdb_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(CVDB_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", txt))
    assert int(consts["CVDB_METRIC_L2"]) == _C.METRIC_L2
    assert int(consts["CVDB_DTYPE_BF16"]) == _C.DTYPE_BF16
    assert int(consts["CVDB_DTYPE_F16"]) == _C.DTYPE_F16
    assert int(consts["CVDB_STORE_EXACT"]) == _C.STORE_EXACT
    assert int(consts["CVDB_MAX_K"]) == _C.MAX_K
    assert int(consts["CVDB_ECUDA"]) == _C.ECUDA


def test_argument_validation_without_gpu():
    lib = _C.lib()
    h = C.c_void_p()
    assert lib.cvdb_index_create(0, 0, 0, 0, C.byref(h)) == _C.ELIMIT
    assert b"d=0" in lib.cvdb_last_error()
    assert lib.cvdb_index_create(8, 7, 0, 0, C.byref(h)) == _C.EINVAL
    assert lib.cvdb_index_create(8, 0, 9, 0, C.byref(h)) == _C.EINVAL
    assert lib.cvdb_index_ntotal(None) == -1
    assert lib.cvdb_index_search(None, None, 1, 0, 1, None, None, 0, None, None) == _C.EINVAL
    assert lib.cvdb_merge_topk(None, None, 4, 0, 1, 1, 0, None, None, 0, None) == _C.EINVAL
    if not torch.cuda.is_available():
        # no CPU fallback: creation must fail loudly without a GPU
        rc = lib.cvdb_index_create(8, 0, 0, 0, C.byref(h))
        assert rc == _C.ECUDA and b"no CPU fallback" in lib.cvdb_last_error()
        with pytest.raises(_C.CvdbError):
            from cloudvectordb_b200 import IndexFlatIP
            IndexFlatIP(8)


def test_slice_planner_properties():
    lib = _C.lib()

    def plan(qt, nt, w, mx=1 << 40):
        a, b = C.c_int(), C.c_int()
        assert lib.cvdb_plan_slices(qt, nt, w, mx, C.byref(a), C.byref(b)) == 0
        return a.value, b.value

    import math
    for qt, nt, w in [(1, 1, 148), (40, 78125, 74), (79, 39063, 148), (1, 48829, 148), (512, 24415, 74), (3, 7, 74),
                      (40, 9766, 74), (1000, 5, 148)]:
        s, tps = plan(qt, nt, w)
        assert 1 <= s <= nt and tps >= 1
        assert (s - 1) * tps < nt <= s * tps                      # the slices cover every tile, none is empty
        waves = math.ceil(qt * s / w)
        cost, cost_one = waves * (tps + 4), math.ceil(qt / w) * (nt + 4)
        assert cost <= 1.01 * cost_one                            # never (noticeably) worse than not slicing at all
        ideal = qt * nt / w
        if qt * nt >= 50 * w:
            assert cost <= 1.08 * ideal + 8                       # close to perfectly balanced for big problems
    # headline shape: 40 query tiles x 74 CTA pairs -> 37 slices, 20 whole waves
    assert plan(40, 78125, 74)[0] == 37
    # long slices are avoided when a plan with shorter ones costs about the same: the mining chunk (256 query tiles of
    # 256 anchors, 10M / 6.25M rows, at most 40 slices of scratch) takes slices of about 2000 tiles, not 13 long ones
    for nt in (78125, 48829):
        s, tps = plan(256, nt, 74, 40)
        assert tps <= 2200 and s <= 40
        assert math.ceil(256 * s / 74) * (tps + 8) <= 1.01 * 256 * nt / 74 + 8     # and stays balanced
    # a scratch budget caps the number of slices
    s, tps = plan(40, 78125, 74, 5)
    assert s <= 5 and (s - 1) * tps < 78125 <= s * tps
    assert lib.cvdb_plan_slices(0, 5, 74, 10, C.byref(C.c_int()), C.byref(C.c_int())) == _C.EINVAL


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 100, 10_000_001):
        for w in (1, 2, 3, 8):
            prev = 0
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = [shard_bounds(n, w, r)[1] - shard_bounds(n, w, r)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------- gloo, world_size 2
class OracleLocalIndex:
    """Stand-in for IndexFlat on CPU ranks: same surface, oracle arithmetic."""

    def __init__(self, d, metric):
        self.d, self.metric = d, metric
        self.x = np.zeros((0, d), np.float32)
        self.groups = None

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, np.float32)])

    def reset(self):
        self.x = self.x[:0]

    def set_groups(self, g):
        self.groups = None if g is None else np.asarray(g, np.int32)

    def search(self, q, k, self_ids=None, group_q=None, id_base=0):
        sid = None if self_ids is None else np.asarray(self_ids, np.int64)
        D, I = O.search_ref(self.x, np.asarray(q, np.float32), k, O.METRIC_IP if self.metric == "ip" else O.METRIC_L2,
                            self_ids=sid, group_db=self.groups if group_q is not None else None,
                            group_q=None if group_q is None else np.asarray(group_q))
        I = np.where(I >= 0, I + id_base, -1)
        return torch.from_numpy(D), torch.from_numpy(I)

    # the exchange format of the product path, restated in NumPy: key = ordered(score) << 32 | ~id, where the score
    # is "larger is better" (IP: the score, L2: minus the squared distance); 0 = no result
    def search_keys(self, q, k, self_ids=None, group_q=None, id_base=0):
        D, I = self.search(q, k, self_ids=self_ids, group_q=group_q, id_base=id_base)
        D, I = D.numpy(), I.numpy()
        s = (D if self.metric == "ip" else -D).astype(np.float32)
        u = np.where(I >= 0, s, np.float32(0)).view(np.uint32).astype(np.uint64)
        u = np.where(u == 0x80000000, 0, u)                                       # -0.0 == +0.0
        o = np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000)
        key = (o << np.uint64(32)) | (~I.astype(np.uint64) & np.uint64(0xFFFFFFFF))
        key = np.where(I >= 0, key, np.uint64(0))
        return torch.from_numpy(key.view(np.int64).copy())

    def merge_keys(self, keys, k):
        key = keys.numpy().view(np.uint64)                                         # [lists, nq, k_in]
        nl, nq, k_in = key.shape
        flat = np.transpose(key, (1, 0, 2)).reshape(nq, nl * k_in)
        top = np.sort(flat, axis=1)[:, ::-1][:, :k]                                # uint64 keys: larger is better
        o = (top >> np.uint64(32)).astype(np.uint32)
        u = np.where(o & 0x80000000, o & 0x7FFFFFFF, ~o).astype(np.uint32)
        s = u.view(np.float32)
        I = np.where(top != 0, (~top & np.uint64(0xFFFFFFFF)).astype(np.int64), -1)
        D = np.where(top != 0, s if self.metric == "ip" else -s, -np.inf if self.metric == "ip" else np.inf)
        return torch.from_numpy(D.astype(np.float32)), torch.from_numpy(I)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, metric, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        n, d, nq, k = 501, 12, 9, 6
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        groups = (np.arange(n) // 4).astype(np.int32)
        self_ids = rng.integers(0, n, nq)
        idx = ShardedIndex(d, metric, local_index=OracleLocalIndex(d, metric))
        idx.add(xb)
        lo, hi = shard_bounds(n, world, rank)
        assert idx.local.ntotal == hi - lo and idx.id_base == lo and idx.ntotal == n
        idx.set_groups_local(groups[lo:hi])
        D, I = idx.search(xq, k)
        D2, I2 = idx.search(xq, k, self_ids=self_ids, group_q=groups[self_ids])
        q.put((rank, D.numpy(), I.numpy(), D2.numpy(), I2.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_sharded_search_world2_gloo(metric):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, metric, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    n, d, nq, k = 501, 12, 9, 6
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    groups = (np.arange(n) // 4).astype(np.int32)
    self_ids = rng.integers(0, n, nq)
    m = O.METRIC_IP if metric == "ip" else O.METRIC_L2
    D_ref, I_ref = O.search_ref(xb, xq, k, m)
    D2_ref, I2_ref = O.search_ref(xb, xq, k, m, self_ids=self_ids, group_db=groups, group_q=groups[self_ids])
    for rank, D, I, D2, I2 in outs:
        assert np.array_equal(I, I_ref) and np.allclose(D, D_ref, atol=1e-5)
        assert np.array_equal(I2, I2_ref) and np.allclose(D2, D2_ref, atol=1e-5)


def _mining_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cloudvectordb_b200.mining import mine_hard_negatives_sharded
        rng = np.random.default_rng(11)
        n, d, k = 501, 10, 7
        emb = rng.standard_normal((n, d), dtype=np.float32)
        groups = (np.arange(n) // 3).astype(np.int32)
        lo, hi = (0, 300) if rank == 0 else (300, n)               # uneven shards, ragged last chunk
        idx = ShardedIndex(d, "ip", local_index=OracleLocalIndex(d, "ip"))
        idx.add_local(emb[lo:hi])
        assert idx.ntotal == n and idx.id_base == lo
        idx.set_groups_local(groups[lo:hi])
        D, I = mine_hard_negatives_sharded(idx, torch.from_numpy(emb[lo:hi]), k, torch.from_numpy(groups[lo:hi]), chunk=128)
        q.put((rank, lo, hi, D.numpy(), I.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_mining_world2_gloo():
    """configs[2] host logic on CPU: anchor chunks broadcast from their owner, global self ids mapped to the
    shard that holds them, groups broadcast with the chunk, per-rank candidates all-gathered and merged."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mining_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(11)
    n, d, k = 501, 10, 7
    emb = rng.standard_normal((n, d), dtype=np.float32)
    groups = (np.arange(n) // 3).astype(np.int32)
    D_ref, I_ref = O.search_ref(emb, emb, k, O.METRIC_IP, self_ids=np.arange(n), group_db=groups, group_q=groups)
    seen = 0
    for rank, lo, hi, D, I in outs:
        assert D.shape == (hi - lo, k)
        assert np.array_equal(I, I_ref[lo:hi]) and np.allclose(D, D_ref[lo:hi], atol=1e-5)
        assert not np.any(I == np.arange(lo, hi)[:, None])
        assert not np.any(groups[I] == groups[lo:hi, None])
        seen += hi - lo
    assert seen == n


def test_selfjoin_schedule_and_default_seed():
    from cloudvectordb_b200.mining import default_seed_rows, selfjoin_schedule
    for n in (1, 255, 1000, 70_000, 1_000_000, 2_000_000, 6_250_000, 50_000_000):
        seed = default_seed_rows(n)
        assert seed == 65536
        sched = selfjoin_schedule(n, 65536, seed)
        assert sched[0] == (0, min(seed, n))
        r = 0
        for r0, m in sched:
            assert r0 == r and 0 < m <= 65536 and (r0 == 0 or m <= r0)     # contiguous, in order, at most doubling
            assert r0 % 256 == 0
            r += m
        assert r == n


def test_sharded_symmetric_join_plan_covers_every_cross_shard_pair_exactly_once():
    """The block plan of the sharded symmetric self-join (mining.selfjoin_block_plan): block (h, begin) of rank g at
    step s scores the anchors of chunk s of shard h against rows [begin, n_g) of shard g in BOTH directions.  Over all
    steps and ranks every unordered pair of rows from two different shards must be scored exactly once, for odd and
    even rank counts, uneven and empty shards, ragged last chunks and shards with different numbers of chunks; and
    the ranks' work per step must be balanced for equal shards."""
    import itertools
    from cloudvectordb_b200.mining import selfjoin_block_plan
    cases = [([1024], 256, 256), ([1024, 1024], 256, 256), ([1000, 700], 256, 256), ([700, 1000], 256, 512),
             ([1024, 1024, 1024], 256, 256), ([900, 300, 1500], 256, 256), ([512, 512, 512, 512], 256, 256),
             ([1500, 260, 777, 1024], 256, 256), ([600, 0, 900, 300], 256, 256), ([640] * 5, 256, 256),
             ([400, 800, 1200, 300, 650, 1024], 256, 256), ([768] * 8, 256, 256), ([2048] * 8, 512, 256),
             ([300, 1300, 700, 256, 1024, 513, 900, 1100], 256, 256)]
    for counts, chunk, first in cases:
        G = len(counts)
        sched, plan = selfjoin_block_plan(counts, chunk, first)
        cover = {(a, b): np.zeros((counts[a], counts[b]), np.int32) for a, b in itertools.combinations(range(G), 2)}
        work = np.zeros((len(plan), G))
        for s_i, row in enumerate(plan):
            for g, blocks in enumerate(row):
                for h, begin in blocks:
                    assert h != g and begin % 128 == 0 and 0 <= begin < counts[g]
                    hr0, hm = sched[h][s_i]
                    assert hm > 0
                    work[s_i, g] += hm * (counts[g] - begin)
                    if h < g:
                        cover[(h, g)][hr0:hr0 + hm, begin:] += 1
                    else:
                        cover[(g, h)][begin:, hr0:hr0 + hm] += 1
        for (a, b), c in cover.items():
            assert (c == 1).all(), (counts, a, b, int(c.min()), int(c.max()))
        if len(set(counts)) == 1 and G > 1:      # equal shards: no rank waits for another inside a step
            for s_i in range(len(plan)):
                assert work[s_i].max() - work[s_i].min() <= chunk * chunk, (counts, s_i, work[s_i])
