"""Hard-negative mining for triplet construction (README.md:2 "building a very
large dataset of triplets"): a self-join top-k over the embedding matrix with
the anchor itself and its known positives (rows of the same group) excluded.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _C
from .index import IndexFlat, _is_torch

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def mine_hard_negatives(emb, k: int, groups=None, *, exclude_self: bool = True, metric: str = "ip",
                        storage: str = "bf16", device: int = 0, chunk: int = 65536, index=None,
                        row_offset: int = 0, queries=None, query_groups=None, symmetric: bool = False,
                        first_chunk: Optional[int] = None):
    """Top-k most similar rows of `emb` for every row (or for `queries`), never
    returning the anchor row itself nor any row with the anchor's group id.

    emb            [n, d] numpy / torch
    groups         [n] int group id per row (<0: no group) or None
    index          a prebuilt IndexFlat / ShardedIndex holding `emb` (optional)
    queries        [m, d] anchors if they are not all of `emb`; `row_offset` is
                   the global id of queries[0] (used for the self exclusion)
    returns        (D [m, k], I [m, k]) like search()
    """
    own = index is None
    if own:
        index = IndexFlat(int(emb.shape[1]), metric, storage, device)
        index.add(emb)
        if groups is not None:
            index.set_groups(groups)
    if symmetric:
        # every tile of X.X^T once, selected in both directions (half the flops); whole-matrix self-join only
        if queries is not None or not exclude_self or metric.lower() != "ip" or storage.lower() != "bf16" or row_offset:
            raise ValueError("symmetric=True is the plain self-join: IP metric, bf16 storage, all rows as anchors")
        D, I = mine_hard_negatives_symmetric(index, k, emb=emb, groups=groups, chunk=chunk, first_chunk=first_chunk)
        if own:
            index.close()
        if _is_torch(emb) and emb.is_cuda:
            return D, I
        if _is_torch(emb):
            return D.cpu(), I.cpu()
        return D.cpu().numpy(), I.cpu().numpy()
    if queries is None:
        queries, query_groups = emb, groups
    m = int(queries.shape[0])
    outs_d, outs_i = [], []
    for q0 in range(0, m, chunk):
        q1 = min(q0 + chunk, m)
        self_ids = np.arange(q0 + row_offset, q1 + row_offset, dtype=np.int64) if exclude_self else None
        gq = query_groups[q0:q1] if query_groups is not None else None
        D, I = index.search(queries[q0:q1], k, self_ids=self_ids, group_q=gq)
        outs_d.append(D)
        outs_i.append(I)
    if own:
        index.close()
    if _is_torch(outs_d[0]):
        return torch.cat(outs_d), torch.cat(outs_i)
    return np.concatenate(outs_d), np.concatenate(outs_i)


def default_seed_rows(n: int) -> int:
    """Seed size of the symmetric self-join when the caller names none: 65 536 rows (the largest the C ABI takes).
    The seed block is computed twice (plain searches in both directions) but at the full kernel rate, while the first
    chunks after a small seed run at 0.1-0.5 of it (cold column lists).  Measured: one GPU, 2M rows 3.94 / 3.77 /
    3.60 s and 6.25M rows - / 28.2 / 26.7 s with 8192 / 32768 / 65536 seed rows; two GPUs, 1M rows each: 2.64 s with
    32768 against 2.09 s with 65536.  (Corpora smaller than the seed are joined by the plain searches alone.)"""
    return 65536


def selfjoin_schedule(n: int, chunk: int = 65536, first: int = 65536):
    """Anchor chunks of the symmetric self-join: (row0, rows) in row order.  Chunk 0 is the SEED region (handled by
    plain searches, cvdb_selfjoin_seed); every later chunk is at most as long as the rows before it (so a row's
    column buffer sees about k new candidates per chunk) and at most `chunk`."""
    out, r = [], 0
    while r < n:
        m = min(first if r == 0 else min(chunk, r), n - r)
        out.append((r, m))
        r += m
    return out


def mine_hard_negatives_symmetric(index: IndexFlat, k: int, *, emb=None, groups=None, chunk: int = 65536,
                                  first_chunk: Optional[int] = None, stats: Optional[dict] = None):
    """Self-join top-k over ALL rows of `index` (IP, bf16 storage, groups already set with set_groups) that computes
    every tile of X.X^T once and selects in both directions (include/cvdb_b200.h, cvdb_selfjoin_*): half the flops
    of mine_hard_negatives().  The first `first_chunk` rows are the seed (plain searches warm the column side up).
    Rows whose column buffer overflowed (adversarial row orders) are recomputed exactly with a plain search; that
    needs `emb` (and `groups`).  Returns (D [n, k] f32, I [n, k] i64) on the GPU."""
    lib = _C.lib()
    n = index.ntotal
    dev = torch.device("cuda", index.device)
    stream = int(torch.cuda.current_stream(index.device).cuda_stream)
    if first_chunk is None:
        first_chunk = default_seed_rows(n)
    if chunk % 256 or first_chunk % 256 or first_chunk > 65536:
        raise ValueError("chunk sizes must be multiples of 256 (the seed at most 65536)")
    _C.check(lib.cvdb_selfjoin_begin(index._h, int(k), stream))
    try:
        keys = torch.empty((n, k), dtype=torch.int64, device=dev)
        sched = selfjoin_schedule(n, chunk, first_chunk)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(sched) + 1)] if stats is not None else None
        if ev:
            ev[0].record()
        _C.check(lib.cvdb_selfjoin_seed(index._h, sched[0][1], 0, keys.data_ptr(), stream))
        if ev:
            ev[1].record()
        for i, (r0, m) in enumerate(sched[1:]):
            _C.check(lib.cvdb_selfjoin_chunk(index._h, r0, m, 0, keys[r0:].data_ptr(), stream))
            if ev:
                ev[i + 2].record()
        D = torch.empty((n, k), dtype=torch.float32, device=dev)
        I = torch.empty((n, k), dtype=torch.int64, device=dev)
        _C.check(lib.cvdb_selfjoin_finish(index._h, 0, n, keys.data_ptr(), D.data_ptr(), I.data_ptr(), stream))
        del keys
        max_dirty = min(n, 1 << 22)
        rows = torch.empty((max_dirty,), dtype=torch.int32, device=dev)
        nd = _C.C.c_int64(0)
        _C.check(lib.cvdb_selfjoin_dirty(index._h, rows.data_ptr(), max_dirty, _C.C.byref(nd), stream))
    finally:
        _C.check(lib.cvdb_selfjoin_end(index._h))
    n_dirty = int(nd.value)
    if stats is not None:
        stats.update(chunks=len(sched), dirty_rows=n_dirty, schedule=[m for _, m in sched],
                     step_ms=[round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(len(sched))])
    if n_dirty:
        # exact recomputation of the rows that lost column candidates (plain row-direction search of the whole index)
        if emb is None:
            raise RuntimeError(f"{n_dirty} rows overflowed their column buffer and no `emb` was given to recompute them")
        if n_dirty > max_dirty:
            rows = torch.arange(n, device=dev, dtype=torch.int32)        # hopeless order: everything the plain way
        else:
            rows = rows[:n_dirty].sort().values
        emb_t = emb if _is_torch(emb) else torch.from_numpy(np.ascontiguousarray(emb))
        g_t = None if groups is None else torch.as_tensor(groups).to(dev).to(torch.int32)
        for q0 in range(0, rows.numel(), 65536):
            rr = rows[q0:q0 + 65536].long()
            q = emb_t[rr.to(emb_t.device)].to(dev)
            Dq, Iq = index.search(q, k, self_ids=rr.to(torch.int32), group_q=None if g_t is None else g_t[rr])
            D[rr], I[rr] = Dq, Iq
    return D, I


def mine_hard_negatives_sharded(index, local_emb, k: int, local_groups=None, *, exclude_self: bool = True,
                                chunk: int = 65536, max_chunks: Optional[int] = None):
    """Self-join over a ShardedIndex (one process per GPU): every rank owns the rows
    `local_emb` it added with add_local().  Anchor chunks are broadcast from their
    owner, searched on every shard, merged (ShardedIndex.search) and kept by the owner.

    Returns (D, I) for this rank's rows (global ids), like search().
    `max_chunks` bounds the number of anchor chunks per owner (benchmarks)."""
    import torch.distributed as dist
    rank, world = index.rank, index.world
    dev = local_emb.device
    counts = index._counts
    outs_d, outs_i = [], []
    for owner in range(world):
        n_owner = counts[owner]
        base = sum(counts[:owner])
        n_chunks = (n_owner + chunk - 1) // chunk
        if max_chunks is not None:
            n_chunks = min(n_chunks, max_chunks)
        for c in range(n_chunks):
            q0, q1 = c * chunk, min((c + 1) * chunk, n_owner)
            if rank == owner:
                q = local_emb[q0:q1].contiguous()
                g = local_groups[q0:q1].to(torch.int32).contiguous() if local_groups is not None else None
            else:
                q = torch.empty((q1 - q0, local_emb.shape[1]), dtype=local_emb.dtype, device=dev)
                g = torch.empty((q1 - q0,), dtype=torch.int32, device=dev) if local_groups is not None else None
            if world > 1:
                src = dist.get_global_rank(index.group, owner) if index.group is not None else owner
                dist.broadcast(q, src=src, group=index.group)
                if g is not None:
                    dist.broadcast(g, src=src, group=index.group)
            self_ids = torch.arange(base + q0, base + q1, device=dev) if exclude_self else None
            D, I = index.search(q, k, self_ids=self_ids, group_q=g)
            if rank == owner:
                outs_d.append(D)
                outs_i.append(I)
    if not outs_d:
        return (torch.empty((0, k), dtype=torch.float32, device=dev), torch.empty((0, k), dtype=torch.int64, device=dev))
    return torch.cat(outs_d), torch.cat(outs_i)


def selfjoin_block_plan(counts, chunk: int = 65536, first_chunk: int = 65536):
    """Host-side plan of the sharded symmetric self-join (pure function of the shard sizes; no GPU needed).

    Returns (sched, plan): sched[h] is shard h's chunk schedule (selfjoin_schedule), plan[s][g] the cross blocks rank
    g computes at step s as (h, row_begin) tuples: the anchors are chunk s of shard h, the database rows are rows
    [row_begin, counts[g]) of shard g, both directions at once.  The diagonal blocks (shard g's own chunks) are not
    listed.  Every unordered pair of rows from two different shards is covered by exactly one block of the plan
    (tests/test_host_logic.py enumerates this for 1..8 ranks and uneven shards); see
    mine_hard_negatives_sharded_symmetric for the rule."""
    G = len(counts)
    sched = [selfjoin_schedule(c, chunk, first_chunk) for c in counts]
    steps = max((len(sc) for sc in sched), default=0)
    full_d = (G - 1) // 2
    plan = []
    for s_i in range(steps):
        ch = [sc[s_i] if s_i < len(sc) else (0, 0) for sc in sched]
        row = []
        for g in range(G):
            blocks = []
            r0, m = ch[g]

            def want(h):
                return ch[h][1] > 0 and counts[g] > 0

            for dd in range(1, full_d + 1):
                h = (g + dd) % G
                if want(h):
                    blocks.append((h, 0))
            if G % 2 == 0 and G > 1 and s_i < len(sched[g]):
                # the distance-G/2 pair, split by chunk index (the chunk-s x chunk-s corner alternates)
                h = (g + G // 2) % G
                begin = r0 if (g > h) == (s_i % 2 == 0) else r0 + m
                if begin < counts[g] and want(h):
                    blocks.append((h, begin))
            row.append(blocks)
        plan.append(row)
    return sched, plan


def mine_hard_negatives_sharded_symmetric(index, local_emb, k: int, local_groups=None, *, chunk: int = 65536,
                                          first_chunk: Optional[int] = None, stats: Optional[dict] = None):
    """The symmetric self-join over a ShardedIndex (one process per GPU, NCCL): every unordered pair of rows is
    scored once in the whole job.

    Shard pairs.  The block (anchors of shard h) x (rows of shard g) and its transpose hold the same scores, so only
    ONE of the two ranks computes it -- in both directions at once (cvdb_selfjoin_cross): the row direction gives
    the anchors of h their candidates among g's rows (sent back to h), the column direction gives g's rows their
    candidates among h's anchors (kept in g's column lists).  Rank g takes the anchors of the owners at distance
    1 .. (G-1)/2 after it (mod G).  With an even number of ranks the pair at distance G/2 is split like the diagonal
    block, by chunk index, so that both ranks of the pair have the same work in every step (the all_to_all at the
    end of a step makes ranks wait for each other): at step s each rank of the pair scores the OTHER rank's chunk s
    against its own rows after its own chunk s, and one of the two (the upper rank at even steps, the lower at odd
    ones) includes its chunk s as well -- a pair (a in chunk s of the lower, b in chunk t of the upper) is scored
    on the upper rank when t > s, on the lower rank when t < s, and by that parity rule when t == s: exactly once.
    The diagonal block is the single-GPU symmetric join (cvdb_selfjoin_chunk).  Every rank does (G/2)/G of the plain
    job's flops.  The plan itself is selfjoin_block_plan (a pure function, checked exhaustively on CPU).

    Steps.  All owners walk the same chunk schedule (selfjoin_schedule: a seed chunk handled by plain searches, then
    chunks that at most double); step s: all-gather chunk s of every
    owner (vectors, groups), each rank runs its 1 + (G-1)/2 (+1/2) blocks, the row-direction keys go to the
    anchors' owners with ONE all_to_all, and the owner merges the G lists into its running row keys.  After the last
    step cvdb_selfjoin_finish merges row keys and column lists.  Rows whose column buffer overflowed anywhere are
    recomputed by the plain sharded search (collectively).

    Returns (D [n_local, k] f32, I [n_local, k] i64 global ids) for this rank's rows, on the GPU."""
    import torch.distributed as dist
    lib = _C.lib()
    G, g = index.world, index.rank
    local = index.local
    dev = local_emb.device
    counts = list(index._counts)
    bases = [sum(counts[:h]) for h in range(G)]
    n_loc, d = counts[g], int(local_emb.shape[1])
    stream = int(torch.cuda.current_stream(dev.index).cuda_stream)
    if first_chunk is None:
        first_chunk = default_seed_rows(max(counts))
    if chunk % 256 or first_chunk % 256 or first_chunk > 65536:
        raise ValueError("chunk sizes must be multiples of 256 (the seed at most 65536)")
    sched, plan = selfjoin_block_plan(counts, chunk, first_chunk)
    seed = [sc[0][1] if sc else 0 for sc in sched]       # rows [0, seed[h]) of shard h: handled by plain searches
    steps = len(plan)
    has_groups = local_groups is not None
    grp_dev = local_groups.to(dev).to(torch.int32).contiguous() if has_groups else None
    _C.check(lib.cvdb_selfjoin_begin(local._h, int(k), stream))
    n_blocks = 0
    marks = []   # (phase, event) pairs when the caller wants timings

    def mark(name):
        if stats is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))
    mark("begin")
    try:
        row_keys = torch.zeros((n_loc, k), dtype=torch.int64, device=dev)
        if n_loc:   # column lists of my later rows start with their top-k among my seed anchors
            _C.check(lib.cvdb_selfjoin_seed(local._h, seed[g], bases[g], None, stream))
        mark("seed_columns")
        for s_i in range(steps):
            ch = [sc[s_i] if s_i < len(sc) else (0, 0) for sc in sched]
            m_max = max(m for _, m in ch)
            r0, m = ch[g]
            q_mine = torch.zeros((m_max, d), dtype=local_emb.dtype, device=dev)
            g_mine = torch.full((m_max,), -1, dtype=torch.int32, device=dev)
            if m:
                q_mine[:m] = local_emb[r0:r0 + m]
                if has_groups:
                    g_mine[:m] = grp_dev[r0:r0 + m]
            Q = torch.empty((G, m_max, d), dtype=local_emb.dtype, device=dev)
            GR = torch.empty((G, m_max), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(Q.view(G * m_max, d), q_mine, group=index.group)
            dist.all_gather_into_tensor(GR.view(-1), g_mine, group=index.group)
            send = torch.zeros((G, m_max, k), dtype=torch.int64, device=dev)
            mark("gather")
            if m and s_i > 0:   # the diagonal block (chunk 0, the seed anchors, gets the plain sharded search below)
                _C.check(lib.cvdb_selfjoin_chunk(local._h, r0, m, bases[g], send[g].data_ptr(), stream))
                n_blocks += 1
            mark("diagonal")

            def cross(h, row_begin, row_end):
                hr0, hm = ch[h]
                if hm == 0 or n_loc == 0:
                    return 0
                ids = (torch.arange(hm, device=dev, dtype=torch.int64) + (bases[h] + hr0)).to(torch.int32)
                dt = {torch.float32: _C.DTYPE_F32, torch.bfloat16: _C.DTYPE_BF16, torch.float16: _C.DTYPE_F16}[Q.dtype]
                qh = Q[h, :hm].contiguous()
                gh = GR[h, :hm].contiguous() if has_groups else None
                _C.check(lib.cvdb_selfjoin_cross(local._h, qh.data_ptr(), hm, dt, ids.data_ptr(),
                                                 gh.data_ptr() if has_groups else None, int(row_begin), int(row_end),
                                                 seed[g], bases[g], send[h].data_ptr(), stream))
                return 1

            for h, begin in plan[s_i][g]:
                n_blocks += cross(h, begin, 0)
            mark("cross")
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv.view(G * m_max, k), send.view(G * m_max, k), group=index.group)
            mark("all_to_all")
            if m and s_i > 0:
                merged = torch.empty((m_max, k), dtype=torch.int64, device=dev)
                _C.check(lib.cvdb_merge_keys(recv.data_ptr(), m_max, G, k, k, merged.data_ptr(), stream))
                row_keys[r0:r0 + m] = merged[:m]
            mark("merge")
        D = torch.empty((n_loc, k), dtype=torch.float32, device=dev)
        I = torch.empty((n_loc, k), dtype=torch.int64, device=dev)
        max_dirty = max(1, min(n_loc, 1 << 22))
        rows = torch.empty((max_dirty,), dtype=torch.int32, device=dev)
        nd = _C.C.c_int64(0)
        if n_loc > seed[g]:
            _C.check(lib.cvdb_selfjoin_finish(local._h, seed[g], n_loc - seed[g], row_keys[seed[g]:].data_ptr(),
                                              D[seed[g]:].data_ptr(), I[seed[g]:].data_ptr(), stream))
        if n_loc:
            _C.check(lib.cvdb_selfjoin_dirty(local._h, rows.data_ptr(), max_dirty, _C.C.byref(nd), stream))
        del row_keys
        mark("finish")
    finally:
        _C.check(lib.cvdb_selfjoin_end(local._h))
    # ---- rows that lost column candidates anywhere: recomputed by the plain sharded search, owner by owner
    n_dirty = int(nd.value)
    if n_dirty > max_dirty:
        rows, n_dirty = torch.arange(n_loc, device=dev, dtype=torch.int32), n_loc
    rows = rows[:n_dirty].sort().values.long()
    # the seed anchors of every shard take the same plain path
    rows = torch.unique(torch.cat([torch.arange(seed[g], device=dev, dtype=torch.int64), rows]))
    n_plain = int(rows.numel())
    tot = torch.tensor([n_plain], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(tot) for _ in range(G)]
    dist.all_gather(all_n, tot, group=index.group)
    all_n = [int(t.item()) for t in all_n]
    if stats is not None:
        stats.update(steps=steps, blocks=n_blocks, dirty_rows=n_dirty, seed_rows=seed[g], plain_rows_all_ranks=sum(all_n))
    for owner in range(G):
        src = dist.get_global_rank(index.group, owner) if index.group is not None else owner
        for q0 in range(0, all_n[owner], 65536):
            mq = min(65536, all_n[owner] - q0)
            if g == owner:
                rr = rows[q0:q0 + mq]
                q = local_emb[rr].contiguous()
                ids = (rr + bases[g]).contiguous()
                gq = grp_dev[rr].contiguous() if has_groups else None
            else:
                q = torch.empty((mq, d), dtype=local_emb.dtype, device=dev)
                ids = torch.empty((mq,), dtype=torch.int64, device=dev)
                gq = torch.empty((mq,), dtype=torch.int32, device=dev) if has_groups else None
            dist.broadcast(q, src=src, group=index.group)
            dist.broadcast(ids, src=src, group=index.group)
            if has_groups:
                dist.broadcast(gq, src=src, group=index.group)
            Dq, Iq = index.search(q, k, self_ids=ids, group_q=gq)
            if g == owner:
                D[rr], I[rr] = Dq, Iq
    if stats is not None:
        mark("plain_rows")
        torch.cuda.synchronize(dev)
        phase_ms = {}
        for (_, e0), (name, e1) in zip(marks, marks[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + e0.elapsed_time(e1)
        stats["phase_ms"] = {n_: round(v, 2) for n_, v in phase_ms.items()}
    return D, I


def build_triplets(D, I, positives, *, skip_top: int = 0, per_anchor: int = 1, metric: str = "ip",
                   limit: Optional[float] = None, anchor_base: int = 0):
    """(anchor, positive, hard negative) triplets from mined neighbours.

    D, I        [n, k] as returned by mine_hard_negatives (positives and the anchor already excluded)
    positives   [n] id of one positive per anchor (< 0: the anchor yields no triplet)
    skip_top    ignore the first ranks (the very closest rows are often unlabeled positives)
    limit       drop rows scoring above it (IP) / closer than it (L2): a margin against false negatives
    returns     int64 [n, per_anchor, 3]; unused slots are -1.  Runs on the GPU (cvdb_build_triplets)."""
    as_numpy = not _is_torch(I)
    dev = I.device if (_is_torch(I) and I.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    It = torch.as_tensor(I).to(dev, torch.int64).contiguous()
    Dt = torch.as_tensor(D).to(dev, torch.float32).contiguous()
    pt = torch.as_tensor(positives).to(dev, torch.int64).contiguous()
    n, k = It.shape
    if pt.numel() != n:
        raise ValueError("positives must have one entry per anchor")
    out = torch.empty((n, per_anchor, 3), dtype=torch.int64, device=dev)
    metric_code = {"ip": _C.METRIC_IP, "l2": _C.METRIC_L2}[metric.lower()]
    _C.check(_C.lib().cvdb_build_triplets(It.data_ptr(), Dt.data_ptr(), n, k, pt.data_ptr(), int(anchor_base),
                                          int(skip_top), int(per_anchor), metric_code,
                                          float(limit if limit is not None else 0.0), int(limit is not None),
                                          out.data_ptr(), int(torch.cuda.current_stream(dev.index).cuda_stream)))
    if as_numpy:
        return out.cpu().numpy()
    return out if (_is_torch(I) and I.is_cuda) else out.cpu()
