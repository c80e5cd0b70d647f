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
                        row_offset: int = 0, queries=None, query_groups=None, symmetric: bool = False):
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
        D, I = mine_hard_negatives_symmetric(index, k, emb=emb, groups=groups, chunk=chunk)
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


def selfjoin_schedule(n: int, chunk: int = 65536, first: int = 256):
    """Anchor chunks of the symmetric self-join: (row0, rows) in row order, sizes 256, 256, 512, 1024, ... (each at
    most the number of rows before it, so a row's column buffer sees about k new candidates per chunk) up to `chunk`."""
    out, r = [], 0
    while r < n:
        m = min(first if r == 0 else min(chunk, r), n - r)
        out.append((r, m))
        r += m
    return out


def mine_hard_negatives_symmetric(index: IndexFlat, k: int, *, emb=None, groups=None, chunk: int = 65536,
                                  first_chunk: int = 256, stats: Optional[dict] = None):
    """Self-join top-k over ALL rows of `index` (IP, bf16 storage, groups already set with set_groups) that computes
    every tile of X.X^T once and selects in both directions (include/cvdb_b200.h, cvdb_selfjoin_*): half the flops
    of mine_hard_negatives().  Rows whose column buffer overflowed (adversarial row orders) are recomputed exactly
    with a plain search; that needs `emb` (and `groups`).  Returns (D [n, k] f32, I [n, k] i64) on the GPU."""
    lib = _C.lib()
    n = index.ntotal
    dev = torch.device("cuda", index.device)
    stream = int(torch.cuda.current_stream(index.device).cuda_stream)
    if chunk % 256 or first_chunk % 256:
        raise ValueError("chunk sizes must be multiples of 256")
    _C.check(lib.cvdb_selfjoin_begin(index._h, int(k), stream))
    try:
        keys = torch.empty((n, k), dtype=torch.int64, device=dev)
        sched = selfjoin_schedule(n, chunk, first_chunk)
        for r0, m in sched:
            _C.check(lib.cvdb_selfjoin_chunk(index._h, r0, m, keys[r0:].data_ptr(), stream))
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
        stats.update(chunks=len(sched), dirty_rows=n_dirty)
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
