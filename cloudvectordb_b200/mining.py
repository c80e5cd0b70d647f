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
                        row_offset: int = 0, queries=None, query_groups=None):
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
