"""Hard-negative mining for triplet construction (README.md:2 "building a very
large dataset of triplets"): a self-join top-k over the embedding matrix with
the anchor itself and its known positives (rows of the same group) excluded.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

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
