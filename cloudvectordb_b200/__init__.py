"""B200-native exact nearest-neighbour engine (FAISS-style surface).

    index = IndexFlatIP(d)          # or IndexFlatL2(d); storage="bf16" | "exact"
    index.add(x)                    # [n, d] float32 / bfloat16, numpy or torch (host or cuda)
    D, I = index.search(q, k)       # D [nq, k] f32, I [nq, k] i64

Everything numeric runs in hand-written sm_100a kernels behind the C ABI in
``include/cvdb_b200.h``; this package only moves pointers.
"""
from .index import IndexFlat, IndexFlatIP, IndexFlatL2, merge_topk  # noqa: F401
from .ivf import IndexIVFFlat  # noqa: F401
from .kmeans import Kmeans  # noqa: F401
from .mining import (build_triplets, mine_hard_negatives, mine_hard_negatives_sharded,  # noqa: F401
                     mine_hard_negatives_sharded_symmetric, mine_hard_negatives_symmetric, selfjoin_schedule)
from .sharded import ShardedIndex, ShardedIVFFlat  # noqa: F401

__all__ = ["IndexIVFFlat", "IndexFlat", "IndexFlatIP", "IndexFlatL2", "merge_topk", "Kmeans", "build_triplets", "mine_hard_negatives",
           "mine_hard_negatives_sharded", "mine_hard_negatives_sharded_symmetric", "mine_hard_negatives_symmetric", "selfjoin_schedule", "ShardedIndex", "ShardedIVFFlat"]
