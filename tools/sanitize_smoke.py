"""Tiny run of every kernel family for compute-sanitizer (memcheck): all four GEMM+top-k variants,
the grouped IVF kernel, exact storage, exclusion, k-means update, merges."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cloudvectordb_b200 import IndexFlat, IndexIVFFlat, Kmeans, merge_topk  # noqa: E402

rng = np.random.default_rng(0)


def rows(n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


for metric in ("ip", "l2"):
    for d in (40, 768):
        xb, xq = rows(1500, d), rows(300, d)
        idx = IndexFlat(d, metric, "bf16")
        idx.add(xb)
        idx.set_groups((np.arange(1500) // 4).astype(np.int32))
        for v in (1, 2, 3, 4):
            if v == 4 and metric == "l2" and d == 768:
                continue
            for k in (1, 10, 50):
                idx.search(xq, k, force_variant=v, self_ids=np.arange(300), group_q=(np.arange(300) // 4).astype(np.int32))
        idx.close()
ex = IndexFlat(100, "l2", "exact")
ex.add(rows(900, 100))
D, I = ex.search(rows(70, 100), 10)
ex.close()
merge_topk(np.stack([D, D]), np.stack([I, I + 1000]), 10, "l2")
ivf = IndexIVFFlat(64, 16, "l2")
xb = rows(3000, 64)
ivf.train(None, centroids=xb[:16])
ivf.add(xb)
ivf.search(rows(200, 64), 10, nprobe=4)
ivf.search(rows(3, 64), 1, nprobe=16)
ivf.close()
km = Kmeans(32, 8, niter=2)
km.train(rows(2000, 32))
torch.cuda.synchronize()
print("SANITIZE_SMOKE_OK")
