"""Generates tests/golden/*.npz from the NumPy oracle (oracle/flat_oracle.py).

The reference ships no code or vectors (README.md only), so these fixtures pin
the ORACLE (against accidental edits) and give the GPU tests fixed inputs with
known answers; they are not outputs of the reference.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flat_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261018)
    n, d, nq, k = 700, 40, 24, 7
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    # values exactly representable in bf16, so the bf16 engine sees the same numbers
    xb, xq = O.bf16_round(xb), O.bf16_round(xq)
    # duplicated rows: exercises the lower-id-first tie rule
    xb[100] = xb[7]
    xb[650] = xb[7]
    xq[3] = xb[7]
    self_ids = rng.integers(0, n, nq).astype(np.int32)
    group_db = (np.arange(n) // 5).astype(np.int32)
    group_q = group_db[self_ids].copy()
    group_q[::4] = -1
    out = dict(xb=xb, xq=xq, k=np.int32(k), self_ids=self_ids, group_db=group_db, group_q=group_q)
    for name, metric in (("ip", O.METRIC_IP), ("l2", O.METRIC_L2)):
        D, I = O.search_ref(xb, xq, k, metric)
        out[f"D_{name}"], out[f"I_{name}"] = D, I
        D, I = O.search_ref(xb, xq, k, metric, self_ids=self_ids, group_db=group_db, group_q=group_q)
        out[f"D_{name}_excl"], out[f"I_{name}_excl"] = D, I
    # k-means: one Lloyd step from fixed centroids
    K = 16
    cent = xb[:K].copy()
    a, dist = O.kmeans_assign_ref(xb, cent)
    newc, counts, _ = O.kmeans_update_ref(xb, a, cent)
    out.update(km_centroids=cent, km_assign=a, km_dist=dist, km_new_centroids=newc, km_counts=counts)
    np.savez_compressed(os.path.join(HERE, "flat_small.npz"), **out)
    print("wrote", os.path.join(HERE, "flat_small.npz"))


if __name__ == "__main__":
    main()
