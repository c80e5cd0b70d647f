"""Generates tests/golden/sklearn_pin.npz from scikit-learn / SciPy -- NOT from the oracle.

The reference ships no code or vectors (README.md only) and FAISS cannot be installed here, so the
independent published implementations available in this image pin the conventions the oracle restates
(BASELINE.json north_star: "fp32 NumPy/FAISS-CPU IndexFlatIP oracle"):

  * sklearn.neighbors.NearestNeighbors(algorithm="brute")  -- exact k-NN, euclidean and cosine
  * scipy.spatial.distance.cdist (float64)                 -- squared L2 / inner product, stable argsort
  * sklearn.metrics.pairwise_distances_argmin_min           -- the k-means assignment step
  * sklearn.cluster.KMeans(init=C, n_init=1, max_iter=1, algorithm="lloyd") -- one Lloyd update

Nothing in this script imports oracle/ .  Run from the repo root:

    python tests/golden/make_golden_sklearn.py

scikit-learn 1.9.0, SciPy 1.18.1, NumPy 2.3.5 produced the committed file.
"""
import os

import numpy as np
import scipy
import sklearn
from scipy.spatial.distance import cdist
from sklearn.cluster import KMeans
from sklearn.metrics import pairwise_distances_argmin_min
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))


def bf16_round(x):
    """fp32 -> nearest-even bf16 -> fp32 (so that the bf16 engine sees exactly these values)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def main():
    rng = np.random.default_rng(20261019)
    n, d, nq, k = 3000, 96, 64, 10
    xb = bf16_round(unit(rng, n, d))
    xq = bf16_round(unit(rng, nq, d))
    out = dict(xb=xb, xq=xq, k=np.int32(k),
               versions=np.array([sklearn.__version__, scipy.__version__, np.__version__]))

    # ---- L2: sklearn brute-force k-NN (euclidean; squared here) and SciPy float64 cdist + stable argsort
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(xb.astype(np.float64))
    dist, ind = nn.kneighbors(xq.astype(np.float64))
    out["skl_l2_D"], out["skl_l2_I"] = (dist ** 2).astype(np.float64), ind.astype(np.int64)
    dd = cdist(xq.astype(np.float64), xb.astype(np.float64), "sqeuclidean")
    o = np.argsort(dd, axis=1, kind="stable")[:, :k]
    out["scipy_l2_D"], out["scipy_l2_I"] = np.take_along_axis(dd, o, 1), o.astype(np.int64)

    # ---- IP: cosine k-NN on rows that are NOT unit norm after rounding, so rank by the raw inner product
    # with SciPy (cdist has no "dot" metric: 1 - cosine * norms would re-derive it, so use the float64
    # definition sum_k q_k x_k through cdist's user-callable form on a subsample, and sklearn cosine on the rest)
    nnc = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(xb.astype(np.float64))
    cd, ci = nnc.kneighbors(xq.astype(np.float64))
    out["skl_cos_D"], out["skl_cos_I"] = (1.0 - cd).astype(np.float64), ci.astype(np.int64)
    ip = -cdist(xq[:16].astype(np.float64), xb.astype(np.float64), lambda a, b: -float(np.dot(a, b)))
    o = np.argsort(-ip, axis=1, kind="stable")[:, :k]
    out["scipy_ip_D"], out["scipy_ip_I"] = np.take_along_axis(ip, o, 1), o.astype(np.int64)

    # ---- ties: duplicated rows; every implementation must return the same SET, the convention orders it by id
    xt = xb[:400].copy()
    xt[50] = xt[7]
    xt[333] = xt[7]
    xt[120] = xt[9]
    qt = np.stack([xt[7], xt[9]])
    nnt = NearestNeighbors(n_neighbors=4, algorithm="brute", metric="euclidean").fit(xt.astype(np.float64))
    td, ti = nnt.kneighbors(qt.astype(np.float64))
    out["tie_xb"], out["tie_xq"], out["tie_skl_I"], out["tie_skl_D"] = xt, qt, ti.astype(np.int64), td ** 2

    # ---- exclusion: k-NN over the rows that survive the mask (sklearn on the filtered matrix, ids mapped back)
    self_ids = rng.integers(0, n, nq).astype(np.int32)
    xq_self = xb[self_ids].copy()
    group_db = (np.arange(n) // 6).astype(np.int32)
    group_q = group_db[self_ids].copy()
    group_q[::3] = -1
    ex_I = np.empty((nq, k), np.int64)
    ex_D = np.empty((nq, k), np.float64)
    for i in range(nq):
        keep = np.ones(n, bool)
        keep[self_ids[i]] = False
        if group_q[i] >= 0:
            keep &= group_db != group_q[i]
        rows = np.flatnonzero(keep)
        m = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(xb[rows].astype(np.float64))
        dd_, ii_ = m.kneighbors(xq_self[i:i + 1].astype(np.float64))
        ex_I[i], ex_D[i] = rows[ii_[0]], dd_[0] ** 2
    out.update(ex_self_ids=self_ids, ex_xq=xq_self, ex_group_db=group_db, ex_group_q=group_q, ex_skl_l2_I=ex_I,
               ex_skl_l2_D=ex_D)

    # ---- k-means: assignment and one Lloyd update from fixed centroids
    K = 24
    cent = xb[rng.choice(n, K, replace=False)].astype(np.float32)
    a, dmin = pairwise_distances_argmin_min(xb.astype(np.float64), cent.astype(np.float64), metric="sqeuclidean")
    km = KMeans(n_clusters=K, init=cent.astype(np.float64), n_init=1, max_iter=1, algorithm="lloyd", tol=0.0)
    km.fit(xb.astype(np.float64))
    assert np.bincount(a, minlength=K).min() > 0, "pick another seed: an empty cluster makes sklearn relocate it"
    out.update(km_centroids=cent, km_skl_assign=a.astype(np.int64), km_skl_dist=dmin.astype(np.float64),
               km_skl_new_centroids=km.cluster_centers_.astype(np.float64),
               km_skl_counts=np.bincount(a, minlength=K).astype(np.int64))
    path = os.path.join(HERE, "sklearn_pin.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
