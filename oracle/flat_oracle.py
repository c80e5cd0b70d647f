"""CPU oracle for the exact nearest-neighbour hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED AGAINST THE REFERENCE, PINNED AGAINST INDEPENDENT IMPLEMENTATIONS.
The reference (dorenwick/CloudVectorDB) ships nothing but ``README.md:1-2`` (a
title and one sentence naming the four pipeline stages: triplets -> encoder ->
embeddings -> vector DB).  There is no reference source, test, golden vector or
fixture for this path, so nothing OF THE REFERENCE can pin this file.  It
restates the *published* algorithm that BASELINE.json's ``north_star`` names as
the oracle - an fp32 NumPy restatement of FAISS ``IndexFlatIP`` /
``IndexFlatL2`` / ``Kmeans`` semantics (FAISS itself is not installable here) -
and is pinned against
  (i)   independent published implementations of the same exact computations:
        scikit-learn ``NearestNeighbors(algorithm="brute")`` (euclidean and
        cosine), SciPy ``cdist`` in float64 with a stable argsort,
        ``pairwise_distances_argmin_min`` and one ``KMeans(init=C, n_init=1,
        max_iter=1, algorithm="lloyd")`` step - both live and through the
        committed fixture ``tests/golden/sklearn_pin.npz`` written by
        ``tests/golden/make_golden_sklearn.py`` (which does not import this
        file): ``tests/test_oracle_thirdparty.py``;
  (ii)  the naive double loop in ``oracle/naive_oracle.c``;
  (iii) the fixtures in ``tests/golden/flat_small.npz`` (generated from this
        file by ``make_golden.py``; they guard against accidental edits only).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module.  Nothing under
``cloudvectordb_b200/`` does; the product path has no CPU fallback.

Conventions restated here:
  * ``search(q, k) -> (D [nq,k] f32, I [nq,k] i64)``
  * IP: k largest scores, descending.  L2: k smallest *squared* distances, ascending.
  * ties broken by lower database index first.
  * fewer than k valid rows -> I = -1 and D = -inf (IP) / +inf (L2).
  * exclusion (hard-negative mining, README.md:2 "dataset of triplets"):
    a database row j is dropped for query i when ``j == self_ids[i]`` or
    ``group_db[j] == group_q[i]`` (group id < 0 means "no group").
  * k-means (README.md:2 "building the vectordb"): assign = argmin squared L2
    (lower centroid id wins ties), update = mean of assigned points, empty
    clusters keep their previous centroid.
"""
from __future__ import annotations

import numpy as np

METRIC_IP = 0
METRIC_L2 = 1


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return the value as fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit pattern (uint16), round-to-nearest-even."""
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def _select_rows_slow(scores, rows_todo, base, kk, vals, ids):
    """Per-row selection with the full tie rule (rows holding NaN, or too few finite candidates)."""
    n = scores.shape[1]
    for i in rows_todo:
        row = scores[i]
        o = np.lexsort((np.arange(n), -row))[:kk]      # NaN sorts last
        vals[i] = row[o]
        ids[i] = o + base


_BLOCK = 256


def _select_topk(scores: np.ndarray, base: int, k: int):
    """Top-k of each row of ``scores`` (larger is better), ties -> lower index.

    Returns (vals [nq,kk], ids [nq,kk]) with kk = min(k, ncols), sorted.

    Vectorised so that the oracle is sgemm-bound, not selection-bound: the kk-th largest of the
    per-256-column block maxima of a row is a lower bound of the row's kk-th best score (kk
    different columns reach it), so one compare pass against that bound leaves a handful of
    candidates per row; those are ordered by (row, score descending, column ascending) in ONE
    lexsort, and the first kk of every row are the answer, ties resolved by index.
    """
    nq, n = scores.shape
    kk = min(k, n)
    vals = np.empty((nq, kk), np.float32)
    ids = np.empty((nq, kk), np.int64)
    if nq == 0 or kk == 0:
        return vals, ids
    nb = n // _BLOCK
    if nb >= 4 * kk:
        blocks = scores[:, :nb * _BLOCK].reshape(nq, nb, _BLOCK)
        bm = blocks.max(axis=2)
        thr = np.partition(bm, nb - kk, axis=1)[:, nb - kk]
        # only blocks whose maximum reaches the bound can hold a candidate: look inside those alone
        br, bc = np.nonzero(bm >= thr[:, None])
        inner = blocks[br, bc] >= thr[br, None]                     # [pairs, _BLOCK]
        pi, pj = np.nonzero(inner)
        rows, cols = br[pi], bc[pi] * _BLOCK + pj
        if nb * _BLOCK < n:                                         # the columns after the last whole block
            tr, tc = np.nonzero(scores[:, nb * _BLOCK:] >= thr[:, None])
            rows, cols = np.concatenate([rows, tr]), np.concatenate([cols, tc + nb * _BLOCK])
    else:
        rows, cols = np.divmod(np.arange(nq * n), n)
    v = scores[rows, cols]
    o = np.lexsort((cols, -v, rows))
    rows, cols, v = rows[o], cols[o], v[o]
    starts = np.searchsorted(rows, np.arange(nq + 1))
    counts = np.diff(starts)
    keep = (np.arange(len(rows)) - starts[rows]) < kk
    full = counts >= kk
    if full.all():
        ids[:] = cols[keep].reshape(nq, kk) + base
        vals[:] = v[keep].reshape(nq, kk)
    else:
        ok = full[rows] & keep
        ids[full] = cols[ok].reshape(-1, kk) + base
        vals[full] = v[ok].reshape(-1, kk)
    # NaN scores poison the bound and the ordering: such rows take the per-row path
    bad = ~full | np.isnan(vals).any(axis=1)
    if bad.any():
        _select_rows_slow(scores, np.flatnonzero(bad), base, kk, vals, ids)
    return vals, ids


def _merge_sorted(vals_a, ids_a, vals_b, ids_b, k):
    vals = np.concatenate([vals_a, vals_b], axis=1)
    ids = np.concatenate([ids_a, ids_b], axis=1)
    kk = min(k, vals.shape[1])
    # invalid entries carry id -1 and value -inf; push them last
    key_id = np.where(ids < 0, np.iinfo(np.int64).max, ids)
    order = np.lexsort((key_id, -vals), axis=1)[:, :kk]
    return (np.take_along_axis(vals, order, axis=1).astype(np.float32, copy=False),
            np.take_along_axis(ids, order, axis=1))


def search_ref(xb, xq, k, metric=METRIC_IP, self_ids=None, group_db=None,
               group_q=None, block_rows=262144, dtype=np.float32):
    """Exact brute-force search.  fp32 sgemm scores, blocked over database rows.

    ``dtype=np.float64`` gives the tie-arbitration pass of SURVEY.md 8(c).
    """
    xb = np.ascontiguousarray(xb, dtype=dtype)
    xq = np.ascontiguousarray(xq, dtype=dtype)
    nq, d = xq.shape
    n = xb.shape[0]
    assert xb.ndim == 2 and (n == 0 or xb.shape[1] == d)
    best_v = np.full((nq, 0), -np.inf, np.float32)
    best_i = np.full((nq, 0), -1, np.int64)
    if metric == METRIC_L2:
        qn = np.einsum("ij,ij->i", xq, xq)
    for r0 in range(0, n, block_rows):
        blk = xb[r0:r0 + block_rows]
        s = xq @ blk.T
        if metric == METRIC_L2:
            bn = np.einsum("ij,ij->i", blk, blk)
            # larger-is-better form of the squared distance
            s = -(qn[:, None] - 2.0 * s + bn[None, :])
        s = s.astype(np.float32, copy=False)
        if self_ids is not None:
            sid = np.asarray(self_ids, np.int64) - r0
            ok = (sid >= 0) & (sid < blk.shape[0])
            s[np.nonzero(ok)[0], sid[ok]] = -np.inf
        if group_db is not None:
            gq = np.asarray(group_q, np.int64)
            gb = np.asarray(group_db, np.int64)[r0:r0 + block_rows]
            s[(gq[:, None] == gb[None, :]) & (gq[:, None] >= 0)] = -np.inf
        v, i = _select_topk(s, r0, k)
        best_v, best_i = _merge_sorted(best_v, best_i, v, i, k)
    # drop excluded rows that leaked in as -inf, then pad to k
    bad = np.isneginf(best_v)
    best_i[bad] = -1
    if best_v.shape[1] < k:
        pad = k - best_v.shape[1]
        best_v = np.concatenate([best_v, np.full((nq, pad), -np.inf, np.float32)], 1)
        best_i = np.concatenate([best_i, np.full((nq, pad), -1, np.int64)], 1)
    if metric == METRIC_L2:
        best_v = -best_v  # back to a distance; padding becomes +inf
    return best_v.astype(np.float32), best_i


def merge_ref(D_parts, I_parts, k, metric=METRIC_IP):
    """k-way select over per-shard results (shard/merge layer)."""
    sign = 1.0 if metric == METRIC_IP else -1.0
    v = np.full((D_parts[0].shape[0], 0), -np.inf, np.float32)
    i = np.full((D_parts[0].shape[0], 0), -1, np.int64)
    for Dp, Ip in zip(D_parts, I_parts):
        vv = (sign * Dp).astype(np.float32)
        vv = np.where(Ip < 0, -np.inf, vv).astype(np.float32)
        v, i = _merge_sorted(v, i, vv, Ip.astype(np.int64), k)
    i[np.isneginf(v)] = -1
    return (sign * v).astype(np.float32), i


def recall_at_k(I, I_ref) -> float:
    """|I intersect I_ref| / |valid I_ref|, averaged over all queries."""
    hit = 0
    tot = 0
    for a, b in zip(I, I_ref):
        b = b[b >= 0]
        hit += len(np.intersect1d(a[a >= 0], b))
        tot += len(b)
    return hit / max(tot, 1)


def check_topk(D, I, D_ref, I_ref, tie_tol, metric=METRIC_IP, rtol=0.0):
    """Index identity up to ties: rank r may differ only if the two scores at
    that rank differ by <= tie_tol + rtol*|ref| (north_star: "indices identical
    except for ties within 1e-5").  Returns the number of offending entries."""
    D = np.asarray(D, np.float64)
    D_ref = np.asarray(D_ref, np.float64)
    diff_idx = I != I_ref
    with np.errstate(invalid="ignore"):
        gap = np.abs(D - D_ref)
        gap = np.where(np.isinf(D) & np.isinf(D_ref) & (D == D_ref), 0.0, gap)
        tol = tie_tol + rtol * np.abs(np.where(np.isfinite(D_ref), D_ref, 0.0))
    bad = diff_idx & ~(gap <= tol)
    return int(bad.sum())


# ---------------------------------------------------------------------------
# k-means (IVF coarse quantizer) - one Lloyd iteration
# ---------------------------------------------------------------------------

def kmeans_assign_ref(x, centroids, block_rows=65536, dtype=np.float32):
    """assign[i] = argmin_c ||x_i - c||^2, lower centroid id wins ties.
    Returns (assign int32 [n], dist f32 [n])."""
    x = np.ascontiguousarray(x, dtype=dtype)
    c = np.ascontiguousarray(centroids, dtype=dtype)
    cn = np.einsum("ij,ij->i", c, c)
    n = x.shape[0]
    assign = np.empty(n, np.int32)
    dist = np.empty(n, np.float32)
    for r0 in range(0, n, block_rows):
        blk = x[r0:r0 + block_rows]
        s = blk @ c.T
        dd = np.einsum("ij,ij->i", blk, blk)[:, None] - 2.0 * s + cn[None, :]
        a = np.argmin(dd, axis=1)  # first minimum == lower id on ties
        assign[r0:r0 + block_rows] = a
        dist[r0:r0 + block_rows] = dd[np.arange(len(a)), a]
    return assign, dist


def kmeans_update_ref(x, assign, centroids):
    """sums/counts per centroid; new centroid = mean, empty keeps the old one.
    Returns (new_centroids f32 [K,d], counts int64 [K], sums f64 [K,d])."""
    x = np.asarray(x, np.float64)
    K, d = centroids.shape
    sums = np.zeros((K, d), np.float64)
    np.add.at(sums, assign, x)
    counts = np.bincount(assign, minlength=K).astype(np.int64)
    new_c = np.array(centroids, np.float32, copy=True)
    nz = counts > 0
    new_c[nz] = (sums[nz] / counts[nz, None]).astype(np.float32)
    return new_c, counts, sums


def kmeans_split_empty_ref(centroids, counts, eps=1.0 / 1024):
    """Re-seed empty clusters the way include/cvdb_b200.h:cvdb_kmeans_split_empty states it (FAISS Kmeans
    convention, with the donor chosen deterministically): in ascending order each empty cluster e takes the
    currently largest cluster j (ties -> lower id; stop when it has < 2 points); centroid[e] = centroid[j] * (1 +- eps)
    alternating by dimension (even up, odd down), centroid[j] the other way round; counts[e] = counts[j] // 2.
    Returns (centroids f32, counts, n_split)."""
    c = np.array(centroids, np.float32, copy=True)
    n = np.array(counts, np.int64, copy=True)
    up, down = np.float32(1.0) + np.float32(eps), np.float32(1.0) - np.float32(eps)
    odd = (np.arange(c.shape[1]) & 1).astype(bool)
    n_split = 0
    for e in np.flatnonzero(n == 0):
        j = int(np.argmax(n))            # first maximum = lowest id
        if n[j] < 2:
            break
        v = c[j].copy()
        c[e] = np.where(odd, v * down, v * up)
        c[j] = np.where(odd, v * up, v * down)
        n[e] = n[j] // 2
        n[j] -= n[e]
        n_split += 1
    return c, n, n_split


# ---------------------------------------------------------------------------
# IVF-Flat: exact search restricted to the probed inverted lists
# ---------------------------------------------------------------------------

def ivf_probe_ref(centroids, xq, nprobe, metric=METRIC_L2):
    """The nprobe best centroids of every query under the index metric (ties -> lower id)."""
    _, I = search_ref(centroids, xq, nprobe, metric)
    return I


def ivf_assign_ref(centroids, xb, metric=METRIC_L2):
    """List of every database row = its best centroid (k = 1 search over the centroids)."""
    _, I = search_ref(centroids, xb, 1, metric)
    return I[:, 0]


def ivf_search_ref(xb, list_of_row, xq, k, probes, metric=METRIC_IP):
    """Exact top-k of every query over the rows whose list is among probes[q] (ids < 0 skipped).
    Same conventions as search_ref (ties -> lower row id, -1 / +-inf padding)."""
    xb = np.ascontiguousarray(xb, np.float32)
    xq = np.ascontiguousarray(xq, np.float32)
    list_of_row = np.asarray(list_of_row)
    nq = xq.shape[0]
    D = np.full((nq, k), -np.inf if metric == METRIC_IP else np.inf, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for i in range(nq):
        pl = np.asarray(probes[i])
        rows = np.nonzero(np.isin(list_of_row, pl[pl >= 0]))[0]
        if rows.size == 0:
            continue
        d_, i_ = search_ref(xb[rows], xq[i:i + 1], k, metric)
        ok = i_[0] >= 0
        D[i, ok] = d_[0, ok]
        I[i, ok] = rows[i_[0, ok]]
    return D, I


# ---------------------------------------------------------------------------
# triplet assembly from mined neighbours
# ---------------------------------------------------------------------------

def build_triplets_ref(D, I, positives, skip_top=0, per_anchor=1, metric=METRIC_IP, limit=None, anchor_base=0):
    """(anchor, positive, negative) rows, fixed stride [n, per_anchor, 3], unused slots -1."""
    n, k = I.shape
    out = np.full((n, per_anchor, 3), -1, np.int64)
    for i in range(n):
        p = int(positives[i])
        if p < 0:
            continue
        w = 0
        for r in range(skip_top, k):
            if w == per_anchor:
                break
            j = int(I[i, r])
            if j < 0:
                break
            if limit is not None and ((D[i, r] < limit) if metric == METRIC_L2 else (D[i, r] > limit)):
                continue
            if j == p:
                continue
            out[i, w] = (anchor_base + i, p, j)
            w += 1
    return out


# ---------------------------------------------------------------------------
# synthetic data: reproducible in row chunks on the host
# ---------------------------------------------------------------------------

CHUNK_ROWS = 65536


def synth_rows(seed, r0, r1, d, dist="iid", centres=None, noise=0.3):
    """Rows [r0, r1) of the synthetic matrix with base seed ``seed``; every
    CHUNK_ROWS-row chunk has its own generator (seed + chunk id) so any row
    range can be regenerated without materialising the whole matrix."""
    out = np.empty((r1 - r0, d), np.float32)
    c0 = r0 // CHUNK_ROWS
    c1 = (r1 - 1) // CHUNK_ROWS if r1 > r0 else c0 - 1
    for c in range(c0, c1 + 1):
        rng = np.random.default_rng(seed + c)
        lo, hi = c * CHUNK_ROWS, (c + 1) * CHUNK_ROWS
        blk = rng.standard_normal((CHUNK_ROWS, d), dtype=np.float32)
        if dist == "clustered":
            which = rng.integers(0, centres.shape[0], CHUNK_ROWS)
            blk = centres[which] + noise * blk
        a, b = max(lo, r0), min(hi, r1)
        out[a - r0:b - r0] = blk[a - lo:b - lo]
    out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def synth_centres(d, n_centres=4096, seed=99):
    return np.random.default_rng(seed).standard_normal((n_centres, d), dtype=np.float32)
