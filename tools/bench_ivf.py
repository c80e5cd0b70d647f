"""IVF-Flat measurement: build (k-means + add + grouping) and nprobe sweep on a clustered synthetic corpus,
with recall against this engine's exact flat search.  1 GPU.

    python tools/bench_ivf.py [--rows 10000000 --dim 768 --nlist 16384 --nq 10000]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cloudvectordb_b200 import IndexFlat, IndexIVFFlat  # noqa: E402

DEV = torch.device("cuda:0")


def clustered(n, d, centres, seed, noise=0.3, chunk=1 << 20):
    out = torch.empty((n, d), dtype=torch.bfloat16, device=DEV)
    for i, r0 in enumerate(range(0, n, chunk)):
        r1 = min(n, r0 + chunk)
        g = torch.Generator(device=DEV).manual_seed(seed + i)
        which = torch.randint(0, centres.shape[0], (r1 - r0,), generator=g, device=DEV)
        blk = centres[which] + noise * torch.randn((r1 - r0, d), generator=g, device=DEV)
        out[r0:r1] = torch.nn.functional.normalize(blk, dim=1).bfloat16()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nlist", type=int, default=16384)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="ip")
    ap.add_argument("--train-rows", type=int, default=2_000_000)
    ap.add_argument("--niter", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ivf.jsonl"))
    a = ap.parse_args()
    f = open(a.out, "a")

    def emit(**kw):
        line = json.dumps(kw, default=float)
        print(line, flush=True)
        f.write(line + "\n")
        f.flush()

    g = torch.Generator(device=DEV).manual_seed(99)
    centres = torch.randn((4096, a.dim), generator=g, device=DEV)
    xb = clustered(a.rows, a.dim, centres, 1234)
    qsrc = xb[torch.randint(0, a.rows, (a.nq,), generator=g, device=DEV)].float()
    xq = torch.nn.functional.normalize(qsrc + 0.1 * torch.randn(qsrc.shape, generator=g, device=DEV), dim=1).bfloat16()

    flat = IndexFlat(a.dim, a.metric, "bf16")
    flat.reserve(a.rows)
    flat.add(xb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D_ex, I_ex = flat.search(xq, a.k)
    torch.cuda.synchronize()
    t_flat = (time.perf_counter() - t0) * 1e3
    flat.close()

    ivf = IndexIVFFlat(a.dim, a.nlist, a.metric)
    t0 = time.perf_counter()
    ivf.train(xb[: a.train_rows], niter=a.niter)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    ivf.add(xb)
    ivf._group()
    torch.cuda.synchronize()
    t_add = time.perf_counter() - t0
    sizes = (ivf.list_offsets[1:] - ivf.list_offsets[:-1]).float()
    emit(event="built", rows=a.rows, dim=a.dim, nlist=a.nlist, train_s=t_train, add_and_group_s=t_add, flat_search_ms=t_flat,
         list_rows_mean=float(sizes.mean()), list_rows_max=float(sizes.max()), empty_lists=int((sizes == 0).sum()))
    for nprobe in (1, 2, 4, 8, 16, 32, 64, 128):
        if nprobe > a.nlist:
            break
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D, I = ivf.search(xq, a.k, nprobe=nprobe)
            e1.record()
            torch.cuda.synchronize()
            if it:
                ts.append(e0.elapsed_time(e1))
        hit = (I[:, :, None] == I_ex[:, None, :]).any(-1).float().sum(1).mean().item() / a.k
        ms = float(np.median(ts))
        emit(event="search", nprobe=nprobe, ms_per_batch=ms, qps=a.nq / ms * 1e3, recall_at_k_vs_exact=hit,
             speedup_vs_flat=t_flat / ms, items=ivf.lists.last_work()["grid"])
    # small batches: end-to-end latency of one search call (coarse search + bucketing + list scan + merge)
    for nq in (1, 16, 64, 256):
        q = xq[:nq]
        for nprobe in (8, 32):
            lat = []
            for it in range(30):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                D, I = ivf.search(q, a.k, nprobe=nprobe)
                torch.cuda.synchronize()
                if it >= 5:
                    lat.append((time.perf_counter() - t0) * 1e3)
            hit = (I[:, :, None] == I_ex[:nq, None, :]).any(-1).float().sum(1).mean().item() / a.k
            emit(event="small_batch", nq=nq, nprobe=nprobe, latency_ms_p50=float(np.percentile(lat, 50)),
                 latency_ms_p99=float(np.percentile(lat, 99)), recall_at_k_vs_exact=hit)
    ivf.close()


if __name__ == "__main__":
    main()
