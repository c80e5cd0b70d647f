"""Secondary measurements for BASELINE.json configs[2..4] at the per-GPU share of
the 8-GPU configuration (1 GPU, synthetic data).  Writes JSON lines.

    python tools/bench_configs.py [--which mining,kmeans,small] [--out gpurun_out/configs.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import gen_rows, load_peaks  # noqa: E402
from cloudvectordb_b200 import IndexFlat, Kmeans  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, iters=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def emit(f, **kw):
    line = json.dumps(kw, default=float)
    print(line, flush=True)
    f.write(line + "\n")
    f.flush()


def mining(f, peaks, rows=6_250_000, d=768, chunk=65536, k=int(os.environ.get("CVDB_K", "50")), variant=int(os.environ.get("CVDB_VARIANT", "0"))):
    """configs[2]: 50M x 768 self-join top-50 with positive exclusion over 8 GPUs -> 6.25M rows per GPU;
    one step = one 65 536-anchor chunk against the local shard."""
    xb = gen_rows(torch, DEV, 1234, 0, rows, d, torch.bfloat16)
    groups = (torch.arange(rows, device=DEV) // 4).to(torch.int32)
    idx = IndexFlat(d, "ip", "bf16")
    idx.reserve(rows)
    idx.add(xb)
    idx.set_groups(groups)
    q = xb[:chunk]
    self_ids = torch.arange(chunk, device=DEV, dtype=torch.int32)
    gq = groups[:chunk]
    excl = os.environ.get("CVDB_NOEXCL") is None
    ms, (D, I) = timed(lambda: idx.search(q, k, self_ids=self_ids if excl else None, group_q=gq if excl else None,
                                          profile=True, force_variant=variant))
    kms = idx.profile_ms()
    w = idx.last_work()
    ok_self = bool((I != self_ids[:, None].long()).all())
    ok_grp = bool((groups[I.clamp(min=0)] != gq[:, None]).all())
    tf = w["flops"] / np.median(kms) / 1e9
    emit(f, config="mining_chunk", rows_local=rows, d=d, anchors=chunk, k=k, ms_per_chunk=ms, kernel_ms=float(np.median(kms)),
         anchors_per_s=chunk / ms * 1e3, tflops=tf, frac_sustained=tf / peaks["bf16_tflops_sustained"],
         frac_burst=tf / peaks["bf16_tflops"], variant=w["variant"], n_slices=w["n_slices"],
         self_excluded=ok_self, group_excluded=ok_grp,
         whole_join_estimate_s_8gpu=(50_000_000 / chunk) * ms / 1e3)
    idx.close()
    del xb


def kmeans(f, peaks, n=12_500_000, d=384, K=65536):
    """configs[3]: 100M x 384 points, 65 536 centroids over 8 GPUs -> 12.5M points per GPU; one Lloyd iteration."""
    x = gen_rows(torch, DEV, 1234, 0, n, d, torch.bfloat16)
    km = Kmeans(d, K, niter=1, seed=42, storage="bf16", device=0)
    km.centroids = x[torch.randperm(n, device=DEV)[:K]].float().contiguous()
    ms, (assign, obj) = timed(lambda: km.step(x, profile=True), iters=2, warmup=1)
    flops = 2.0 * n * K * d
    tf = flops / ms / 1e9
    # assignment agreement with a torch fp32 reference on a subsample
    sub = torch.randperm(n, device=DEV)[:2048]
    km._set_centroids(km.centroids)
    a_sub, _ = km._index.assign(x[sub])
    c = km.centroids.bfloat16().float()
    xs = x[sub].float()
    d2 = (xs * xs).sum(1)[:, None] - 2 * xs @ c.T + (c * c).sum(1)[None, :]
    ref = d2.argmin(1)
    agree = float((ref == a_sub.long()).float().mean())
    gap_ok = True
    if agree < 1.0:
        bad = ref != a_sub.long()
        gap = (d2[bad, a_sub.long()[bad]] - d2[bad, ref[bad]]).abs().max()
        gap_ok = bool(gap < 1e-3)
    emit(f, config="kmeans_iter", points_local=n, d=d, K=K, ms_per_iter=ms, iters_per_s=1e3 / ms, tflops_whole_iter=tf,
         frac_sustained=tf / peaks["bf16_tflops_sustained"], frac_burst=tf / peaks["bf16_tflops"],
         phases_ms=km.last_timing, assign_agreement_subsample=agree, disagreements_are_near_ties=gap_ok,
         nonempty_clusters=int((km.last_counts > 0).sum()))
    del x


def small_batch(f, peaks, rows=12_500_000, d=768, k=10):
    """configs[4]: 100M x 768 over 8 GPUs -> 12.5M rows (19.2 GB) per GPU, 1..64 queries: HBM-bound."""
    xb = gen_rows(torch, DEV, 1234, 0, rows, d, torch.bfloat16)
    idx = IndexFlat(d, "ip", "bf16")
    idx.reserve(rows)
    idx.add(xb)
    del xb
    torch.cuda.empty_cache()
    for nq in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        q = gen_rows(torch, DEV, 5678, 0, nq, d, torch.bfloat16)
        ms, _ = timed(lambda: idx.search(q, k, profile=True), iters=5, warmup=2)
        kms = float(np.median(idx.profile_ms()))
        w = idx.last_work()
        emit(f, config="small_batch", rows_local=rows, d=d, nq=nq, k=k, ms_per_batch=ms, kernel_ms=kms, qps=nq / ms * 1e3,
             hbm_gbs_kernel=w["db_bytes"] / kms / 1e6, frac_hbm=w["db_bytes"] / kms / 1e6 / peaks["hbm_gbs"],
             hbm_gbs_call=w["db_bytes"] / ms / 1e6, variant=w["variant"], n_slices=w["n_slices"])
    idx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="mining,kmeans,small")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    peaks = load_peaks()
    with open(a.out, "a") as f:
        for name in a.which.split(","):
            try:
                {"mining": mining, "kmeans": kmeans, "small": small_batch}[name](f, peaks)
            except Exception as e:
                emit(f, config=name, error=repr(e))
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
