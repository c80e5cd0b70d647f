"""Multi-GPU measurements of BASELINE.json configs[2..4] at (a bounded part of) their full size.
Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_mgpu.py

  mining : 50M x 768 over 8 ranks (6.25M rows each), self-join top-50 with self + group exclusion,
           `--chunks` anchor chunks of 65 536 per owner rank timed, whole join extrapolated
  kmeans : 100M x 384 points over 8 ranks, 65 536 centroids, Lloyd iterations (assign + update + all-reduce)
  small  : 100M x 768 over 8 ranks, 1..64 queries, latency p50/p99
Rows per rank scale with 8 / world so that the per-GPU share stays the 8-GPU one.
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
import torch.distributed as dist  # noqa: E402

from bench import gen_rows, load_peaks  # noqa: E402
from cloudvectordb_b200 import Kmeans, ShardedIndex, mine_hard_negatives_sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="mining,kmeans,small")
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "mgpu.jsonl"))
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    f = open(a.out, "a") if rank == 0 else None

    def emit(**kw):
        if rank == 0:
            line = json.dumps(kw, default=float)
            print(line, flush=True)
            f.write(line + "\n")
            f.flush()

    def sync_time(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    which = a.which.split(",")
    if "mining" in which:
        rows, d, k, chunk = 6_250_000, 768, 50, 65536
        xb = gen_rows(torch, dev, 1234, rank * rows, (rank + 1) * rows, d, torch.bfloat16)
        groups = ((torch.arange(rows, device=dev) + rank * rows) // 4).to(torch.int32)
        idx = ShardedIndex(d, "ip", "bf16", device=local)
        idx.local.reserve(rows)
        idx.add_local(xb)
        idx.set_groups_local(groups)
        fn = lambda: mine_hard_negatives_sharded(idx, xb, k, groups, chunk=chunk, max_chunks=a.chunks)  # noqa: E731
        sync_time(lambda: mine_hard_negatives_sharded(idx, xb[:8192], k, groups[:8192], chunk=8192, max_chunks=1))
        ms, (D, I) = sync_time(fn)
        anchors = world * a.chunks * chunk
        flops = 2.0 * anchors * rows * world * d
        own = torch.arange(rank * rows, rank * rows + I.shape[0], device=dev)
        ok_self = bool((I != own[:, None]).all())
        ok_grp = bool(((I // 4) != (own // 4)[:, None]).all())
        emit(config="mining_sharded", world=world, rows_total=rows * world, d=d, k=k, anchors=anchors, ms=ms,
             anchors_per_s=anchors / ms * 1e3, tflops_aggregate=flops / ms / 1e9,
             frac_sustained_per_gpu=flops / ms / 1e9 / world / peaks["bf16_tflops_sustained"],
             whole_join_estimate_s=rows * world / (anchors / ms * 1e3), self_excluded=ok_self, group_excluded=ok_grp)
        idx.local.close()
        del xb, idx
        torch.cuda.empty_cache()
    if "kmeans" in which:
        n, d, K = 12_500_000, 384, 65536
        x = gen_rows(torch, dev, 4321, rank * n, (rank + 1) * n, d, torch.bfloat16)
        km = Kmeans(d, K, niter=1, seed=42, device=local)
        c0 = x[torch.randperm(n, device=dev)[:K]].float().contiguous()
        dist.broadcast(c0, src=0)
        km.centroids = c0
        sync_time(lambda: km.step(x[:262144]))
        km.centroids = c0.clone()
        objs, times = [], []
        for _ in range(a.iters):
            ms, (_, obj) = sync_time(lambda: km.step(x, profile=True))
            times.append(ms)
            objs.append(float(obj))
        flops = 2.0 * n * world * K * d
        ms = float(np.median(times))
        emit(config="kmeans_sharded", world=world, points_total=n * world, d=d, K=K, ms_per_iter=ms, iters_per_s=1e3 / ms,
             tflops_aggregate=flops / ms / 1e9, frac_sustained_per_gpu=flops / ms / 1e9 / world / peaks["bf16_tflops_sustained"],
             rank0_phases_ms=km.last_timing, objective=objs, objective_non_increasing=all(b <= a_ * (1 + 1e-6) for a_, b in zip(objs, objs[1:])))
        del x, km
        torch.cuda.empty_cache()
    if "small" in which:
        rows, d, k = 12_500_000, 768, 10
        xb = gen_rows(torch, dev, 1234, rank * rows, (rank + 1) * rows, d, torch.bfloat16)
        idx = ShardedIndex(d, "ip", "bf16", device=local)
        idx.local.reserve(rows)
        idx.add_local(xb)
        del xb
        torch.cuda.empty_cache()
        for nq in (1, 4, 16, 64):
            q = gen_rows(torch, dev, 5678, 0, nq, d, torch.bfloat16)
            lat = []
            for it in range(25):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                D, I = idx.search(q, k)
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                if it >= 5:
                    lat.append(float(dt.item()) * 1e3)
            p50, p99 = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
            gbs = rows * d * 2 / (p50 / 1e3) / 1e9
            emit(config="small_batch_sharded", world=world, rows_total=rows * world, d=d, nq=nq, k=k, latency_ms_p50=p50,
                 latency_ms_p99=p99, qps=nq / p50 * 1e3, hbm_gbs_per_gpu_at_p50=gbs, frac_hbm=gbs / peaks["hbm_gbs"])
        idx.local.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
