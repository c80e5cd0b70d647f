"""Headline benchmark: queries/sec of exact top-10 inner-product search,
10M x 768 bf16 database, 10k-query batches (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one search of the whole query batch against the whole database.
With N > 1 the database is row-sharded over the ranks (strong scaling: total
work fixed), every rank searches its shard, one NCCL all-gather exchanges the
per-rank candidates and each rank does the final k-way select.

Prints ONE JSON line (rank 0).  `value` is timed with the queries already in
HBM; `e2e` goes through the public API with pinned HOST buffers (host->device
copy of the queries and device->host copy of (D, I) inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 65536          # rows per seeded generator chunk (same data for every N)
DB_SEED, Q_SEED = 1234, 5678


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="ip", choices=["ip", "l2"])
    ap.add_argument("--dist", default="iid", choices=["iid", "clustered"],
                    help="iid: unit-norm Gaussian rows; clustered: 4096 centres + 0.3 noise, queries = rows + 0.1 noise")
    ap.add_argument("--cpu-rows", type=int, default=400_000, help="database rows of the bounded CPU sample")
    ap.add_argument("--cpu-nq", type=int, default=2048, help="queries of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dbg", type=int, default=0, help="kernel debug flags (tuning experiments)")
    ap.add_argument("--slices", type=int, default=0, help="override the database-slice heuristic")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 streaming kernel, 2 CTA-pair/TMEM kernel")
    return ap.parse_args()


METRIC = "queries/sec at 10M x 768 k=10"   # BASELINE.json "metric"


def workload_name(a):
    return (f"exact top-{a.k} {a.metric.upper()} search, {a.rows}x{a.dim} bf16 database, "
            f"{a.nq}-query batches (BASELINE.json configs[1])")


def workload_config(a, world):
    return {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "queries_per_step": a.nq,
            "k": a.k, "metric": a.metric,
            "distribution": ("iid unit-norm Gaussian rows" if a.dist == "iid" else
                             "clustered: 4096 centres + 0.3 noise, queries = rows + 0.1 noise") + ", seeds 1234/5678",
            "sharding": f"rows split over {world} rank(s), one all-gather of (D,I) + k-way select",
            "cache": "inputs larger than L2 (database 15.4 GB vs 126 MB L2), no explicit flush"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"bf16_tflops": j["bf16_tflops"], "bf16_tflops_sustained": j.get("bf16_tflops_sustained"),
                "hbm_gbs": j["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------- CPU arm
def cpu_sample_qps(a, steps, warmup):
    """The oracle (NumPy fp32 sgemm + select, all host threads) on a bounded sample of the
    workload; scaled to full-size queries/sec by rows_sample / rows (cost is linear in rows)."""
    from oracle import flat_oracle as O
    rows, nq = min(a.cpu_rows, a.rows), min(a.cpu_nq, a.nq)
    xb = O.bf16_round(O.synth_rows(DB_SEED, 0, rows, a.dim))
    xq = O.bf16_round(O.synth_rows(Q_SEED, 0, nq, a.dim))
    metric = O.METRIC_IP if a.metric == "ip" else O.METRIC_L2
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1 to its workers: give the BLAS every core this process may run on, and
    # report the thread count it really uses
    limiter = None
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        limiter = threadpool_limits(limits=cores, user_api="blas")
        used = [i["num_threads"] for i in threadpool_info() if i.get("user_api") == "blas"]
        if used:
            cores = max(used)
    except Exception:
        pass
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.search_ref(xb, xq, a.k, metric)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    if limiter is not None:
        limiter.restore_original_limits()
    t = float(np.mean(times))
    qps_full = nq / t * (rows / a.rows)
    return {"value": qps_full, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{nq} queries x {rows} rows x {a.dim} (NumPy/OpenBLAS fp32 oracle, {t*1e3:.0f} ms per pass), "
                      f"scaled by rows to the {a.rows}-row database"}, t


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, t = cpu_sample_qps(a, max(1, a.steps), max(0, a.warmup))
    out = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(a, max(1, a.gpus)), note="reference ships no code (README.md only): the CPU arm is "
                       "the NumPy IndexFlat oracle port on a bounded sample of this workload"),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled every 200 ms DURING the timed region, through NVML
    (the library nvidia-smi itself reads; in-process so that no subprocess stalls the launch queue)."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.t, self.h = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.h = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.2)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[3] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                "power_w_max": max(r[2] for r in self.rows) if sm else None, "samples": len(sm), "reasons": reasons,
                "source": "NVML, 200 ms period, during the timed region"}


# --------------------------------------------------------------------------- GPU arm
def gen_rows(torch, dev, seed, lo, hi, d, dtype, centres=None, noise=0.3):
    """Rows [lo, hi) of the synthetic unit-norm matrix; chunk c uses generator seed+c.
    With `centres` the rows are centre[random] + noise * N(0, I) (the clustered distribution of SURVEY.md 8(d))."""
    out = torch.empty((hi - lo, d), dtype=dtype, device=dev)
    c0, c1 = lo // CHUNK, (hi - 1) // CHUNK
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=dev).manual_seed(seed + c)
        blk = torch.randn((CHUNK, d), generator=g, device=dev)
        if centres is not None:
            which = torch.randint(0, centres.shape[0], (CHUNK,), generator=g, device=dev)
            blk = centres[which] + noise * blk
        blk = torch.nn.functional.normalize(blk, dim=1)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        out[a - lo:b - lo] = blk[a - c * CHUNK:b - c * CHUNK].to(dtype)
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist

    from cloudvectordb_b200 import IndexFlat, ShardedIndex, _C
    from cloudvectordb_b200.sharded import shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _C.lib()
    peaks = load_peaks()

    lo, hi = shard_bounds(a.rows, world, rank)
    centres = None
    if a.dist == "clustered":
        centres = torch.randn((4096, a.dim), generator=torch.Generator(device=dev).manual_seed(99), device=dev)
    xb = gen_rows(torch, dev, DB_SEED, lo, hi, a.dim, torch.bfloat16, centres)
    if a.dist == "clustered":   # queries: random database rows (of the whole matrix) plus a little noise
        gq = torch.Generator(device=dev).manual_seed(Q_SEED)
        picks = torch.randint(0, a.rows, (a.nq,), generator=gq, device=dev)
        # regenerate the picked rows chunk-wise (any rank can do it: the generator is seeded per chunk)
        rows_f32 = torch.empty((a.nq, a.dim), device=dev)
        order = torch.argsort(picks)
        sp = picks[order]
        for c in torch.unique(sp // CHUNK).tolist():
            blk = gen_rows(torch, dev, DB_SEED, c * CHUNK, (c + 1) * CHUNK, a.dim, torch.float32, centres)
            m = (sp // CHUNK) == c
            rows_f32[order[m]] = blk[sp[m] - c * CHUNK]
        xq32 = torch.nn.functional.normalize(rows_f32 + 0.1 * torch.randn((a.nq, a.dim), generator=gq, device=dev), dim=1)
        xq = xq32.bfloat16()
    else:
        xq32 = None
        xq = gen_rows(torch, dev, Q_SEED, 0, a.nq, a.dim, torch.bfloat16)
    if world > 1:
        index = ShardedIndex(a.dim, a.metric, "bf16", device=local_rank)
        index.local.reserve(hi - lo)
        index.add_local(xb)
        local = index.local
    else:
        index = IndexFlat(a.dim, a.metric, "bf16", device=local_rank)
        index.reserve(hi - lo)
        index.add(xb)
        local = index
    q_host = xq.cpu().pin_memory()
    torch.cuda.synchronize()

    vkw = {"force_variant": a.variant} if (a.variant and world == 1) else {}
    if a.slices and world == 1:
        vkw["force_slices"] = a.slices
    if a.dbg and world == 1:
        vkw["debug_flags"] = a.dbg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return index.search(xq, a.k, profile=True, **vkw)

    def step_e2e():
        return index.search(q_host, a.k, **vkw)

    def timed(fn, steps, warmup, collect_kernel=False):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record()
        if collect_kernel:
            local.profile_ms()  # drop warm-up launches
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if collect_kernel:
            kms = local.profile_ms()[-steps:]
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), kms, out

    sampler = ClockSampler(local_rank)
    launches0 = lib.cvdb_kernel_launches()
    if rank == 0:
        sampler.start()
    total_ms, kms, (D, I) = timed(step_device, a.steps, a.warmup, collect_kernel=True)
    clocks = sampler.stop() if rank == 0 else None
    launches = (lib.cvdb_kernel_launches() - launches0) * a.steps // (a.steps + a.warmup)
    ms_per_step = total_ms / a.steps
    value = a.nq / ms_per_step * 1e3

    # end to end through the public API with pinned host buffers
    e2e_ms, _, (Dh, Ih) = timed(step_e2e, max(2, a.steps // 2), 2)
    e2e_ms /= max(2, a.steps // 2)
    h2d = q_host.numel() * q_host.element_size()
    d2h = a.nq * a.k * (4 + 8)

    # dominant kernel roofline (tensor-bound: 2*nq*rows_local*d flops per launch)
    w = local.last_work()
    kernel_ms = float(np.mean(kms))
    achieved = w["flops"] / kernel_ms / 1e9
    peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm_topk", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "peak_kind": "sustained cuBLAS bf16, " + peaks["source"],
                "peak_burst": peaks["bf16_tflops"], "frac_burst": achieved / peaks["bf16_tflops"],
                "frac_nominal_2250": achieved / 2250.0, "kernel_ms": kernel_ms,
                "kernel_share_of_step": kernel_ms / ms_per_step, "flops_per_launch": w["flops"],
                "db_bytes_per_launch": w["db_bytes"], "hbm_gbs_algorithmic": w["db_bytes"] / kernel_ms / 1e6,
                "traffic": traffic, "n_slices": w["n_slices"], "grid": w["grid"], "variant": w["variant"]}

    # recall of the bf16 engine against a torch fp32 matmul on a query subsample (rank-local rows -> global merge
    # is already done by the engine, so gather the reference over all ranks' rows)
    nchk = 64
    qs = xq[:nchk].float()
    best_v = torch.full((nchk, a.k), -float("inf"), device=dev)
    best_i = torch.full((nchk, a.k), -1, dtype=torch.int64, device=dev)
    for r0 in range(0, hi - lo, 1 << 20):
        blk = xb[r0:r0 + (1 << 20)].float()
        s = qs @ blk.T
        if a.metric == "l2":
            s = -((qs * qs).sum(1)[:, None] - 2 * s + (blk * blk).sum(1)[None, :])
        v, i = torch.topk(s, min(a.k, blk.shape[0]), dim=1)
        cv, ci = torch.cat([best_v, v], 1), torch.cat([best_i, i + r0 + lo], 1)
        o = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :a.k]
        best_v, best_i = torch.gather(cv, 1, o), torch.gather(ci, 1, o)
    if world > 1:
        gv = [torch.empty_like(best_v) for _ in range(world)]
        gi = [torch.empty_like(best_i) for _ in range(world)]
        dist.all_gather(gv, best_v)
        dist.all_gather(gi, best_i)
        cv, ci = torch.cat(gv, 1), torch.cat(gi, 1)
        o = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :a.k]
        best_i = torch.gather(ci, 1, o)
    ref_i = best_i.cpu().numpy()
    got_i = torch.as_tensor(I)[:nchk].cpu().numpy()
    recall = float(np.mean([len(np.intersect1d(x, y)) / a.k for x, y in zip(got_i, ref_i)]))

    # the same against fp32 scores of the UNROUNDED fp32 rows and queries (regenerated chunk by chunk)
    qs32 = xq32[:nchk] if xq32 is not None else gen_rows(torch, dev, Q_SEED, 0, nchk, a.dim, torch.float32)
    bv = torch.full((nchk, a.k), -float("inf"), device=dev)
    bi = torch.full((nchk, a.k), -1, dtype=torch.int64, device=dev)
    for r0 in range(lo, hi, 1 << 20):
        blk = gen_rows(torch, dev, DB_SEED, r0, min(hi, r0 + (1 << 20)), a.dim, torch.float32, centres)
        s = qs32 @ blk.T
        if a.metric == "l2":
            s = -((qs32 * qs32).sum(1)[:, None] - 2 * s + (blk * blk).sum(1)[None, :])
        v, i = torch.topk(s, min(a.k, blk.shape[0]), dim=1)
        cv, ci = torch.cat([bv, v], 1), torch.cat([bi, i + r0], 1)
        o = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :a.k]
        bv, bi = torch.gather(cv, 1, o), torch.gather(ci, 1, o)
    if world > 1:
        gv = [torch.empty_like(bv) for _ in range(world)]
        gi = [torch.empty_like(bi) for _ in range(world)]
        dist.all_gather(gv, bv)
        dist.all_gather(gi, bi)
        cv, ci = torch.cat(gv, 1), torch.cat(gi, 1)
        o = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :a.k]
        bi = torch.gather(ci, 1, o)
    recall32 = float(np.mean([len(np.intersect1d(x, y)) / a.k for x, y in zip(got_i, bi.cpu().numpy())]))
    same_e2e = bool(np.array_equal(torch.as_tensor(Ih)[:nchk].cpu().numpy(), got_i))

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base, _ = cpu_sample_qps(a, steps=2, warmup=1)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(a, world),
            "e2e": {"value": a.nq / e2e_ms * 1e3, "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "same_ids_as_device_path": same_e2e},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_base,
            "recall_at_k_vs_fp32_torch_on_same_bf16_values": recall,
            "recall_at_k_vs_fp32_torch_on_unrounded_fp32_inputs": recall32,
            "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
