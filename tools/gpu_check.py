"""Developer battery run on the GPU box: parity cases against the NumPy oracle
plus a few timed shapes.  Not part of the product; tests/ holds the real suite.

    python tools/gpu_check.py [--quick] [--perf] [--out gpurun_out/check.log]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cloudvectordb_b200 import IndexFlat  # noqa: E402
from oracle import flat_oracle as O  # noqa: E402

LOG = None


def emit(**kw):
    line = json.dumps(kw, default=float)
    print(line, flush=True)
    if LOG:
        LOG.write(line + "\n")
        LOG.flush()


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def parity_case(name, n, d, nq, k, metric="ip", storage="bf16", excl=False, force_slices=0, seed=0, device_io=False,
                variant=0):
    rng = np.random.default_rng(seed)
    xb = unit_rows(rng, n, d)
    xq = unit_rows(rng, nq, d)
    if storage == "bf16":
        xb, xq = O.bf16_round(xb), O.bf16_round(xq)  # isolate kernel error from input rounding
    self_ids = group_db = group_q = None
    if excl:
        self_ids = rng.integers(0, n, nq).astype(np.int32)
        group_db = (np.arange(n) // 4).astype(np.int32)
        group_q = group_db[self_ids].copy()
        group_q[::5] = -1
    t0 = time.time()
    D_ref, I_ref = O.search_ref(xb, xq, k, O.METRIC_IP if metric == "ip" else O.METRIC_L2, self_ids=self_ids,
                                group_db=group_db, group_q=group_q)
    t_ref = time.time() - t0
    idx = IndexFlat(d, metric, storage)
    try:
        if device_io:
            idx.add(torch.from_numpy(xb).cuda())
            if excl:
                idx.set_groups(group_db)
            D, I = idx.search(torch.from_numpy(xq).cuda(), k, self_ids=self_ids, group_q=group_q,
                              force_slices=force_slices, force_variant=variant)
            torch.cuda.synchronize()
            D, I = D.cpu().numpy(), I.cpu().numpy()
        else:
            idx.add(xb)
            if excl:
                idx.set_groups(group_db)
            D, I = idx.search(xq, k, self_ids=self_ids, group_q=group_q, force_slices=force_slices,
                              force_variant=variant)
        used = idx.last_work()["variant"]
    finally:
        idx.close()
    tol = 2e-5 if storage == "bf16" else 1e-5
    bad = O.check_topk(D, I, D_ref, I_ref, tie_tol=tol, metric=0 if metric == "ip" else 1)
    fin = np.isfinite(D_ref) & np.isfinite(D)
    maxerr = float(np.abs(D - D_ref)[fin].max()) if fin.any() else 0.0
    inf_mismatch = int((np.isfinite(D_ref) != np.isfinite(D)).sum())
    rec = O.recall_at_k(I, I_ref)
    ok = bad == 0 and inf_mismatch == 0 and maxerr < 1e-3
    emit(case=name, ok=bool(ok), variant=used, n=n, d=d, nq=nq, k=k, metric=metric, storage=storage, excl=excl,
         force_slices=force_slices, bad=bad, inf_mismatch=inf_mismatch, maxerr=maxerr, recall=rec,
         idx_equal=float((I == I_ref).mean()), oracle_s=round(t_ref, 3))
    return ok


def perf_case(name, n, d, nq, k, metric="ip", iters=3, check_q=64, variant=0, dbg=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    xb = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for r0 in range(0, n, 1 << 20):
        r1 = min(n, r0 + (1 << 20))
        blk = torch.randn((r1 - r0, d), generator=g, device=dev, dtype=torch.float32)
        xb[r0:r1] = torch.nn.functional.normalize(blk, dim=1).to(torch.bfloat16)
    xq = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device=dev), dim=1).to(torch.bfloat16)
    idx = IndexFlat(d, metric, "bf16")
    try:
        idx.reserve(n)
        idx.add(xb)
        torch.cuda.synchronize()
        times, kms = [], []
        for it in range(iters + 1):
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            D, I = idx.search(xq, k, profile=True, force_variant=variant, debug_flags=dbg)
            t1.record()
            torch.cuda.synchronize()
            if it > 0:
                times.append(t0.elapsed_time(t1))
                kms.append(idx.last_kernel_ms())
        w = idx.last_work()
        # reference on a query subsample: fp32 scores of the same bf16 values
        qs = xq[:check_q].float()
        best_v = torch.full((check_q, k), -float("inf"), device=dev)
        best_i = torch.full((check_q, k), -1, dtype=torch.int64, device=dev)
        for r0 in range(0, n, 1 << 20):
            r1 = min(n, r0 + (1 << 20))
            blk = xb[r0:r1].float()
            s = qs @ blk.T
            if metric == "l2":
                s = -((qs * qs).sum(1)[:, None] - 2 * s + (blk * blk).sum(1)[None, :])
            v, i = torch.topk(s, min(k, r1 - r0), dim=1)
            cv = torch.cat([best_v, v], 1)
            ci = torch.cat([best_i, i + r0], 1)
            o = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :k]
            best_v = torch.gather(cv, 1, o)
            best_i = torch.gather(ci, 1, o)
        rec = O.recall_at_k(I[:check_q].cpu().numpy(), best_i.cpu().numpy())
        ref_d = best_v if metric == "ip" else -best_v
        derr = float((D[:check_q] - ref_d).abs().max())
        ms = float(np.median(times))
        kms_med = float(np.median(kms))
        emit(case=name, dbg=dbg, n=n, d=d, nq=nq, k=k, metric=metric, ms_search=ms, ms_kernel=kms_med,
             qps=nq / ms * 1e3, tflops_kernel=w["flops"] / kms_med / 1e9, n_slices=w["n_slices"], grid=w["grid"], variant=w["variant"],
             recall_vs_torch=rec, max_d_err=derr)
    finally:
        idx.close()
        del xb
        torch.cuda.empty_cache()


def main():
    global LOG
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--dbg", type=int, default=0)
    ap.add_argument("--only", default="", help="comma list of perf case names; skips parity")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "check.log"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    LOG = open(a.out, "a")
    emit(event="start", gpu=torch.cuda.get_device_name(0), argv=sys.argv[1:])
    cases = [
        dict(name="tiny_ip", n=1000, d=64, nq=7, k=10),
        dict(name="tiny_ip_dev", n=1000, d=64, nq=7, k=10, device_io=True),
        dict(name="d768", n=5000, d=768, nq=129, k=10),
        dict(name="d100_k1", n=3000, d=100, nq=64, k=1),
        dict(name="d100_k50", n=3000, d=100, nq=64, k=50),
        dict(name="d100_k128", n=3000, d=100, nq=33, k=128),
        dict(name="l2_d64", n=4000, d=64, nq=100, k=10, metric="l2"),
        dict(name="l2_d384_k1", n=4097, d=384, nq=300, k=1, metric="l2"),
        dict(name="slices3", n=20000, d=128, nq=200, k=10, force_slices=3),
        dict(name="slices7_k20", n=20000, d=128, nq=200, k=20, force_slices=7),
        dict(name="excl", n=6000, d=96, nq=150, k=10, excl=True),
        dict(name="excl_k50_l2", n=6000, d=96, nq=150, k=50, excl=True, metric="l2"),
        dict(name="k_gt_n", n=5, d=32, nq=3, k=10),
        dict(name="exact_ip", n=20000, d=384, nq=256, k=10, storage="exact"),
        dict(name="exact_l2", n=9000, d=100, nq=130, k=10, storage="exact", metric="l2"),
        dict(name="k300", n=3000, d=64, nq=40, k=300),
    ]
    if not a.quick:
        cases += [
            dict(name="c1_exact", n=100000, d=384, nq=2000, k=10, storage="exact"),
            dict(name="mid_bf16", n=200000, d=768, nq=1000, k=10),
        ]
    if a.variant:
        cases = [dict(c, variant=a.variant) for c in cases if c.get("storage", "bf16") == "bf16" or a.variant in (1, 3)]
    if a.only:
        cases = []
    n_ok = 0
    for c in cases:
        try:
            n_ok += bool(parity_case(**c))
        except Exception as e:  # keep going: one call should tell us as much as possible
            emit(case=c["name"], ok=False, error=repr(e), tb=traceback.format_exc()[-800:])
            if "CUDA" in repr(e) or "cuda" in repr(e):
                emit(event="abort", reason="cuda failure, context is likely dead")
                break
    emit(event="parity_done", ok=n_ok, total=len(cases))
    if a.perf:
        try:
            v = a.variant
            pc = [("perf_1M", (1_000_000, 768, 10000, 10), {}),
                  ("perf_1M_l2", (1_000_000, 765, 10000, 10), dict(metric="l2")),
                  ("perf_1M_d384_k1", (1_000_000, 384, 100000, 1), dict(metric="l2")),
                  ("perf_1M_d512", (1_000_000, 512, 10000, 10), {}),
                  ("perf_small_batch", (4_000_000, 768, 64, 10), {}),
                  ("perf_small_256", (4_000_000, 768, 256, 10), {}),
                  ("perf_small_200", (4_000_000, 768, 200, 10), {}),
                  ("perf_mid_1024", (4_000_000, 768, 1024, 10), {}),
                  ("perf_mid_512", (4_000_000, 768, 512, 10), {})]
            if a.big:
                pc.append(("perf_10M", (10_000_000, 768, 10000, 10), {}))
            for nm, args, kw in pc:
                if a.only and nm not in a.only.split(","):
                    continue
                perf_case(nm, *args, variant=v, dbg=a.dbg, **kw)
        except Exception as e:
            emit(case="perf", ok=False, error=repr(e), tb=traceback.format_exc()[-800:])
    emit(event="done")


if __name__ == "__main__":
    main()
