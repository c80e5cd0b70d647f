"""Top-k cost sweep on the headline shape (10M x 768 bf16, 10k queries) and the mining chunk (6.25M rows, 65 536
anchors, self + group exclusion), one JSON line per measurement.  A/B of two builds of the library:

    CVDB_LIB_PATH=cloudvectordb_b200/ab/libcvdb_sort.so python tools/k_sweep.py --tag sort
    python tools/k_sweep.py --tag select
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

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="default")
    ap.add_argument("--ks", default="10,50,100,200")
    ap.add_argument("--mine-ks", default="50,100")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--flags", default="0", help="comma list of debug_flags to A/B in-process (128 = no slice inheritance)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k_sweep.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    d, nq = 768, 10_000
    xb = gen_rows(torch, dev, 1234, 0, a.rows, d, torch.bfloat16)
    xq = gen_rows(torch, dev, 5678, 0, nq, d, torch.bfloat16)
    idx = IndexFlat(d, "ip", "bf16")
    idx.reserve(a.rows)
    idx.add(xb)
    f = open(a.out, "a")

    def emit(**kw):
        line = json.dumps(dict(tag=a.tag, lib=os.environ.get("CVDB_LIB_PATH", "default"), **kw), default=float)
        print(line, flush=True)
        f.write(line + "\n")
        f.flush()

    def run(fn, iters=4):
        fn()
        torch.cuda.synchronize()
        idx.profile_ms()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, float(np.median(idx.profile_ms())), out

    flags = [int(v) for v in a.flags.split(",")]
    for k in [int(v) for v in a.ks.split(",")]:
        for fl in flags:
            ms, kms, _ = run(lambda: idx.search(xq, k, profile=True, debug_flags=fl))
            w = idx.last_work()
            emit(shape=f"{a.rows}x{d}, {nq} queries", k=k, debug_flags=fl, ms=ms, kernel_ms=kms, qps=nq / ms * 1e3,
                 tflops=w["flops"] / kms / 1e9, variant=w["variant"], n_slices=w["n_slices"])
    rows_m = min(a.rows, 6_250_000)
    idx.truncate(rows_m)
    groups = (torch.arange(rows_m, device=dev) // 4).to(torch.int32)
    idx.set_groups(groups)
    anchors = xb[:65536]
    self_ids = torch.arange(65536, device=dev, dtype=torch.int32)
    gq = groups[:65536]
    for k in [int(v) for v in a.mine_ks.split(",")]:
        for fl in flags + flags:     # twice, interleaved: shows the run-to-run spread next to the A/B difference
            ms, kms, _ = run(lambda: idx.search(anchors, k, self_ids=self_ids, group_q=gq, profile=True, debug_flags=fl), iters=3)
            w = idx.last_work()
            emit(shape=f"mining chunk {rows_m}x{d}, 65536 anchors, self+group exclusion", k=k, debug_flags=fl, ms=ms,
                 kernel_ms=kms, tflops=w["flops"] / kms / 1e9, variant=w["variant"], n_slices=w["n_slices"])
    idx.close()


if __name__ == "__main__":
    main()
