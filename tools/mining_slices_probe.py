"""Mining chunk (65 536 anchors, k = 50, self + group exclusion) against 6.25M and 10M rows with the database-slice
count forced: does a long slice (CTAs drifting apart over thousands of tiles) cost throughput?  One JSON line each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402

dev = torch.device("cuda:0")
rows_all = 10_000_000
k = int(os.environ.get("CVDB_K", "50"))
xb = gen_rows(torch, dev, 1234, 0, rows_all, 768, torch.bfloat16)
idx = IndexFlat(768, "ip", "bf16")
idx.reserve(rows_all)
idx.add(xb)
anchors = xb[:65536].clone()
del xb
self_ids = torch.arange(65536, device=dev, dtype=torch.int32)
for rows in (10_000_000, 6_250_000):
    idx.truncate(rows)
    groups = (torch.arange(rows, device=dev) // 4).to(torch.int32)
    idx.set_groups(groups)
    gq = groups[:65536]
    for rep in range(2):
        for s in [int(v) for v in os.environ.get("CVDB_SLICES", "0,13,20,26,37,40").split(",")]:
            kw = {"force_slices": s} if s else {}
            idx.search(anchors, k, self_ids=self_ids, group_q=gq, profile=True, **kw)
            torch.cuda.synchronize()
            idx.profile_ms()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                idx.search(anchors, k, self_ids=self_ids, group_q=gq, profile=True, **kw)
            e1.record()
            torch.cuda.synchronize()
            kms = float(np.median(idx.profile_ms()))
            w = idx.last_work()
            print(json.dumps({"rows": rows, "k": k, "force_slices": s, "n_slices": w["n_slices"], "ms": e0.elapsed_time(e1) / 3,
                              "kernel_ms": kms, "tflops": w["flops"] / kms / 1e9}), flush=True)
