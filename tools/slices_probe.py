"""Headline shape (10M x 768, 10k queries) with forced slice counts, k = 10 / 50 / 200: with pooled thresholds more
(shorter) slices tighten the bound earlier.  One JSON line per measurement."""
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
rows, d, nq = 10_000_000, 768, 10_000
xb = gen_rows(torch, dev, 1234, 0, rows, d, torch.bfloat16)
xq = gen_rows(torch, dev, 5678, 0, nq, d, torch.bfloat16)
idx = IndexFlat(d, "ip", "bf16")
idx.reserve(rows)
idx.add(xb)
del xb
for k in (10, 50, 200):
    for rep in range(2):
        for s in (0, 37, 74, 111, 148):
            kw = {"force_slices": s} if s else {}
            idx.search(xq, k, profile=True, **kw)
            torch.cuda.synchronize()
            idx.profile_ms()
            for _ in range(4):
                idx.search(xq, k, profile=True, **kw)
            torch.cuda.synchronize()
            kms = float(np.median(idx.profile_ms()))
            w = idx.last_work()
            print(json.dumps({"k": k, "force_slices": s, "n_slices": w["n_slices"], "kernel_ms": round(kms, 3),
                              "tflops": round(w["flops"] / kms / 1e9, 1)}), flush=True)
