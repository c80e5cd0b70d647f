"""Small (HBM-bound) batches on 12.5M x 768: kernel time with the top-k scan on (debug flag 0), skipped (1) and with the
TMEM read skipped as well (2) -- separates the epilogue's share from the MMA / memory side.  One JSON line each."""
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
rows = 12_500_000
xb = gen_rows(torch, dev, 1234, 0, rows, 768, torch.bfloat16)
idx = IndexFlat(768, "ip", "bf16")
idx.add(xb)
del xb
zero_q = os.environ.get("ZERO_Q") == "1"
for nq in (1, 16, 64, 128):
    q = gen_rows(torch, dev, 5678, 0, nq, 768, torch.bfloat16)
    for rep in range(2):
        for dbg in (0, 1, 2):
            for _ in range(3):
                idx.search(q, 10, debug_flags=dbg, profile=True)
            torch.cuda.synchronize()
            idx.profile_ms()
            for _ in range(20):
                idx.search(q, 10, debug_flags=dbg, profile=True)
            torch.cuda.synchronize()
            kms = float(np.median(idx.profile_ms()))
            print(json.dumps({"nq": nq, "dbg": dbg, "kernel_ms": round(kms, 4), "gbs": round(rows * 1536 / kms / 1e6, 1)}), flush=True)
