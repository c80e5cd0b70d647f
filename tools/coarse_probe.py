"""IVF coarse search shape: 10k queries against 8192 / 16 384 / 65 536 centroids x 768, k = nprobe = 8, every kernel variant
and a few slice counts.  One JSON line each (kernel time from the library's events)."""
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
d, nq, k = 768, 10_000, 8
xq = gen_rows(torch, dev, 5678, 0, nq, d, torch.bfloat16)
for rows in (8192, 16384, 65536):
    xb = gen_rows(torch, dev, 1234, 0, rows, d, torch.bfloat16)
    idx = IndexFlat(d, "ip", "bf16")
    idx.add(xb)
    ref = None
    for variant, slices in ((0, 0), (1, 0), (2, 0), (3, 0), (1, 1), (1, 2), (1, 4), (1, 8), (2, 1), (2, 3), (2, 5), (3, 1), (3, 2), (3, 4)):
        kw = {}
        if variant:
            kw["force_variant"] = variant
        if slices:
            kw["force_slices"] = slices
        for _ in range(3):
            D, I = idx.search(xq, k, profile=True, **kw)
        torch.cuda.synchronize()
        idx.profile_ms()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            D, I = idx.search(xq, k, profile=True, **kw)
        e1.record()
        torch.cuda.synchronize()
        kms = float(np.median(idx.profile_ms()))
        w = idx.last_work()
        if ref is None:
            ref = I.clone()
        print(json.dumps({"rows": rows, "force_variant": variant, "force_slices": slices, "variant": w["variant"],
                          "n_slices": w["n_slices"], "call_ms": e0.elapsed_time(e1) / 20, "kernel_ms": kms,
                          "tflops": w["flops"] / kms / 1e9, "same_ids": bool(torch.equal(ref, I))}), flush=True)
    idx.close()
