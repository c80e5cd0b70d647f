"""Variant 1 (single-CTA streaming) against variant 3 (CTA-pair streaming, M = 256) where the resident-query kernel
does not apply: exact (three-plane) storage and padded K > 832.  One JSON line per measurement."""
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


def run(idx, xq, k, variant):
    kw = {"force_variant": variant} if variant else {}
    for _ in range(3):
        D, I = idx.search(xq, k, profile=True, **kw)
    torch.cuda.synchronize()
    idx.profile_ms()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        D, I = idx.search(xq, k, profile=True, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10, float(np.median(idx.profile_ms())), idx.last_work(), I


cases = [("exact", 100_000, 384, 10_000, 10), ("exact", 100_000, 384, 256, 10), ("exact", 2_000_000, 384, 10_000, 10),
         ("exact", 2_000_000, 768, 10_000, 10), ("exact", 2_000_000, 768, 256, 10),
         ("bf16", 4_000_000, 1024, 10_000, 10), ("bf16", 4_000_000, 1024, 256, 10), ("bf16", 2_000_000, 1536, 10_000, 10),
         ("bf16", 2_000_000, 1536, 200, 10)]
for storage, rows, d, nq, k in cases:
    xb = gen_rows(torch, dev, 1234, 0, rows, d, torch.float32 if storage == "exact" else torch.bfloat16)
    xq = gen_rows(torch, dev, 5678, 0, nq, d, xb.dtype)
    idx = IndexFlat(d, "ip", storage)
    idx.add(xb)
    del xb
    ref = None
    for variant in (0, 1, 3, 1, 3):
        ms, kms, w, I = run(idx, xq, k, variant)
        if ref is None:
            ref = I.clone()
        print(json.dumps({"storage": storage, "rows": rows, "d": d, "nq": nq, "k": k, "force_variant": variant,
                          "variant": w["variant"], "n_slices": w["n_slices"], "call_ms": ms, "kernel_ms": kms,
                          "tflops": w["flops"] / kms / 1e9, "same_ids": bool(torch.equal(ref, I))}), flush=True)
    idx.close()
    torch.cuda.empty_cache()
