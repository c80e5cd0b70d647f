"""The 128 < nq <= 256 ridge: time every kernel variant on an HBM-bound shape (12.5M x 768, the per-GPU share
of configs[4]) for a few query counts.  Usage: python tools/ridge_probe.py [rows]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
xb = gen_rows(torch, dev, 1234, 0, rows, 768, torch.bfloat16)
idx = IndexFlat(768, "ip", "bf16")
idx.add(xb)
del xb
for nq in (64, 128, 129, 160, 192, 256, 257, 384, 512):
    q = gen_rows(torch, dev, 5678, 0, nq, 768, torch.bfloat16)
    ref = None
    for variant in (0, 1, 2, 3):
        try:
            for _ in range(2):
                D, I = idx.search(q, 10, force_variant=variant)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                D, I = idx.search(q, 10, force_variant=variant)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            if ref is None:
                ref = I.clone()
            same = bool(torch.equal(ref, I))
            print(json.dumps({"nq": nq, "variant": variant, "used": idx.last_work().get("variant"), "ms": round(ms, 3),
                              "gbs": round(rows * 768 * 2 / ms / 1e6, 1), "same_ids": same}), flush=True)
        except Exception as e:  # a variant may not support the shape
            print(json.dumps({"nq": nq, "variant": variant, "error": str(e)[:100]}), flush=True)
