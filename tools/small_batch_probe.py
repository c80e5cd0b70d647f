"""Small (HBM-bound) batches on 12.5M x 768, the per-GPU share of configs[4]: latency per batch with the queries
spread over the four epilogue warps (default) and packed from lane 0 up (debug flag 64)."""
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
for nq in (1, 4, 16, 32, 64, 100, 127, 128):
    q = gen_rows(torch, dev, 5678, 0, nq, 768, torch.bfloat16)
    ref = None
    for dbg in (0, 64, 0, 64):
        for _ in range(3):
            D, I = idx.search(q, 10, debug_flags=dbg)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D, I = idx.search(q, 10, debug_flags=dbg)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        if ref is None:
            ref = I.clone()
        print(json.dumps({"nq": nq, "dbg": dbg, "ms_p50": round(ts[7], 3), "ms_min": round(ts[0], 3),
                          "gbs_p50": round(rows * 768 * 2 / ts[7] / 1e6, 1), "same_ids": bool(torch.equal(ref, I))}), flush=True)
