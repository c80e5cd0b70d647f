"""The 129..256-query ridge under ncu: 12.5M x 768 bf16 (19.2 GB), nq = 256, k = 10; a few warm-up calls of each
CTA-pair variant, then ONE launch of each between cudaProfilerStart/Stop.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_topk \
        -o gpurun_out/r2_ridge python tools/ridge_ncu.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402

dev = torch.device("cuda:0")
rows = int(os.environ.get("RIDGE_ROWS", 12_500_000))
nq = int(os.environ.get("RIDGE_NQ", 256))
xb = gen_rows(torch, dev, 1234, 0, rows, 768, torch.bfloat16)
idx = IndexFlat(768, "ip", "bf16")
idx.reserve(rows)
idx.add(xb)
del xb
q = gen_rows(torch, dev, 5678, 0, nq, 768, torch.bfloat16)
for variant in (2, 3, 1):
    ts = []
    for i in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx.search(q, 10, force_variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = np.array(ts[5:])
    print(json.dumps({"nq": nq, "rows": rows, "variant": variant, "ms_p50": float(np.median(ts)), "ms_min": float(ts.min()),
                      "ms_max": float(ts.max()), "spread": float((ts.max() - ts.min()) / np.median(ts)),
                      "n_slices": idx.last_work()["n_slices"], "all_ms": [round(float(t), 3) for t in ts]}), flush=True)
torch.cuda.profiler.start()
for variant in (2, 3):
    idx.search(q, 10, force_variant=variant)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
