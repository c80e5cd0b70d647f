"""configs[0]: exact top-10 IP search, 100k x 384 fp32 database, 10k queries: GPU time of the exact-storage path
(three bf16 planes, six plane combinations on the tensor cores) next to the NumPy oracle on the host cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cloudvectordb_b200 import IndexFlat  # noqa: E402
from oracle import flat_oracle as O  # noqa: E402

xb = O.synth_rows(1234, 0, 100_000, 384)
xq = O.synth_rows(5678, 0, 10_000, 384)
t0 = time.perf_counter()
D_ref, I_ref = O.search_ref(xb, xq, 10, O.METRIC_IP)
t_cpu = time.perf_counter() - t0
out = {"config": "configs[0] exact 100k x 384 fp32, 10k queries, k=10", "cpu_oracle_s": t_cpu, "cpu_cores": os.cpu_count()}
for storage in ("exact", "bf16"):
    idx = IndexFlat(384, "ip", storage)
    idx.add(xb)
    xq_d = torch.from_numpy(xq).cuda()
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, I = idx.search(xq_d, 10, profile=True)
        e1.record()
        torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    D, I = D.cpu().numpy(), I.cpu().numpy()
    out[storage] = {"ms_per_batch": float(np.median(ts)), "kernel_ms": float(np.median(idx.profile_ms())),
                    "variant": idx.last_work()["variant"], "idx_identical": float((I == I_ref).mean()),
                    "violations_beyond_1e-5_ties": O.check_topk(D, I, D_ref, I_ref, tie_tol=1e-5),
                    "recall_at_10": O.recall_at_k(I, I_ref), "max_abs_D_err": float(np.abs(D - D_ref).max())}
    # host path (numpy in, numpy out) end to end
    t0 = time.perf_counter()
    idx.search(xq, 10)
    out[storage]["host_call_ms"] = (time.perf_counter() - t0) * 1e3
    idx.close()
print(json.dumps(out))
