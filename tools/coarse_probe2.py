import json, os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from bench import gen_rows
from cloudvectordb_b200 import IndexFlat
dev = torch.device("cuda:0")
d, nq = 768, 10_000
xq = gen_rows(torch, dev, 5678, 0, nq, d, torch.bfloat16)
for rows in (8192, 16384, 65536):
    xb = gen_rows(torch, dev, 1234, 0, rows, d, torch.bfloat16)
    idx = IndexFlat(d, "ip", "bf16")
    idx.add(xb)
    for k in (1, 8, 50):
        for dbg in (0, 1, 4):
            for sl in (0, 1):
                kw = {"debug_flags": dbg}
                if sl: kw["force_slices"] = sl
                for _ in range(3):
                    idx.search(xq, k, profile=True, **kw)
                torch.cuda.synchronize(); idx.profile_ms()
                for _ in range(10):
                    idx.search(xq, k, profile=True, **kw)
                torch.cuda.synchronize()
                kms = float(np.median(idx.profile_ms())); w = idx.last_work()
                print(json.dumps({"rows": rows, "k": k, "dbg": dbg, "force_slices": sl, "n_slices": w["n_slices"], "variant": w["variant"], "kernel_ms": round(kms, 4)}), flush=True)
    idx.close()
