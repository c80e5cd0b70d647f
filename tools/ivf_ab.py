"""A/B of the two inverted-list scan kernels on one index: the grouped kernel (queries on the M side, CVDB_IVF_KERNEL=0)
against the transposed kernel (list rows on the M side, CVDB_IVF_KERNEL=1).  Results must be identical; prints
the time of a whole search() call (coarse probe + bucketing + list scan + merge) for both.

    python tools/ivf_ab.py [--rows 10000000 --nlist 16384 --nq 10000 --k 10] [--out gpurun_out/ivf_ab.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cloudvectordb_b200 import IndexIVFFlat  # noqa: E402
from tools.bench_ivf import clustered  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nlist", type=int, default=16384)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="ip")
    ap.add_argument("--nprobes", default="1,8,32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ivf_ab.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(99)
    centres = torch.randn((4096, a.dim), generator=g, device=dev)
    xb = clustered(a.rows, a.dim, centres, 1234)
    q = xb[torch.randint(0, a.rows, (a.nq,), generator=g, device=dev)].float()
    xq = torch.nn.functional.normalize(q + 0.1 * torch.randn(q.shape, generator=g, device=dev), dim=1).bfloat16()
    ivf = IndexIVFFlat(a.dim, a.nlist, a.metric)
    ivf.train(xb[: min(a.rows, 2_000_000)], niter=3)
    ivf.add(xb)
    del xb
    db_bytes = a.rows * a.dim * 2
    with open(a.out, "a") as f:
        for nprobe in [int(v) for v in a.nprobes.split(",")]:
            res = {}
            for kern in (0, 1):
                os.environ["CVDB_IVF_KERNEL"] = str(kern)
                for _ in range(2):
                    D, I = ivf.search(xq, a.k, nprobe=nprobe)
                torch.cuda.synchronize()
                ts = []
                for _ in range(7):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    D, I = ivf.search(xq, a.k, nprobe=nprobe)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                res[kern] = (float(np.median(ts)), float(np.min(ts)), D.clone(), I.clone(), ivf.lists.last_work()["variant"])
            same_i = bool(torch.equal(res[0][3], res[1][3]))
            same_d = bool(torch.equal(res[0][2], res[1][2]))
            line = json.dumps({"rows": a.rows, "dim": a.dim, "nlist": a.nlist, "nq": a.nq, "k": a.k, "nprobe": nprobe,
                               "grouped_ms": res[0][0], "grouped_ms_min": res[0][1], "transposed_ms": res[1][0],
                               "transposed_ms_min": res[1][1], "variants": [res[0][4], res[1][4]],
                               "ids_identical": same_i, "distances_identical": same_d,
                               "ids_differing": int((res[0][3] != res[1][3]).sum()),
                               "db_bytes": db_bytes, "hbm_floor_ms_if_every_list_is_read_once": db_bytes / 6550.4e6})
            print(line, flush=True)
            f.write(line + "\n")
    ivf.close()


if __name__ == "__main__":
    main()
