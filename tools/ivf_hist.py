"""Where an IVF search spends its time: pairs-per-list histogram of the probes, the tile work the two scan kernels
see (sum over lists of items x tiles), and device time of each phase of one search() call.

    python tools/ivf_hist.py [--rows 10000000 --nlist 16384 --nq 10000 --nprobes 8,32]
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
    ap.add_argument("--nprobes", default="8,32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ivf_hist.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(99)
    centres = torch.randn((4096, a.dim), generator=g, device=dev)
    xb = clustered(a.rows, a.dim, centres, 1234)
    q = xb[torch.randint(0, a.rows, (a.nq,), generator=g, device=dev)].float()
    xq = torch.nn.functional.normalize(q + 0.1 * torch.randn(q.shape, generator=g, device=dev), dim=1).bfloat16()
    ivf = IndexIVFFlat(a.dim, a.nlist, "ip")
    ivf.train(xb[: min(a.rows, 2_000_000)], niter=3)
    ivf.add(xb)
    del xb
    ivf.search(xq, a.k, nprobe=1)
    sizes = (ivf.list_offsets[1:] - ivf.list_offsets[:-1]).long()
    with open(a.out, "a") as f:
        for nprobe in [int(v) for v in a.nprobes.split(",")]:
            probes = ivf.probe(xq, nprobe)
            c = torch.bincount(probes.flatten().long(), minlength=a.nlist)
            tiles = (sizes + 127) // 128
            rec = {"nprobe": nprobe, "pairs": int(c.sum()), "lists_probed": int((c > 0).sum()),
                   "pairs_per_list_mean": float(c.float().mean()), "pairs_per_list_p50": float(c.float().median()),
                   "pairs_per_list_p99": float(torch.quantile(c.float(), 0.99)), "pairs_per_list_max": int(c.max()),
                   "rows_probed_once": int(sizes[c > 0].sum()), "tiles_once": int(tiles[c > 0].sum())}
            for cap in (16, 32, 64, 128):
                items = (c + cap - 1) // cap
                rec[f"items_cap{cap}"] = int(items.sum())
                rec[f"tile_work_cap{cap}"] = int((items * tiles).sum())
            # hybrid: lists with <= T pairs in 16-query items, the rest in 128-query items
            for T in (16, 32, 48, 64):
                small = c <= T
                rec[f"hybrid_T{T}_tiles16"] = int((((c + 15) // 16) * tiles)[small].sum())
                rec[f"hybrid_T{T}_tiles128"] = int((((c + 127) // 128) * tiles)[~small].sum())
            # device time of the phases of one call
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for kern in (0, 1):
                os.environ["CVDB_IVF_KERNEL"] = str(kern)
                ts = []
                for _ in range(5):
                    ev[0].record()
                    pr = ivf.probe(xq, nprobe)
                    ev[1].record()
                    ivf.search(xq, a.k, nprobe=nprobe)
                    ev[2].record()
                    torch.cuda.synchronize()
                    ts.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
                rec[f"kernel{kern}_probe_ms"] = float(np.median([t[0] for t in ts]))
                rec[f"kernel{kern}_search_ms_incl_probe"] = float(np.median([t[1] for t in ts]))
            # the coarse search alone, per kernel variant (10k queries x nlist centroids, k = nprobe)
            for v in (0, 1, 2, 3):
                try:
                    for _ in range(2):
                        ivf.quantizer.search(xq, nprobe, force_variant=v)
                    torch.cuda.synchronize()
                    ev[0].record()
                    for _ in range(5):
                        ivf.quantizer.search(xq, nprobe, force_variant=v, profile=True)
                    ev[1].record()
                    torch.cuda.synchronize()
                    rec[f"coarse_variant{v}_ms"] = ev[0].elapsed_time(ev[1]) / 5
                    rec[f"coarse_variant{v}_kernel_ms"] = float(np.median(ivf.quantizer.profile_ms()))
                    rec[f"coarse_variant{v}_used"] = ivf.quantizer.last_work()
                except Exception as e:
                    rec[f"coarse_variant{v}_error"] = str(e)[:80]
            line = json.dumps(rec)
            print(line, flush=True)
            f.write(line + "\n")
    ivf.close()


if __name__ == "__main__":
    main()
