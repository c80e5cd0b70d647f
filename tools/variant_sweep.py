"""Kernel-variant sweep on a few (d, k) shapes: burst throughput of variants 1, 2, 3 (same box, one call)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cloudvectordb_b200 import IndexFlat  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for (n, d, nq) in ((4_000_000, 128, 20_000), (2_000_000, 256, 10_000), (2_000_000, 384, 10_000), (1_000_000, 512, 10_000)):
    xb = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1).bfloat16()
    xq = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device="cuda"), dim=1).bfloat16()
    idx = IndexFlat(d, "ip")
    idx.add(xb)
    for k in (1, 10, 50):
        row = []
        for v in (1, 2, 3):
            for _ in range(3):
                idx.search(xq, k, profile=True, force_variant=v)
            torch.cuda.synchronize()
            ms = idx.profile_ms()[-1]
            row.append(f"v{v}: {ms:6.2f} ms {2.0*n*d*nq/ms/1e9:5.0f} TF")
        print(f"d={d:4d} k={k:3d}  " + "   ".join(row), flush=True)
    idx.close()
    del xb
