import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from cloudvectordb_b200 import IndexFlat
g = torch.Generator(device="cuda").manual_seed(0)
for (n, d, nq) in ((2_000_000, 256, 10_000), (4_000_000, 128, 20_000)):
    xb = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1).bfloat16()
    xq = torch.nn.functional.normalize(torch.randn((nq, d), generator=g, device="cuda"), dim=1).bfloat16()
    idx = IndexFlat(d, "ip"); idx.add(xb)
    for k in (1, 10, 100):
        for it in range(3):
            D, I = idx.search(xq, k, profile=True)
        torch.cuda.synchronize()
        ms = idx.profile_ms()[-1]
        print(f"n={n} d={d} nq={nq} k={k} kernel_ms={ms:.2f} tflops={2.0*n*d*nq/ms/1e9:.0f}")
    idx.close(); del xb
