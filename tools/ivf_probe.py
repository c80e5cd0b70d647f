"""IVF-Flat search under ncu: build a 4M x 768 clustered index (nlist 8192), then bracket ONE search of 10k queries
with cudaProfilerStart/Stop so that `ncu --profile-from-start off` lists only the launches of that search.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv python tools/ivf_probe.py [nprobe]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cloudvectordb_b200 import IndexIVFFlat  # noqa: E402
from tools.bench_ivf import clustered  # noqa: E402

dev = torch.device("cuda:0")
nprobe = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows, d, nlist, nq = int(os.environ.get("IVF_ROWS", 4_000_000)), 768, int(os.environ.get("IVF_NLIST", 8192)), 10_000
g = torch.Generator(device=dev).manual_seed(99)
centres = torch.randn((4096, d), generator=g, device=dev)
xb = clustered(rows, d, centres, 1234)
q = xb[torch.randint(0, rows, (nq,), generator=g, device=dev)].float()
xq = torch.nn.functional.normalize(q + 0.1 * torch.randn(q.shape, generator=g, device=dev), dim=1).bfloat16()
ivf = IndexIVFFlat(d, nlist, "ip")
ivf.train(xb[:500_000], niter=3)
ivf.add(xb)
for np_ in sorted({1, 8, 32, nprobe}):
    ivf.nprobe = np_
    for _ in range(2):
        ivf.search(xq, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        D, I = ivf.search(xq, 10)
    e1.record()
    torch.cuda.synchronize()
    print(f"nprobe {np_}: search ms {e0.elapsed_time(e1) / 5:.3f}  checksum {int(I.sum())}", flush=True)
ivf.nprobe = nprobe
torch.cuda.profiler.start()
ivf.search(xq, 10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
