"""Small k-means assignment run for ncu: 2M x 384 points against 65 536 centroids (L2 metric, k = 1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402

dev = torch.device("cuda:0")
x = gen_rows(torch, dev, 1, 0, 2_000_000, 384, torch.bfloat16)
c = x[torch.randperm(x.shape[0], device=dev)[:65536]].float().contiguous()
idx = IndexFlat(384, "l2", "bf16")
idx.add(c)
for _ in range(3):
    a, d = idx.assign(x)
torch.cuda.synchronize()
print("ok", int(a[0]))
