"""Mining-shaped run for ncu: 16 384 anchors x 2M rows x 768, k = 50, self + group exclusion."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat  # noqa: E402

dev = torch.device("cuda:0")
rows, chunk, k = 2_000_000, 16384, int(os.environ.get("CVDB_K", "50"))
xb = gen_rows(torch, dev, 1234, 0, rows, 768, torch.bfloat16)
groups = (torch.arange(rows, device=dev) // 4).to(torch.int32)
idx = IndexFlat(768, "ip", "bf16")
idx.add(xb)
idx.set_groups(groups)
q, self_ids, gq = xb[:chunk], torch.arange(chunk, device=dev, dtype=torch.int32), groups[:chunk]
for _ in range(3):
    D, I = idx.search(q, k, self_ids=self_ids, group_q=gq, profile=True)
torch.cuda.synchronize()
print("ok", idx.profile_ms()[-1], idx.last_work())
