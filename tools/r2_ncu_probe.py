"""Two HBM-side launches for ncu (round-2 evidence): the single-CTA streaming kernel on a small batch
(nq = 64 against 12.5M x 768 bf16, the per-GPU share of configs[4]) and the k-means scatter-add update
(12.5M x 384 bf16 points into 65 536 centroids).  Warm-up calls run unprofiled; ONE launch of each sits between
cudaProfilerStart/Stop.

    ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_topk_ss_kernel|kmeans_update" \
        -o gpurun_out/r2_hbm_side python tools/r2_ncu_probe.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat, _C  # noqa: E402

dev = torch.device("cuda:0")
rows = int(os.environ.get("PROBE_ROWS", 12_500_000))
xb = gen_rows(torch, dev, 1234, 0, rows, 768, torch.bfloat16)
idx = IndexFlat(768, "ip", "bf16")
idx.reserve(rows)
idx.add(xb)
del xb
q = gen_rows(torch, dev, 5678, 0, 64, 768, torch.bfloat16)
for _ in range(5):
    idx.search(q, 10, profile=True)
torch.cuda.synchronize()
print(json.dumps({"small_batch_kernel_ms": idx.profile_ms()[-3:], "work": idx.last_work()}), flush=True)
# k-means update inputs
n, d, K = rows, 384, 65536
pts = gen_rows(torch, dev, 4242, 0, n, d, torch.bfloat16)
assign = torch.randint(0, K, (n,), device=dev, dtype=torch.int32)
sums = torch.zeros((K, d), device=dev)
counts = torch.zeros((K,), device=dev, dtype=torch.int32)
st = int(torch.cuda.current_stream().cuda_stream)


def update():
    _C.check(_C.lib().cvdb_kmeans_accumulate(pts.data_ptr(), n, d, _C.DTYPE_BF16, assign.data_ptr(), sums.data_ptr(),
                                             counts.data_ptr(), st))


for _ in range(2):
    update()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
update()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"kmeans_update_ms": e0.elapsed_time(e1), "points": n, "d": d, "K": K,
                  "point_bytes": n * d * 2, "atomic_bytes_fp32": n * d * 4}), flush=True)
torch.cuda.profiler.start()
idx.search(q, 10)
update()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
