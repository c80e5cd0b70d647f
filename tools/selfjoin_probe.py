"""One chunk of the symmetric self-join under ncu: 1M x 768 bf16, k = 50, groups of four; seed 32768 rows, one
unprofiled chunk, then ONE 65 536-anchor chunk (thresholds from 65 536 anchors: the push-heavy early regime)
between cudaProfilerStart/Stop.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_topk_ts2 \
        -o gpurun_out/r2_sj_chunk python tools/selfjoin_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat, _C  # noqa: E402

dev = torch.device("cuda:0")
n, d, k = int(os.environ.get("SJ_ROWS", 1_000_000)), 768, 50
emb = gen_rows(torch, dev, 1234, 0, n, d, torch.bfloat16)
idx = IndexFlat(d, "ip", "bf16")
idx.add(emb)
idx.set_groups((torch.arange(n, device=dev) // 4).to(torch.int32))
lib = _C.lib()
st = int(torch.cuda.current_stream().cuda_stream)
keys = torch.empty((n, k), dtype=torch.int64, device=dev)
_C.check(lib.cvdb_selfjoin_begin(idx._h, k, st))
_C.check(lib.cvdb_selfjoin_seed(idx._h, 32768, 0, keys.data_ptr(), st))
_C.check(lib.cvdb_selfjoin_chunk(idx._h, 32768, 32768, 0, keys[32768:].data_ptr(), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
_C.check(lib.cvdb_selfjoin_chunk(idx._h, 65536, 65536, 0, keys[65536:].data_ptr(), st))
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("chunk ms", e0.elapsed_time(e1), idx.last_work())
_C.check(lib.cvdb_selfjoin_end(idx._h))
