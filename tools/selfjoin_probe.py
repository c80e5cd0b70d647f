"""One symmetric-join chunk bracketed by cudaProfilerStart/Stop for ncu (--profile-from-start off).

    SJ_ROWS=1000000 SJ_R0=65536   the candidate-rich chunk right after the seed (default)
    SJ_ROWS=6250000 SJ_R0=3014656 a steady-state chunk of the 6.25M-row shard
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import gen_rows  # noqa: E402
from cloudvectordb_b200 import IndexFlat, _C  # noqa: E402
from cloudvectordb_b200.mining import selfjoin_schedule  # noqa: E402

dev = torch.device("cuda:0")
n, d, k = int(os.environ.get("SJ_ROWS", 1_000_000)), 768, 50
target = int(os.environ.get("SJ_R0", 65536))
emb = gen_rows(torch, dev, 1234, 0, n, d, torch.bfloat16)
idx = IndexFlat(d, "ip", "bf16")
idx.add(emb)
idx.set_groups((torch.arange(n, device=dev) // 4).to(torch.int32))
lib = _C.lib()
st = int(torch.cuda.current_stream().cuda_stream)
keys = torch.empty((n, k), dtype=torch.int64, device=dev)
sched = selfjoin_schedule(n, 65536, min(65536, target))
_C.check(lib.cvdb_selfjoin_begin(idx._h, k, st))
_C.check(lib.cvdb_selfjoin_seed(idx._h, sched[0][1], 0, keys.data_ptr(), st))
for r0, m in sched[1:]:
    prof = r0 == target
    if prof:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()
        e0.record()
    _C.check(lib.cvdb_selfjoin_chunk(idx._h, r0, m, 0, keys[r0:].data_ptr(), st))
    if prof:
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("chunk at row", r0, "anchors", m, "ms", e0.elapsed_time(e1), idx.last_work())
        break
_C.check(lib.cvdb_selfjoin_end(idx._h))
