"""Summarise an `ncu --page source --csv` dump: the most-sampled SASS instructions,
their executed counts and dominant stall reason.  Usage:
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_top.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith("0x")]
S = ci["# Samples"]
IE = ci["Instructions Executed"]
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in body)
tot_inst = sum(int(r[IE]) for r in body)
print(f"kernel: {rows[0][1][:120]}")
print(f"total samples {tot}, warp instructions executed {tot_inst}")
for rank, idx in enumerate(sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:n_top]):
    r = body[idx]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    print(f"{int(r[S]):7d} {100*int(r[S])/max(tot,1):5.1f}%  exec={int(r[IE]):9d}  #{idx:5d}  {r[1].strip()[:70]:70s} {st}")
