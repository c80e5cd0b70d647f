"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: samples, share,
executed warp instructions.   python tools/ncu_lines.py dump.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, cur, hdr = [], "", None
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif r and hdr and r[0].isdigit():
        try:
            s_i = [i for i, h in enumerate(hdr) if h == "# Samples"][0]
            e_i = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
            out.append((cur, int(r[0]), r[1].strip()[:110], int(r[s_i] or 0), int(r[e_i] or 0)))
        except (ValueError, IndexError):
            pass
tot = sum(o[3] for o in out) or 1
print("total samples", tot)
for o in sorted(out, key=lambda x: -x[3])[:n_top]:
    print(f"{o[3]:8d} {100*o[3]/tot:5.1f}% exec={o[4]:11d} {o[0]}:{o[1]:5d} {o[2]}")
