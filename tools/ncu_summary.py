"""Summarise an .ncu-rep (raw page) into a small JSON: per captured launch the metrics the judge reads
(duration, dram bytes, dram / tensor / L2 utilisation, clocks, registers).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.json ["note"]
"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:160]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[w] = {"value": r[i], "unit": units[i]}
        launches.append(d)
    json.dump({"source": rep, "how": "ncu --set full --clock-control none (cold-cache, serialised replays)", "note": note,
               "launches": launches}, open(out, "w"), indent=1)
    print("wrote", out, len(launches), "launches")


if __name__ == "__main__":
    main()
