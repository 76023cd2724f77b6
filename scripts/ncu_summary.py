#!/usr/bin/env python3
"""Turns an ncu report into the short text summary kept under profiles/.

  python scripts/ncu_summary.py gpurun_out/r02/k2_gemm.ncu-rep "header line" > profiles/r02_k2_gemm_ncu_full.txt

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and prints, per captured launch, the metrics
the roofline discussion in DESIGN.md / BASELINE.md uses.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main():
    if len(sys.argv) < 2:
        print(__doc__); return 2
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(head)}
    if len(sys.argv) > 2:
        print("# " + sys.argv[2] + "\n")
    for r in data:
        if len(r) < len(head):
            continue
        print(f"Kernel Name: {r[col['Kernel Name']]} ")
        for k in ("Grid Size", "Block Size"):
            if k in col:
                print(f"{k}: {r[col[k]]} ")
        for m in METRICS:
            if m in col:
                print(f"{m}: {r[col[m]]} {units[col[m]]}")
        print()
    return 0


if __name__ == "__main__":
    sys.exit(main())
