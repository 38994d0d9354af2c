#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the small text summaries committed here.

    python profiles/extract.py gpurun_out/prof_r1d.ncu-rep profiles/r01d_ncu_full.csv

Writes one CSV row per profiled launch with the metrics the roofline discussion uses.
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{hdr[c]} [{units[c]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index("Kernel Name")]] + [r[c] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
