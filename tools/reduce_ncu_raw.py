#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns profiles/ keeps
(one row per launch).  Usage: reduce_ncu_raw.py raw.csv > small.csv"""
import csv
import re
import sys

KEEP = [
    "ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_.*_per_issue_active"
                   r"|smsp__average_warp_latency_issue_stalled_.*")


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    cols += [i for i, h in enumerate(hdr) if STALL.match(h) and i not in cols]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in cols])
    w.writerow([units[i] for i in cols])
    for r in data:
        if len(r) == len(hdr):
            w.writerow([r[i] for i in cols])


if __name__ == "__main__":
    main()
