#!/usr/bin/env python
"""The instructions with the most warp-stall samples from
`ncu -i X.ncu-rep --page source --csv` (one kernel).  Usage:
    top_stalls.py source.csv [n]"""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if r and "Address" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ei = hdr.index("Instructions Executed")
    total = sum(num(r[si]) for r in data)
    execd = sum(num(r[ei]) for r in data)
    print("samples %d, warp instructions %d" % (total, execd))
    print("%8s %6s %12s  %s" % ("samples", "%", "executed", "instruction"))
    top = sorted(data, key=lambda r: -num(r[si]))[:n]
    for r in top:
        print("%8d %6.2f %12d  %s" % (num(r[si]), 100 * num(r[si]) / total,
                                      num(r[ei]), r[1].strip()[:80]))


if __name__ == "__main__":
    main()
