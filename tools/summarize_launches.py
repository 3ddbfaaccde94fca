#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch
list (one row per launch).  Usage: summarize_launches.py launches.csv [top]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], newline=""))
            if len(r) > 10]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "")
        name = name.replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    total = sum(v[1] for v in agg.values())
    print("%-64s %6s %10s %9s" % ("kernel", "n", "total ms", "per ms"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-64s %6d %10.3f %9.4f" % (name[:64], n, t, t / n))
    print("%d launches, %.1f ms of device time" % (len(rows) - 1, total))


if __name__ == "__main__":
    main()
