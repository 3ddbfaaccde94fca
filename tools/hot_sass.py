#!/usr/bin/env python
"""The hot instructions of one kernel from `ncu -i X.ncu-rep --page source
--csv [-k regex:NAME]` (first launch in the file): every instruction executed
at least FRAC x the most executed one, with its stall samples.  Usage:
    hot_sass.py source.csv [frac=0.3] [units]
`units` (e.g. nonzeros / 32) turns the total into warp instructions per unit."""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = heads[0]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = ix["Instructions Executed"]
    st = ix["Warp Stall Sampling (All Samples)"]
    tot = sum(num(r[ex]) for r in data)
    samples = sum(num(r[st]) for r in data)
    print("kernel:", rows[hi - 1][1] if hi > 0 else "?")
    print("warp instructions %d, stall samples %d" % (tot, samples))
    if units > 0:
        print("warp instructions per unit: %.2f" % (tot / units))
    mx = max(num(r[ex]) for r in data)
    hot = 0.0
    for r in data:
        n = num(r[ex])
        if n >= frac * mx:
            hot += n
            print("%12d %6.2f%% %s" % (n, 100 * num(r[st]) / max(samples, 1),
                                       r[1].strip()[:100]))
    print("hot instructions: %.1f%% of all" % (100 * hot / tot))
    reasons = {}
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            reasons[h] = sum(num(r[ix[h]]) for r in data)
    t = sum(reasons.values()) or 1.0
    print("stalls:", ", ".join("%s %.1f%%" % (k, 100 * v / t) for k, v in
                               sorted(reasons.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
