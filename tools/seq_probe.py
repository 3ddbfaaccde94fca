#!/usr/bin/env python
"""Does an op's time depend on what ran before it?  Event-timed op pairs on
the 33,538 x COLS count shard."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsearray_b200.device import DeviceSVT
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
s = DeviceSVT.generate_poisson(33538, cols, 0.07, seed=2, na_rate=1e-6)
ops = {
    "colSums": lambda: s.colstats("sum", na_rm=True),
    "colMeans": lambda: s.colstats("mean", na_rm=True),
    "rowSums": lambda: s.rowstats("sum", na_rm=True),
    "rowVars": lambda: s.rowmoments(na_rm=True),
    "rowMaxs": lambda: s.rowstats("max", na_rm=True),
    "sleep": lambda: torch.cuda._sleep(2000000),
    "fill1G": lambda: big.zero_(),
    "copy1G": lambda: big2.copy_(big),
}
big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
big2 = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
def timed(seq, reps=10):
    for o in seq: ops[o]()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)] for _ in range(reps)]
    for r in range(reps):
        ev[r][0].record()
        for i, o in enumerate(seq):
            ops[o]()
            ev[r][i + 1].record()
    torch.cuda.synchronize()
    return [sum(ev[r][i].elapsed_time(ev[r][i + 1]) for r in range(reps)) / reps for i in range(len(seq))]
for seq in (["rowSums"], ["rowVars"], ["colMeans", "rowSums"], ["rowVars", "rowSums"], ["colMeans", "rowVars"],
            ["colSums", "colMeans", "rowSums", "rowVars"], ["colSums", "colMeans", "rowVars", "rowSums"],
            ["colMeans", "rowMaxs"], ["colMeans", "sleep", "rowSums"], ["colMeans", "fill1G", "rowSums"],
            ["copy1G", "rowSums"], ["sleep", "rowSums"], ["rowSums", "colMeans"], ["rowSums", "sleep", "colMeans"]):
    t = timed(seq)
    print(" + ".join("%s %.3f" % (o, x) for o, x in zip(seq, t)), flush=True)
