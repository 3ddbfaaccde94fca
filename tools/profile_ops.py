#!/usr/bin/env python
"""Small fixed workload for ncu: each requested op `reps` times on a resident
33,538 x COLS count shard.  Usage:
    python tools/profile_ops.py --cols 50000 --ops colSums,rowSums,rowVars,crossprod,matmul
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from sparsearray_b200.device import DeviceSVT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cols", type=int, default=50000)
    ap.add_argument("--nrow", type=int, default=33538)
    ap.add_argument("--density", type=float, default=0.07)
    ap.add_argument("--ops", default="colSums,colVars,rowSums,rowVars")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--lacunar", action="store_true")
    a = ap.parse_args()
    ops = a.ops.split(",")
    need_dbl = any(o in ("crossprod", "matmul", "colSumsD", "rowSumsD",
                         "colVarsD", "rowVarsD") for o in ops)
    s = DeviceSVT.generate_poisson(a.nrow, a.cols, a.density, seed=2,
                                   na_rate=1e-6, lacunar=a.lacunar)
    d = None
    if need_dbl:
        d = DeviceSVT.generate_poisson(a.nrow, a.cols, a.density, seed=2,
                                       na_rate=0.0, val_type="double")
        Y = torch.randn(a.nrow, a.K, dtype=torch.float64, device="cuda")
        D = torch.randn(a.cols, a.K, dtype=torch.float64, device="cuda")
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(3))
    RG = rng.integers(1, 13, size=a.nrow).astype(np.int32)
    CG = rng.integers(1, 9, size=a.cols).astype(np.int32)
    for op in ops:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        for r in range(a.reps + 1):
            if r == 1:
                ev0.record()
            if op == "colSums":
                s.colstats("sum", na_rm=True)
            elif op == "colVars":
                s.colstats("var1", na_rm=True)
            elif op == "colMaxs":
                s.colstats("max", na_rm=True)
            elif op == "rowSums":
                s.rowstats("sum", na_rm=True)
            elif op == "rowMaxs":
                s.rowstats("max", na_rm=True)
            elif op == "rowVars":
                s.rowmoments(na_rm=True)
            elif op == "rowsum":
                kms = s.rowsum(RG, 12, na_rm=True)[2]
            elif op == "colsum":
                kms = s.colsum(CG, 8, na_rm=True)[2]
            elif op == "sum":
                s.summarize("sum", na_rm=True)
            elif op == "var":
                s.summarize("var1", na_rm=True)
            elif op == "colVarsD":
                d.colstats("var1", na_rm=True)
            elif op == "rowVarsD":
                d.rowmoments(na_rm=True)
            elif op == "colSumsD":
                d.colstats("sum", na_rm=True)
            elif op == "rowSumsD":
                d.rowstats("sum", na_rm=True)
            elif op == "crossprod":
                d.crossprod(Y)
            elif op == "matmul":
                d.matmul(D)
            else:
                raise SystemExit("unknown op " + op)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / a.reps
        if op in ("rowsum", "colsum"):
            ms = kms     # kernels only: the result matrix goes to the host
        nnz = (d if op in ("crossprod", "matmul", "colSumsD", "rowSumsD",
                           "colVarsD", "rowVarsD") else s).nnz
        print("%-10s %8.3f ms  %.3e nnz/s" % (op, ms, nnz / ms * 1e3),
              flush=True)


if __name__ == "__main__":
    main()
