#!/usr/bin/env python
"""Randomised parity run: many small-to-mid random SVT matrices (shapes,
densities, types, NA rates, lacunar / regular, value ranges) through the
.Call entry points on the GPU against the oracle port (TEST
INFRASTRUCTURE: oracle/ is only the checker).  Usage:
    python tests/fuzz_gpu.py [iterations=200] [seed=1]
Prints the failing case and exits 1 at the first mismatch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # (this directory)

import numpy as np

import sparsearray_b200 as sa
import runners
import conditioning as C
import fixtures as fx
from rcompare import assert_identical, assert_close


def make(rng):
    kind = rng.choice(["integer", "double", "lacunar", "counts"])
    nrow = int(2 ** rng.uniform(0, 17.2))
    ncol = int(2 ** rng.uniform(0, 9.5))
    dens = float(10 ** rng.uniform(-3, -0.05)) if rng.random() < 0.85 else 1.0
    cols = []
    for j in range(ncol):
        d = dens * (3.0 if rng.random() < 0.05 else 1.0)
        if rng.random() < 0.05:
            d = 0.0
        if d >= 1.0:
            o = np.arange(nrow, dtype=np.int32)
        else:
            n = rng.binomial(nrow, min(d, 1.0))
            o = np.sort(rng.choice(nrow, size=n, replace=False)).astype(
                np.int32) if n > 0 else np.zeros(0, np.int32)
        cols.append(o)
    ptr = np.zeros(ncol + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([o.size for o in cols])
    offs = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    nnz = offs.size
    na = float(10 ** rng.uniform(-4, -1)) if rng.random() < 0.6 else 0.0
    if kind == "lacunar":
        return sa.SVT_SparseArray((nrow, ncol), "integer", ptr, offs, None), kind
    if kind == "double":
        v = rng.standard_normal(nnz) * float(10 ** rng.uniform(-3, 3))
        v[v == 0] = 1.0
        sp = rng.random(nnz)
        v[sp < na] = fx.NA_R
        if rng.random() < 0.3:
            v[(sp >= na) & (sp < 1.5 * na)] = np.nan
        if rng.random() < 0.15 and nnz:
            v[rng.integers(0, nnz)] = np.inf
        return sa.SVT_SparseArray((nrow, ncol), "double", ptr, offs, v), kind
    if kind == "counts":
        v = rng.integers(1, 13, size=nnz).astype(np.int32)
    else:
        hi = int(10 ** rng.uniform(0.5, 9.2))
        v = rng.integers(-hi, hi + 1, size=nnz).astype(np.int32)
        v[v == 0] = 1
    v[rng.random(nnz) < na] = fx.NA_I
    return sa.SVT_SparseArray((nrow, ncol), "integer", ptr, offs, v), kind


def check(x, kind, rng):
    exact = kind != "double"
    for na_rm in (False, True):
        for op in ("sum", "mean", "var1", "max", "min", "countNAs"):
            v, w = runners.api_col(x, op, na_rm, None, 1)
            e, ew = runners.port_col(x, op, na_rm, None, 1)
            if exact and op in ("sum", "max", "min", "countNAs"):
                assert_identical(v, e, ("col", op, na_rm))
            else:
                assert_close(v, e, rtol=1e-12, what=("col", op, na_rm),
                             cond=C.cond(x, "col", op))
            assert w == ew, ("col warn", op, na_rm)
        for op in ("sum", "max", "min", "countNAs"):
            v, w = runners.api_row(x, op, na_rm, None)
            e, ew = runners.port_row(x, op, na_rm, None)
            if exact or op != "sum":
                assert_identical(v, e, ("row", op, na_rm))
            else:
                assert_close(v, e, rtol=1e-12, what=("row", op, na_rm),
                             cond=C.cond(x, "row", op))
            assert w == ew, ("row warn", op, na_rm)
    for na_rm in (False, True):
        for op in ("sum", "mean", "var1", "range", "countNAs", "anyNA"):
            v, w = runners.api_summarize(x, op, na_rm, None)
            e, ew = runners.port_summarize(x, op, na_rm, None)
            e = np.asarray(e).reshape(-1)
            if exact and op in ("sum", "range", "countNAs", "anyNA"):
                assert_identical(v, e, ("summarize", op, na_rm))
            elif exact and op == "var1":
                # integer var(): exact sums here, the reference's sequential
                # passes there (n * eps)
                assert_close(v, e, rtol=1e-10, what=("summarize", op, na_rm))
            else:
                assert_close(v, e, rtol=1e-12, what=("summarize", op, na_rm),
                             cond=C.cond(x, "col", op, None, len(x.dim)))
            assert w == ew, ("summarize warn", op, na_rm)
    if x.dim[0] > 0 and x.dim[1] > 0:
        nrow, ncol = x.dim
        if kind != "double":
            ng = int(rng.integers(1, 20))
            rg = rng.integers(1, ng + 1, size=nrow).astype(np.int32)
            cg = rng.integers(1, ng + 1, size=ncol).astype(np.int32)
            for na_rm in (False, True):
                v, w = runners.api_rowsum(x, rg, ng, na_rm)
                e, ew = runners.port_rowsum(x, rg, ng, na_rm)
                assert_identical(v, e, ("rowsum", ng, na_rm))
                v, w = runners.api_colsum(x, cg, ng, na_rm)
                e, ew = runners.port_colsum(x, cg, ng, na_rm)
                assert_identical(v, e, ("colsum", ng, na_rm))
        # the device transpose, through a non-native row operation
        if kind != "lacunar" and nrow <= 20000:
            from oracle import port
            tp, to, tv = port.transpose(nrow, ncol, x.ptr, x.offs, x.vals,
                                        x.type, x.lacunar)
            e, _ = port.colstats(ncol, nrow, tp, to, tv, x.type, "mean", True)
            v = np.asarray(sa.svt._rowStats("mean", x, na_rm=True,
                                            useNames=False)).reshape(-1)
            assert_close(v, e, rtol=1e-12, what="row mean via transpose",
                         cond=C.cond(x, "row", "mean"))
        if kind in ("double", "lacunar") and nrow <= 40000:
            K = int(rng.choice([1, 3, 8, 33, 50]))
            xd = x if kind == "double" else x.with_type("double")
            y = rng.standard_normal((nrow, K))
            with np.errstate(all="ignore"):
                assert_close(np.asarray(sa.crossprod(xd, y)),
                             runners.port_crossprod(xd, y, False, True),
                             rtol=1e-12, what=("crossprod", K),
                             cond=C.dot_cond(xd, y))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.Generator(np.random.PCG64(seed))
    t0 = time.time()
    for it in range(iters):
        state = rng.bit_generator.state
        x, kind = make(rng)
        try:
            check(x, kind, rng)
        except Exception:
            print("FUZZ FAIL at iteration %d: kind=%s dim=%s nnz=%d seed=%d"
                  % (it, kind, x.dim, x.nnz, seed), flush=True)
            import pickle
            with open(os.path.join(ROOT, "gpurun_out", "fuzz_fail.pkl"),
                      "wb") as f:
                pickle.dump({"state": state, "iteration": it}, f)
            raise
    print("FUZZ PASS: %d matrices, %.1f s" % (iters, time.time() - t0))


if __name__ == "__main__":
    main()
