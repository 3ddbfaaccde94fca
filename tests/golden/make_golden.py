"""Generate tests/golden/golden.npz: outputs of the REFERENCE ITSELF (its own C
sources compiled in place into oracle/_ref/libsvtref.so by oracle/Makefile)
on every case of tests/cases.py.  Run in the build container, where
/root/reference exists:

    python tests/golden/make_golden.py

The .npz is committed; the tests never need /root/reference.
Keys:  stat|<case>|col|<op>|<na_rm>|<center>|<dims>          value
       stat|<case>|row|<op>|<na_rm>|<center kind>            value
       stat|<case>|rowd<dims>|<op>|<na_rm>                   row*(x, dims >= 2)
       stat|<case>|rowMeans|<na_rm>, rowVars, rowSds         R compositions
       summ|<case>|<op>|<na_rm>|<center>                     C_summarize_SVT
       gs|<case>|rowsum|<na_rm>, gs|<case>|colsum|<na_rm>    C_rowsum/colsum_SVT
       cps|<case>|xy / yx / xx                               C_crossprod2_SVT_SVT,
                                                             C_crossprod1_SVT
       cp|<case>|left / cp|<case>|right                      crossprod
       mm|<case>                                             %*% via t(x)
       each with a companion '<key>|warn' (number of R warnings raised).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, ROOT)
sys.path.insert(0, TESTS)

import cases  # noqa: E402
from oracle import refcall, port  # noqa: E402
from sparsearray_b200.svt import SVT_SparseArray  # noqa: E402


def key_col(name, op, na_rm, center, dims):
    return "stat|%s|col|%s|%d|%s|%d" % (name, op, int(na_rm),
                                        "NULL" if center is None
                                        else repr(center), dims)


def key_row(name, op, na_rm, kind):
    return "stat|%s|row|%s|%d|%s" % (name, op, int(na_rm), kind or "NULL")


def key_summ(name, op, na_rm, center):
    return "summ|%s|%s|%d|%s" % (name, op, int(na_rm),
                                 "NULL" if center is None else repr(center))


def transpose_svt(x):
    """t(x) through the oracle's restatement of transpose_2D_SVT()."""
    tp, to, tv = port.transpose(x.dim[0], x.dim[1], x.ptr, x.offs, x.vals,
                                x.type, x.lacunar)
    return SVT_SparseArray((x.dim[1], x.dim[0]), x.type, tp, to, tv)


def main():
    assert refcall.available(), "build oracle/_ref first (make -C oracle)"
    out = {}

    def put(key, res):
        out[key] = np.asarray(res.value)
        out[key + "|warn"] = np.array(len(res.warnings))

    for name, x in cases.stat_cases().items():
        for op, na_rm, center, dims in cases.col_requests(x):
            try:
                res = refcall.colStats(x, op, na_rm=na_rm, center=center,
                                       dims=dims)
            except Exception as e:   # op rejected by the reference
                out[key_col(name, op, na_rm, center, dims) + "|error"] = \
                    np.array(str(e))
                continue
            put(key_col(name, op, na_rm, center, dims), res)
        for op, na_rm, kind in cases.row_requests(x):
            res = refcall.rowStats(x, op, na_rm=na_rm,
                                   center=cases.row_center(x, kind))
            put(key_row(name, op, na_rm, kind), res)
        for op, na_rm, dims in cases.row_requests_nd(x):
            put("stat|%s|rowd%d|%s|%d" % (name, dims, op, int(na_rm)),
                refcall.rowStats(x, op, na_rm=na_rm, dims=dims))
        for op, na_rm, center in cases.summarize_requests(x):
            put(key_summ(name, op, na_rm, center),
                refcall.summarize(x, op, na_rm=na_rm, center=center))
        if len(x.dim) >= 2:
            for na_rm in (False, True):
                out["stat|%s|rowMeans|%d" % (name, na_rm)] = \
                    np.asarray(refcall.rowMeans(x, na_rm=na_rm))
                out["stat|%s|rowVars|%d" % (name, na_rm)] = \
                    np.asarray(refcall.rowVars(x, na_rm=na_rm))
                out["stat|%s|rowSds|%d" % (name, na_rm)] = \
                    np.asarray(refcall.rowSds(x, na_rm=na_rm))
    for name, (x, rg, nrg, cg, ncg) in cases.groupsum_cases().items():
        for na_rm in (False, True):
            put("gs|%s|rowsum|%d" % (name, na_rm),
                refcall.rowsum(x, rg, nrg, na_rm=na_rm))
            put("gs|%s|colsum|%d" % (name, na_rm),
                refcall.colsum(x, cg, ncg, na_rm=na_rm))
    for name, (x, y, ty) in cases.crossprod_cases().items():
        put("cp|%s|left" % name,
            refcall.crossprod2_SVT_mat(x, y, transpose_y=ty))
        put("cp|%s|right" % name,
            refcall.crossprod2_mat_SVT(y, x, transpose_x=ty))
    for name, (x, y) in cases.sparse_crossprod_cases().items():
        put("cps|%s|xy" % name, refcall.crossprod2_SVT_SVT(x, y))
        put("cps|%s|yx" % name, refcall.crossprod2_SVT_SVT(y, x))
        put("cps|%s|xx" % name, refcall.crossprod1_SVT(x))
    for name, (x, d) in cases.matmul_cases().items():
        # `x %*% d` in the reference: crossprod(t(x), d)
        put("mm|%s" % name, refcall.crossprod2_SVT_mat(transpose_svt(x), d))
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d entries, %d bytes" % (path, len(out),
                                              os.path.getsize(path)))


if __name__ == "__main__":
    main()
