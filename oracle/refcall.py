"""TEST INFRASTRUCTURE -- drives the reference's own C code.

oracle/_ref/libsvtref.so is Bioconductor/SparseArray's hot-path C compiled
unmodified (oracle/Makefile) against the R-API shim.  The functions below
perform the same `.Call`s the reference's R code performs
(R/SparseArray-matrixStats.R:104-106, :256-258; R/SparseMatrix-mult.R:50-52,
:83-85) and compose rowMeans/rowVars/rowSds the way the R methods do
(R/SparseArray-matrixStats.R:511-517, :645-684).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  Nothing here is on the product path.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from rshim import rshim  # noqa: E402

_REF_PATH = os.path.join(_HERE, "_ref", "libsvtref.so")
_ref = None


def available():
    return os.path.exists(_REF_PATH)


def ref():
    global _ref
    if _ref is None:
        rshim.lib()
        _ref = ctypes.CDLL(_REF_PATH, mode=ctypes.RTLD_LOCAL)
    return _ref


def _fn(name):
    return ctypes.cast(getattr(ref(), name), ctypes.c_void_p).value


def set_threads(n):
    """C_set_max_threads (src/thread_control.c:59-64); returns previous."""
    a = rshim.integer([int(n)])
    ans, _ = rshim.dot_call(_fn("C_set_max_threads"), [a])
    prev = int(rshim.to_numpy(ans)[0][0])
    rshim.lib().rshim_release_tree(ans)
    return prev


def num_procs():
    ans, _ = rshim.dot_call(_fn("C_get_num_procs"), [])
    n = int(rshim.to_numpy(ans)[0][0])
    rshim.lib().rshim_release_tree(ans)
    return n


class Result:
    def __init__(self, value, names, dimnames, rtype, warnings):
        self.value = value
        self.names = names
        self.dimnames = dimnames
        self.rtype = rtype
        self.warnings = warnings


def _finish(ans, warns):
    value, names = rshim.to_numpy(ans)
    res = Result(value, names, rshim.dimnames(ans), rshim.sexptype(ans), warns)
    rshim.release_result(ans)
    return res


def colStats(x, op, na_rm=False, center=None, dims=1, na_background=False):
    """.Call("C_colStats_SVT", x@dim, dimnames, x@type, x@SVT, FALSE, op,
    na.rm, center, dims).  `x` provides r_dim/r_dimnames/r_type/r_SVT."""
    c = rshim.NA_REAL if center is None else float(center)
    args = [x.r_dim, x.r_dimnames, x.r_type, x.r_SVT,
            rshim.logical([int(na_background)]), rshim.string(op),
            rshim.logical([int(na_rm)]), rshim.real([c]),
            rshim.integer([dims])]
    ans, warns = rshim.dot_call(_fn("C_colStats_SVT"), args)
    return _finish(ans, warns)


def summarize(x, op, na_rm=False, center=None, na_background=False):
    """.Call("C_summarize_SVT", x@dim, x@type, x@SVT, FALSE, op, na.rm,
    center) -- summarize_SVT(), R/SparseArray-summarization.R:19-46."""
    c = rshim.NA_REAL if center is None else float(center)
    args = [x.r_dim, x.r_type, x.r_SVT,
            rshim.logical([int(na_background)]), rshim.string(op),
            rshim.logical([int(na_rm)]), rshim.real([c])]
    ans, warns = rshim.dot_call(_fn("C_summarize_SVT"), args)
    return _finish(ans, warns)


def rowsum(x, group, ngroup, na_rm=False):
    """.Call("C_rowsum_SVT", x@dim, x@type, x@SVT, group, ngroup, na.rm):
    `group` = match(group, ugroup), 1-based ints, NA allowed
    (R/rowsum-methods.R:7-27)."""
    args = [x.r_dim, x.r_type, x.r_SVT, rshim.integer(group),
            rshim.integer([ngroup]), rshim.logical([int(na_rm)])]
    ans, warns = rshim.dot_call(_fn("C_rowsum_SVT"), args)
    return _finish(ans, warns)


def colsum(x, group, ngroup, na_rm=False):
    """.Call("C_colsum_SVT", ...), R/rowsum-methods.R:29-49."""
    args = [x.r_dim, x.r_type, x.r_SVT, rshim.integer(group),
            rshim.integer([ngroup]), rshim.logical([int(na_rm)])]
    ans, warns = rshim.dot_call(_fn("C_colsum_SVT"), args)
    return _finish(ans, warns)


def rowStats(x, op, na_rm=False, center=None, dims=1, na_background=False):
    """.Call("C_rowStats_SVT", ...); center: None or array of length
    prod(head(dim, dims))."""
    cen = None if center is None else rshim.real(center)
    args = [x.r_dim, x.r_dimnames, x.r_type, x.r_SVT,
            rshim.logical([int(na_background)]), rshim.string(op),
            rshim.logical([int(na_rm)]), cen, rshim.integer([dims])]
    ans, warns = rshim.dot_call(_fn("C_rowStats_SVT"), args)
    return _finish(ans, warns)


def crossprod2_SVT_mat(x, y, transpose_y=False, ans_dimnames=None):
    """.Call("C_crossprod2_SVT_mat", x@dim, x@type, x@SVT, y, transpose.y,
    "double", ans_dimnames); y is a numpy matrix of x's type."""
    rtype = rshim.INTSXP if np.asarray(y).dtype.kind in "iu" else rshim.REALSXP
    ym = rshim.matrix(y, rtype)
    args = [x.r_dim, x.r_type, x.r_SVT, ym,
            rshim.logical([int(transpose_y)]), rshim.string("double"),
            ans_dimnames]
    ans, warns = rshim.dot_call(_fn("C_crossprod2_SVT_mat"), args)
    return _finish(ans, warns)


def crossprod2_mat_SVT(x, y, transpose_x=False, ans_dimnames=None):
    """.Call("C_crossprod2_mat_SVT", x, y@dim, y@type, y@SVT, transpose.x,
    "double", ans_dimnames); x is a numpy matrix of y's type."""
    rtype = rshim.INTSXP if np.asarray(x).dtype.kind in "iu" else rshim.REALSXP
    xm = rshim.matrix(x, rtype)
    args = [xm, y.r_dim, y.r_type, y.r_SVT,
            rshim.logical([int(transpose_x)]), rshim.string("double"),
            ans_dimnames]
    ans, warns = rshim.dot_call(_fn("C_crossprod2_mat_SVT"), args)
    return _finish(ans, warns)


class RefSVT:
    """An SVT_SparseMatrix living in the shim heap as the reference built it
    (the result of C_transpose_2D_SVT / C_build_SVT_from_CSC): provides the
    r_dim / r_type / r_SVT the calls above take.  release() frees it."""

    def __init__(self, dim, type_, svt_sexp):
        self.dim = tuple(int(d) for d in dim)
        self.type = type_
        self.r_dim = rshim.integer(list(self.dim))
        self.r_dimnames = None
        self.r_type = rshim.string(type_)
        self.r_SVT = None if rshim._is_nil(svt_sexp) else rshim.RObj(svt_sexp)

    def release(self):
        if self.r_SVT is not None:
            self.r_SVT.release()
            self.r_SVT = None


def transpose_2D_SVT(x):
    """t(x): .Call("C_transpose_2D_SVT", x@dim, x@type, x@SVT),
    R/SparseArray-aperm.R:15 (src/SparseArray_aperm.c:405-423)."""
    ans, _ = rshim.dot_call(_fn("C_transpose_2D_SVT"),
                            [x.r_dim, x.r_type, x.r_SVT])
    dim = [int(d) for d in rshim.to_numpy(x.r_dim.sexp)[0]]
    type_ = x.type if hasattr(x, "type") else None
    return RefSVT((dim[1], dim[0]), type_, ans)


def matmul_SVT_mat(x, y, ans_dimnames=None):
    """x %*% y for an SVT_SparseMatrix x and an ordinary matrix y:
    .crossprod2_SparseMatrix_matrix(t(x), y), R/SparseMatrix-mult.R:196-198."""
    tx = transpose_2D_SVT(x)
    try:
        return crossprod2_SVT_mat(tx, y, ans_dimnames=ans_dimnames)
    finally:
        tx.release()


def build_SVT_from_CSC(dim, indptr, data, indices, indices_are_1based=False):
    """.Call("C_build_SVT_from_CSC", dim, indptr, data, indices,
    indices_are_1based) -- src/SVT_SparseArray_class.c:833-861 (the
    constructor TENxMatrix / HDF5 loaders use, R/SVT_SparseArray-class.R:
    261-275).  data: int32 or float64 numpy array; returns a RefSVT."""
    data = np.asarray(data)
    rtype = rshim.REALSXP if data.dtype.kind == "f" else rshim.INTSXP
    indptr = np.asarray(indptr)
    ip = rshim.real(indptr.astype(np.float64)) if indptr.dtype.kind == "f" \
        or int(indptr[-1]) > 2**31 - 1 else \
        rshim.integer([int(v) for v in indptr])
    args = [rshim.integer([int(d) for d in dim]), ip,
            rshim.wrap(data, rtype),
            rshim.wrap(np.asarray(indices, dtype=np.int32), rshim.INTSXP),
            rshim.logical([int(indices_are_1based)])]
    ans, _ = rshim.dot_call(_fn("C_build_SVT_from_CSC"), args)
    return RefSVT(dim, "double" if rtype == rshim.REALSXP else "integer", ans)


def from_SVT_to_CSC(x, as_ngCMatrix=False):
    """.Call("C_from_SVT_SparseMatrix_to_CsparseMatrix", x@dim, x@type,
    x@SVT, as.ngCMatrix) -- src/SVT_SparseArray_class.c:636-679.  Returns
    (p, i, x-or-None) as numpy arrays."""
    ans, _ = rshim.dot_call(
        _fn("C_from_SVT_SparseMatrix_to_CsparseMatrix"),
        [x.r_dim, x.r_type, x.r_SVT, rshim.logical([int(as_ngCMatrix)])])
    import ctypes
    elts = ctypes.cast(ans.contents.data, ctypes.POINTER(rshim.SEXP))
    out = []
    for k in range(3):
        if rshim._is_nil(elts[k]):
            out.append(None)
        else:
            out.append(np.array(rshim.to_numpy(elts[k])[0]))
    rshim.lib().rshim_release_tree(ans)
    return tuple(out)


# -- R-level compositions (R/SparseArray-matrixStats.R) ----------------------

def crossprod2_SVT_SVT(x, y, ans_dimnames=None):
    """.Call("C_crossprod2_SVT_SVT", x@dim, x@type, x@SVT, y@dim, y@type,
    y@SVT, "double", ans_dimnames) -- R/SparseMatrix-mult.R:103-118."""
    args = [x.r_dim, x.r_type, x.r_SVT, y.r_dim, y.r_type, y.r_SVT,
            rshim.string("double"), ans_dimnames]
    ans, warns = rshim.dot_call(_fn("C_crossprod2_SVT_SVT"), args)
    return _finish(ans, warns)


def crossprod1_SVT(x, ans_dimnames=None):
    """.Call("C_crossprod1_SVT", x@dim, x@type, x@SVT, "double",
    ans_dimnames) -- R/SparseMatrix-mult.R:120-133."""
    args = [x.r_dim, x.r_type, x.r_SVT, rshim.string("double"),
            ans_dimnames]
    ans, warns = rshim.dot_call(_fn("C_crossprod1_SVT"), args)
    return _finish(ans, warns)


def _rowCountVals(x, na_rm, dims=1):
    """.rowCountVals_SparseArray(), :300-310."""
    dim = [int(d) for d in rshim.to_numpy(x.r_dim.sexp)[0]]
    ans = float(np.prod(dim[dims:]))
    if na_rm:
        ans = ans - rowStats(x, "countNAs", dims=dims).value
    return ans


def rowMeans(x, na_rm=False, dims=1):
    """:511-517"""
    sums = rowStats(x, "sum", na_rm=na_rm, dims=dims).value
    with np.errstate(all="ignore"):
        return sums / _rowCountVals(x, na_rm, dims)


def rowVars(x, na_rm=False, center=None, dims=1):
    """:645-661"""
    nvals = _rowCountVals(x, na_rm, dims)
    dim = [int(d) for d in rshim.to_numpy(x.r_dim.sexp)[0]]
    ans_len = int(np.prod(dim[:dims]))
    with np.errstate(all="ignore"):
        if center is None:
            sums = rowStats(x, "sum", na_rm=na_rm, dims=dims).value
            center = sums / nvals
        center = np.asarray(center, dtype=np.float64).reshape(-1)
        if center.size == 1:  # .rowStats_SparseArray() :236-247
            center = np.full(ans_len, center[0])
        x2 = rowStats(x, "centered_X2_sum", na_rm=na_rm, center=center,
                      dims=dims).value
        return x2 / (nvals - 1)


def rowSds(x, na_rm=False, center=None, dims=1):
    """:674-684"""
    with np.errstate(all="ignore"):
        return np.sqrt(rowVars(x, na_rm=na_rm, center=center, dims=dims))
