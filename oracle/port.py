"""TEST INFRASTRUCTURE -- ctypes front-end to oracle/libsvtoracle.so, the CPU
restatement (svt_oracle.c) of the reference's SVT statistics / crossprod
algorithms on flat CSC arrays.  Checked against the reference's compiled C in
tests/test_golden.py (golden vectors written from the compiled reference by
tests/golden/make_golden.py).  Never imported by the product.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsvtoracle.so")

OPCODES = {"anyNA": 1, "countNAs": 2, "any": 3, "all": 4, "min": 5, "max": 6,
           "range": 7, "sum": 8, "prod": 9, "mean": 10, "centered_X2_sum": 11,
           "sum_X_X2": 12, "var1": 13, "var2": 14, "sd1": 15, "sd2": 16}
RTYPE = {"logical": 10, "integer": 13, "double": 14}
NA_REAL = np.array([0x7FF00000000007A2], dtype=np.uint64).view(np.float64)[0]


class _Csc(ctypes.Structure):
    _fields_ = [("nrow", ctypes.c_int64), ("nleaf", ctypes.c_int64),
                ("leaf_ptr", ctypes.c_void_p), ("offs", ctypes.c_void_p),
                ("vals", ctypes.c_void_p), ("val_type", ctypes.c_int),
                ("lacunar", ctypes.c_void_p)]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(
            ["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH,
             os.path.join(_HERE, "svt_oracle.c"), "-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar):
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    keep = [ptr, offs]
    c = _Csc()
    c.nrow, c.nleaf = int(nrow), int(nleaf)
    c.leaf_ptr = ptr.ctypes.data
    c.offs = offs.ctypes.data
    c.val_type = RTYPE[type_]
    c.vals = None
    if vals is not None:
        vals = np.ascontiguousarray(
            vals, dtype=np.float64 if type_ == "double" else np.int32)
        keep.append(vals)
        c.vals = vals.ctypes.data
    c.lacunar = None
    if lacunar is not None:
        lacunar = np.ascontiguousarray(lacunar, dtype=np.uint8)
        keep.append(lacunar)
        c.lacunar = lacunar.ctypes.data
    return c, keep


def _out_is_int(op, type_):
    return op in ("anyNA", "any", "all") or \
        (op in ("min", "max") and type_ != "double")


def colstats(nrow, nleaf, ptr, offs, vals, type_, op, na_rm=False,
             center=None, group=1, lacunar=None):
    """Returns (values, warn)."""
    c, keep = _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar)
    nout = nleaf // group
    out = np.zeros(nout, dtype=np.int32 if _out_is_int(op, type_)
                   else np.float64)
    warn = ctypes.c_int(0)
    cen = NA_REAL if center is None else float(center)
    rc = lib().svt_oracle_colstats(
        ctypes.byref(c), OPCODES[op], int(na_rm), ctypes.c_double(cen),
        ctypes.c_int64(group), out.ctypes.data_as(ctypes.c_void_p),
        ctypes.byref(warn))
    if rc != 0:
        raise ValueError("svt_oracle_colstats: unsupported op/type (%d)" % rc)
    return out, bool(warn.value)


INT_MAX = 2147483647
NA_INT = -2147483648


def summarize(nrow, nleaf, ptr, offs, vals, type_, op, na_rm=False,
              center=None, lacunar=None):
    """C_summarize_SVT (src/SparseArray_summarization.c:89-142): the whole
    array as one vector = the column engine over a single segment of all
    leaves, then the result typing of res2nakedSEXP()
    (src/Rvector_summarization.c:1245-1301).  Returns (values, warn); values
    has length 1 (2 for "range") and dtype bool-as-int32 / int32 / float64."""
    if nleaf == 0 or nrow == 0:
        # an empty vector: one empty segment
        nrow, nleaf = 0, 1
        ptr = np.zeros(2, dtype=np.int64)
        offs = np.zeros(0, dtype=np.int32)
        vals = None if vals is None else vals[:0]
        lacunar = None
    parts = ["min", "max"] if op == "range" else [op]
    vs, warn = [], False
    for o in parts:
        v, w = colstats(nrow, nleaf, ptr, offs, vals, type_, o, na_rm,
                        center, nleaf, lacunar)
        vs.append(v[0])
        warn = warn or w
    if op in ("anyNA", "any", "all"):
        return np.array(vs, dtype=np.int32), warn
    if op == "countNAs":
        if vs[0] > INT_MAX:
            return np.array(vs, dtype=np.float64), warn
        return np.array([int(vs[0] + 0.5)], dtype=np.int32), warn
    if op in ("min", "max", "range") and type_ != "double":
        return np.array(vs, dtype=np.int32), warn
    if op in ("sum", "prod") and type_ != "double":
        v = float(vs[0])
        if np.isnan(v):
            return np.array([NA_INT], dtype=np.int32), warn
        if v < -INT_MAX or v > INT_MAX:
            return np.array([v], dtype=np.float64), warn
        return np.array([int(v + 0.5 if v >= 0 else v - 0.5)],
                        dtype=np.int32), warn
    return np.array(vs, dtype=np.float64), warn


def _groupsum(fn, nrow, nleaf, ptr, offs, vals, type_, group, ngroup, na_rm,
              lacunar, shape):
    c, keep = _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar)
    group = np.ascontiguousarray(group, dtype=np.int32)
    out = np.zeros(shape[0] * shape[1],
                   dtype=np.float64 if type_ == "double" else np.int32)
    ov = ctypes.c_int(0)
    rc = fn(ctypes.byref(c), group.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_int32(ngroup), int(na_rm),
            out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ov))
    if rc != 0:
        raise ValueError("rowsum()/colsum(): unsupported type (%d)" % rc)
    return out.reshape(shape, order="F"), bool(ov.value)


def rowsum(nrow, nleaf, ptr, offs, vals, type_, group, ngroup, na_rm=False,
           lacunar=None):
    """C_rowsum_SVT: (ngroup x ncol matrix, overflow warning)."""
    return _groupsum(lib().svt_oracle_rowsum, nrow, nleaf, ptr, offs, vals,
                     type_, group, ngroup, na_rm, lacunar, (ngroup, nleaf))


def colsum(nrow, nleaf, ptr, offs, vals, type_, group, ngroup, na_rm=False,
           lacunar=None):
    """C_colsum_SVT: (nrow x ngroup matrix, overflow warning)."""
    return _groupsum(lib().svt_oracle_colsum, nrow, nleaf, ptr, offs, vals,
                     type_, group, ngroup, na_rm, lacunar, (nrow, ngroup))


def rowstats(nrow, nleaf, ptr, offs, vals, type_, op, na_rm=False,
             center=None, lacunar=None):
    c, keep = _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar)
    out = np.zeros(nrow, dtype=np.int32
                   if (op == "anyNA" or _out_is_int(op, type_))
                   else np.float64)
    warn = ctypes.c_int(0)
    cp = None
    if center is not None:
        center = np.ascontiguousarray(center, dtype=np.float64)
        cp = center.ctypes.data_as(ctypes.c_void_p)
    rc = lib().svt_oracle_rowstats(
        ctypes.byref(c), OPCODES[op], int(na_rm), cp,
        out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(warn))
    if rc != 0:
        raise ValueError("svt_oracle_rowstats: unsupported op (%d)" % rc)
    return out, bool(warn.value)


def crossprod(nrow, nleaf, ptr, offs, vals, type_, y, transpose_y=False,
              svt_on_left=True, lacunar=None):
    """ans = crossprod(svt, y) (nleaf x K) or crossprod(y, svt) (K x nleaf)."""
    c, keep = _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar)
    y = np.asarray(y)
    yf = np.asfortranarray(y, dtype=np.float64 if type_ == "double"
                           else np.int32)
    K = y.shape[0] if transpose_y else y.shape[1]
    shape = (nleaf, K) if svt_on_left else (K, nleaf)
    ans = np.zeros(shape, dtype=np.float64, order="F")
    rc = lib().svt_oracle_crossprod(
        ctypes.byref(c), yf.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int64(y.shape[0]), ctypes.c_int64(y.shape[1]),
        int(transpose_y), int(svt_on_left),
        ans.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("svt_oracle_crossprod: non-conformable/unsupported")
    return ans


def transpose(nrow, nleaf, ptr, offs, vals, type_, lacunar=None):
    """CSC of t(x): (ptr[nrow+1], offs, vals)."""
    c, keep = _csc(nrow, nleaf, ptr, offs, vals, type_, lacunar)
    nnz = int(np.asarray(ptr)[-1])
    L = lib()
    tp, to, tv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    L.svt_oracle_transpose(ctypes.byref(c), ctypes.byref(tp),
                           ctypes.byref(to), ctypes.byref(tv))
    vt = np.float64 if type_ == "double" else np.int32

    def take(p, n, dt):
        if n == 0:
            return np.zeros(0, dtype=dt)
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt).copy()

    out = (take(tp, nrow + 1, np.int64), take(to, nnz, np.int32),
           take(tv, nnz, vt))
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    for p in (tp, to, tv):
        libc.free(p)
    return out
