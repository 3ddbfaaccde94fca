"""ctypes front-end to the R-API shim (rshim/librshim.so).

R is not installed where this repository is built and measured, so `.Call`
entry points -- the reference's (oracle/_ref/libsvtref.so) and our own glue
(sparsearray_b200/rglue) -- are driven from Python with SEXPs built here.
This module only constructs/reads R objects and performs the `.Call`; it does
no arithmetic.  See rshim/include/Rinternals.h.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librshim.so")

NILSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, VECSXP, RAWSXP = \
    0, 9, 10, 13, 14, 15, 16, 19, 24

NA_INTEGER = -2**31
# R's NA_real_: a NaN whose low word is 1954.
NA_REAL = np.array([0x7FF00000000007A2], dtype=np.uint64).view(np.float64)[0]
_TYPE_OF_STRING = {"logical": LGLSXP, "integer": INTSXP, "double": REALSXP}
_NP_OF_TYPE = {LGLSXP: np.int32, INTSXP: np.int32, REALSXP: np.float64}


def is_na_real(x):
    """Elementwise R_IsNA(): NaN with low word 1954."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    bits = x.view(np.uint64)
    return np.isnan(x) & ((bits & np.uint64(0xFFFFFFFF)) == np.uint64(1954))


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(
            ["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH,
             os.path.join(_HERE, "rshim.c"), "-lm"])
    return _LIB_PATH


class SEXPREC(ctypes.Structure):
    pass


SEXP = ctypes.POINTER(SEXPREC)
SEXPREC._fields_ = [
    ("type", ctypes.c_uint),
    ("owns_data", ctypes.c_int),
    ("length", ctypes.c_ssize_t),
    ("data", ctypes.c_void_p),
    ("dim", SEXP),
    ("names", SEXP),
    ("dimnames", SEXP),
    ("finalizer", ctypes.c_void_p),
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        L.Rf_allocVector.restype = SEXP
        L.Rf_allocVector.argtypes = [ctypes.c_uint, ctypes.c_ssize_t]
        L.rshim_wrap_vector.restype = SEXP
        L.rshim_wrap_vector.argtypes = [ctypes.c_uint, ctypes.c_ssize_t,
                                        ctypes.c_void_p]
        L.Rf_mkChar.restype = SEXP
        L.Rf_mkChar.argtypes = [ctypes.c_char_p]
        L.SET_VECTOR_ELT.restype = SEXP
        L.SET_VECTOR_ELT.argtypes = [SEXP, ctypes.c_ssize_t, SEXP]
        L.SET_STRING_ELT.restype = None
        L.SET_STRING_ELT.argtypes = [SEXP, ctypes.c_ssize_t, SEXP]
        L.rshim_release.argtypes = [SEXP]
        L.rshim_release_tree.argtypes = [SEXP]
        L.rshim_svt_from_csc.restype = SEXP
        L.rshim_svt_from_csc.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint, ctypes.c_void_p]
        L.rshim_try_call.restype = SEXP
        L.rshim_try_call.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                     ctypes.POINTER(SEXP),
                                     ctypes.POINTER(ctypes.c_int)]
        L.rshim_last_error.restype = ctypes.c_char_p
        L.rshim_warning_count.restype = ctypes.c_int
        L.rshim_warning_message.restype = ctypes.c_char_p
        L.rshim_warning_message.argtypes = [ctypes.c_int]
        L.rshim_live_objects.restype = ctypes.c_long
        _lib = L
    return _lib


def nil():
    return SEXP.in_dll(lib(), "R_NilValue")


def _is_nil(s):
    return ctypes.addressof(s.contents) == ctypes.addressof(nil().contents)


class RError(RuntimeError):
    """An R-level error() raised inside a `.Call` entry point."""


class RObj:
    """An R object living in the shim heap, plus the numpy buffers it wraps.

    `keep` pins any Python-owned memory that the SEXP tree points into.
    """

    def __init__(self, sexp, keep=()):
        self.sexp = sexp
        self.keep = list(keep)

    def release(self):
        if self.sexp is not None and not _is_nil(self.sexp):
            lib().rshim_release_tree(self.sexp)
        self.sexp = None
        self.keep = []


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def wrap(a, rtype):
    """Zero-copy R vector over a C-contiguous numpy array."""
    a = np.ascontiguousarray(a, dtype=_NP_OF_TYPE[rtype])
    return RObj(lib().rshim_wrap_vector(rtype, a.size, _ptr(a)), [a])


def integer(values):
    return wrap(np.asarray(values, dtype=np.int32).reshape(-1), INTSXP)


def logical(values):
    return wrap(np.asarray(values, dtype=np.int32).reshape(-1), LGLSXP)


def real(values):
    return wrap(np.asarray(values, dtype=np.float64).reshape(-1), REALSXP)


def string(values):
    """character vector; None elements become NA_character_."""
    if isinstance(values, str):
        values = [values]
    L = lib()
    s = L.Rf_allocVector(STRSXP, len(values))
    na_string = SEXP.in_dll(L, "R_NaString")
    for i, v in enumerate(values):
        L.SET_STRING_ELT(s, i, na_string if v is None
                         else L.Rf_mkChar(v.encode()))
    return RObj(s)


def rlist(elts):
    """list(...) of RObj (None -> NULL). Takes ownership of the elements."""
    L = lib()
    s = L.Rf_allocVector(VECSXP, len(elts))
    keep = []
    for i, e in enumerate(elts):
        if e is None:
            continue
        L.SET_VECTOR_ELT(s, i, e.sexp)
        keep.extend(e.keep)
        e.sexp, e.keep = None, []
    return RObj(s, keep)


def matrix(a, rtype):
    """R matrix (column-major) from a 2-D numpy array."""
    a = np.asarray(a)
    assert a.ndim == 2
    flat = np.asfortranarray(a, dtype=_NP_OF_TYPE[rtype]).reshape(-1, order="F")
    flat = np.ascontiguousarray(flat)
    obj = RObj(lib().rshim_wrap_vector(rtype, flat.size, _ptr(flat)), [flat])
    dim = integer(list(a.shape))
    obj.sexp.contents.dim = dim.sexp
    obj.keep.extend(dim.keep)
    dim.sexp = None
    return obj


def svt_from_csc(ncol, ptr, offs, vals, rtype, lacunar=None):
    """SVT list for a 2-D matrix from CSC arrays (wrapped in place).

    vals=None builds an all-lacunar SVT; lacunar (uint8 per column) marks
    individual lacunar leaves.
    """
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    keep = [ptr, offs]
    vp = None
    if vals is not None:
        vals = np.ascontiguousarray(vals, dtype=_NP_OF_TYPE[rtype])
        keep.append(vals)
        vp = _ptr(vals)
    lp = None
    if lacunar is not None:
        lacunar = np.ascontiguousarray(lacunar, dtype=np.uint8)
        keep.append(lacunar)
        lp = _ptr(lacunar)
    s = lib().rshim_svt_from_csc(int(ncol), _ptr(ptr), _ptr(offs), vp,
                                 rtype, lp)
    return RObj(s, keep)


def to_numpy(s):
    """Copy an R atomic vector/array out of the shim heap.

    Returns (array, names) where array has R's dim (column-major order
    preserved) and names is a list of str/None or None.
    """
    rec = s.contents
    t = rec.type
    if t == NILSXP:
        return None, None
    n = rec.length
    if t in (LGLSXP, INTSXP):
        ct = ctypes.c_int32
    elif t == REALSXP:
        ct = ctypes.c_double
    else:
        raise TypeError("to_numpy(): unsupported SEXPTYPE %d" % t)
    if n:
        buf = ctypes.cast(rec.data, ctypes.POINTER(ct * n)).contents
        a = np.frombuffer(buf, dtype=_NP_OF_TYPE[t]).copy()
    else:
        a = np.zeros(0, dtype=_NP_OF_TYPE[t])
    if not _is_nil(rec.dim):
        d, _ = to_numpy(rec.dim)
        a = a.reshape(tuple(int(x) for x in d), order="F")
    names = None
    if not _is_nil(rec.names):
        names = strings(rec.names)
    return a, names


def strings(s):
    rec = s.contents
    if rec.type == NILSXP:
        return None
    out = []
    elts = ctypes.cast(rec.data, ctypes.POINTER(SEXP))
    na_addr = ctypes.addressof(SEXP.in_dll(lib(), "R_NaString").contents)
    for i in range(rec.length):
        e = elts[i]
        if ctypes.addressof(e.contents) == na_addr:
            out.append(None)
        else:
            out.append(ctypes.string_at(e.contents.data).decode())
    return out


def dimnames(s):
    """dimnames attribute as a list of (list of str | None), or None."""
    rec = s.contents
    if _is_nil(rec.dimnames):
        return None
    dn = rec.dimnames.contents
    elts = ctypes.cast(dn.data, ctypes.POINTER(SEXP))
    return [strings(elts[i]) for i in range(dn.length)]


def sexptype(s):
    return s.contents.type


def dot_call(fn, args):
    """`.Call(fn, ...)`: returns (result SEXP, [warning messages]).

    `fn` is a C function pointer (ctypes function or address); `args` RObj /
    None.  Raises RError when the routine calls error().  The caller owns the
    result (rshim_release_tree).
    """
    L = lib()
    L.rshim_clear_warnings()
    arr = (SEXP * max(len(args), 1))()
    for i, a in enumerate(args):
        arr[i] = nil() if a is None else a.sexp
    status = ctypes.c_int(0)
    addr = fn if isinstance(fn, int) else ctypes.cast(fn, ctypes.c_void_p).value
    ans = L.rshim_try_call(addr, len(args), arr, ctypes.byref(status))
    warns = [L.rshim_warning_message(i).decode()
             for i in range(L.rshim_warning_count())]
    if status.value != 0:
        raise RError(L.rshim_last_error().decode())
    return ans, warns


def release_result(ans):
    """Free a `.Call` result.  Its names/dimnames may be shared with the
    inputs (R shares them; the shim has no GC), so those attributes are
    detached rather than freed -- whoever created them releases them."""
    if ans is None or _is_nil(ans):
        return

    def detach(s):
        rec = s.contents
        if rec.type == VECSXP:
            elts = ctypes.cast(rec.data, ctypes.POINTER(SEXP))
            for i in range(rec.length):
                if not _is_nil(elts[i]):
                    detach(elts[i])
        else:
            rec.names = nil()   # a list's names are freshly allocated
        rec.dimnames = nil()

    detach(ans)
    lib().rshim_release_tree(ans)
