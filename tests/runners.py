"""Evaluate the requests of tests/cases.py through the oracle port (CPU
restatement) or through the product's reference-facing API (CUDA)."""
import os

import numpy as np

import cases
from oracle import port
import sparsearray_b200 as sa

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                           "golden", "golden.npz")
_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = dict(np.load(GOLDEN_PATH, allow_pickle=False))
    return _golden


def key_col(name, op, na_rm, center, dims):
    return "stat|%s|col|%s|%d|%s|%d" % (name, op, int(na_rm),
                                        "NULL" if center is None
                                        else repr(center), dims)


def key_row(name, op, na_rm, kind):
    return "stat|%s|row|%s|%d|%s" % (name, op, int(na_rm), kind or "NULL")


def key_summ(name, op, na_rm, center):
    return "summ|%s|%s|%d|%s" % (name, op, int(na_rm),
                                 "NULL" if center is None else repr(center))


def _nleaf(x):
    return int(np.prod(x.dim[1:], dtype=np.int64))


def col_out_shape(x, dims):
    return tuple(x.dim[dims:])


# ---- oracle port ---------------------------------------------------------

def port_col(x, op, na_rm, center, dims):
    group = int(np.prod(x.dim[1:dims], dtype=np.int64))
    nleaf = _nleaf(x)
    if group == 0 or nleaf == 0:
        # zero-extent geometry: ask the port for one empty segment per output
        nout = int(np.prod(x.dim[dims:], dtype=np.int64))
        ptr = np.zeros(nout + 1, dtype=np.int64)
        v, w = port.colstats(0, nout, ptr, np.zeros(0, np.int32), None,
                             x.type, op, na_rm, center, 1)
        # an empty vector of length prod(dim[:dims]) == 0
    else:
        v, w = port.colstats(x.dim[0], nleaf, x.ptr, x.offs, x.vals, x.type,
                             op, na_rm, center, group, x.lacunar)
    shape = col_out_shape(x, dims)
    if len(shape) >= 2:
        v = v.reshape(shape, order="F")
    return v, w


def port_summarize(x, op, na_rm, center):
    return port.summarize(x.dim[0], _nleaf(x), x.ptr, x.offs, x.vals, x.type,
                          op, na_rm, center, x.lacunar)


def port_rowsum(x, group, ngroup, na_rm):
    return port.rowsum(x.dim[0], x.dim[1], x.ptr, x.offs, x.vals, x.type,
                       group, ngroup, na_rm, x.lacunar)


def port_colsum(x, group, ngroup, na_rm):
    return port.colsum(x.dim[0], x.dim[1], x.ptr, x.offs, x.vals, x.type,
                       group, ngroup, na_rm, x.lacunar)


def api_rowsum(x, group, ngroup, na_rm):
    r = sa.svt._groupsum("C_rowsum_SVT", x, group, ngroup, na_rm)
    return np.asarray(r), len(r.warnings) > 0


def api_colsum(x, group, ngroup, na_rm):
    r = sa.svt._groupsum("C_colsum_SVT", x, group, ngroup, na_rm)
    return np.asarray(r), len(r.warnings) > 0


def key_row_nd(name, op, na_rm, dims):
    return "stat|%s|rowd%d|%s|%d" % (name, dims, op, int(na_rm))


def port_row_nd(x, op, na_rm, dims):
    """row*(x, dims >= 2): dimensions 2..dims folded into the rows"""
    fold = int(np.prod(x.dim[1:dims], dtype=np.int64))
    nleaf = _nleaf(x)
    nrow = x.dim[0]
    if fold == 0 or nrow == 0:
        # zero-extent strata: nothing to compute
        shape = tuple(x.dim[:dims])
        return np.zeros(shape, dtype=np.float64), False
    leaf_of = np.repeat(np.arange(nleaf), np.diff(x.ptr))
    offs = x.offs + (nrow * (leaf_of % fold)).astype(np.int32)
    ptr = x.ptr[::fold]
    lac = None
    if x.lacunar is not None:
        # mixed lacunar leaves: materialise the ones (as the flattener does)
        vals = np.array(x.vals, copy=True)
        for l in np.flatnonzero(x.lacunar):
            vals[x.ptr[l]:x.ptr[l + 1]] = 1
    else:
        vals = x.vals
    v, w = port.rowstats(nrow * fold, nleaf // fold, ptr, offs, vals, x.type,
                         op, na_rm, None, lac)
    return v.reshape(tuple(x.dim[:dims]), order="F"), w


def api_row_nd(x, op, na_rm, dims):
    r = sa.svt._rowStats(op, x, na_rm=na_rm, center=None, dims=dims,
                         useNames=False)
    return np.asarray(r), len(r.warnings) > 0


def port_row(x, op, na_rm, center):
    return port.rowstats(x.dim[0], _nleaf(x), x.ptr, x.offs, x.vals, x.type,
                         op, na_rm, center, x.lacunar)


def port_crossprod(x, y, transpose_y, left):
    return port.crossprod(x.dim[0], x.dim[1], x.ptr, x.offs, x.vals, x.type,
                          y, transpose_y, left, x.lacunar)


def _dense_of(x):
    """as.matrix(x) with the payload type of x"""
    d = np.asarray(x.to_dense())
    return np.ascontiguousarray(
        d, dtype=np.float64 if x.type == "double" else np.int32)


def port_crossprod_svt(x, y):
    """crossprod(x, y), both sparse = crossprod(x, as.matrix(y)): what the
    reference's pre-processing of one operand amounts to (held against its
    own C_crossprod2_SVT_SVT / C_crossprod1_SVT outputs in test_golden.py).
    An all-zero x goes through the mirror formulation, as in the reference
    (crossprod2_mat0_SVT_*())."""
    if x is not y and (x.nnz == 0) != (y.nnz == 0):
        # one operand is a NULL SVT: the reference's "fictive matrix of
        # zeros" routines (src/SparseMatrix_mult.c:558-629,
        # _dotprod_doubles_zero() src/SparseVec_dotprod.c:116-127) --
        # a dense column of zeros, except that a leaf holding an NA gives NA
        # wherever the NA stands among its NaNs
        sp = y if x.nnz == 0 else x
        ans = port_crossprod(sp, np.zeros((sp.dim[0],
                                           (x if sp is y else y).dim[1]),
                                          dtype=np.float64
                                          if sp.type == "double"
                                          else np.int32), False, sp is x)
        if sp.type == "double" and sp.vals is not None:
            na = sa.is_na_real(sp.vals)
            for l in range(sp.dim[1]):
                if na[sp.ptr[l]:sp.ptr[l + 1]].any():
                    if sp is x:
                        ans[l, :] = sa.NA_REAL
                    else:
                        ans[:, l] = sa.NA_REAL
        return ans
    if x is not y:
        # C_crossprod2_SVT_SVT pre-processes the operand that costs fewer
        # operations (src/SparseMatrix_mult.c:1075-1098); which side is dense
        # decides NA-vs-NaN when a dot product meets both
        # (a NULL SVT is always the dense, all-zero side: :707-711,730-734)
        if y.nnz > 0 and (x.nnz == 0 or
                          y.nnz * x.dim[1] < x.nnz * y.dim[1]):
            return port_crossprod(y, _dense_of(x), False, False)
    ans = port_crossprod(x, _dense_of(y), False, True)
    if x is y:
        # crossprod(x): each pair (j < i) is computed once, with leaf j
        # pre-processed, and mirrored (compute_sym_dotprods_double(),
        # src/SparseMatrix_mult.c:826-852) = the lower triangle of
        # crossprod(x, as.matrix(x))
        il = np.tril_indices(ans.shape[0], -1)
        ans[il[1], il[0]] = ans[il]
    return ans


def api_crossprod_svt(x, y=None):
    return np.asarray(sa.crossprod(x, y))


def port_matmul(x, d):
    tp, to, tv = port.transpose(x.dim[0], x.dim[1], x.ptr, x.offs, x.vals,
                                x.type, x.lacunar)
    return port.crossprod(x.dim[1], x.dim[0], tp, to, tv, x.type, d, False,
                          True, None)


# ---- product API ---------------------------------------------------------

def api_col(x, op, na_rm, center, dims):
    r = sa.svt._colStats(op, x, na_rm=na_rm, center=center, dims=dims,
                         useNames=False)
    return np.asarray(r), len(r.warnings) > 0


def api_summarize(x, op, na_rm, center):
    r = sa.svt.summarize_SVT(op, x, na_rm=na_rm, center=center)
    return np.asarray(r).reshape(-1), len(r.warnings) > 0


def api_row(x, op, na_rm, center):
    r = sa.svt._rowStats(op, x, na_rm=na_rm, center=center, dims=1,
                         useNames=False)
    return np.asarray(r), len(r.warnings) > 0
