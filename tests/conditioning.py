"""Condition numbers for the 1e-12 bar.

north_star: "within 1e-12 relative for double sums and variances whose
summation order differs".  A sum of signed terms added in another order is
only defined to n * eps * sum|terms|: when the terms cancel, |result| is far
smaller than that and "relative to the result" would demand more digits than
either implementation has.  The tests therefore hold every double result to

    |cur - exp| <= 1e-12 * max(|exp|, cond)

with `cond` = the sum of the MAGNITUDES of the terms the reference adds:

    sum / mean          sum|x|                       (/ n for the mean)
    centered_X2_sum     sum x^2 + 2|c| sum|x| + c^2 n
                        (the reference's own  c^2 * n + sum x(x - 2c),
                        src/SparseArray_matrixStats.c:1052-1058, or
                        sum (x - c)^2 + c^2 * #zeros for columns)
    var1                the above / (n - 1);  sd1: d sqrt = d var / (2 sd)
    dot products        sum |x| |y|

computed here from the dense form of the (small) fixtures.  For sums of
non-negative terms cond == |exp| and the bound is the plain relative one.
"""
import numpy as np

NA_I = -2**31


def _abs_dense(x):
    d = x.to_dense()
    if d.dtype.kind != "f":
        na = d == NA_I
        d = d.astype(np.float64)
        d[na] = 0.0
    else:
        d = d.astype(np.float64)
        d[~np.isfinite(d)] = 0.0
    return np.abs(d)


def _reduce(a, axes):
    return a.sum(axis=axes) if axes else a


def moments(x, by, dims=1):
    """(sum|x|, sum x^2, n) over the reduced axes: by="col" reduces the first
    `dims` axes (colStats), by="row" the remaining ones (rowStats)."""
    a = _abs_dense(x)
    nd = a.ndim
    axes = tuple(range(dims)) if by == "col" else tuple(range(dims, nd))
    n = float(np.prod([a.shape[i] for i in axes])) if axes else 1.0
    return _reduce(a, axes).reshape(-1), _reduce(a * a, axes).reshape(-1), n


def cond(x, by, op, center=None, dims=1, exp=None):
    """Sum of term magnitudes behind op ("sum", "mean", "centered_X2_sum",
    "var1", "sd1") for every result element; None for exact ops."""
    s1, s2, n = moments(x, by, dims)
    if op == "sum":
        return s1
    if op == "mean":
        return s1 / max(n, 1.0)
    if op in ("centered_X2_sum", "var1", "sd1"):
        if center is None:
            c = s1 / max(n, 1.0)          # |mean| <= sum|x| / n
        else:
            c = np.abs(np.nan_to_num(np.asarray(center, dtype=np.float64)
                                     .reshape(-1)))
            if c.size == 1:
                c = np.full(s1.shape, c[0])
        x2 = s2 + 2.0 * c * s1 + c * c * n
        if op == "centered_X2_sum":
            return x2
        var = x2 / max(n - 1.0, 1.0)
        if op == "var1":
            return var
        sd = np.sqrt(np.abs(np.nan_to_num(np.asarray(exp, dtype=np.float64)
                                          .reshape(-1)))) \
            if exp is not None else np.sqrt(var)
        # d sqrt(v) = d v / (2 sqrt(v)); near v = 0 only sqrt(d v) holds
        with np.errstate(all="ignore"):
            return np.where(sd * sd > 1e-12 * var, var / (2.0 * sd),
                            np.sqrt(var) * 1e6)
    return None


def dot_cond(x, y, transpose_y=False, svt_left=True):
    """sum_i |x[i, l]| |y[i, k]| for crossprod(x, y) (l x k) or its
    transpose"""
    a = _abs_dense(x)
    y = np.asarray(y)
    if y.dtype.kind != "f":
        na = y == NA_I
        y = y.astype(np.float64)
        y[na] = 0.0
    b = np.abs(np.nan_to_num(y.astype(np.float64), nan=0.0, posinf=0.0,
                             neginf=0.0))
    if transpose_y:
        b = b.T
    c = a.T @ b
    return c if svt_left else c.T


def groupsum_cond(x, what, group, ngroup):
    """sum of |x| per (group, column) for rowsum() / per (row, group) for
    colsum(); `group` 1-based, NA_integer_ = the last group"""
    a = _abs_dense(x)
    g = np.asarray(group, dtype=np.int64).copy()
    g[g == NA_I] = ngroup
    g -= 1
    if what == "rowsum":
        out = np.zeros((ngroup, a.shape[1]))
        np.add.at(out, g, a)
    else:
        out = np.zeros((a.shape[0], ngroup))
        np.add.at(out.T, g, a.T)
    return out


def matmul_cond(x, d):
    """sum_l |x[i, l]| |d[l, k]| for x %*% d"""
    d = np.asarray(d)
    if d.dtype.kind != "f":
        na = d == NA_I
        d = d.astype(np.float64)
        d[na] = 0.0
    b = np.abs(np.nan_to_num(d.astype(np.float64), nan=0.0, posinf=0.0,
                             neginf=0.0))
    return _abs_dense(x) @ b
