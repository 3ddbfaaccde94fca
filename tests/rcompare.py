"""Comparison helpers with R's notion of identity: NA_real_ and NaN are
different values; integers/logicals compare exactly."""
import numpy as np

from rshim.rshim import is_na_real


def classes(a):
    """0 regular, 1 NA, 2 NaN per element of a double array."""
    a = np.asarray(a, dtype=np.float64)
    c = np.zeros(a.shape, dtype=np.int8)
    na = is_na_real(a).reshape(a.shape)
    c[na] = 1
    c[np.isnan(a) & ~na] = 2
    return c


def assert_identical(cur, exp, what=""):
    """expect_identical(): same type class, shape, NA/NaN pattern, bits."""
    cur, exp = np.asarray(cur), np.asarray(exp)
    assert cur.shape == exp.shape, "%s: shape %s != %s" % (what, cur.shape,
                                                            exp.shape)
    if exp.dtype.kind == "f" or cur.dtype.kind == "f":
        assert cur.dtype.kind == "f" and exp.dtype.kind == "f", \
            "%s: type mismatch %s vs %s" % (what, cur.dtype, exp.dtype)
        cc, ce = classes(cur), classes(exp)
        assert np.array_equal(cc, ce), \
            "%s: NA/NaN pattern differs\n cur=%r\n exp=%r" % (what, cur, exp)
        reg = ce == 0
        assert np.array_equal(cur[reg], exp[reg]), \
            "%s: values differ\n cur=%r\n exp=%r" % (what, cur, exp)
    else:
        assert cur.dtype.kind == exp.dtype.kind, \
            "%s: type mismatch %s vs %s" % (what, cur.dtype, exp.dtype)
        assert np.array_equal(cur, exp), \
            "%s: values differ\n cur=%r\n exp=%r" % (what, cur, exp)


def assert_close(cur, exp, rtol=1e-12, atol=0.0, what="", na_nan_strict=True,
                 cond=None):
    """`cond` (optional, same shape): sum of the magnitudes of the terms
    behind each result (tests/conditioning.py); the bound becomes
    rtol * max(|exp|, cond).
    Same NA/NaN/Inf pattern; regular values within rtol (the tolerance the
    north star states for double sums/variances whose summation order
    differs)."""
    cur = np.asarray(cur, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    assert cur.shape == exp.shape, "%s: shape %s != %s" % (what, cur.shape,
                                                            exp.shape)
    cc, ce = classes(cur), classes(exp)
    if not na_nan_strict:
        cc, ce = np.minimum(cc, 1), np.minimum(ce, 1)
    assert np.array_equal(cc, ce), \
        "%s: NA/NaN pattern differs at %r" % (what,
                                              np.flatnonzero(cc != ce)[:10])
    reg = ce == 0
    a, b = cur[reg], exp[reg]
    inf = np.isinf(b)
    assert np.array_equal(a[inf], b[inf]), "%s: infinities differ" % what
    a, b = a[~inf], b[~inf]
    err = np.abs(a - b)
    scale = np.abs(b)
    if cond is not None:
        c = np.broadcast_to(np.asarray(cond, dtype=np.float64).reshape(
            exp.shape) if np.size(cond) == exp.size else
            np.asarray(cond, dtype=np.float64), exp.shape)[reg][~inf]
        scale = np.maximum(scale, np.nan_to_num(c))
    tol = atol + rtol * scale
    bad = err > tol
    assert not bad.any(), "%s: max rel err %.3e (tol %.1e) at %r" % (
        what, float((err / np.maximum(np.abs(b), 1e-300)).max()), rtol,
        np.flatnonzero(bad)[:10])
