"""Fixture matrices restated from the reference's own tests (inline R literals
there); citations are paths under Bioconductor/SparseArray.  NA_integer_ is
INT_MIN, NA_real_ the NaN with low word 1954, as in R."""
import numpy as np

NA_I = -2**31
NA_R = np.array([0x7FF00000000007A2], dtype=np.uint64).view(np.float64)[0]
NaN = float("nan")
Inf = float("inf")


def rset(m, idx, values):
    """R's `m[idx] <- values` with 1-based column-major linear indices."""
    flat = m.reshape(-1, order="F")
    flat[np.asarray(idx) - 1] = values
    return flat.reshape(m.shape, order="F")


def ms_m1():
    """tests/testthat/test-SparseArray-matrixStats.R:143-147 (int, NAs,
    a zero row, an all-zero column)."""
    m = np.array([[0, 0, NA_I, 0, NA_I],
                  [NA_I, 0, -3, 1, NA_I],
                  [0, 0, 0, 0, 0],
                  [15, 0, 0, 0, NA_I]], dtype=np.int32)
    return m, [list("ABCD"), list("abcde")]


def ms_m2_logical():
    """:195 `m2 <- is.na(m1)` -- all TRUE leaves => lacunar."""
    m, dn = ms_m1()
    return (m == NA_I).astype(np.int32), dn


def ms_m0_man():
    """man/SparseArray-matrixStats.Rd:189-191"""
    m = np.zeros((6, 4), dtype=np.int32)
    m[:, 0] = [10, 20, 0, 0, 0, 0]
    m[:, 1] = [0, 30, 0, 40, NA_I, 0]
    m[:, 2] = [0, 0, 50, 60, 70, 0]
    m[:, 3] = [0, 0, 0, 0, 0, 80]
    return m


def ms_a3d():
    """:244-252, 3-D double 6x5x4 with 1e12, 1e-10, NA, NaN."""
    a = np.zeros((6, 5, 4), dtype=np.float64)
    a[0, :, 1] = [1e12, -1234.55, -2.1, -1, -0.55]
    a[2, :, 1] = [-0.55, 0, 1e-10, 0.88, 1]
    a[4, :, 1] = [np.pi, 10.33, 3.4567895e8, 300, 2009.01]
    a[5, 2:4, 1] = [NA_R, NaN]
    return a


def ms_torture_2d():
    """:276-277"""
    m1 = np.array([[NA_I, -8, 0], [0, 0, 1]], dtype=np.int32)
    m2 = np.array([[0, NA_I, 0, 0], [8, 9, 1, 1], [-8, -9, -10, -11]],
                  dtype=np.int32)
    return [m1, m2]


def ms_torture_3d(double=False):
    """:293-297 (integer) and :325 (double: svt0[39:40] <- NaN)."""
    a = np.zeros((5, 4, 3), dtype=np.int32)
    a = rset(a, [1, 6, 16, 20, 21, 22, 36, 39, 40, 60],
             [2, -5, NA_I, NA_I, -11, 99, -8, NA_I, NA_I, NA_I])
    if not double:
        return a
    d = a.astype(np.float64)
    d[a == NA_I] = NA_R
    d = rset(d, [39, 40], [NaN, NaN])
    return d


def cp_double():
    """tests/testthat/test-SparseMatrix-mult.R:206-229: m0, m1, m2, m3."""
    m0 = np.zeros((5, 3))
    m0[2, 0] = Inf
    m0[1, 2] = -11.99
    m1 = np.array([[0, -4.5, 7, NA_R, 0, NaN, Inf, -Inf]])
    m2 = rset(np.zeros((6, 4)), [24, 1, 2, 8, 10, 15, 16, 17],
              np.arange(1, 9) - 3.5)
    m3 = np.zeros((6, 7))
    m3 = rset(m3, 3 + 4 * np.arange(10), 2.4 ** np.arange(1, 11))
    m3 = rset(m3, 4 + 4 * np.arange(10), -np.arange(101, 111).astype(float))
    m3[0, 4] = NaN
    m3[4, 2] = Inf
    return m0, m1, m2, m3


def cp_int(with_na=False):
    """:283-294: m2 6x4, m3 6x7 (+ NAs at m2[2,4], m3[1,5])."""
    m2 = rset(np.zeros((6, 4), dtype=np.int32),
              [24, 1, 2, 8, 10, 15, 16, 17], np.arange(1, 9) * 10 - 35)
    m3 = np.zeros((6, 7), dtype=np.int32)
    m3 = rset(m3, 3 + 4 * np.arange(10), np.arange(1, 11))
    m3 = rset(m3, 4 + 4 * np.arange(10), -np.arange(101, 111))
    if with_na:
        m2[1, 3] = NA_I
        m3[0, 4] = NA_I
    return m2, m3


def cp_int_m1():
    """:280"""
    return np.array([[0, -4, 7, NA_I, 0, NA_I]], dtype=np.int32)


def mm_m1():
    """:307-308, `%*%` test."""
    return rset(np.zeros((15, 6), dtype=np.int32),
                [2, 6] + list(range(12, 18)) + list(range(22, 34)) +
                [55] + list(range(59, 63)) + [90], np.arange(101, 127))


# Known answers (base R results quoted in SURVEY.md appendix B)
B1_EXPECTED = {
    ("colSums", False): [NA_R, 0, NA_R, 1, NA_R],
    ("colSums", True): [15, 0, -3, 1, 0],
    ("rowSums", False): [NA_R, NA_R, 0, NA_R],
    ("rowSums", True): [0, -2, 0, 15],
    ("colMeans", False): [NA_R, 0, NA_R, 0.25, NA_R],
    ("colMeans", True): [5, 0, -1, 0.25, 0],
    ("rowMeans", False): [NA_R, NA_R, 0, NA_R],
    ("rowMeans", True): [0, -2 / 3, 0, 3.75],
    ("colVars", False): [NA_R, 0, NA_R, 0.25, NA_R],
    ("colVars", True): [75, 0, 3, 0.25, NA_R],
    ("rowVars", False): [NA_R, NA_R, 0, NA_R],
    ("rowVars", True): [0, 13 / 3, 0, 56.25],
    ("colMaxs", False): [NA_I, 0, NA_I, 1, NA_I],
    ("colMaxs", True): [15, 0, 0, 1, 0],
    ("colMins", False): [NA_I, 0, NA_I, 0, NA_I],
    ("colMins", True): [0, 0, -3, 0, 0],
    ("rowMaxs", False): [NA_I, NA_I, 0, NA_I],
    ("rowMaxs", True): [0, 1, 0, 15],
}

B2_EXPECTED = {
    ("colSums", False): [30, NA_R, 180, 80],
    ("colSums", True): [30, 70, 180, 80],
    ("rowSums", False): [10, 50, 50, 100, NA_R, 80],
    ("rowSums", True): [10, 50, 50, 100, 70, 80],
    ("colVars", False): [70, NA_R, 1120, 3200 / 3],
}

B3_CROSSPROD_M2 = [[250, -25, 0, 0], [-25, 250, 525, 0],
                   [0, 525, 3875, 0], [0, 0, 0, 625]]
B3_CROSSPROD_M2_M3 = [[0, 480, 0, 450, 0, 420, 0],
                      [-1515, -510, -1560, -525, -1605, -540, -1650],
                      [-3510, 135, -3540, 270, -3570, 405, -3600],
                      [0, 2575, 0, 2650, 0, 2725, 0]]
