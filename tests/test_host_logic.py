"""Host-side logic: SVT construction, argument checks of the R-method mirror
(raised before any device work), synthetic generators, shard planning."""
import numpy as np
import pytest

import fixtures as fx
import sparsearray_b200 as sa
from sparsearray_b200 import synth
from sparsearray_b200.device import plan_column_shards
from sparsearray_b200.rcall import rshim


def test_from_dense_roundtrip_and_lacunar():
    m1, dn = fx.ms_m1()
    x = sa.SVT_SparseArray.from_dense(m1, "integer", dimnames=dn)
    assert x.dim == (4, 5) and x.nnz == 8
    assert np.array_equal(x.to_dense(), m1)
    # column 4 is a single 1 -> lacunar leaf; NAs count as nonzero
    assert x.lacunar is not None and x.lacunar[3] == 1
    m2, _ = fx.ms_m2_logical()
    y = sa.SVT_SparseArray.from_dense(m2, "logical")
    assert y.vals is None          # every leaf lacunar
    assert np.array_equal(y.to_dense(), m2)
    a = fx.ms_a3d()
    z = sa.SVT_SparseArray.from_dense(a)
    assert z.type == "double" and z.ptr.size == 21
    d = z.to_dense()
    assert np.array_equal(np.isnan(d), np.isnan(a))
    assert np.array_equal(d[~np.isnan(a)], a[~np.isnan(a)])


def test_with_type_keeps_na():
    m1, _ = fx.ms_m1()
    x = sa.SVT_SparseArray.from_dense(m1, "integer", lacunar=False)
    d = x.with_type("double")
    assert d.type == "double"
    assert rshim.is_na_real(d.vals).sum() == (m1 == fx.NA_I).sum()


def test_nested_svt_layout_matches_reference_walk():
    """3-D SVT built as nested lists is what REC_colStats_SVT() walks:
    checked by running the reference on it (when available)."""
    from oracle import refcall
    if not refcall.available():
        pytest.skip("oracle/_ref not built")
    a = fx.ms_torture_3d(False)
    x = sa.SVT_SparseArray.from_dense(a, "integer")
    r = refcall.colStats(x, "sum", na_rm=True, dims=1)
    exp = np.where(a == fx.NA_I, 0, a).sum(axis=0).astype(np.float64)
    assert np.array_equal(np.asarray(r.value), exp)
    r = refcall.colStats(x, "sum", na_rm=True, dims=2)
    assert np.array_equal(np.asarray(r.value), exp.sum(axis=0))


@pytest.mark.parametrize("call,msg", [
    (lambda x: sa.colSums(x, dims=0), "'dims' must be a single integer"),
    (lambda x: sa.colSums(x, dims=3), "'dims' must be a single integer"),
    (lambda x: sa.rowSums(x, dims=2), "'dims' must be a single integer"),
    (lambda x: sa.colSums(x, na_rm=None), "'na.rm' must be TRUE or FALSE"),
    (lambda x: sa.colVars(x, center=[1, 2]), "'center' must be NULL or a"),
    (lambda x: sa.rowVars(x, center=np.ones(3)), "unexpected 'center' len"),
    (lambda x: sa.crossprod(x, np.ones((3, 2))), "non-conformable"),
    (lambda x: sa.matmul(x, np.ones((3, 2))), "non-conformable"),
])
def test_argument_errors_before_device(call, msg):
    x = sa.SVT_SparseArray.from_dense(np.eye(4, 5, dtype=np.int32))
    with pytest.raises(ValueError, match=msg):
        call(x)


def test_crossprod_type_check():
    x = sa.SVT_SparseArray.from_dense(np.eye(4, dtype=np.int32), "logical")
    with pytest.raises(ValueError, match="must be of type"):
        sa.crossprod(x, np.ones((4, 2), dtype=bool))


def test_non_native_row_op_needs_matrix():
    """Non-native row ops go through the device transpose: matrices only."""
    a = np.zeros((3, 4, 2), dtype=np.int32)
    a[1, 2, 1] = 5
    x = sa.SVT_SparseArray.from_dense(a)
    with pytest.raises(NotImplementedError):
        sa.svt._rowStats("prod", x)


def test_poisson_generator_distribution():
    x = synth.poisson_svt(2000, 50, 0.07, seed=5)
    dens = x.nnz / (2000 * 50)
    assert abs(dens - 0.07) < 0.004
    # zero-truncated Poisson(lambda = -log(1 - d)): P(1) = lam e^-lam / d
    lam = -np.log1p(-0.07)
    p1 = lam * np.exp(-lam) / 0.07
    assert abs((x.vals == 1).mean() - p1) < 0.01
    assert x.vals.min() >= 1
    # strictly ascending offsets inside every leaf
    for l in range(50):
        o = x.offs[x.ptr[l]:x.ptr[l + 1]]
        assert np.all(np.diff(o) > 0)
    # counter-based: a column shard regenerates identically
    p, o, v = synth.poisson_csc(2000, 10, 0.07, seed=5, leaf0=20)
    a, b = x.ptr[20], x.ptr[30]
    assert np.array_equal(o, x.offs[a:b]) and np.array_equal(v, x.vals[a:b])


def test_poisson_na_injection():
    x = synth.poisson_svt(1000, 40, 0.1, seed=7, na_rate=0.05)
    frac = (x.vals == sa.NA_INTEGER).mean()
    assert 0.03 < frac < 0.07


def test_random_generator_exact_count():
    x = synth.random_svt(2000, 50, 0.05, seed=1)
    assert x.nnz == 5000 and x.type == "double"
    assert np.all(x.vals != 0)


def test_plan_column_shards():
    assert plan_column_shards(10, 1) == [(0, 10)]
    assert plan_column_shards(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    ptr = np.array([0, 100, 100, 100, 110, 120, 200], dtype=np.int64)
    shards = plan_column_shards(6, 2, ptr)
    assert shards[0][0] == 0 and shards[-1][1] == 6
    assert shards[0][1] == shards[1][0]
    # balanced by nonzeros: the cut lands right after the heavy first leaf
    assert shards[0] == (0, 1)
    for w in (3, 8):
        s = plan_column_shards(6, w, ptr)
        assert len(s) == w and s[0][0] == 0 and s[-1][1] == 6
        assert all(a <= b for a, b in s)
        assert all(s[i][1] == s[i + 1][0] for i in range(w - 1))


def test_invalid_leaf_is_an_r_error_before_any_device_work():
    """Leaf validation (unzip_leaf()/toSparseVec() in the reference) runs in
    the parallel indexing pass and is reported from the main thread."""
    from sparsearray_b200 import rcall
    x = sa.SVT_SparseArray.from_dense(
        np.arange(1, 13, dtype=np.int32).reshape(3, 4), "integer",
        lacunar=False)
    args = [x.r_dim, None, rshim.string("double"), x.r_SVT,
            rshim.logical([0]), rshim.string("sum"), rshim.logical([0]),
            rshim.real([sa.NA_REAL]), rshim.integer([1])]
    with pytest.raises(rshim.RError, match=r"TYPEOF\(nzvals\) != Rtype"):
        rcall.SparseArray_Call("C_colStats_SVT", *args)
