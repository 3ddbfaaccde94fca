"""The composition rules of svt_semantics.h (shared by every CUDA kernel),
compiled for the host and checked against the outputs of the reference's C
(golden.npz): partials are formed here with numpy exactly the way the kernels
form them (counts of NA / NaN, sums over regular values, min/max over regular
values), then finalised by the C functions under test."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases
import fixtures as fx
import runners
from rcompare import assert_identical, assert_close
from rshim.rshim import is_na_real
from oracle.port import OPCODES

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "semantics_host.c")
LIB = os.path.join(HERE, "libsemantics_host.so")
STAT = cases.stat_cases()


@pytest.fixture(scope="module")
def sem():
    hdr = os.path.join(HERE, "..", "sparsearray_b200", "csrc",
                       "svt_semantics.h")
    if (not os.path.exists(LIB) or
            os.path.getmtime(LIB) < max(os.path.getmtime(SRC),
                                        os.path.getmtime(hdr))):
        subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared",
                               "-o", LIB, SRC, "-lm"])
    L = ctypes.CDLL(LIB)
    D, I, I64 = ctypes.c_double, ctypes.c_int, ctypes.c_int64
    P = ctypes.c_void_p
    L.sem_col_finalize.argtypes = [I, I, I, I64, D, P, P, P, P]
    L.sem_col_mean.argtypes = [I, I, I64, P]
    L.sem_col_mean.restype = D
    L.sem_row_finalize.argtypes = [I, I, I, I64, I, D, P, P, P, P]
    L.sem_row_moments.argtypes = [I, I64, P, P, P]   # states: 6 doubles
    L.sem_dot_finalize.argtypes = [I, D, I, I64, I, I]
    L.sem_dot_finalize.restype = D
    return L


def _classify(x, v):
    """(regular mask, na mask, nan mask) for a leaf's stored values."""
    if x.type == "double":
        nan = np.isnan(v)
        na = is_na_real(v)
        return ~nan, na, nan & ~na
    na = v == fx.NA_I
    return ~na, na, np.zeros_like(na)


def _segment_values(x, seg, group):
    """Stored values of the `group` leaves of one output cell, lacunar leaves
    materialised as ones (what the flattener uploads for a mixed SVT)."""
    out = []
    for l in range(seg * group, (seg + 1) * group):
        a, b = int(x.ptr[l]), int(x.ptr[l + 1])
        lac = x.vals is None or (x.lacunar is not None and x.lacunar[l])
        if lac:
            out.append(np.ones(b - a, dtype=np.float64 if x.type == "double"
                               else np.int32))
        else:
            out.append(x.vals[a:b])
    if not out:
        return np.zeros(0, dtype=np.float64 if x.type == "double"
                        else np.int32)
    return np.concatenate(out)


def _col_via_semantics(sem, x, op, na_rm, center, dims):
    group = int(np.prod(x.dim[1:dims], dtype=np.int64))
    nout = int(np.prod(x.dim[dims:], dtype=np.int64))
    in_length = x.dim[0] * group
    code = OPCODES[op]
    is_double = x.type == "double"
    out_is_int = sem.sem_col_out_is_int(code, 14 if is_double else 13)
    out = np.zeros(nout, dtype=np.int32 if out_is_int else np.float64)
    warn = False
    for s in range(nout):
        v = _segment_values(x, s, group) if group > 0 else \
            np.zeros(0, np.int32)
        reg, na, nan = _classify(x, v)
        r = v[reg].astype(np.float64)
        part = np.zeros(9)
        part[0], part[1], part[2] = v.size, na.sum(), nan.sum()
        part[3] = (r == 0).sum()
        with np.errstate(all="ignore"):
            part[4] = r.sum()
            part[6] = r.prod()
        part[7] = r.min() if r.size else np.inf
        part[8] = r.max() if r.size else -np.inf
        c = fx.NA_R if center is None else center
        if op in ("centered_X2_sum", "var1", "sd1"):
            if np.isnan(c):
                c = sem.sem_col_mean(int(is_double), int(na_rm), in_length,
                                     part.ctypes.data)
            with np.errstate(all="ignore"):
                part[5] = ((r - c) ** 2).sum()
        od, oi, w = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int()
        sem.sem_col_finalize(code, int(is_double), int(na_rm), in_length, c,
                             part.ctypes.data, ctypes.byref(od),
                             ctypes.byref(oi), ctypes.byref(w))
        out[s] = oi.value if out_is_int else od.value
        warn = warn or bool(w.value)
    shape = tuple(x.dim[dims:])
    if len(shape) >= 2:
        out = out.reshape(shape, order="F")
    return out, warn


@pytest.mark.parametrize("name", sorted(STAT))
def test_col_semantics_vs_reference(sem, name):
    G = runners.golden()
    x = STAT[name]
    for op, na_rm, center, dims in cases.col_requests(x):
        k = runners.key_col(name, op, na_rm, center, dims)
        if k + "|error" in G:
            continue
        v, w = _col_via_semantics(sem, x, op, na_rm, center, dims)
        exp = G[k]
        if exp.dtype.kind != "f" or x.type != "double" and \
                op in ("sum", "countNAs"):
            assert_identical(v, exp, k)
        else:
            # numpy's pairwise sums vs the reference's sequential ones
            assert_close(v, exp, rtol=1e-12, what=k)
        assert bool(G[k + "|warn"]) == w, k


def _row_state(x, want_minmax, is_min):
    """Per-row state as the kernels build it (svt_semantics.h): slots
    {sum | coverage, #NA, #NaN, sum2 | extreme} and, for the sums, the last
    leaf that put an NA / a NaN into the row."""
    nrow = x.dim[0]
    nleaf = x.ptr.size - 1
    state = np.zeros((6, nrow))
    state[3, :] = (np.inf if is_min else -np.inf) if want_minmax else 0.0
    state[4:, :] = -np.inf
    for l in range(nleaf):
        a, b = int(x.ptr[l]), int(x.ptr[l + 1])
        if a == b:
            continue
        lac = x.vals is None or (x.lacunar is not None and x.lacunar[l])
        offs = x.offs[a:b]
        v = np.ones(b - a) if lac else x.vals[a:b]
        if lac:
            reg = np.ones(b - a, bool)
            na = nan = np.zeros(b - a, bool)
        else:
            reg, na, nan = _classify(x, v)
        vv = v.astype(np.float64)
        state[1, offs[na]] += 1
        state[2, offs[nan]] += 1
        if want_minmax:
            state[0, offs] += 1
            cur = state[3, offs[reg]]
            state[3, offs[reg]] = np.minimum(cur, vv[reg]) if is_min \
                else np.maximum(cur, vv[reg])
        else:
            with np.errstate(all="ignore"):
                state[0, offs[reg]] += vv[reg]
                state[3, offs[reg]] += vv[reg] ** 2
            for slot, m in ((4, na), (5, nan)):
                state[slot, offs[m]] = np.maximum(state[slot, offs[m]], l)
    return state


def _rows_with_na_and_inf(x):
    """Rows whose centered_X2_sum the reference decides by floating-point
    accident: an NA or NaN together with an infinity (Inf * (Inf - 2c) terms
    and Inf - Inf inside the running value).  Plain row sums of such rows are
    exact (the first-leaf slots order NA / NaN / Inf events)."""
    if x.type != "double" or x.vals is None:
        return np.zeros(x.dim[0], bool)
    st = _row_state(x, False, False)
    special = (st[1] > 0) | (st[2] > 0)
    d = x.to_dense()
    infs = np.isinf(d).any(axis=tuple(range(1, d.ndim)))
    return special & infs


@pytest.mark.parametrize("name", sorted(STAT))
def test_row_semantics_vs_reference(sem, name):
    G = runners.golden()
    x = STAT[name]
    if len(x.dim) < 2:
        return
    nrow = x.dim[0]
    nstrata = x.ptr.size - 1
    is_double = x.type == "double"
    mix = _rows_with_na_and_inf(x)
    for op, na_rm, kind in cases.row_requests(x):
        k = runners.key_row(name, op, na_rm, kind)
        exp = G[k].reshape(-1)
        code = OPCODES[op]
        st = _row_state(x, op in ("min", "max"), op == "min")
        center = cases.row_center(x, kind)
        out_is_int = exp.dtype.kind != "f"
        out = np.zeros(nrow, dtype=np.int32 if out_is_int else np.float64)
        warn = False
        for i in range(nrow):
            s4 = np.ascontiguousarray(st[:, i])
            od, oi, w = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int()
            sem.sem_row_finalize(code, int(is_double), int(na_rm), nstrata,
                                 int(center is not None),
                                 0.0 if center is None else float(center[i]),
                                 s4.ctypes.data, ctypes.byref(od),
                                 ctypes.byref(oi), ctypes.byref(w))
            out[i] = oi.value if out_is_int else od.value
            warn = warn or bool(w.value)
        if out_is_int or (not is_double and op in ("sum", "countNAs")):
            assert_identical(out, exp, k)
        else:
            keep = ~mix if (op == "centered_X2_sum" and not na_rm) \
                else np.ones(nrow, bool)
            assert_close(out[keep], exp[keep], rtol=1e-12, atol=1e-9
                         if op == "centered_X2_sum" else 0.0, what=k)
        assert bool(G[k + "|warn"]) == warn, k


@pytest.mark.parametrize("name", sorted(STAT))
def test_row_moments_vs_reference_composition(sem, name):
    """One-pass {sum, sum2, nNA} -> rowMeans/rowVars == the reference's
    three-pass R composition, to 1e-12 relative (of the data scale)."""
    G = runners.golden()
    x = STAT[name]
    if len(x.dim) != 2 or x.dim[0] == 0:
        return
    nstrata = x.dim[1]
    st = _row_state(x, False, False)
    mix = _rows_with_na_and_inf(x)
    for na_rm in (False, True):
        mean = np.zeros(x.dim[0])
        var = np.zeros(x.dim[0])
        for i in range(x.dim[0]):
            s4 = np.ascontiguousarray(st[:, i])
            m, v = ctypes.c_double(), ctypes.c_double()
            sem.sem_row_moments(int(na_rm), nstrata, s4.ctypes.data,
                                ctypes.byref(m), ctypes.byref(v))
            mean[i], var[i] = m.value, v.value
        keep = np.ones(x.dim[0], bool)
        em = G["stat|%s|rowMeans|%d" % (name, na_rm)].reshape(-1)
        ev = G["stat|%s|rowVars|%d" % (name, na_rm)].reshape(-1)
        assert_close(mean[keep], em[keep], rtol=1e-12, what=name + " mean")
        finite = np.isfinite(ev) & keep
        scale = np.abs(st[3, :]).max() if finite.any() else 1.0
        assert_close(var[finite], ev[finite], rtol=1e-10,
                     atol=1e-12 * max(scale, 1.0), what=name + " var")


def test_dot_finalize_rules(sem):
    NA, NaN = fx.NA_R, fx.NaN
    f = sem.sem_dot_finalize
    assert f(0, 12.0, 0, 0, 0, 0) == 12.0
    assert is_na_real(f(0, 12.0, 1, 0, 0, 0))       # int leaf holds NA
    assert is_na_real(f(0, 12.0, 0, 0, 1, 1))       # int dense col holds NA
    assert f(1, 3.5, 0, 0, 0, 0) == 3.5
    assert is_na_real(f(1, NaN, 1, 0, 0, 0))        # NA in leaf, finite col
    r = f(1, NaN, 0, 0, 0, 0)
    assert np.isnan(r) and not is_na_real(r)
    assert is_na_real(f(1, 1.0, 0, 0, 2, 1))        # NA in dense col
    r = f(1, 1.0, 0, 1, 2, 0)                       # a 0 * Inf somewhere
    assert np.isnan(r) and not is_na_real(r)
    assert f(1, np.inf, 0, 2, 2, 0) == np.inf       # every Inf hit by a nz
