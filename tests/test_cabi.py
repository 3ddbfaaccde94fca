"""The drop-in boundary loads without a GPU and exports what it declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from sparsearray_b200 import _native, rcall, build
import sparsearray_b200 as sa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svtgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svtgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    names = _declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(build.LIBSVTGPU)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_binding_covers_header():
    assert sorted(_native.SIGNATURES) == _declared_symbols()
    _native.lib()


def test_glue_registers_reference_entry_points():
    """names/arities of src/R_init_SparseArray.c:41-43,121-122,131-132"""
    r = rcall.registered_routines()
    assert r["C_colStats_SVT"] == 9
    assert r["C_rowStats_SVT"] == 9
    assert r["C_crossprod2_SVT_mat"] == 7
    assert r["C_crossprod2_mat_SVT"] == 7
    assert r["C_crossprod2_SVT_SVT"] == 8  # src/R_init_SparseArray.c:133-134
    assert r["C_crossprod1_SVT"] == 5
    assert r["C_get_num_procs"] == 0
    assert r["C_get_max_threads"] == 0
    assert r["C_set_max_threads"] == 1
    # first widening: src/R_init_SparseArray.c:94
    assert r["C_summarize_SVT"] == 7
    assert r["C_rowsum_SVT"] == 6          # src/R_init_SparseArray.c:125,127
    assert r["C_colsum_SVT"] == 6
    # extensions of the GPU path
    assert r["C_svtgpu_resident_SVT"] == 3
    assert r["C_svtgpu_release"] == 1


def test_unregistered_routine_is_an_error():
    with pytest.raises(rcall.rshim.RError):
        rcall.dot_call("C_not_there", [])


@pytest.mark.skipif(_native.device_count() > 0, reason="needs a GPU-less box")
def test_no_cpu_fallback():
    """Without a device the product path fails loudly -- even with the
    reference build (same symbol names) loaded in the same process."""
    from oracle import refcall
    if refcall.available():
        refcall.ref()
    x = sa.SVT_SparseArray.from_dense(np.eye(3, dtype=np.int32))
    with pytest.raises(rcall.rshim.RError, match="no usable CUDA device"):
        sa.colSums(x)
    with pytest.raises(rcall.rshim.RError, match="no usable CUDA device"):
        sa.svt_sum(x)
    with pytest.raises(rcall.rshim.RError, match="no usable CUDA device"):
        sa.to_device(x)
    with pytest.raises(rcall.rshim.RError, match="no usable CUDA device"):
        sa.rowSums(x)
    with pytest.raises(rcall.rshim.RError, match="no usable CUDA device"):
        sa.crossprod(x, np.ones((3, 2), dtype=np.int32))
    h = ctypes.c_void_p()
    rc = _native.lib().svtgpu_matrix_create(ctypes.byref(h), 3, 3, 3,
                                            _native.INT, 3)
    assert rc == _native.ERR_NO_DEVICE
