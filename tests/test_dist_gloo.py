"""The N>1 composition on CPU: world_size-2 gloo run of the column-shard plan
and of the row-state allreduce (the same combine_row_state() the GPU ranks run
over NCCL), finalised with the host build of svt_semantics.h and compared with
the reference's outputs on the whole matrix."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_name, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import test_semantics_host as tsh
    from sparsearray_b200.device import plan_column_shards, combine_row_state
    from sparsearray_b200.svt import SVT_SparseArray
    x = cases.stat_cases()[case_name]
    nrow, ncol = x.dim
    l0, l1 = plan_column_shards(ncol, world, x.ptr)[rank]
    e0, e1 = int(x.ptr[l0]), int(x.ptr[l1])
    shard = SVT_SparseArray(
        (nrow, l1 - l0), x.type, x.ptr[l0:l1 + 1] - e0, x.offs[e0:e1],
        None if x.vals is None else x.vals[e0:e1],
        None if x.lacunar is None else x.lacunar[l0:l1])
    res = {}
    for op, is_min in (("sum", False), ("min", True), ("max", False)):
        mm = op in ("min", "max")
        st = tsh._row_state(shard, mm, is_min)
        t = torch.from_numpy(np.ascontiguousarray(st.reshape(-1)))
        # MIN / MAX: 3 summed slots + the extreme; sums: 4 summed slots + 2
        # MAX-combined "last leaf" slots (leaf indices made global first)
        if mm:
            combine_row_state(t[:4 * nrow], nrow, 3, 1, is_min,
                              dist.group.WORLD)
        else:
            t.view(6, nrow)[4:] += l0
            combine_row_state(t, nrow, 4, 2, False, dist.group.WORLD)
        res[op] = t.numpy().reshape(6, nrow).copy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "states.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["rand_int_na", "rand_dbl_clean",
                                       "rand_lacunar_int"])
def test_row_state_allreduce_world2(tmp_path, case_name):
    import cases
    import runners
    import test_semantics_host as tsh
    from rcompare import assert_identical, assert_close
    from oracle.port import OPCODES
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case_name, str(tmp_path)), nprocs=2,
             join=True)
    states = np.load(os.path.join(str(tmp_path), "states.npz"))
    x = cases.stat_cases()[case_name]
    nrow, ncol = x.dim
    # the combined state equals the state of the whole matrix
    whole = tsh._row_state(x, False, False)
    assert np.allclose(states["sum"][:4], whole[:4], rtol=1e-13, atol=0)
    assert np.array_equal(states["sum"][4:], whole[4:])
    # and finalises to the reference's answers
    sem = tsh.sem.__wrapped__() if hasattr(tsh.sem, "__wrapped__") else None
    if sem is None:
        import subprocess
        if not os.path.exists(tsh.LIB):
            subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fPIC",
                                   "-shared", "-o", tsh.LIB, tsh.SRC, "-lm"])
        sem = ctypes.CDLL(tsh.LIB)
        D, I, I64, P = (ctypes.c_double, ctypes.c_int, ctypes.c_int64,
                        ctypes.c_void_p)
        sem.sem_row_finalize.argtypes = [I, I, I, I64, I, D, P, P, P, P]
    G = runners.golden()
    is_double = x.type == "double"
    for op in ("sum", "min", "max"):
        for na_rm in (False, True):
            exp = G[runners.key_row(case_name, op, na_rm, None)].reshape(-1)
            out_is_int = exp.dtype.kind != "f"
            out = np.zeros(nrow, dtype=np.int32 if out_is_int else np.float64)
            for i in range(nrow):
                s4 = np.ascontiguousarray(states[op][:, i])
                od, oi, w = (ctypes.c_double(), ctypes.c_int32(),
                             ctypes.c_int())
                sem.sem_row_finalize(OPCODES[op], int(is_double), int(na_rm),
                                     ncol, 0, 0.0, s4.ctypes.data,
                                     ctypes.byref(od), ctypes.byref(oi),
                                     ctypes.byref(w))
                out[i] = oi.value if out_is_int else od.value
            if out_is_int or not is_double:
                assert_identical(out, exp, "%s %s" % (op, na_rm))
            else:
                assert_close(out, exp, rtol=1e-12, what="%s %s" % (op, na_rm))


# ---- whole-array summaries and colsum across two column shards -------------

def _summary_worker(rank, world, port, case_name, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import runners
    from sparsearray_b200 import sharded
    from sparsearray_b200.device import plan_column_shards
    from sparsearray_b200.svt import SVT_SparseArray, RArray

    # the ranks have no GPU here: the per-shard `.Call`s are answered by the
    # oracle port, so what is tested is the cross-shard composition
    class Backend:
        NA_REAL = sharded.S.NA_REAL
        NA_INTEGER = sharded.S.NA_INTEGER
        is_na_real = staticmethod(sharded.S.is_na_real)

        @staticmethod
        def summarize_SVT(op, x, na_rm=False, center=None):
            v, w = runners.port_summarize(x, op, na_rm, center)
            return RArray(v, warnings=["w"] if w else [])

        @staticmethod
        def _groupsum(name, x, group, ngroup, na_rm):
            f = runners.port_rowsum if "rowsum" in name else \
                runners.port_colsum
            return RArray(f(x, group, ngroup, na_rm)[0])
    sharded.S = Backend

    if case_name in cases.groupsum_cases():
        x, rg, nrg, cg, ncg = cases.groupsum_cases()[case_name]
    else:
        x, cg, ncg = cases.stat_cases()[case_name], None, 0
    nrow, ncol = x.dim
    l0, l1 = plan_column_shards(ncol, world, x.ptr)[rank]
    e0, e1 = int(x.ptr[l0]), int(x.ptr[l1])
    shard = SVT_SparseArray(
        (nrow, l1 - l0), x.type, x.ptr[l0:l1 + 1] - e0, x.offs[e0:e1],
        None if x.vals is None else x.vals[e0:e1],
        None if x.lacunar is None else x.lacunar[l0:l1])
    g = dist.group.WORLD
    res = {}
    for na_rm in (False, True):
        for name, fn in (("sum", sharded.svt_sum), ("mean", sharded.mean),
                         ("var1", sharded.var), ("sd1", sharded.sd),
                         ("min", sharded.svt_min), ("max", sharded.svt_max)):
            res["%s|%d" % (name, na_rm)] = np.array([fn(shard, na_rm, g)])
        if cg is not None:
            res["colsum|%d" % na_rm] = np.asarray(
                sharded.colsum(shard, cg[l0:l1], ncg, na_rm, g))
    res["countNAs"] = np.array([sharded.countNAs(shard, g)])
    if rank == 0:
        np.savez(os.path.join(out_dir, "summary.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["rand_int_na_g3", "rand_dbl_special_g3",
                                       "rand_lacunar_int_g3",
                                       "poisson_small_g40", "rand_dbl_clean"])
def test_summaries_and_colsum_world2(tmp_path, case_name):
    import runners
    from rcompare import assert_close
    port = _free_port()
    mp.spawn(_summary_worker, args=(2, port, case_name, str(tmp_path)),
             nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "summary.npz"))
    G = runners.golden()
    base = case_name.rsplit("_g", 1)[0] if "_g" in case_name else case_name

    def as_double(a):
        a = np.asarray(a)
        if a.dtype.kind in "ib":
            d = a.astype(np.float64)
            d[a == -2147483648] = runners.port.NA_REAL
            return d
        return a

    for na_rm in (0, 1):
        for op in ("sum", "mean", "var1", "sd1", "min", "max"):
            exp = as_double(G[runners.key_summ(base, op, na_rm, None)])
            assert_close(got["%s|%d" % (op, na_rm)].reshape(-1),
                         exp.reshape(-1), rtol=1e-12, atol=1e-9,
                         what="%s %s na_rm=%d" % (case_name, op, na_rm))
        k = "gs|%s|colsum|%d" % (case_name, na_rm)
        if k in G:
            assert_close(as_double(got["colsum|%d" % na_rm]),
                         as_double(G[k]), rtol=1e-12, atol=1e-9, what=k)
    exp = as_double(G[runners.key_summ(base, "countNAs", 0, None)])
    assert got["countNAs"][0] == exp.reshape(-1)[0]
