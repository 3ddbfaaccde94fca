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
