#!/usr/bin/env python
"""Multi-GPU parity: column-sharded row statistics (NCCL allreduce of the row
states) against the same matrix on one GPU.  Launch with torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
from sparsearray_b200 import _native as N
from sparsearray_b200.device import DeviceSVT


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    N.check(N.lib().svtgpu_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nrow, per = 33538, 20000
    shard = DeviceSVT.generate_poisson(nrow, per, 0.07, seed=5, na_rate=1e-4,
                                       leaf0=rank * per,
                                       nleaf_total=world * per)
    res = {}
    for op in ("sum", "max", "min", "countNAs"):
        for na_rm in (False, True):
            res[(op, na_rm)] = shard.rowstats(op, na_rm=na_rm,
                                              group=dist.group.WORLD)[0]
    mean, var = shard.rowmoments(na_rm=True, group=dist.group.WORLD)
    K = 8
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    D_all = torch.randn(world * per, K, dtype=torch.float64, device="cuda",
                        generator=g)
    dsh = DeviceSVT.generate_poisson(nrow, per, 0.07, seed=5, na_rate=0.0,
                                     val_type="double", leaf0=rank * per,
                                     nleaf_total=world * per)
    mm = dsh.matmul(D_all[rank * per:(rank + 1) * per].contiguous(),
                    group=dist.group.WORLD)
    ok = True
    if rank == 0:
        full = DeviceSVT.generate_poisson(nrow, world * per, 0.07, seed=5,
                                          na_rate=1e-4)
        for (op, na_rm), v in res.items():
            e = full.rowstats(op, na_rm=na_rm)[0]
            same = torch.equal(torch.nan_to_num(v.double(), nan=-7.0),
                               torch.nan_to_num(e.double(), nan=-7.0))
            print("row %-8s na_rm=%-5s %s" % (op, na_rm,
                                              "identical" if same else "DIFF"))
            ok = ok and same
        em, ev = full.rowmoments(na_rm=True)
        same = torch.equal(mean, em) and torch.allclose(var, ev, rtol=1e-12,
                                                        atol=0)
        print("rowmoments", "ok" if same else "DIFF")
        ok = ok and same
        fulld = DeviceSVT.generate_poisson(nrow, world * per, 0.07, seed=5,
                                           na_rate=0.0, val_type="double")
        emm = fulld.matmul(D_all)
        same = torch.allclose(mm, emm, rtol=1e-12, atol=1e-9)
        print("matmul", "ok" if same else "DIFF",
              float((mm - emm).abs().max()))
        ok = ok and same
        print("MULTIGPU_PARITY", "PASS" if ok else "FAIL", "world", world)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
