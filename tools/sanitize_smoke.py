#!/usr/bin/env python
"""Every kernel once at small sizes (for compute-sanitizer memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sparsearray_b200 as sa
from sparsearray_b200 import synth
from sparsearray_b200.device import DeviceSVT

os.environ["SVTGPU_CP_IMPL"] = "force"
for vt, lac in (("integer", False), ("double", False), ("integer", True)):
    d = DeviceSVT.generate_poisson(3000, 257, 0.09, seed=3, na_rate=1e-3,
                                   val_type=vt, lacunar=lac)
    for op in ("sum", "mean", "var1", "max", "min", "countNAs", "prod"):
        d.colstats(op, na_rm=True)
    for op in ("sum", "max", "min", "countNAs"):
        d.rowstats(op, na_rm=True)
    d.rowmoments(na_rm=True)
    if vt == "double" or lac:
        y = torch.randn(3000, 50, dtype=torch.float64, device="cuda")
        dd = torch.randn(257, 50, dtype=torch.float64, device="cuda")
        d.crossprod(y)
        d.matmul(dd)
    torch.cuda.synchronize()
for impl in ("tiles", "flat", "f64acc", "acc32"):
    os.environ["SVTGPU_ROW_IMPL"] = impl
    d = DeviceSVT.generate_poisson(3000, 257, 0.09, seed=3, na_rate=1e-3)
    d.rowstats("sum", na_rm=True); d.rowstats("max"); d.rowmoments()
    torch.cuda.synchronize()
os.environ.pop("SVTGPU_ROW_IMPL")
os.environ["SVTGPU_COLSTATS_IMPL"] = "tma"
d = DeviceSVT.generate_poisson(3000, 257, 0.09, seed=3, na_rate=1e-3)
d.colstats("var1"); d.colstats("sum")
os.environ.pop("SVTGPU_COLSTATS_IMPL")
x = synth.poisson_svt(2000, 120, 0.1, seed=5, na_rate=1e-3)
sa.colSums(x); sa.rowSums(x); sa.rowVars(x, na_rm=True); sa.rowProds(x)
xd = x.with_type("double")
sa.crossprod(xd, np.random.randn(2000, 9)); sa.matmul(xd, np.random.randn(120, 9))
torch.cuda.synchronize()
# round-2 kernels: lacunar rowsum in packed registers, pipelined colsum pieces,
# row maxima as a histogram, the element-parallel transpose, wrapped 16-bit
# offsets, double colVars in one pass
rng = np.random.Generator(np.random.PCG64(4))
xl = synth.poisson_svt(70000, 90, 0.02, seed=6, lacunar=True)
rg = rng.integers(1, 13, size=xl.dim[0]).astype(np.int32)
cg = rng.integers(1, 5, size=xl.dim[1]).astype(np.int32)
sa.rowsum(xl, rg); sa.colsum(xl, cg); sa.rowSums(xl); sa.colSums(xl)
rg2 = rng.integers(1, 13, size=x.dim[0]).astype(np.int32)
cg2 = rng.integers(1, 5, size=x.dim[1]).astype(np.int32)
sa.rowsum(x, rg2); sa.colsum(x, cg2); sa.rowMaxs(x, na_rm=True); sa.rowMins(x)
h = sa.to_device(x)
sa.rowSums(h); sa.colVars(h, na_rm=True); sa.rowMaxs(h)
h.release()
sa.colVars(xd); sa.colSds(xd, na_rm=True)
for vt in ("integer", "double"):
    d = DeviceSVT.generate_poisson(3000, 700, 0.09, seed=3, na_rate=1e-3,
                                   val_type=vt)
    d.rowstats_via_transpose("prod", na_rm=True) if hasattr(
        d, "rowstats_via_transpose") else None
    d.matmul(torch.randn(700, 40, dtype=torch.float64, device="cuda")) \
        if vt == "double" else None
torch.cuda.synchronize()
print("sanitize smoke done")
