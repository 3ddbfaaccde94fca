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
print("sanitize smoke done")
