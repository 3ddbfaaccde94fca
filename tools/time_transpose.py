#!/usr/bin/env python
"""Cost of the device transpose (second build: memory pool warm)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsearray_b200 import _native as N
from sparsearray_b200.device import DeviceSVT
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
d = DeviceSVT.generate_poisson(33538, cols, 0.07, seed=2, na_rate=0.0,
                               val_type="double")
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    w = DeviceSVT(d.nrow, d.nleaf, d.nnz, "double", d.leaf_ptr, d.offs, d.vals)
    torch.cuda.synchronize()
    t = ctypes.c_void_p()
    l0 = N.launch_count()
    t0 = time.perf_counter()
    N.check(N.lib().svtgpu_matrix_transposed(w._h, ctypes.byref(t)))
    torch.cuda.synchronize()
    print("rep %d: transpose of %d cols (nnz %.3e): %.1f ms, %d launches"
          % (rep, cols, d.nnz, (time.perf_counter() - t0) * 1e3,
             N.launch_count() - l0), flush=True)
    w.free()
