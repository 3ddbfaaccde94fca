#!/usr/bin/env python
"""Phase trace of the stateless .Call path on a host SVT (SVTGPU_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVTGPU_TRACE", "1")
import bench
import sparsearray_b200 as sa
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nthread = int(sys.argv[2]) if len(sys.argv) > 2 else 0
x = bench.host_sample(ncols)
x.r_SVT
if nthread:
    sa.set_SparseArray_nthread(nthread)
print("nthread", sa.get_SparseArray_nthread(), "nnz", x.nnz, flush=True)
for rep in range(2):
    for name, fn in (("colSums", lambda: sa.colSums(x, na_rm=True)),
                     ("rowSums", lambda: sa.rowSums(x, na_rm=True)),
                     ("rowVars", lambda: sa.rowVars(x, na_rm=True))):
        t0 = time.perf_counter()
        fn()
        print("%s wall %.1f ms" % (name, (time.perf_counter() - t0) * 1e3),
              sa.last_timings(), flush=True)
