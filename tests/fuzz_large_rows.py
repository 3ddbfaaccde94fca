#!/usr/bin/env python
"""A few LARGE random matrices (>= 20 million entries, so that the cyclic
tile form of the row histogram runs) with very different leaf sizes, integer
and lacunar, against numpy: rowSums / rowMaxs / countNAs / row moments.
    python tests/fuzz_large_rows.py [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sparsearray_b200.device import DeviceSVT

NA = -2**31
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.Generator(np.random.PCG64(seed))
for case in range(8):
    nrow = int(rng.choice([64, 3000, 33538, 50000, 100000]))
    mean_leaf = float(rng.choice([3, 40, 700, 5000]))
    mean_leaf = min(mean_leaf, nrow * 0.6)
    target = int(rng.integers(20, 40)) * 1000000
    ncol = int(target / mean_leaf)
    lac = bool(rng.random() < 0.3)
    cnt = rng.poisson(mean_leaf, size=ncol).clip(0, nrow).astype(np.int64)
    if rng.random() < 0.5:
        cnt[rng.integers(0, ncol, size=ncol // 50)] = 0
    ptr = np.zeros(ncol + 1, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    nnz = int(ptr[-1])
    # ascending distinct rows per leaf: sorted random keys per leaf
    leaf = np.repeat(np.arange(ncol), cnt)
    r = rng.integers(0, nrow, size=nnz)
    order = np.lexsort((r, leaf))
    r = r[order]
    dup = np.zeros(nnz, dtype=bool)
    dup[1:] = (r[1:] == r[:-1]) & (leaf[1:] == leaf[:-1])
    keep = ~dup
    r, leaf = r[keep].astype(np.int32), leaf[keep]
    cnt = np.bincount(leaf, minlength=ncol)
    ptr = np.zeros(ncol + 1, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    nnz = int(ptr[-1])
    M = int(rng.choice([3, 12, 200, 40000]))
    vals = None
    if not lac:
        vals = rng.integers(0 if rng.random() < 0.7 else -M, M + 1,
                            size=nnz).astype(np.int32)
        vals[rng.random(nnz) < 1e-5] = NA
    d = DeviceSVT(nrow, ncol, nnz, "integer",
                  torch.from_numpy(ptr).cuda(),
                  torch.cat([torch.from_numpy(r).cuda(),
                             torch.zeros(64, dtype=torch.int32, device="cuda")]),
                  None if lac else torch.cat([torch.from_numpy(vals).cuda(),
                             torch.zeros(64, dtype=torch.int32, device="cuda")]))
    v = np.ones(nnz, dtype=np.int64) if lac else vals.astype(np.int64)
    ok = v != NA
    s1 = np.bincount(r[ok], weights=v[ok].astype(np.float64), minlength=nrow)
    nna = np.bincount(r[~ok], minlength=nrow)
    got = d.rowstats("sum", na_rm=True)[0].cpu().numpy()
    assert np.array_equal(got, s1), ("sum", case)
    got = d.rowstats("countNAs")[0].cpu().numpy()
    assert np.array_equal(got, nna.astype(np.float64)), ("countNAs", case)
    # max over stored regular values and the implicit zero (if any)
    mx = np.full(nrow, -2**40, dtype=np.int64)
    np.maximum.at(mx, r[ok], v[ok])
    cvg = np.bincount(r, minlength=nrow)
    has_zero = cvg < ncol
    exp = np.where(has_zero, np.maximum(mx, 0), mx)
    allna = (~has_zero) & (np.bincount(r[ok], minlength=nrow) == 0)
    got = d.rowstats("max", na_rm=True)[0].cpu().numpy()
    assert np.array_equal(got[~allna], exp[~allna].astype(got.dtype)), ("max", case)
    mean, var = d.rowmoments(na_rm=True)
    n = ncol - nna
    s2 = np.bincount(r[ok], weights=v[ok].astype(np.float64) ** 2, minlength=nrow)
    with np.errstate(all="ignore"):
        e_mean = s1 / n
        e_var = (s2 - s1 * s1 / n) / (n - 1)
    m_ = mean.cpu().numpy(); v_ = var.cpu().numpy()
    good = n > 1
    assert np.allclose(m_[good], e_mean[good], rtol=1e-12, atol=0), ("mean", case)
    scale = (s2 + s1 * s1 / np.maximum(n, 1)) / np.maximum(n - 1, 1)
    assert np.all(np.abs(v_[good] - e_var[good]) <= 1e-11 * scale[good] + 1e-300), ("var", case)
    print("case %d ok: nrow %d ncol %d nnz %.1fM mean leaf %.0f lacunar %s max %d"
          % (case, nrow, ncol, nnz / 1e6, mean_leaf, lac, M), flush=True)
print("LARGE FUZZ PASS")
