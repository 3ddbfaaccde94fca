"""Column-sharded use of the reference-facing API, one process per GPU.

Every rank holds the columns [l0, l1) of the matrix as its own host
SVT_SparseMatrix and calls the ordinary methods (`.Call` -> GPU) on it.
Column-shaped results are final per shard.  Row-shaped results compose across
shards exactly as the reference's R code composes them within one matrix
(R/SparseArray-matrixStats.R:300-310, 511-517, 645-661), with one allreduce of
a length-nrow vector where the R code has a whole-matrix value:

    nvals  = ncol_total - sum_over_shards(rowCountNAs)
    sums   = sum_over_shards(rowSums)
    X2     = sum_over_shards(centered_X2_sum(shard, center))   # each shard
             # starts from center^2 * ncol_shard, so the sum starts from
             # center^2 * ncol_total (src/SparseArray_matrixStats.c:1052-1058)
    rowVars = X2 / (nvals - 1)

The vectors are host arrays (they come out of `.Call`), so the reduction runs
on a CPU-capable process group (gloo) when one is given; nrow doubles only.
"""
import numpy as np

from . import svt as S


def _allreduce_sum(a, group):
    import torch
    import torch.distributed as dist
    if group is None or not dist.is_initialized() or \
            dist.get_world_size(group) == 1:
        return np.asarray(a, dtype=np.float64)
    t = torch.from_numpy(np.array(a, dtype=np.float64))
    # NA_real_ must survive the sum as NA: carry an NA count beside the data
    na = torch.from_numpy(S.is_na_real(np.asarray(a)).astype(np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(na, op=dist.ReduceOp.SUM, group=group)
    out = t.numpy()
    out[na.numpy() > 0] = S.NA_REAL
    return out


def colSums(x, na_rm=False):
    return S.colSums(x, na_rm=na_rm)


def colMeans(x, na_rm=False):
    return S.colMeans(x, na_rm=na_rm)


def colVars(x, na_rm=False, center=None):
    return S.colVars(x, na_rm=na_rm, center=center)


def rowSums(x, na_rm=False, group=None):
    return _allreduce_sum(S.rowSums(x, na_rm=na_rm), group)


def rowCountNAs(x, group=None):
    return _allreduce_sum(S.rowCountNAs(x, useNames=False), group)


def _ncol_total(x, group):
    import torch.distributed as dist
    if group is None or not dist.is_initialized() or \
            dist.get_world_size(group) == 1:
        return float(x.dim[1])
    return float(_allreduce_sum(np.array([x.dim[1]], dtype=np.float64),
                                group)[0])


def rowMeans(x, na_rm=False, group=None):
    nvals = _ncol_total(x, group)
    if na_rm:
        nvals = nvals - rowCountNAs(x, group)
    sums = rowSums(x, na_rm=na_rm, group=group)
    with np.errstate(all="ignore"):
        return np.asarray(S._keep_na(sums, sums / nvals))


def rowVars(x, na_rm=False, center=None, group=None):
    nvals = _ncol_total(x, group)
    if na_rm:
        nvals = nvals - rowCountNAs(x, group)
    with np.errstate(all="ignore"):
        if center is None:
            sums = rowSums(x, na_rm=na_rm, group=group)
            center = np.asarray(S._keep_na(sums, sums / nvals))
        x2 = _allreduce_sum(
            S._rowStats("centered_X2_sum", x, na_rm=na_rm,
                        center=np.asarray(center), useNames=False), group)
        return np.asarray(S._keep_na(x2, x2 / (nvals - 1)))
