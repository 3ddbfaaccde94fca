"""Randomised GPU-vs-oracle parity (tests/fuzz_gpu.py): 80 random matrices --
shapes from 1 x 1 to ~150,000 rows, densities 0.001 .. 1, integer / double /
lacunar / count values, NA / NaN / Inf -- through the .Call entry points
(column and row statistics, rowsum / colsum, the device transpose, crossprod)
against the oracle port.  The oracle is only the checker."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_matrices_against_the_oracle(seed):
    spec = importlib.util.spec_from_file_location(
        "fuzz_gpu", os.path.join(ROOT, "tests", "fuzz_gpu.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.Generator(np.random.PCG64(seed))
    for it in range(40):
        x, kind = fz.make(rng)
        try:
            fz.check(x, kind, rng)
        except Exception as e:
            raise AssertionError("seed %d, matrix %d: kind=%s dim=%s nnz=%d: %s"
                                 % (seed, it, kind, x.dim, x.nnz, e)) from e
