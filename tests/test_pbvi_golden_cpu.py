"""CPU tier: internal consistency of the committed outputs of the reference's
offline solvers (tests/golden/pbvi_ref_*.npz, produced on a B200 by
tools/ref_offline.py from the unmodified reference translation units).  The
GPU tier (tests/test_pbvi_gpu.py) requires the product to reproduce them bit
for bit; here the fixtures themselves are sanity-checked so that a corrupted
or mis-generated file cannot silently become the bar."""
import glob
import os

import numpy as np
import pytest

import cases
import pomdp_oracle_py as po

SMALL = sorted(p for p in glob.glob(os.path.join(cases.GOLDEN, "pbvi_ref_*.npz"))
               if not p.endswith("_crc.npz"))


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[9:-4] for p in SMALL])
def test_reference_pbvi_outputs_are_consistent(path):
    g = np.load(path)
    grid, goal, gamma = g["grid"], tuple(int(v) for v in g["goal"]), float(g["gamma"])
    bs, al, ac, fib = g["belief_set"], g["pbvi"], g["pbvi_actions"], g["fib"]
    n, hw = bs.shape
    assert hw == grid.size and al.shape == bs.shape and ac.shape == (n,)
    # beliefs: the first one is the initial belief, all are distributions over free cells
    assert np.array_equal(bs[0], g["b0"])
    assert np.all(bs >= 0) and np.allclose(bs.sum(axis=1, dtype=np.float64), 1.0, atol=1e-4)
    assert np.all(bs[:, grid.reshape(-1) == 1] == 0)
    assert np.all(ac <= 8)
    # the FIB alphas of the fixture are the oracle's (same kernel arithmetic)
    m = po.Model(grid, goal)
    ofib, _ = m.fib(gamma)
    assert np.array_equal(np.ascontiguousarray(ofib).view(np.uint32), fib.view(np.uint32))
    # lower bound <= upper bound at every belief point (float64 dots)
    lower = (bs.astype(np.float64) @ al.astype(np.float64).T).max(axis=1)
    upper = (bs.astype(np.float64) @ fib.astype(np.float64)).max(axis=1)
    assert np.all(lower <= upper + 1e-3 * np.abs(upper))
    # rewards are <= 0, so every alpha vector is <= 0 and >= -2/(1-gamma)
    assert al.max() <= 1e-6 and al.min() >= -2.0 / (1.0 - gamma) - 1e-3


def test_bundled_map_crc_fixture_shape():
    g = np.load(os.path.join(cases.GOLDEN, "pbvi_ref_sparse_map_100x40_g0.95_n500_crc.npz"))
    assert g["belief_set_crc"].shape == (500,) and g["pbvi_crc"].shape == (500,)
    assert g["pbvi_actions"].shape == (500,) and np.all(g["pbvi_actions"] <= 8)
    assert len(np.unique(g["belief_set_crc"])) > 450       # the beliefs are (almost all) distinct
    grid = cases.load_bundled("sparse_map_100x40")
    free = (grid.reshape(-1) == 0).astype(np.float32)
    assert np.array_equal(g["b0"], free / free.sum(dtype=np.float32))
    assert np.array_equal(g["belief_set_head"][0], g["b0"])
