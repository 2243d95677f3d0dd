"""GPU tier: pin the CPU restatement (oracle/mdp_oracle.c) against the
UNMODIFIED reference kernels compiled from /root/reference into oracle/_ref.
This is what makes the oracle trustworthy; the same comparison, frozen, is
tests/golden/ref_*.npz for the CPU tier."""
import os
import sys

import numpy as np
import pytest

import cases
import oracle_py

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpp2d_ref_mdp.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import make_golden
    return make_golden, make_golden.ref_lib()


def test_reference_kernels_vs_restatement(ref):
    mg, lib = ref
    for name, grid, goal, gamma, max_batches, with_tables in mg.golden_cases():
        J, A, n, res = mg.ref_solve(lib, grid, goal, gamma, max_batches)
        oJ, oA, on, ores = oracle_py.value_iteration(grid, goal, gamma, max_batches)
        assert n == on, name
        assert np.array_equal(J.view(np.uint32), oJ.view(np.uint32)), name
        assert np.array_equal(A, oA), name
        assert np.array_equal(res, ores), name
        if with_tables:
            tp, sc = mg.ref_tables(lib, grid, goal)
            otp, osc = oracle_py.tables(grid, goal)
            assert np.array_equal(tp.view(np.uint32), otp.view(np.uint32)), name
            assert np.array_equal(sc.view(np.uint32), osc.view(np.uint32)), name


def test_reference_kernels_vs_restatement_random_512(ref):
    mg, lib = ref
    grid, goal = cases.synthetic_map(512, 384, 0.2, seed=99)
    J, A, n, res = mg.ref_solve(lib, grid, goal, cases.GAMMA, 1)
    oJ, oA, on, ores = oracle_py.value_iteration(grid, goal, cases.GAMMA, 1)
    assert np.array_equal(J.view(np.uint32), oJ.view(np.uint32))
    assert np.array_equal(A, oA)
    assert np.array_equal(res, ores)
