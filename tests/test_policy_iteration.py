"""Policy iteration (dead code in the reference:
src/mdp/path_planning_2d.cu:271-357 with the kernels of
path_planning_2d_cuda.cu:266-355).

tests/golden/pi_ref_*.npz: the reference's loop around its OWN kernels run on a
B200 (oracle/_ref/libpp2d_ref_mdp.so, `make_golden.py pi`).  CPU tier: the C
oracle equals them bit for bit.  GPU tier: pp2d_mdp_policy_iteration equals
them and the oracle (J, policy, sweep count, residual and changed-action
sequences)."""
import glob
import os

import numpy as np
import pytest

import cases
import oracle_py

GOLD = sorted(glob.glob(os.path.join(cases.GOLDEN, "pi_ref_*.npz")))
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-4] for p in GOLD])
def test_oracle_policy_iteration_equals_reference(path):
    g = np.load(path)
    goal = tuple(int(v) for v in g["goal"])
    J, A, n, res, chg = oracle_py.policy_iteration(g["grid"], goal, float(g["gamma"]))
    assert n == int(g["sweeps"])
    assert np.array_equal(bits(J), bits(g["J"]))
    assert np.array_equal(A, g["action"])
    assert np.array_equal(res, g["residuals"])
    assert np.array_equal(chg, g["changed"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-4] for p in GOLD])
def test_product_policy_iteration_equals_reference(path):
    from path_planning_2d_b200 import MdpPathPlanning2d
    g = np.load(path)
    goal = tuple(int(v) for v in g["goal"])
    with MdpPathPlanning2d(g["grid"], goal, float(g["gamma"])) as mdp:
        n, res, chg = mdp.policyIteration()
        cost, action = mdp.download()
    assert n == int(g["sweeps"])
    assert np.array_equal(bits(cost), bits(g["J"]))
    assert np.array_equal(action, g["action"])
    assert np.array_equal(res, g["residuals"])
    assert np.array_equal(chg, g["changed"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,p_occ", [((97, 143), 0.2), ((40, 100), 0.0), ((64, 64), 0.45)])
def test_product_policy_iteration_equals_oracle(shape, p_occ):
    from path_planning_2d_b200 import MdpPathPlanning2d, _lib
    grid, goal = cases.synthetic_map(shape[0], shape[1], p_occ, seed=shape[0] + shape[1])
    J, A, n, res, chg = oracle_py.policy_iteration(grid, goal, cases.GAMMA, max_rounds=6)
    with MdpPathPlanning2d(grid, goal, cases.GAMMA) as mdp:
        gn, gres, gchg = mdp.policyIteration(max_rounds=6)
        cost, action = mdp.download()
        assert gn == n
        assert np.array_equal(bits(cost), bits(J))
        assert np.array_equal(action, A)
        assert np.array_equal(gres, res) and np.array_equal(gchg, chg)
        # value-iteration sweeps need a reset after policy iteration, and policy
        # iteration needs a fresh handle
        with pytest.raises(_lib.Pp2dError):
            mdp.sweeps(2)
        mdp.reset(grid, goal)
        mdp.sweeps(3)
        with pytest.raises(_lib.Pp2dError):
            mdp.policyIteration(max_rounds=1)
