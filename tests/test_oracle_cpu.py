"""CPU tier: the oracle restatement against closed-form known answers and the
committed golden vectors (outputs of the UNMODIFIED reference kernels run on a
B200 through oracle/_ref, see tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import cases
import oracle_py


def test_bundled_map_occupancy_counts():
    # occupancy after threshold(250, BINARY_INV).
    want = {"map_3x3": 4, "map_5x5": 4, "map_10x10": 39, "map_100x40": 2229,
            "sparse_map_100x40": 1642}
    for name, n in want.items():
        assert int(cases.load_bundled(name).sum()) == n


def test_free_space_model_row():
    # path_planning_2d_cuda.cu:89-125: 0.7 main, 0.1 x3 (incl. stay).
    grid = np.zeros((5, 5), np.uint8)
    tp, sc = oracle_py.tables(grid, (4, 4))
    centre = 2 * 5 + 2
    for u in range(9):
        row = tp[centre, u]
        assert np.isclose(row.sum(), 1.0)
        if u == 4:
            assert row[4] == 1.0
        else:
            assert row[u] == np.float32(0.7) and row[4] == np.float32(0.1)
            assert np.count_nonzero(row) == 4
    # stage cost: free neighbourhood -> 1 for moves, 2 for stay (0 at goal).
    assert np.allclose(sc[centre, [0, 1, 2, 3, 5, 6, 7, 8]], 1.0)
    assert sc[centre, 4] == 2.0
    assert sc[4 * 5 + 4, 4] == 0.0


def test_blocked_mass_moves_to_stay_and_border_is_occupied():
    grid = np.zeros((3, 3), np.uint8)
    tp, sc = oracle_py.tables(grid, (1, 1))
    corner = 0  # (0,0): slots 0,1,2,3,6 are out of the map
    # action 0 (up-left): everything except "stay" is outside -> P[4] = 1
    assert tp[corner, 0, 4] == np.float32(np.float32(np.float32(
        np.float32(0.1) + np.float32(0.7)) + np.float32(0.1)) + np.float32(0.1))
    assert np.count_nonzero(tp[corner, 0]) == 1
    # occupied cell: trapped for every action, cost 2
    grid[1, 2] = 1
    tp, sc = oracle_py.tables(grid, (1, 1))
    occ = 1 * 3 + 2
    for u in range(9):
        assert tp[occ, u, 4] == 1.0 and np.count_nonzero(tp[occ, u]) == 1
        assert sc[occ, u] == 2.0


@pytest.mark.parametrize("name", list(cases.BUNDLED))
def test_value_iteration_known_answers(name):
    goal, start = cases.BUNDLED[name]
    grid = cases.load_bundled(name)
    J, A, n, res = oracle_py.value_iteration(grid, goal, cases.GAMMA)
    # every bundled map stops after exactly 300 sweeps.
    assert n == 300
    assert np.allclose(res, [39.7631683, 0.235420227, 0.00136566162], rtol=1e-6)
    assert J[goal[1], goal[0]] == 0.0
    assert A[goal[1], goal[0]] == 4
    assert (A == 4).sum() == 1           # "stay" only at the goal
    occ = grid == 1
    assert np.all(A[occ] == 0)           # trapped cells tie -> first action
    assert np.allclose(J[occ], 2.0 / (1.0 - 0.95), rtol=1e-5)
    if name == "sparse_map_100x40":
        assert np.bincount(A.ravel(), minlength=9).tolist() == \
            [1963, 170, 302, 246, 1, 764, 112, 134, 308]
    # greedy rollout from the start reaches the goal
    path = oracle_py.waypoints(A, start)
    assert path[-1] == goal[1] * grid.shape[1] + goal[0]


def _golden_files():
    return sorted(glob.glob(os.path.join(cases.GOLDEN, "ref_*.npz")))


@pytest.mark.parametrize("path", _golden_files() or [None])
def test_oracle_matches_reference_golden(path):
    """Bit-exact: the restatement vs the reference kernels' recorded output."""
    if path is None:
        pytest.skip("tests/golden/ref_*.npz not generated yet")
    g = np.load(path)
    grid, goal = g["grid"], tuple(int(v) for v in g["goal"])
    gamma = float(g["gamma"])
    batches = int(g["sweeps"]) // 100
    J, A, n, res = oracle_py.value_iteration(grid, goal, gamma, batches)
    assert n == int(g["sweeps"])
    assert np.array_equal(J.view(np.uint32), g["J"].view(np.uint32))
    assert np.array_equal(A, g["action"])
    assert np.array_equal(res, g["residuals"])
    if "trans_prob" in g:
        tp, sc = oracle_py.tables(grid, goal)
        assert np.array_equal(tp.view(np.uint32), g["trans_prob"].view(np.uint32))
        assert np.array_equal(sc.view(np.uint32), g["stage_cost"].view(np.uint32))


def test_plan_first_strict_maximum():
    action = np.arange(12, dtype=np.uint8).reshape(3, 4) % 9
    b = np.zeros(12, np.float32)
    assert oracle_py.plan(b, action) == action.ravel()[0]   # all zero -> idx 0
    b[[5, 9]] = 0.3
    assert oracle_py.plan(b, action) == action.ravel()[5]   # first of the ties
    b[10] = 0.31
    assert oracle_py.plan(b, action) == action.ravel()[10]
