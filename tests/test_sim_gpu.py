"""GPU tier: pp2d_sim_* (the dummy_simulator's Bayes filter on the GPU,
SURVEY.md section 8f row 4b) through the C ABI against the oracle and the
reference record, bit for bit; plus the cross-check SURVEY section 2 row 17
suggests: the simulator's filter and the planner's cudaBayesBeliefUpdate are
two statements of the same Bayes rule."""
import os

import numpy as np
import pytest

import cases
import pomdp_fixtures as pf
import sim_oracle_py as so
from path_planning_2d_b200 import DummySimulator, PomdpPathPlanning2d, _lib

pytestmark = pytest.mark.gpu
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", so.SCENARIOS)
def test_scenarios_bit_exact(name):
    """Every intermediate belief of the scenario, all start beliefs in one
    batch (each with its own action / measurement stream position)."""
    g = np.load(os.path.join(cases.GOLDEN, f"sim_{name}.npz"))
    grid, beliefs, steps = so.scenario(name)
    out = np.zeros((len(beliefs), 2 * len(steps), grid.size), np.float32)
    with DummySimulator(grid) as sim:
        cur = np.stack(beliefs)
        for k, (a, m) in enumerate(steps):
            cur = sim.updateBelief(cur, action=a)
            out[:, 2 * k] = cur
            cur = sim.updateBelief(cur, measurement=m)
            out[:, 2 * k + 1] = cur
        # controlCallback's order in one call
        both = sim.updateBelief(np.stack(beliefs), action=steps[0][0], measurement=steps[0][1])
    assert np.array_equal(so.crc_rows(out), g["crc"])
    assert np.array_equal(bits(out[:, -1]), bits(g["last"]))
    assert np.array_equal(bits(both), bits(out[:, 1]))


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (9, 1), (2, 2), (17, 33), (64, 129)])
def test_ragged_shapes_and_mixed_batches(shape):
    h, w = shape
    grid, _ = cases.synthetic_map(h, w, 0.3, seed=h * 7 + w)
    rng = np.random.default_rng(h + w)
    n = 23
    beliefs = (rng.random((n, h * w), dtype=np.float32) ** 6).astype(np.float32)
    beliefs[::4, ::3] = 0
    beliefs[1] = 0
    beliefs[1, rng.integers(h * w)] = 1.0                    # a certain robot
    acts = rng.integers(9, size=n).astype(np.uint8)
    meas = rng.integers(0, 2, size=(n, 4)).astype(np.uint8)
    with DummySimulator(grid) as sim:
        got_a = sim.updateBelief(beliefs, action=acts)
        got_m = sim.updateBelief(beliefs, measurement=meas)
    for i in range(n):
        want_a = so.update(grid, beliefs[i], action=acts[i])
        want_m = so.update(grid, beliefs[i], measurement=meas[i])
        nan = np.isnan(want_a)
        assert np.array_equal(bits(got_a[i])[~nan], bits(want_a)[~nan]), i
        assert np.array_equal(np.isnan(got_a[i]), nan), i
        nan = np.isnan(want_m)
        assert np.array_equal(bits(got_m[i])[~nan], bits(want_m)[~nan]), i
        assert np.array_equal(np.isnan(got_m[i]), nan), i


def test_bad_arguments():
    grid = np.zeros((4, 4), np.uint8)
    with DummySimulator(grid) as sim:
        with pytest.raises(_lib.Pp2dError) as e:
            sim.updateBelief(np.full(16, 1 / 16, np.float32), action=9)
        assert e.value.code == _lib.PP2D_ERR_INVALID


def test_simulator_filter_agrees_with_the_planners_bayes_update():
    """On beliefs without mass on occupied cells the simulator's model (no
    trapped-cell override) and the planner's tables describe the same motion,
    so predict + correct + normalise must agree up to rounding (different
    summation orders and FMA use: tolerance 1e-6 relative to the peak)."""
    name, goal = "sparse_map_100x40", (95, 34)
    grid = cases.load_bundled(name)
    beliefs = pf.gaussian_beliefs(grid, 16, sigma=2.0, seed=3)
    rng = np.random.default_rng(0)
    acts = rng.integers(9, size=16).astype(np.uint8)
    meas = rng.integers(0, 2, size=(16, 4)).astype(np.uint8)
    obs = (meas[:, 3] << 3 | meas[:, 2] << 2 | meas[:, 1] << 1 | meas[:, 0]).astype(np.uint8)
    with DummySimulator(grid) as sim:
        s = sim.updateBelief(beliefs, action=acts, measurement=meas)
    with PomdpPathPlanning2d(grid, goal, cases.GAMMA) as p:
        q, _ = p.bayes_update(beliefs, acts, obs, normalize=True)
    ok = ~np.isnan(q).any(axis=1)
    assert ok.sum() >= 8
    assert np.abs(s[ok] - q[ok]).max() <= 1e-6 * max(1.0, float(q[ok].max()))
