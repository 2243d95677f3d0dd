"""CPU tier: the restatement of the dummy_simulator's Bayes filter
(oracle/sim_oracle.c) against the committed outputs of the reference's own
methods (tests/golden/sim_*.npz, from oracle/_ref/libpp2d_ref_sim.so = lines
440-522 and 671-773 of dummy_simulator.cpp compiled unmodified), and against
that library itself where it is built.  Bit for bit."""
import os

import numpy as np
import pytest

import cases
import sim_oracle_py as so

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", so.SCENARIOS)
def test_oracle_equals_reference_record(name):
    g = np.load(os.path.join(cases.GOLDEN, f"sim_{name}.npz"))
    out = so.run_scenario(name, "oracle")
    assert np.array_equal(so.crc_rows(out), g["crc"])
    assert np.array_equal(bits(out[:, -1]), bits(g["last"]))
    if "beliefs" in g:
        assert np.array_equal(bits(out), bits(g["beliefs"]))
    assert not np.isnan(out).any()
    assert np.allclose(out.sum(axis=2), 1.0, atol=1e-4)


@pytest.mark.parametrize("name", so.SCENARIOS)
def test_oracle_equals_live_reference_methods(name):
    if not so.have_ref():
        pytest.skip("oracle/_ref/libpp2d_ref_sim.so not built")
    assert np.array_equal(bits(so.run_scenario(name, "oracle")),
                          bits(so.run_scenario(name, "ref")))


def test_filter_properties():
    """Known answers: "stay" leaves a belief unchanged up to normalisation; a
    certain robot next to a wall keeps the blocked mass; a measurement that
    matches the surroundings sharpens the belief there."""
    grid = np.zeros((5, 5), np.uint8)
    grid[2, 3] = 1
    b = np.zeros(25, np.float32)
    b[2 * 5 + 2] = 1.0                       # robot at (2, 2), wall to the right
    stay = so.update(grid, b, action=4)
    assert np.array_equal(bits(stay), bits(b))
    right = so.update(grid, b, action=5)     # blocked: 0.7 stays, 0.1 up-right / down-right
    assert right[2 * 5 + 3] == 0.0
    assert abs(right[2 * 5 + 2] - 0.8) < 1e-6
    assert abs(right[1 * 5 + 3] - 0.1) < 1e-6 and abs(right[3 * 5 + 3] - 0.1) < 1e-6
    flat = np.full(25, 1 / 25, np.float32)
    post = so.update(grid, flat, measurement=[0, 0, 1, 0])   # occupied to the right only
    assert post[2 * 5 + 2] == post.max() > post[0]   # (cells at the right border tie)
