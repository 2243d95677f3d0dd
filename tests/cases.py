"""Shared test/bench case definitions (maps, goals, synthetic generator)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)
GAMMA = 0.95  # launch/mdp_path_planning_2d.launch:21 (discount_factor)

# name -> (goal (x, y), start (x, y)).  sparse_map_100x40 with goal (95,34) and
# start (11,6) are the launch-file defaults
# (launch/mdp_path_planning_2d.launch:7-9, dummy_simulator.launch:9-13);
# map_5x5 goal (3,2) / start (1,2) are the commented alternatives in the same
# files; the other goals are free cells chosen by this repo.
BUNDLED = {
    "map_3x3": ((1, 1), (0, 2)),
    "map_5x5": ((3, 2), (1, 2)),
    "map_10x10": ((8, 7), (1, 1)),
    "map_100x40": ((95, 34), (11, 6)),
    "sparse_map_100x40": ((95, 34), (11, 6)),
}


def load_bundled(name):
    """Occupancy grid exported from the reference PNG by make_golden.py."""
    return np.load(os.path.join(GOLDEN, "maps", name + ".npy"))


def synthetic_map(height, width, p_occupied=0.20, seed=12345, goal=None):
    """i.i.d. occupancy: numpy PCG64,
    cell occupied iff u < p.  The goal cell is forced free."""
    rng = np.random.default_rng(seed)
    grid = np.empty((height, width), dtype=np.uint8)
    step = max(1, (1 << 24) // max(width, 1))
    for r in range(0, height, step):   # chunked: bounded temporary memory
        n = min(step, height - r)
        grid[r:r + n] = rng.random((n, width), dtype=np.float32) < p_occupied
    if goal is None:
        goal = (width // 2, height // 2)
    grid[goal[1], goal[0]] = 0
    return grid, goal
