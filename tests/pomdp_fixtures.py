"""Inputs for the POMDP tests and bench: beliefs and alpha vectors.

The alpha vectors the tree consumes come from the FIB and PBVI offline
solvers (reference: fast_informed_bound_cuda.cu, point_based_value_iteration_
cuda.cu; product: pp2d_pomdp_solve_fib / pp2d_pomdp_solve_pbvi, which the
bench uses).  The tree tests need small, CPU-computable sets: FIB alphas from
the oracle's restatement of the FIB solver, and as lower-bound set blind-policy
value vectors (valid lower bounds) plus convex mixtures of them -- the tree
code accepts any alpha set."""
import os

import numpy as np

import cases
import pomdp_oracle_py as po

_cache = {}


def alphas(name_or_grid, goal, gamma=cases.GAMMA, n_pbvi=24, fib_sweeps=0, seed=0):
    key = (name_or_grid if isinstance(name_or_grid, str) else id(name_or_grid),
           goal, gamma, n_pbvi, fib_sweeps)
    if key in _cache:
        return _cache[key]
    grid = cases.load_bundled(name_or_grid) if isinstance(name_or_grid, str) else name_or_grid
    m = po.Model(grid, goal)
    fib, _ = m.fib(gamma, fib_sweeps)
    blind = np.stack([m.blind(gamma, a, 120) for a in range(9)])
    rng = np.random.default_rng(seed)
    rows = [blind[a] for a in range(9)]
    acts = list(range(9))
    while len(rows) < n_pbvi:
        w = rng.dirichlet(np.ones(9)).astype(np.float32)
        rows.append((w[:, None] * blind).sum(0).astype(np.float32))
        acts.append(int(np.argmax(w)))
    pbvi = np.ascontiguousarray(np.stack(rows[:n_pbvi]), dtype=np.float32)
    out = (m, np.ascontiguousarray(fib, np.float32), pbvi,
           np.arange(9, dtype=np.uint8), np.array(acts[:n_pbvi], np.uint8))
    _cache[key] = out
    return out


def gaussian_beliefs(grid, n, sigma=2.0, seed=0):
    """isotropic Gaussian bump on a seeded free
    cell, zeroed on occupied cells, normalised (float32)."""
    h, w = grid.shape
    rng = np.random.default_rng(seed)
    free = np.argwhere(grid == 0)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h * w), np.float32)
    for i in range(n):
        cy, cx = free[rng.integers(len(free))]
        b = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / np.float32(2 * sigma * sigma))
        b = (b * (grid == 0)).astype(np.float32)
        out[i] = (b / b.sum(dtype=np.float32)).reshape(-1)
    return out


def uniforms():
    """The 100 cuRAND numbers (tests/golden/curand_xorwow_1234.npy, produced on
    a B200 by make_golden.py pomdp)."""
    return np.load(os.path.join(cases.GOLDEN, "curand_xorwow_1234.npy"))


def bundled_alphas(n_pbvi=500, seed=0):
    """Config 5 inputs: sparse_map_100x40, goal (95,34):
    9 FIB alpha vectors (oracle FIB solver run to its stopping rule, committed
    as tests/golden/alphas_sparse_map_100x40.npz) and n_pbvi lower-bound
    vectors = the 9 blind-policy values plus seeded convex mixtures of them
    (a CPU-computable lower-bound set of the same size as the reference's PBVI
    set; used by profiling runs that must skip the 29 000 solver launches)."""
    g = np.load(os.path.join(cases.GOLDEN, "alphas_sparse_map_100x40.npz"))
    fib, blind = g["fib"], g["blind"]
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.ones(9), size=max(n_pbvi - 9, 0)).astype(np.float32)
    pbvi = np.concatenate([blind, w @ blind]).astype(np.float32)[:n_pbvi]
    acts = np.concatenate([np.arange(9), w.argmax(1)]).astype(np.uint8)[:n_pbvi]
    return (np.ascontiguousarray(fib), np.ascontiguousarray(pbvi),
            np.arange(9, dtype=np.uint8), acts)


def write_text_rows(path, arr):
    """One row per line, "%15.8f" per value: the format of the reference's
    save*DataToFile (model_generation_cuda.cu:74-107, fast_informed_bound_cuda.cu:
    343-352, point_based_value_iteration_cuda.cu:747-756)."""
    arr = np.asarray(arr, np.float32)
    arr = arr.reshape(arr.shape[0], -1)
    with open(path, "w") as f:
        for row in arr:
            f.write("".join(["%15.8f" % v for v in row.tolist()]) + "\n")


def text_round(arr):
    """What fscanf("%f") reads back from "%15.8f" text: the decimal string
    converted to float32 by libc's strtof (one rounding, not via double)."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    a = np.asarray(arr, np.float32)
    flat = a.reshape(-1).tolist()
    out = np.array([libc.strtof(("%15.8f" % v).encode(), None) for v in flat], np.float32)
    return out.reshape(a.shape)


def write_actions(path, acts):
    with open(path, "w") as f:
        f.write("".join("%10u\n" % int(a) for a in acts))
