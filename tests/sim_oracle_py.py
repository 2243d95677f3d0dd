"""ctypes access to the TEST-ONLY simulator-filter oracle (oracle/liboracle_sim.so)
and to the reference's own three methods (oracle/_ref/libpp2d_ref_sim.so, CPU
only, built from the reference source lines by oracle/Makefile: ref_sim)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_i32, _vp, _u8 = ctypes.c_int32, ctypes.c_void_p, ctypes.c_uint8
_libs = {}


def _load(which):
    if which not in _libs:
        path = (os.path.join(ROOT, "oracle", "liboracle_sim.so") if which == "oracle"
                else os.path.join(ROOT, "oracle", "_ref", "libpp2d_ref_sim.so"))
        L = ctypes.CDLL(path)
        pre = "oracle_sim" if which == "oracle" else "ref_sim"
        fa = getattr(L, pre + "_update_action")
        fm = getattr(L, pre + "_update_measurement")
        fa.argtypes = [_i32, _i32, _vp, _vp, _u8]
        fm.argtypes = [_i32, _i32, _vp, _vp, _vp]
        _libs[which] = (fa, fm)
    return _libs[which]


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpp2d_ref_sim.so"))


def update(grid, belief, action=None, measurement=None, which="oracle"):
    """One belief through updateBelief(u) and/or updateBelief(meas)."""
    fa, fm = _load(which)
    grid = np.ascontiguousarray(grid, np.uint8)
    h, w = grid.shape
    b = np.array(belief, dtype=np.float32, copy=True).reshape(-1)
    if action is not None:
        fa(h, w, grid.ctypes.data, b.ctypes.data, int(action))
    if measurement is not None:
        m = np.ascontiguousarray(measurement, np.uint8)
        fm(h, w, grid.ctypes.data, b.ctypes.data, m.ctypes.data)
    return b


def scenario(name):
    """(grid, [start beliefs], [(action, measurement), ...]) of a fixture."""
    import cases
    import pomdp_fixtures as pf
    rng = np.random.default_rng(len(name))
    if name.startswith("syn"):
        grid, _ = cases.synthetic_map(23, 31, 0.3, seed=9)
    else:
        grid = cases.load_bundled(name)
    free = (grid.reshape(-1) == 0).astype(np.float32)
    beliefs = [free / free.sum(dtype=np.float32),
               pf.gaussian_beliefs(grid, 1, seed=4)[0],
               # mass on occupied cells too, some exact zeros, tiny values
               (rng.random(grid.size, dtype=np.float32) ** 8 *
                (rng.random(grid.size) < 0.7)).astype(np.float32)]
    beliefs[2] /= beliefs[2].sum(dtype=np.float32)
    steps = [(int(a), [int(v) for v in m]) for a, m in
             zip([0, 1, 2, 3, 4, 5, 6, 7, 8, 5, 5, 7],
                 rng.integers(0, 2, size=(12, 4)))]
    return grid, beliefs, steps


SCENARIOS = ["map_3x3", "map_10x10", "sparse_map_100x40", "syn_23x31"]


def run_scenario(name, which):
    """All intermediate beliefs: out[b][2*k] after the action of step k,
    out[b][2*k+1] after its measurement."""
    grid, beliefs, steps = scenario(name)
    out = np.zeros((len(beliefs), 2 * len(steps), grid.size), np.float32)
    for i, b in enumerate(beliefs):
        cur = b
        for k, (a, m) in enumerate(steps):
            cur = update(grid, cur, action=a, which=which)
            out[i, 2 * k] = cur
            cur = update(grid, cur, measurement=m, which=which)
            out[i, 2 * k + 1] = cur
    return out


def crc_rows(out):
    """CRC-32 of the float bits of every belief of run_scenario's output."""
    import zlib
    return np.array([[zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in b] for b in out],
                    np.uint32)
