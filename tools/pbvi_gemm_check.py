"""GPU box: the PBVI solve with the hand-written contraction kernel (product) against the
same solve with the reference's cublasSgemm call (PP2D_PBVI_CUBLAS=1, checker) and against
the reference's own outputs (tests/golden/pbvi_ref_*.npz).  Prints how far the alpha
vectors and the lower bound they define move when near-tied arg-max decisions flip.
usage: python tools/pbvi_gemm_check.py"""
import os
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def solve(name, cublas):
    code = f"""
import sys, time, numpy as np
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import cases
from path_planning_2d_b200 import PomdpPathPlanning2d
g = np.load({os.path.join(cases.GOLDEN, name + '.npz')!r})
grid = g['grid'] if 'grid' in g else cases.load_bundled('sparse_map_100x40')
goal = tuple(int(v) for v in g['goal'])
n = g['belief_set'].shape[0] if 'belief_set' in g else 500
with PomdpPathPlanning2d(grid, goal, float(g['gamma'])) as p:
    p.pointBasedValueIteration(g['b0'], min(n, 16), rand_seed=1, iterations=2)   # warm-up
    t0 = time.perf_counter()
    bs, al, ac = p.pointBasedValueIteration(g['b0'], n, rand_seed=1)
    dt = time.perf_counter() - t0
np.savez(sys.argv[1], bs=bs, al=al, ac=ac, dt=dt)
"""
    out = f"/tmp/pbvi_{name}_{int(cublas)}.npz"
    env = dict(os.environ, PP2D_PBVI_CUBLAS="1" if cublas else "0")
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env)
    return np.load(out)


for name in ["pbvi_ref_map_3x3_g0.5_n12", "pbvi_ref_map_10x10_g0.8_n40",
             "pbvi_ref_map_10x10_g0.95_n60", "pbvi_ref_sparse_map_100x40_g0.95_n500_crc"]:
    k, c = solve(name, False), solve(name, True)
    g = np.load(os.path.join(cases.GOLDEN, name + ".npz"))
    same_bs = np.array_equal(k["bs"].view(np.uint32), c["bs"].view(np.uint32))
    if "pbvi" in g:
        ref_equal = np.array_equal(c["al"].view(np.uint32), g["pbvi"].view(np.uint32))
    else:
        crc = np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in c["al"]], np.uint32)
        ref_equal = np.array_equal(crc, g["pbvi_crc"])
    al_k, al_c = k["al"].astype(np.float64), c["al"].astype(np.float64)
    rows_same = int((k["al"].view(np.uint32) == c["al"].view(np.uint32)).all(axis=1).sum())
    scale = np.abs(al_c).max()
    # the lower bound both sets define at the belief points (float64 dots)
    lb_k = (k["bs"].astype(np.float64) @ al_k.T).max(axis=1)
    lb_c = (c["bs"].astype(np.float64) @ al_c.T).max(axis=1)
    print(f"{name}: kernel {float(k['dt']):.2f} s, cublas {float(c['dt']):.2f} s; belief sets equal "
          f"{same_bs}; cublas run == reference fixture {ref_equal}; alpha rows bit-equal "
          f"{rows_same}/{al_k.shape[0]}; actions equal {int((k['ac'] == c['ac']).sum())}/{len(k['ac'])}; "
          f"max |d alpha| {np.abs(al_k - al_c).max():.3e} (scale {scale:.3g}, rel "
          f"{np.abs(al_k - al_c).max() / scale:.2e}); max |d lower bound at belief points| "
          f"{np.abs(lb_k - lb_c).max():.3e} (rel {(np.abs(lb_k - lb_c) / np.abs(lb_c)).max():.2e})",
          flush=True)
