"""GPU box: the GPU PBVI solver against what the reference's own solver
produced (tools/ref_offline.py outputs).  usage: python tools/pbvi_check.py <ref.npz> ..."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from path_planning_2d_b200 import PomdpPathPlanning2d  # noqa: E402

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)

for path in sys.argv[1:]:
    g = np.load(path)
    grid, goal, gamma = g["grid"], tuple(int(v) for v in g["goal"]), float(g["gamma"])
    n = g["belief_set"].shape[0]
    with PomdpPathPlanning2d(grid, goal, gamma) as p:
        p.generateBeliefSet(g["b0"], min(n, 8))            # warm-up
        t0 = time.perf_counter()
        bs = p.generateBeliefSet(g["b0"], n, rand_seed=1)
        t1 = time.perf_counter()
        same_rows = (bits(bs) == bits(g["belief_set"])).all(axis=1)
        print(os.path.basename(path), f"belief set {t1-t0:.3f}s (reference {g['seconds'][1]:.2f}s)",
              "identical rows", int(same_rows.sum()), "/", n,
              "first diff", int(np.argmin(same_rows)) if not same_rows.all() else None)
        t0 = time.perf_counter()
        al, ac = p.backupAlphaVectors(g["belief_set"])
        t1 = time.perf_counter()
        same_al = (bits(al) == bits(g["pbvi"])).all(axis=1)
        err = np.abs(al - g["pbvi"]).max() / max(np.abs(g["pbvi"]).max(), 1e-30)
        print(f"   backup {t1-t0:.3f}s (reference {g['seconds'][2]:.2f}s)",
              "identical alphas", int(same_al.sum()), "/", n,
              "identical actions", int((ac == g["pbvi_actions"]).sum()), "/", n,
              f"max rel err {err:.3e}")
        # lower-bound values at the belief points
        v_ref = (g["belief_set"] * g["pbvi"][:, None, :].swapaxes(0, 1)).sum(-1).max(1) \
            if n <= 64 else None
