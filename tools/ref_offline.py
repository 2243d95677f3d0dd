"""GPU box: run the reference's own offline solvers (fastInformedBound,
generateBeliefSet, backupAlphaVectors -- the unmodified translation units in
oracle/_ref/libpp2d_ref_pomdp_full.so) on one map and save what they produce.
usage: python tools/ref_offline.py <map> <gx> <gy> <gamma> <n_pbvi> <out.npz>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import pomdp_oracle_py as po  # noqa: E402


def main():
    name, gx, gy, gamma, n, out = sys.argv[1:7]
    grid = cases.load_bundled(name)
    free = 1.0 - grid.astype(np.float32).reshape(-1)
    s = np.float32(0)
    for v in free:                    # src/pomdp/path_planning_2d.cu:100-107
        s = np.float32(s + v)
    b0 = (free / s).astype(np.float32)
    ref = po.RefFull(grid, (int(gx), int(gy)), float(gamma), int(n))
    t0 = time.perf_counter()
    fib, fa = ref.solve_fib()
    t1 = time.perf_counter()
    bs = ref.belief_set(b0, seed=1)
    t2 = time.perf_counter()
    al, ac = ref.backup(bs)
    t3 = time.perf_counter()
    ref.close()
    print(f"{name}: fib {t1-t0:.2f}s  belief set {t2-t1:.2f}s  backup {t3-t2:.2f}s")
    np.savez_compressed(out, grid=grid, goal=np.array([int(gx), int(gy)]),
                        gamma=np.float32(gamma), b0=b0, fib=fib, fib_actions=fa,
                        belief_set=bs, pbvi=al, pbvi_actions=ac,
                        seconds=np.array([t1 - t0, t2 - t1, t3 - t2]))


if __name__ == "__main__":
    main()
