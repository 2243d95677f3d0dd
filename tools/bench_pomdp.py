"""QV-tree plans/s on one GPU (BASELINE.json configs[4] per-GPU share):
sparse_map_100x40, goal (95,34), Gaussian start beliefs (sigma 2 cells),
9 FIB + 500 lower-bound alpha vectors, depth cap 50, 15 expansions, 50
samples per Q node.  usage: python tools/bench_pomdp.py [n_queries] [--cpu]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import pomdp_fixtures as pf  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1250
    grid = cases.load_bundled("sparse_map_100x40")
    goal = (95, 34)
    fib = None
    beliefs = pf.gaussian_beliefs(grid, n, sigma=2.0, seed=0)
    from path_planning_2d_b200 import PomdpPathPlanning2d, _lib
    lib = _lib.load()
    with PomdpPathPlanning2d(grid, goal, cases.GAMMA) as p:
        free = (grid.reshape(-1) == 0).astype(np.float32)
        t0 = time.perf_counter()
        if "--fixture" in sys.argv:       # profiling runs: skip the 29 000 solver launches
            fib, pbvi, fa, pa = pf.bundled_alphas(500)
        else:
            fib, fa, _ = p.fastInformedBound()
            _, pbvi, pa = p.pointBasedValueIteration(free / free.sum(dtype=np.float32), 500)
        print(f"offline FIB + PBVI(500): {time.perf_counter() - t0:.2f} s")
        p.set_alphas(fib, pbvi, fa, pa)
        p.plan_batch(beliefs[:32])                     # warm-up, pool growth
        p.plan_batch(beliefs)
        l0 = lib.pp2d_kernel_launches()
        w0 = p.work_counters()
        t0 = time.perf_counter()
        acts, vals, stats = p.plan_batch(beliefs, with_stats=True)
        dt = time.perf_counter() - t0
        launches = lib.pp2d_kernel_launches() - l0
        w1 = p.work_counters()
        walked = (w1[2] - w0[2]) / max(1, w1[0] - w0[0])
    vn = stats[:, 0].sum()
    macs = float(vn) * grid.size * (18 + 500)
    print(f"queries {n}  time {dt*1e3:.1f} ms  plans/s {n/dt:.1f}  V-nodes {vn} "
          f"({vn/n:.1f}/plan)  cells/belief {walked:.0f}  launches {launches}  bounds {2*macs/dt/1e12:.2f} Tflop/s "
          f"(mul+add)  actions hist {np.bincount(acts, minlength=9).tolist()}")
    if "--cpu" in sys.argv:
        import pomdp_oracle_py as po
        m = po.Model(grid, goal)
        k = min(n, 4)
        t0 = time.perf_counter()
        for i in range(k):
            t = po.Tree(m, cases.GAMMA, fib, pbvi, pf.uniforms(), beliefs[i], fa, pa)
            a, r, st, rc = t.plan(50, 15)
            assert a == acts[i] and np.float32(r) == vals[i], (i, a, acts[i], r, vals[i])
            t.close()
        dc = time.perf_counter() - t0
        print(f"oracle (1 thread): {k/dc:.2f} plans/s; first {k} plans identical to the GPU's")


if __name__ == "__main__":
    main()
