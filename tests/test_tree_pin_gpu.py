"""GPU tier: the product's QV-tree (pp2d_tree_* through the C ABI) against
(1) the committed records of the reference's own SearchTree host code
(tests/golden/tree_<case>.npz) and (2) the reference stack itself, run live
in a subprocess from oracle/_ref/libpp2d_ref_pomdp_full.so.  Bit-exact."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import cases
import tree_scenario as ts
from path_planning_2d_b200 import PomdpPathPlanning2d

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(cases.ROOT, "oracle", "_ref", "libpp2d_ref_pomdp_full.so")


def product_records(case):
    grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand = ts.inputs(case)
    out = []
    with PomdpPathPlanning2d(grid, goal, cases.GAMMA) as p:
        p.set_alphas(fib, pbvi, fa, pa)
        up, ua, lo, la = p.evaluate(np.stack(beliefs))
        ev = np.array([[up[i].view(np.uint32), ua[i], lo[i].view(np.uint32), la[i]]
                       for i in range(len(beliefs))], np.uint32)
        be = ts.ProductBackend(p)
        for b in beliefs:
            out.append(ts.run(be, b, n_expand))
        be.t.close()
    return ev, out


def check(case, g):
    ev, recs = product_records(case)
    assert np.array_equal(ev, g["evaluate"])
    for i, rec in enumerate(recs):
        want = {k[len(f"b{i}_"):]: v for k, v in g.items() if k.startswith(f"b{i}_")}
        assert set(rec) == set(want), (sorted(rec), sorted(want))
        assert ts.same_record(rec, want) is None, (case, i)


@pytest.mark.parametrize("case", list(ts.CASES))
def test_product_tree_equals_reference_record(case):
    path = os.path.join(cases.GOLDEN, f"tree_{case}.npz")
    if not os.path.exists(path):
        pytest.skip("no reference record for " + case)
    check(case, dict(np.load(path)))


@pytest.mark.parametrize("case", ["map_10x10", "sparse_map_100x40"])
def test_product_tree_equals_live_reference(case):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libpp2d_ref_pomdp_full.so not built")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([sys.executable, os.path.join(cases.GOLDEN, "make_golden.py"),
                        "tree", d, case], check=True, timeout=600)
        g = dict(np.load(os.path.join(d, f"tree_{case}.npz")))
    check(case, g)
