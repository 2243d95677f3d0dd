"""CPU tier: the oracle's QV-tree restatement (oracle/pomdp_oracle.c) against
records of the reference's own SearchTree / VNode / QNode host code.

tests/golden/tree_<case>.npz were produced on a B200 by
`tests/golden/make_golden.py tree` from oracle/_ref/libpp2d_ref_pomdp_full.so
(the four reference POMDP translation units compiled unmodified).  Bar:
every bound, heuristic, depth, weight, chosen action, expansion target and
the tree shape itself identical, bit for bit, after every step of the
scenario in tests/tree_scenario.py (both branches of SearchTree::update)."""
import os

import numpy as np
import pytest

import cases
import pomdp_oracle_py as po
import tree_scenario as ts


def golden(case):
    path = os.path.join(cases.GOLDEN, f"tree_{case}.npz")
    if not os.path.exists(path):
        pytest.skip("no reference record for " + case)
    return dict(np.load(path))


@pytest.mark.parametrize("case", list(ts.CASES))
def test_oracle_tree_equals_reference_record(case):
    g = golden(case)
    grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand = ts.inputs(case)
    assert ts.checksum(grid, fib, pbvi, fa, pa, *beliefs) == g["inputs_crc"], \
        "fixture inputs drifted from the ones the reference record was made with"
    assert ts.checksum(m.tp.reshape(-1, 9, 9), m.mp.reshape(-1, 16),
                       m.sr.reshape(-1, 9)) == g["model_crc"]
    ev = [po.evaluate(b, fib, pbvi, fa, pa) for b in beliefs]
    got = np.array([[np.float32(e[0]).view(np.uint32), e[1],
                     np.float32(e[2]).view(np.uint32), e[3]] for e in ev], np.uint32)
    assert np.array_equal(got, g["evaluate"])
    ob = ts.OracleBackend(m, cases.GAMMA, fib, pbvi, fa, pa)
    for i, b in enumerate(beliefs):
        rec = ts.run(ob, b, n_expand)
        want = {k[len(f"b{i}_"):]: v for k, v in g.items() if k.startswith(f"b{i}_")}
        assert set(rec) == set(want), (sorted(rec), sorted(want))
        assert ts.same_record(rec, want) is None, (case, i)
