"""GPU tier: the reference's DEFAULT start-up (read_data_from_file=true,
launch/pomdp_path_planning_2d.launch:11; src/pomdp/path_planning_2d.cu:127-143)
through pp2d_pomdp_set_model_tables and the C++ host mirror's
loadModelDataFromFile / loadFibDataFromFile / loadPbviDataFromFile.

The text format ("%15.8f") is lossy: 0.02^4 = 1.6e-7 is read back as
0.00000016, so a planner started from files plans on different likelihoods
than one that generated its model.  Both sides therefore start from files the
reference itself wrote (saveModelDataToFile ... savePbviDataToFile of
oracle/_ref/libpp2d_ref_pomdp_full.so) and the reference reads them back with
its own loaders.  Bar: bit-exact tables, alpha vectors, tree records."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import tree_scenario as ts
from path_planning_2d_b200 import PomdpPathPlanning2d
from test_host_mirror import exe, fnv  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(cases.ROOT, "oracle", "_ref", "libpp2d_ref_pomdp_full.so")
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module", params=["map_10x10", "sparse_map_100x40"])
def ref_from_files(request, tmp_path_factory):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libpp2d_ref_pomdp_full.so not built")
    case = request.param
    d = tmp_path_factory.mktemp("ckpt_" + case)
    subprocess.run([sys.executable, os.path.join(cases.GOLDEN, "make_golden.py"),
                    "tree_files", str(d), case], check=True, timeout=900)
    return case, d, dict(np.load(os.path.join(d, f"tree_files_{case}.npz")))


def test_text_checkpoint_changes_the_model(ref_from_files):
    """The premise: the loaded tables are NOT the generated ones."""
    case, d, g = ref_from_files
    grid, goal = ts.inputs(case)[:2]
    with PomdpPathPlanning2d(grid, goal, cases.GAMMA) as p:
        tp, mp, sr = p.model_tables()
    assert not np.array_equal(bits(mp), bits(g["meas_prob"]))
    assert np.abs(mp - g["meas_prob"]).max() < 1e-8
    # every loaded value is the float nearest to an 8-decimal string
    back = np.array([float("%15.8f" % v) for v in g["meas_prob"].reshape(-1)[:4096]])
    assert np.array_equal(back.astype(np.float32), g["meas_prob"].reshape(-1)[:4096])


def test_tree_on_loaded_tables_equals_the_reference(ref_from_files):
    """pp2d_pomdp_set_model_tables + set_alphas with what the reference's own
    loaders produced: evaluate() and every tree record of the scenario."""
    case, d, g = ref_from_files
    grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand = ts.inputs(case)
    with PomdpPathPlanning2d(grid, goal, cases.GAMMA) as p:
        p.set_model_tables(g["trans_prob"], g["meas_prob"], g["stage_reward"])
        tp, mp, sr = p.model_tables()
        assert np.array_equal(bits(tp), bits(g["trans_prob"]))
        assert np.array_equal(bits(mp), bits(g["meas_prob"]))
        assert np.array_equal(bits(sr), bits(g["stage_reward"]))
        p.set_alphas(g["fib"], g["pbvi"], g["fib_actions"], g["pbvi_actions"])
        up, ua, lo, la = p.evaluate(np.stack(beliefs))
        ev = np.array([[up[i].view(np.uint32), ua[i], lo[i].view(np.uint32), la[i]]
                       for i in range(len(beliefs))], np.uint32)
        assert np.array_equal(ev, g["evaluate"])
        be = ts.ProductBackend(p)
        for i, b in enumerate(beliefs):
            rec = ts.run(be, b, n_expand)
            want = {k[len(f"b{i}_"):]: v for k, v in g.items() if k.startswith(f"b{i}_")}
            assert set(rec) == set(want)
            assert ts.same_record(rec, want) is None, (case, i)
        be.t.close()


def test_cpp_planner_started_from_the_references_files(ref_from_files, exe, tmp_path):  # noqa: F811
    """PomdpPathPlanning2d::initialize with read_data_from_file=true parses the
    reference-written files itself: tables and alpha vectors as the
    reference's loaders read them, same two belief callbacks."""
    import cv2
    case, d, g = ref_from_files
    grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand = ts.inputs(case)
    png = str(tmp_path / "map.png")
    cv2.imwrite(png, np.where(grid == 1, 0, 255).astype(np.uint8))
    (tmp_path / "belief.bin").write_bytes(beliefs[0].tobytes())
    out = subprocess.run([exe, "pomdp", png, str(goal[0]), str(goal[1]), "0.95",
                          str(d / "data"), str(tmp_path / "belief.bin"), str(pbvi.shape[0]),
                          str(n_expand)], capture_output=True, text=True)
    lines = out.stdout.splitlines()
    tab = dict(t.split("=") for t in [l for l in lines if l.startswith("TABLES")][0].split()[1:])
    assert tab["tp"] == fnv(g["trans_prob"].tobytes())
    assert tab["mp"] == fnv(g["meas_prob"].tobytes())
    assert tab["sr"] == fnv(g["stage_reward"].tobytes())
    assert tab["fib"] == fnv(g["fib"].tobytes()) and tab["pbvi"] == fnv(g["pbvi"].tobytes())
    assert tab["acts"] == fnv(g["pbvi_actions"].tobytes())
    kv = dict(t.split("=") for t in [l for l in lines if l.startswith("RESULT")][0].split()[1:])
    a0, r0, a1, r1 = (int(v) for v in g["callbacks"])
    assert (int(kv["a0"]), int(kv["a1"])) == (a0, a1)
    assert (kv["r0"], kv["r1"]) == ("%08x" % r0, "%08x" % r1)


def test_binary_checkpoint_is_lossless(exe, tmp_path):  # noqa: F811
    """data_format=binary: a planner restarted from pp2d_data.bin holds the
    bits the solving planner held (tables, alpha vectors) and answers the same
    belief callback; the text checkpoint of the same planner does not."""
    import cv2
    g = np.load(os.path.join(cases.GOLDEN, "pbvi_ref_map_10x10_g0.8_n40.npz"))
    grid, goal = g["grid"], (int(g["goal"][0]), int(g["goal"][1]))
    png = str(tmp_path / "map.png")
    cv2.imwrite(png, np.where(grid == 1, 0, 255).astype(np.uint8))
    (tmp_path / "bin").mkdir()
    (tmp_path / "txt").mkdir()
    res = {}
    for fmt in ("bin", "txt"):
        out = subprocess.run([exe, "pomdp_solve", png, str(goal[0]), str(goal[1]), "0.8", "40",
                              "4", str(tmp_path / fmt), "binary" if fmt == "bin" else "text"],
                             capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        assert line, out.stdout + out.stderr
        res[fmt] = dict(t.split("=") for t in line[0].split()[1:])
    assert res["bin"] == res["txt"]
    assert res["bin"]["pbvi"] == fnv(g["pbvi"].tobytes())
    free = (grid.reshape(-1) == 0).astype(np.float32)
    s = np.float32(0)
    for v in free:
        s = np.float32(s + v)
    (tmp_path / "belief.bin").write_bytes((free / s).astype(np.float32).tobytes())
    with PomdpPathPlanning2d(grid, goal, 0.8) as p:
        tp, mp, sr = p.model_tables()
    tabs = {}
    for fmt in ("bin", "txt"):
        out = subprocess.run([exe, "pomdp", png, str(goal[0]), str(goal[1]), "0.8",
                              str(tmp_path / fmt), str(tmp_path / "belief.bin"), "40", "4",
                              "binary" if fmt == "bin" else "text"], capture_output=True, text=True)
        lines = out.stdout.splitlines()
        tabs[fmt] = dict(t.split("=") for t in
                         [l for l in lines if l.startswith("TABLES")][0].split()[1:])
        kv = dict(t.split("=") for t in [l for l in lines if l.startswith("RESULT")][0].split()[1:])
        if fmt == "bin":
            assert kv["a0"] == res["bin"]["a0"]
    assert tabs["bin"]["tp"] == fnv(tp.tobytes()) and tabs["bin"]["mp"] == fnv(mp.tobytes())
    assert tabs["bin"]["sr"] == fnv(sr.tobytes())
    assert tabs["bin"]["fib"] == fnv(g["fib"].tobytes())
    assert tabs["bin"]["pbvi"] == fnv(g["pbvi"].tobytes())
    assert tabs["txt"]["mp"] != tabs["bin"]["mp"]        # "%15.8f" is lossy, the binary is not
