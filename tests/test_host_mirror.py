"""The C++ host mirror (include/pp2d/planners.hpp, map_io.hpp): compiled with
g++ against libpp2d.so and driven through tests/cpp/host_mirror_main.cpp.
CPU tier: the map loader against cv2.  GPU tier: MdpPathPlanning2d and
PomdpPathPlanning2d against the oracle."""
import glob
import os
import subprocess
import zlib

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror")


def fnv(data):
    h = 1469598103934665603
    for b in np.frombuffer(bytes(data), np.uint8).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.fixture(scope="module")
def exe():
    libdir = os.path.join(ROOT, "path_planning_2d_b200")
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_main.cpp")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", EXE, src,
           "-L", libdir, "-lpp2d", "-lz", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return EXE


@pytest.mark.parametrize("png", sorted(glob.glob(os.path.join(cases.GOLDEN, "maps", "*.png"))))
def test_map_loader_matches_opencv(exe, png):
    import cv2
    img = cv2.imread(png, cv2.IMREAD_GRAYSCALE)
    _, grid = cv2.threshold(img, 250.0, 1.0, cv2.THRESH_BINARY_INV)
    out = subprocess.run([exe, "maps", png], capture_output=True, text=True).stdout.split()
    assert out[:3] == [str(img.shape[1]), str(img.shape[0]), str(int(grid.sum()))]
    assert out[3] == fnv(np.ascontiguousarray(grid, np.uint8).tobytes())


def test_map_loader_rejects_garbage(exe, tmp_path):
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    assert subprocess.run([exe, "maps", str(bad)], capture_output=True,
                          text=True).stdout.strip() == "FAIL"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["map_10x10", "sparse_map_100x40"])
def test_cpp_mdp_planner_matches_oracle(exe, name):
    import cv2
    import oracle_py
    goal, start = cases.BUNDLED[name]
    png = os.path.join(cases.GOLDEN, "maps", name + "_rgb.png")
    img = cv2.imread(png, cv2.IMREAD_GRAYSCALE)
    _, grid = cv2.threshold(img, 250.0, 1.0, cv2.THRESH_BINARY_INV)
    grid = np.ascontiguousarray(grid, np.uint8)
    if grid[goal[1], goal[0]] or grid[start[1], start[0]]:
        pytest.skip("noise fixture blocked the goal/start")
    J, A, n, res = oracle_py.value_iteration(grid, goal, cases.GAMMA)
    path = oracle_py.waypoints(A, start)
    out = subprocess.run([exe, "mdp", png, str(goal[0]), str(goal[1]), "0.95",
                          str(start[0]), str(start[1])], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    kv = dict(t.split("=") for t in line.split()[1:])
    assert int(kv["sweeps"]) == n
    assert kv["cost"] == fnv(J.astype(np.float32).tobytes())
    assert kv["action"] == fnv(A.tobytes())
    assert int(kv["path_len"]) == len(path)
    assert kv["path"] == fnv(path.astype(np.uint32).tobytes())
    assert int(kv["cb"]) == A[start[1], start[0]]


@pytest.mark.gpu
def test_cpp_mdp_planner_on_several_gpus_in_one_process(exe):
    """MdpPathPlanning2d::initialize with num_gpus / gpu_devices binds
    pp2d_mdp_create_multi: one process drives all the row shards.  With 2+
    visible GPUs the shards sit on different devices (peer-to-peer ghost rows);
    a 1-GPU box runs three shards on device 0 (copied ghost rows).  Every
    printed hash must equal the single-GPU run's."""
    import torch
    name = "sparse_map_100x40"
    goal, start = cases.BUNDLED[name]
    png = os.path.join(cases.GOLDEN, "maps", name + "_rgb.png")
    base = [exe, "mdp", png, str(goal[0]), str(goal[1]), "0.95", str(start[0]), str(start[1])]

    def run(*extra):
        out = subprocess.run(base + list(extra), capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        assert line, out.stdout + out.stderr
        return line[0]

    single = run()
    assert run("devices=0,0,0") == single
    if torch.cuda.device_count() >= 2:
        assert run("gpus=2") == single
        assert run("devices=1,0") == single


@pytest.mark.gpu
def test_cpp_mdp_planner_rejects_occupied_goal(exe):
    png = os.path.join(cases.GOLDEN, "maps", "map_10x10_gray.png")
    out = subprocess.run([exe, "mdp", png, "0", "0", "0.95", "1", "1"],
                         capture_output=True, text=True)
    assert "INIT_FAILED" in out.stdout


@pytest.mark.gpu
def test_cpp_pomdp_planner_matches_oracle(exe, tmp_path):
    """read_data_from_file=true path: alpha vectors in the reference's text
    format ("%15.8f"), two belief callbacks (fresh tree, then re-root)."""
    import cv2
    import pomdp_fixtures as pf
    import pomdp_oracle_py as po
    name, goal = "map_10x10", (8, 7)
    png = os.path.join(cases.GOLDEN, "maps", name + "_gray.png")
    img = cv2.imread(png, cv2.IMREAD_GRAYSCALE)
    _, grid = cv2.threshold(img, 250.0, 1.0, cv2.THRESH_BINARY_INV)
    grid = np.ascontiguousarray(grid, np.uint8)
    if grid[goal[1], goal[0]]:
        pytest.skip("noise fixture blocked the goal")
    m, fib, pbvi, fa, pa = pf.alphas(grid, goal, n_pbvi=12)
    # the text round trip is lossy by format: the oracle gets the SAME rounded
    # numbers the C++ planner reads back -- alpha vectors AND model tables
    # (read_data_from_file=true loads the model too, path_planning_2d.cu:127-143)
    hw = grid.size
    for fname, arr in (("fib_alphas", fib), ("pbvi_alphas", pbvi),
                       ("model_data_trans_prob", m.tp.reshape(hw * 9, 9)),
                       ("model_data_meas_prob", m.mp.reshape(hw, 16)),
                       ("model_data_stage_reward", m.sr.reshape(hw, 9))):
        pf.write_text_rows(tmp_path / fname, arr)
    fib_r, pbvi_r = pf.text_round(fib), pf.text_round(pbvi)
    m = po.Model(grid, goal)               # private copy with the rounded tables
    m.tp[:] = pf.text_round(m.tp)
    m.mp[:] = pf.text_round(m.mp)
    m.sr[:] = pf.text_round(m.sr)
    pf.write_actions(tmp_path / "fib_actions", fa)
    pf.write_actions(tmp_path / "pbvi_actions", pa)
    b = pf.gaussian_beliefs(grid, 1, seed=5)[0]
    (tmp_path / "belief.bin").write_bytes(b.tobytes())
    out = subprocess.run([exe, "pomdp", png, str(goal[0]), str(goal[1]), "0.95",
                          str(tmp_path), str(tmp_path / "belief.bin"), "12", "6"],
                         capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    kv = dict(t.split("=") for t in line.split()[1:])
    tab = dict(t.split("=") for t in
               [l for l in out.stdout.splitlines() if l.startswith("TABLES")][0].split()[1:])
    assert tab["tp"] == fnv(m.tp.tobytes()) and tab["mp"] == fnv(m.mp.tobytes())
    assert tab["sr"] == fnv(m.sr.tobytes())
    assert tab["fib"] == fnv(fib_r.tobytes()) and tab["pbvi"] == fnv(pbvi_r.tobytes())
    t = po.Tree(m, cases.GAMMA, fib_r, pbvi_r, pf.uniforms(), b, fa, pa)
    a0, r0, _, _ = t.plan(50, 6)
    assert t.update(a0, 0) == 0
    a1, r1, _, _ = t.plan(50, 6)
    t.close()
    assert int(kv["a0"]) == a0 and int(kv["a1"]) == a1
    assert kv["r0"] == "%08x" % np.float32(r0).view(np.uint32)
    assert kv["r1"] == "%08x" % np.float32(r1).view(np.uint32)


@pytest.mark.gpu
def test_cpp_pomdp_planner_solves_offline_like_the_reference(exe, tmp_path):
    """read_data_from_file=false path: PomdpPathPlanning2d::initialize runs the
    FIB and PBVI solvers (GPU); the alpha vectors must be the ones the
    reference's own solvers produced (tests/golden/pbvi_ref_*.npz), and
    save_data must write them in the reference's text format."""
    import cv2
    g = np.load(os.path.join(cases.GOLDEN, "pbvi_ref_map_10x10_g0.8_n40.npz"))
    png = str(tmp_path / "map.png")
    cv2.imwrite(png, np.where(g["grid"] == 1, 0, 255).astype(np.uint8))
    out = subprocess.run([exe, "pomdp_solve", png, str(int(g["goal"][0])), str(int(g["goal"][1])),
                          "0.8", "40", "4", str(tmp_path)], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stdout + out.stderr
    kv = dict(t.split("=") for t in line[0].split()[1:])
    assert kv["fib"] == fnv(g["fib"].tobytes())
    assert kv["pbvi"] == fnv(g["pbvi"].tobytes())
    assert kv["acts"] == fnv(g["pbvi_actions"].tobytes())
    saved = np.loadtxt(tmp_path / "pbvi_alphas", dtype=np.float64).reshape(g["pbvi"].shape)
    want = np.array([[float("%15.8f" % v) for v in row] for row in g["pbvi"]])
    assert np.array_equal(saved, want)
    assert np.array_equal(np.loadtxt(tmp_path / "pbvi_actions", dtype=np.int64),
                          g["pbvi_actions"].astype(np.int64))
    assert np.loadtxt(tmp_path / "fib_alphas").shape == (g["grid"].size, 9)
    assert (tmp_path / "model_data_trans_prob").exists()
