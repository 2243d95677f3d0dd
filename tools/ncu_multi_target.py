"""ncu target (>= 2 GPUs): a few fused launches of a single-process 2-GPU handle
(pp2d_mdp_create_multi: ghost rows written by the kernel into the neighbour's HBM over
NVLink).  One process, so ncu can read per-launch NVLink byte counters:
  ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum \
      -k regex:mdp_sweep_kernel -c 24 python tools/ncu_multi_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from path_planning_2d_b200 import MdpPathPlanning2d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = 4096 * n
grid, goal = cases.synthetic_map(rows, 4096, 0.20, seed=12345, goal=(2048, 2048))
with MdpPathPlanning2d(grid, goal, cases.GAMMA, devices=n) as m:
    assert m.peer_to_peer, "needs peer access between the devices"
    m.sweeps(24, want_action=False)          # 12 fused launches per device
    print("ok", m.sweep_count, m.residual())
