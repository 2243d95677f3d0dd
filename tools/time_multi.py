"""GPU box with >= 2 GPUs: fused-launch time of the single-process multi-GPU handle
(weak scaling: 4096 x 4096 per device).  usage: python tools/time_multi.py [n_devices]"""
import os
import sys
import time
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from path_planning_2d_b200 import MdpPathPlanning2d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
width = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
# 4096 wide: weak scaling, 4096 rows per device; 16384 wide: 2048-row shards as in the
# 8-GPU split of the 16384^2 grid
rows = 4096 * n if width == 4096 else 2048 * n
grid, goal = cases.synthetic_map(rows, width, 0.20, seed=12345, goal=(width // 2, 1024))
with MdpPathPlanning2d(grid, goal, cases.GAMMA, devices=n) as m:
    m.sweeps(100, want_action=False)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        m.sweeps(100, want_action=False)      # synchronous on return
        best = min(best, time.perf_counter() - t0)
    cost, _ = m.download()
    print(f"{rows}x{width}  devices {n}  p2p {m.peer_to_peer}  edge_short {os.environ.get('PP2D_P2P_EDGE_SHORT', 'default')}  "
          f"{best / 50 * 1e6:7.2f} us/launch (wall, 50 fused launches per device)  "
          f"crc {zlib.crc32(cost.tobytes()):08x}", flush=True)
