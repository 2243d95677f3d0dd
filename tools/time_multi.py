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
grid, goal = cases.synthetic_map(4096 * n, 4096, 0.20, seed=12345, goal=(2048, 2048))
with MdpPathPlanning2d(grid, goal, cases.GAMMA, devices=n) as m:
    m.sweeps(100, want_action=False)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        m.sweeps(100, want_action=False)      # synchronous on return
        best = min(best, time.perf_counter() - t0)
    cost, _ = m.download()
    print(f"devices {n}  p2p {m.peer_to_peer}  edge_short {os.environ.get('PP2D_P2P_EDGE_SHORT', 'default')}  "
          f"{best / 50 * 1e6:7.2f} us/launch (wall, 50 fused launches per device)  "
          f"crc {zlib.crc32(cost.tobytes()):08x}", flush=True)
