"""Small ncu target: a few launches of the fused kernel at 4096^2 (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from path_planning_2d_b200 import MdpPathPlanning2d  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
grid, goal = cases.synthetic_map(size, size, 0.20, seed=12345)
with MdpPathPlanning2d(grid, goal, cases.GAMMA) as m:
    m.sweeps(n, want_action=False)      # n/2 fused launches
    m.sweeps(2)                         # plain + policy
    r = m.residual()
    torch.cuda.synchronize()
    print("ok", m.sweep_count, r)
