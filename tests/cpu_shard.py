"""TEST-ONLY CPU stand-in for the GPU shard: the oracle sweep on a rank-local
view of the grid in which every row outside [begin-2, end+2) is poisoned with
NaN after each exchange, so any mistake in the ghost-row protocol of
path_planning_2d_b200/distributed.py shows up as NaN in owned rows."""
import numpy as np
import torch

import oracle_py


class OracleShard:
    def __init__(self, grid, goal, gamma, rows):
        self.begin, self.end = rows
        self.h, self.w = grid.shape
        self.ora = oracle_py.OracleMdp(grid, goal, gamma)
        self.prev = np.zeros((self.h, self.w), np.float32)
        self._halo = None
        self._poison()

    def reset(self, grid, goal):
        # (new map and goal on the same shard, as GpuShard.reset)
        self.ora = oracle_py.OracleMdp(grid, goal, self.ora.gamma)
        self.prev = np.zeros((self.h, self.w), np.float32)
        self._halo = None
        self.staged = getattr(self, "staged", None)
        self._poison()

    def stage_map(self, grid):
        # (the GPU shard starts an upload here; the stand-in only records the call)
        self.staged = grid

    def _poison(self):
        J = self.ora.J[self.ora.cur].reshape(self.h, self.w)
        lo, hi = max(0, self.begin - 2), min(self.h, self.end + 2)
        J[:lo] = np.nan
        J[hi:] = np.nan

    def sweeps(self, n, want_action):
        assert n <= 2
        self._commit_halo()
        self.ora.sweeps(n)

    # ghost rows: tensors that alias nothing; committed on the next call
    def halo_tensors(self):
        J = self.ora.J[self.ora.cur].reshape(self.h, self.w)
        b, e = self.begin, self.end
        z = lambda: torch.zeros(2 * self.w, dtype=torch.float32)
        send_top = torch.from_numpy(J[b:b + 2].copy().reshape(-1))
        send_bottom = torch.from_numpy(J[e - 2:e].copy().reshape(-1))
        self._halo = (z(), z())
        return send_top, send_bottom, self._halo[0], self._halo[1]

    def _commit_halo(self):
        if self._halo is None:
            return
        J = self.ora.J[self.ora.cur].reshape(self.h, self.w)
        b, e = self.begin, self.end
        if b >= 2:
            J[b - 2:b] = self._halo[0].numpy().reshape(2, self.w)
        if e + 2 <= self.h:
            J[e:e + 2] = self._halo[1].numpy().reshape(2, self.w)
        self._halo = None
        self._poison()

    def residual_tensor(self):
        self._commit_halo()
        J = self.ora.cost[self.begin:self.end]
        r = np.abs(J - self.prev[self.begin:self.end]).max()
        self.prev[self.begin:self.end] = J
        return torch.tensor([r], dtype=torch.float32)

    def download(self):
        self._commit_halo()
        return (self.ora.cost[self.begin:self.end].copy(),
                self.ora.act[self.begin:self.end].copy())

    def close(self):
        pass
