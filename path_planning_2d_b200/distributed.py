"""Row-sharded value iteration: one process per GPU, torch.distributed.

The reference is single-GPU; this is the new
multi-GPU driver of section 8e.  The H x W grid is cut into contiguous row
blocks, one per rank.  Jacobi sweeps only need the neighbours' boundary rows:
every shard keeps 2 ghost rows of J above and below and advances 2 sweeps per
fused kernel launch.  On GPUs of one node the fused kernel itself writes its
first / last two rows into the neighbours' ghost rows through CUDA-IPC peer
mappings (NVLink stores) and synchronises with them through flags in device
memory: no exchange call and no collective between fused launches.  Single
sweeps, arg-min sweeps and the CPU tests swap the 2 rows with
send/recv (NCCL / gloo) instead.  Every 100 sweeps the
per-rank inf-norm is combined with a one-float MAX all-reduce and compared
with the reference's threshold (src/mdp/path_planning_2d.cu:221,263).
Jacobi iteration is partition invariant, so the result is bit-identical to
the single-GPU run.

The shard itself is pluggable (`shard_factory`) so the orchestration can be
tested on CPU with the oracle standing in for the GPU shard (tests/ only).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

HALO_ROWS = 2


def partition_rows(height, world_size):
    """Contiguous, near-equal row blocks; every block has >= HALO_ROWS rows."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if world_size > 1 and height < HALO_ROWS * world_size:
        raise ValueError(
            f"{height} rows cannot be split over {world_size} ranks "
            f"(each shard needs >= {HALO_ROWS} rows)")
    base, extra = divmod(height, world_size)
    bounds, r = [], 0
    for i in range(world_size):
        n = base + (1 if i < extra else 0)
        bounds.append((r, r + n))
        r += n
    return bounds


class _DevMem:
    """Expose raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {
            "shape": (count,), "typestr": typestr, "data": (int(ptr), False),
            "version": 2}


def device_tensor(ptr, nbytes, dtype=torch.float32):
    typestr = {torch.float32: "<f4", torch.uint8: "|u1"}[dtype]
    count = nbytes // {torch.float32: 4, torch.uint8: 1}[dtype]
    return torch.as_tensor(_DevMem(ptr, count, typestr), device="cuda")


class GpuShard:
    """One rank's rows on its GPU: thin adapter over the C-ABI handle."""

    def __init__(self, grid, goal, gamma, rows):
        from .mdp import MdpPathPlanning2d
        self.mdp = MdpPathPlanning2d(grid, goal, gamma, rows=rows)
        # run on torch's current stream so NCCL and our kernels are ordered
        self.mdp.set_stream(torch.cuda.current_stream().cuda_stream,
                            asynchronous=True)

    def sweeps(self, n, want_action):
        self.mdp.sweeps(n, want_action)

    def reset(self, grid, goal):
        torch.cuda.current_stream().synchronize()
        self.mdp.reset(grid, goal)

    # peer-to-peer ghost rows (same node): descriptor = CUDA IPC handles
    def p2p_descriptor(self):
        return self.mdp.ipc_export()

    def p2p_connect(self, up_desc, down_desc):
        self.mdp.ipc_connect(up_desc, down_desc)

    def p2p_timed_out(self):
        return self.mdp.p2p_timed_out()

    def halo_tensors(self):
        h = self.mdp.halo()
        t = lambda p: device_tensor(p, h.bytes)
        return (t(h.send_top), t(h.send_bottom), t(h.recv_top), t(h.recv_bottom))

    def residual_tensor(self):
        return device_tensor(self.mdp.residual_device(), 4)

    def download(self):
        torch.cuda.current_stream().synchronize()
        return self.mdp.download()

    def close(self):
        self.mdp.close()


class ShardedValueIteration:
    def __init__(self, grid, goal, gamma, rank=None, world_size=None,
                 group=None, shard_factory=GpuShard, p2p=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world_size is None else world_size
        self.gamma = np.float32(gamma)
        self.height, self.width = grid.shape
        self.bounds = partition_rows(self.height, self.world)
        self.rows = self.bounds[self.rank]
        self.shard = shard_factory(grid, goal, gamma, self.rows)
        self.n_sweeps = 0
        # Peer-to-peer ghost rows: default on when the shard supports it.
        if p2p is None:
            p2p = (self.world > 1 and hasattr(self.shard, "p2p_descriptor")
                   and os.environ.get("PP2D_P2P", "1") != "0")
        self.p2p = bool(p2p) and self.world > 1
        self._fused_pending = False     # fused P2P launches since the last barrier
        if self.p2p:
            descs = [None] * self.world
            dist.all_gather_object(descs, self.shard.p2p_descriptor(), group=self.group)
            up = descs[self.rank - 1] if self.rank > 0 else None
            down = descs[self.rank + 1] if self.rank + 1 < self.world else None
            self.shard.p2p_connect(up, down)
            self._token = torch.zeros(1, device="cuda")

    def reset(self, grid=None, goal=None):
        """Re-solve from J = 0 with a new map (same shape) and/or goal."""
        if self.p2p:
            self._cross_rank_barrier()
        self.shard.reset(grid, goal)
        if self.p2p:                 # nobody starts over before everyone reset
            dist.all_reduce(self._token, group=self.group)
            torch.cuda.current_stream().synchronize()
        self.n_sweeps = 0

    # -- ghost rows --------------------------------------------------------
    def exchange(self):
        if self.world == 1:
            return
        send_top, send_bottom, recv_top, recv_bottom = self.shard.halo_tensors()
        ops = []
        up, down = self.rank - 1, self.rank + 1
        if up >= 0:
            ops.append(dist.P2POp(dist.isend, send_top, self._global(up), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_top, self._global(up), self.group))
        if down < self.world:
            ops.append(dist.P2POp(dist.isend, send_bottom, self._global(down), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_bottom, self._global(down), self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def _global(self, group_rank):
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)

    # -- sweeps ------------------------------------------------------------
    def _cross_rank_barrier(self):
        """Stream-ordered: completes on a rank only after every rank's earlier
        kernels (and their peer stores) have completed."""
        if self._fused_pending:
            dist.all_reduce(self._token, group=self.group)
            self._fused_pending = False

    def sweeps(self, n, want_action=True):
        """n Jacobi sweeps of the whole grid; ghost rows refreshed every 2."""
        left = n
        while left > 0:
            k = min(HALO_ROWS, left)
            last = (left - k == 0)
            if self.p2p and k == 2:
                # ghost rows travel inside the fused kernel (value-only pair, or
                # a pair whose second sweep also writes the greedy action)
                self.shard.sweeps(2, want_action and last)
                self._fused_pending = True
            else:
                if self.p2p:
                    self._cross_rank_barrier()
                self.shard.sweeps(k, want_action and last)
                self.exchange()
            left -= k
        self.n_sweeps += n

    def residual(self):
        t = self.shard.residual_tensor()
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def value_iteration(self, max_batches=64):
        """The reference's loop (path_planning_2d.cu:219-263) on the shards."""
        max_optimal_cost = 5.0 / (1.0 - float(self.gamma))
        residuals = []
        while True:
            self.sweeps(100)
            residuals.append(self.residual())
            if not (residuals[-1] > max_optimal_cost * 1e-3):
                break
            if len(residuals) >= max_batches:
                break
        return self.n_sweeps, residuals

    def download(self):
        if self.p2p:
            self._cross_rank_barrier()
            if self.shard.p2p_timed_out():
                raise RuntimeError("peer-to-peer ghost-row exchange timed out")
        return self.shard.download()

    def gather(self):
        """Rank 0 receives the whole J and action grids (others get None)."""
        cost, action = self.download()
        if self.world == 1:
            return cost, action
        parts = [None] * self.world if self.rank == 0 else None
        dist.gather_object((cost, action), parts, dst=self._global(0),
                           group=self.group)
        if self.rank != 0:
            return None, None
        return (np.concatenate([p[0] for p in parts]),
                np.concatenate([p[1] for p in parts]))

    def close(self):
        self.shard.close()
