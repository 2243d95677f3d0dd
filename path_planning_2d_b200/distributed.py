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


def grid_checksum(cost, action, row_begin=0):
    """Partition-invariant checksum of (rows of) a solution: CRC-32 of every
    row of J (float bits) and of the action grid, each weighted by its global
    row number, summed modulo 2^64.  The sum over the shards of a row-sharded
    solve equals the checksum of the unsharded grids exactly when every row is
    bit-identical (up to CRC collisions), so bench.py can print ONE number per
    GPU count for the 16384^2 grid instead of shipping 1.25 GiB around."""
    import zlib
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    action = np.ascontiguousarray(action, dtype=np.uint8)
    total = 0
    for r in range(cost.shape[0]):
        g = row_begin + r
        total += zlib.crc32(cost[r].tobytes()) * (2 * g + 1)
        total += zlib.crc32(action[r].tobytes()) * (2 * g + 2) * 0x9E3779B1
    return total & 0xFFFFFFFFFFFFFFFF


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

    def stage_map(self, grid):
        self.mdp.stage_map(grid)

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
        self.converged = False
        # Peer-to-peer ghost rows: on when the shard supports it AND every pair
        # of neighbouring ranks sits on one host with peer access between the
        # two devices; otherwise (other host, no NVLink/PCIe peer path, IPC
        # mapping refused) every rank uses the send/recv exchange.  The decision
        # is collective: a rank never waits for a flag its neighbour will not set.
        want = p2p
        if want is None:
            want = (hasattr(self.shard, "p2p_descriptor")
                    and os.environ.get("PP2D_P2P", "1") != "0")
        self.p2p = False
        self._fused_pending = False     # fused P2P launches since the last barrier
        if want and self.world > 1:
            import socket
            info = [None] * self.world
            mine = (socket.gethostname(), torch.cuda.current_device(),
                    self.shard.p2p_descriptor())
            dist.all_gather_object(info, mine, group=self.group)
            ok = True
            for nb in (self.rank - 1, self.rank + 1):
                if 0 <= nb < self.world:
                    same_host = info[nb][0] == mine[0]
                    ok = ok and same_host and (
                        info[nb][1] == mine[1] or
                        torch.cuda.can_device_access_peer(mine[1], info[nb][1]))
            up = info[self.rank - 1][2] if self.rank > 0 else None
            down = info[self.rank + 1][2] if self.rank + 1 < self.world else None
            if self._all_agree(ok):
                try:
                    self.shard.p2p_connect(up, down)
                    connected = True
                except Exception:
                    connected = False
                if not self._all_agree(connected):
                    if connected:
                        raise RuntimeError(
                            "peer-to-peer ghost rows: a neighbour could not map this "
                            "shard; rebuild the solver with p2p=False")
                    if p2p:
                        raise
                else:
                    self.p2p = True
                    self._token = torch.zeros(1, device="cuda")
            elif p2p:
                raise RuntimeError("peer-to-peer ghost rows requested but the ranks are not "
                                   "peer-capable neighbours on one host")

    def _all_agree(self, flag):
        """Logical AND of a per-rank flag (object all-gather: works on any backend)."""
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(flag), group=self.group)
        return all(flags)

    def reset(self, grid=None, goal=None):
        """Re-solve from J = 0 with a new map (same shape) and/or goal."""
        if self.p2p:
            self._cross_rank_barrier()
        self.shard.reset(grid, goal)
        if self.p2p:                 # nobody starts over before everyone reset
            dist.all_reduce(self._token, group=self.group)
            torch.cuda.current_stream().synchronize()
        self.n_sweeps = 0

    def stage_map(self, grid):
        """Start uploading the map of the next reset(grid, ...) while the current
        solve is still running (pp2d_mdp_stage_map); a no-op for shard back ends
        without it."""
        stage = getattr(self.shard, "stage_map", None)
        if stage is not None:
            stage(grid)

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
        """max |J - J at the previous call| over all shards.  A shard whose
        peer-to-peer hand-shake timed out reports +inf (mdp_residual_kernel):
        every rank sees it after the MAX all-reduce and raises."""
        t = self.shard.residual_tensor()
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        r = float(t.item())
        if self.p2p and r == float("inf"):
            raise RuntimeError("peer-to-peer ghost-row exchange timed out on some rank; "
                               "J is invalid (reset() clears the condition)")
        return r

    def value_iteration(self, max_batches=None):
        """The reference's loop (path_planning_2d.cu:219-263) on the shards:
        batches of 100 sweeps until the inf-norm of a batch is <=
        5/(1-gamma)*1e-3.  max_batches bounds the loop (tests); `converged`
        tells whether the stopping rule was met."""
        max_optimal_cost = 5.0 / (1.0 - float(self.gamma))
        residuals = []
        self.converged = False
        while max_batches is None or len(residuals) < max_batches:
            self.sweeps(100)
            residuals.append(self.residual())
            if not (residuals[-1] > max_optimal_cost * 1e-3):
                self.converged = True
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
        # Tensor point-to-point transfers, shard by shard (no pickling of
        # gigabyte arrays): device staging for NCCL, host tensors for gloo.
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        if self.rank != 0:
            for a in (cost, action):
                dist.send(torch.from_numpy(np.ascontiguousarray(a)).to(dev),
                          self._global(0), group=self.group)
            return None, None
        costs, actions = [cost], [action]
        for r in range(1, self.world):
            rows = self.bounds[r][1] - self.bounds[r][0]
            c = torch.empty((rows, self.width), dtype=torch.float32, device=dev)
            a = torch.empty((rows, self.width), dtype=torch.uint8, device=dev)
            dist.recv(c, self._global(r), group=self.group)
            dist.recv(a, self._global(r), group=self.group)
            costs.append(c.cpu().numpy())
            actions.append(a.cpu().numpy())
        return np.concatenate(costs), np.concatenate(actions)

    def checksum(self):
        """grid_checksum of the whole solution, identical on every rank and for
        every number of shards."""
        cost, action = self.download()
        c = grid_checksum(cost, action, self.rows[0])
        if self.world > 1:
            parts = [None] * self.world
            dist.all_gather_object(parts, c, group=self.group)
            c = sum(parts) & 0xFFFFFFFFFFFFFFFF
        return c

    def close(self):
        self.shard.close()
