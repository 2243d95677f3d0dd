"""CPU tier: the multi-rank orchestration (row partition, ghost-row exchange
every 2 sweeps, MAX all-reduce of the residual, the reference's stopping
rule) over gloo with world_size 2 and 3, the oracle standing in for the GPU
shard.  The result must equal the unsharded oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import oracle_py
from path_planning_2d_b200.distributed import (ShardedValueIteration,
                                               grid_checksum, partition_rows)


def test_grid_checksum_is_partition_invariant_and_bit_sensitive():
    rng = np.random.default_rng(0)
    cost = rng.random((37, 53), dtype=np.float32)
    act = rng.integers(9, size=(37, 53)).astype(np.uint8)
    whole = grid_checksum(cost, act)
    for cuts in ([0, 37], [0, 5, 37], [0, 2, 4, 30, 37]):
        parts = sum(grid_checksum(cost[a:b], act[a:b], a) for a, b in zip(cuts[:-1], cuts[1:]))
        assert parts & 0xFFFFFFFFFFFFFFFF == whole
    c2 = cost.copy()
    c2.view(np.uint32)[20, 7] ^= 1                  # one mantissa bit
    assert grid_checksum(c2, act) != whole
    a2 = act.copy()
    a2[3, 3] = (a2[3, 3] + 1) % 9
    assert grid_checksum(cost, a2) != whole
    swapped = cost.copy()
    swapped[[4, 5]] = swapped[[5, 4]]               # rows exchanged
    assert grid_checksum(swapped, act) != whole


def test_partition_rows():
    assert partition_rows(10, 1) == [(0, 10)]
    assert partition_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    b = partition_rows(16384, 8)
    assert b[0] == (0, 2048) and b[-1] == (14336, 16384)
    assert all(e - s == 2048 for s, e in b)
    with pytest.raises(ValueError):
        partition_rows(3, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        goal, _ = cases.BUNDLED[name]
        grid = cases.load_bundled(name)
        vi = ShardedValueIteration(grid, goal, cases.GAMMA,
                                   shard_factory=cpu_shard.OracleShard)
        vi.sweeps(7, want_action=False)      # odd count: 2+2+2+1
        vi.sweeps(1)
        sweeps, residuals = vi.value_iteration()
        assert vi.converged
        checksum = vi.checksum()
        cost, action = vi.gather()
        # a second problem on the same shards: the next map is staged while the
        # first solution is still there, then reset() starts over from J = 0
        grid2 = np.ascontiguousarray(grid[::-1])
        goal2 = (goal[0], grid.shape[0] - 1 - goal[1])
        vi.stage_map(grid2)
        assert vi.shard.staged is grid2
        vi.reset(grid2, goal2)
        assert vi.n_sweeps == 0
        vi.sweeps(9)
        res2 = vi.residual()
        cost2, action2 = vi.gather()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), cost=cost, action=action,
                     sweeps=sweeps, residuals=np.array(residuals),
                     checksum=np.uint64(checksum), cost2=cost2, action2=action2,
                     res2=np.float32(res2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_value_iteration_gloo(tmp_path, world):
    name = "sparse_map_100x40"
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world,
             join=True)
    got = np.load(tmp_path / "out.npz")
    goal, _ = cases.BUNDLED[name]
    grid = cases.load_bundled(name)
    ora = oracle_py.OracleMdp(grid, goal, cases.GAMMA)
    ora.sweeps(8)
    prev = np.zeros_like(ora.cost)   # residual is since the last check point
    residuals = []
    thr = 5.0 / (1.0 - float(np.float32(cases.GAMMA))) * 1e-3
    while True:
        ora.sweeps(100)
        residuals.append(float(np.abs(ora.cost - prev).max()))
        prev = ora.cost.copy()
        if not residuals[-1] > thr:
            break
    assert int(got["sweeps"]) == ora.n
    assert got["residuals"].tolist() == [float(np.float32(r)) for r in residuals]
    assert np.array_equal(got["cost"].view(np.uint32), ora.cost.view(np.uint32))
    assert np.array_equal(got["action"], ora.act)
    assert not np.isnan(got["cost"]).any()
    assert int(got["checksum"]) == grid_checksum(ora.cost, ora.act)
    # the staged / reset second problem (the map upside down)
    grid2 = np.ascontiguousarray(grid[::-1])
    ora2 = oracle_py.OracleMdp(grid2, (goal[0], grid.shape[0] - 1 - goal[1]), cases.GAMMA)
    ora2.sweeps(9)
    assert np.array_equal(got["cost2"].view(np.uint32), ora2.cost.view(np.uint32))
    assert np.array_equal(got["action2"], ora2.act)
    assert float(got["res2"]) == float(np.float32(np.abs(ora2.cost).max()))
