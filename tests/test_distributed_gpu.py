"""GPU tier, needs >= 2 GPUs (skipped otherwise): the NCCL row-sharded driver
against the single-GPU solve, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import cases
import oracle_py

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, p2p):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from path_planning_2d_b200.distributed import ShardedValueIteration
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        grid, goal = cases.synthetic_map(301, 517, 0.2, seed=77)
        vi = ShardedValueIteration(grid, goal, cases.GAMMA, p2p=p2p)
        assert vi.p2p == p2p
        vi.sweeps(5, want_action=False)
        sweeps, residuals = vi.value_iteration(max_batches=2)
        cost, action = vi.gather()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), cost=cost, action=action,
                     sweeps=sweeps, residuals=np.array(residuals))
        vi.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p,lin", [(True, -1), (False, -1), (True, 1), (True, 0)])
def test_nccl_sharded_solve_matches_single_gpu(tmp_path, monkeypatch, p2p, lin):
    """p2p=True: ghost rows written by the fused kernel into the neighbours'
    HBM (CUDA IPC + device flags); p2p=False: NCCL send/recv after every launch."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("PP2D_MDP_LINEAR_UNITS", str(lin))   # unit scheme of the sweep kernels
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), p2p), nprocs=world,
             join=True)
    got = np.load(tmp_path / "out.npz")
    grid, goal = cases.synthetic_map(301, 517, 0.2, seed=77)
    ora = oracle_py.OracleMdp(grid, goal, cases.GAMMA)
    ora.sweeps(205)
    assert int(got["sweeps"]) == 205
    assert np.array_equal(got["cost"].view(np.uint32), ora.cost.view(np.uint32))
    assert np.array_equal(got["action"], ora.act)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_single_process_multi_gpu_handle_matches_the_oracle(n):
    """pp2d_mdp_create_multi on n distinct devices: one host thread, ghost rows
    through the fused kernel's peer stores; bit-identical to the oracle."""
    from path_planning_2d_b200 import MdpPathPlanning2d
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs >= {n} GPUs")
    grid, goal = cases.synthetic_map(301, 517, 0.2, seed=77)
    ora = oracle_py.OracleMdp(grid, goal, cases.GAMMA)
    with MdpPathPlanning2d(grid, goal, cases.GAMMA, devices=n) as mdp:
        assert mdp.device_count == n and mdp.peer_to_peer
        for k, wa in [(5, False), (100, True), (1, True), (100, True)]:
            mdp.sweeps(k, wa)
            ora.sweeps(k)
        cost, action = mdp.download()
        assert np.array_equal(cost.view(np.uint32), ora.cost.view(np.uint32))
        assert np.array_equal(action, ora.act)
        mdp.reset()
        sweeps, residuals = mdp.initialize()
        J, A, m, res = oracle_py.value_iteration(grid, goal, cases.GAMMA)
        assert sweeps == m and np.array_equal(residuals, res)
        assert np.array_equal(mdp.optimal_cost.view(np.uint32), J.view(np.uint32))
        assert np.array_equal(mdp.optimal_action, A)


@pytest.mark.parametrize("shape,short,lin,rpu", [
    ((3000, 300), 0, 1, 0), ((3000, 300), 20, 1, 0), ((3000, 300), 60, 1, 0),
    ((1200, 1000), 20, 1, 0), ((160, 4000), 20, 1, 0),          # unit table (strip-major runs)
    ((3000, 300), 12, 0, 100), ((3000, 300), 40, 0, 100)])       # shorter boundary row blocks
def test_cost_balanced_units_keep_the_bits(monkeypatch, shape, short, lin, rpu):
    """The units of the peer-to-peer kernels that do the hand-shake get fewer
    rows (PP2D_P2P_EDGE_SHORT; 0 = equal rows): a table of unit boundaries over
    the strip-major row sequence, or shorter first / last row blocks.  The
    partition must not change a bit, whatever the shape (many short units,
    units spanning several strips)."""
    from path_planning_2d_b200 import MdpPathPlanning2d
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("PP2D_P2P_EDGE_SHORT", str(short))
    monkeypatch.setenv("PP2D_MDP_LINEAR_UNITS", str(lin))
    monkeypatch.setenv("PP2D_MDP_ROWS_PER_UNIT", str(rpu))
    grid, goal = cases.synthetic_map(shape[0], shape[1], 0.2, seed=78)
    ora = oracle_py.OracleMdp(grid, goal, cases.GAMMA)
    with MdpPathPlanning2d(grid, goal, cases.GAMMA, devices=2) as mdp:
        assert mdp.peer_to_peer
        for k, wa in [(6, False), (20, True), (3, True)]:
            mdp.sweeps(k, wa)
            ora.sweeps(k)
        cost, action = mdp.download()
    assert np.array_equal(cost.view(np.uint32), ora.cost.view(np.uint32))
    assert np.array_equal(action, ora.act)
