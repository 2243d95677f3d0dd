"""CPU tier: the C-ABI library loads, exports every symbol include/pp2d.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import path_planning_2d_b200 as pp
from path_planning_2d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pp2d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pp2d_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in pp2d.h but not exported"


def test_binding_table_covers_the_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.pp2d_abi_version() == 2
    assert isinstance(lib.pp2d_last_error(), bytes)


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    grid = np.zeros((4, 4), np.uint8)
    grid[1, 1] = 1
    rc = lib.pp2d_mdp_create(4, 4, grid.ctypes.data, 1, 1, 0.95, ctypes.byref(h))
    assert rc == _lib.PP2D_ERR_GOAL_OCCUPIED     # path_planning_2d.cu:84-88
    assert b"occupied" in lib.pp2d_last_error()
    rc = lib.pp2d_mdp_create(4, 4, grid.ctypes.data, 7, 0, 0.95, ctypes.byref(h))
    assert rc == _lib.PP2D_ERR_INVALID
    rc = lib.pp2d_mdp_create(0, 4, grid.ctypes.data, 0, 0, 0.95, ctypes.byref(h))
    assert rc == _lib.PP2D_ERR_INVALID
    rc = lib.pp2d_mdp_create_shard(4, 4, grid.ctypes.data, 0, 0, 0.95, 3, 2,
                                   ctypes.byref(h))
    assert rc == _lib.PP2D_ERR_INVALID
    # NULL handles / buffers are refused before any CUDA call
    assert lib.pp2d_mdp_stage_map(None, grid.ctypes.data) == _lib.PP2D_ERR_INVALID
    assert lib.pp2d_mdp_reset(None, grid.ctypes.data, 0, 0) == _lib.PP2D_ERR_INVALID
    assert lib.pp2d_mdp_download_begin(None, None, None) == _lib.PP2D_ERR_INVALID
    assert lib.pp2d_mdp_download_wait(None) == _lib.PP2D_ERR_INVALID


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Pp2dError) as e:
        pp.MdpPathPlanning2d(np.zeros((4, 4), np.uint8), (1, 1), 0.95)
    assert e.value.code == _lib.PP2D_ERR_CUDA
