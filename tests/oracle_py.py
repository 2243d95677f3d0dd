"""ctypes access to the TEST-ONLY CPU oracle (oracle/liboracle_mdp.so) and,
on a GPU box, to the reference kernels (oracle/_ref).  Only tests/, smoke()
and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_vp, _u32, _f = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle_mdp.so")
        L = ctypes.CDLL(path)
        L.oracle_mdp_generate_model.restype = None
        L.oracle_mdp_generate_model.argtypes = [_u32, _u32, _u32, _u32, _vp, _vp, _vp]
        L.oracle_mdp_sweep.restype = None
        L.oracle_mdp_sweep.argtypes = [_u32, _u32, _f, _vp, _vp, _vp, _vp, _vp]
        L.oracle_mdp_inf_norm.restype = ctypes.c_double
        L.oracle_mdp_inf_norm.argtypes = [ctypes.c_uint64, _vp, _vp]
        L.oracle_mdp_value_iteration.restype = ctypes.c_int
        L.oracle_mdp_value_iteration.argtypes = [_u32, _u32, _u32, _u32, _f, _vp,
                                                 _vp, _vp, _vp, ctypes.c_int]
        L.oracle_mdp_plan.restype = ctypes.c_uint8
        L.oracle_mdp_plan.argtypes = [ctypes.c_uint64, _vp, _vp]
        L.oracle_mdp_waypoints.restype = _u32
        L.oracle_mdp_waypoints.argtypes = [_u32, _u32, _vp, _u32, _u32, _vp, _u32]
        L.oracle_set_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def set_threads(n):
    """OpenMP threads of the oracle's loops; returns the count really used."""
    return lib().oracle_set_threads(int(n))


class OracleMdp:
    """Step-by-step oracle: tables + ping-pong J, like the reference's
    device state (src/mdp/path_planning_2d_cuda.cu:26-64)."""

    def __init__(self, grid, goal, gamma):
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.h, self.w = self.grid.shape
        self.gamma = float(np.float32(gamma))
        n = self.h * self.w
        self.tp = np.zeros(n * 81, np.float32)
        self.sc = np.zeros(n * 9, np.float32)
        lib().oracle_mdp_generate_model(self.h, self.w, goal[0], goal[1],
                                        self.grid.ctypes.data,
                                        self.tp.ctypes.data, self.sc.ctypes.data)
        self.J = [np.zeros(n, np.float32), np.zeros(n, np.float32)]
        self.action = np.zeros(n, np.uint8)
        self.cur = 0
        self.n = 0

    def sweeps(self, k):
        for _ in range(k):
            lib().oracle_mdp_sweep(self.h, self.w, self.gamma,
                                   self.tp.ctypes.data, self.sc.ctypes.data,
                                   self.J[self.cur].ctypes.data,
                                   self.J[self.cur ^ 1].ctypes.data,
                                   self.action.ctypes.data)
            self.cur ^= 1
            self.n += 1

    @property
    def cost(self):
        return self.J[self.cur].reshape(self.h, self.w)

    @property
    def act(self):
        return self.action.reshape(self.h, self.w)


def value_iteration(grid, goal, gamma, max_batches=0):
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    h, w = grid.shape
    J = np.zeros(h * w, np.float32)
    A = np.zeros(h * w, np.uint8)
    res = np.zeros(256, np.float64)
    n = lib().oracle_mdp_value_iteration(h, w, goal[0], goal[1], gamma,
                                         grid.ctypes.data, J.ctypes.data,
                                         A.ctypes.data, res.ctypes.data,
                                         max_batches)
    return J.reshape(h, w), A.reshape(h, w), n, res[:n // 100].copy()


def policy_iteration(grid, goal, gamma, max_rounds=0):
    """oracle_mdp_policy_iteration -> (J, action, evaluation sweeps, residuals, changed)."""
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    h, w = grid.shape
    J = np.zeros(h * w, np.float32)
    A = np.zeros(h * w, np.uint8)
    res = np.zeros(256, np.float64)
    chg = np.zeros(256, np.uint32)
    f = lib().oracle_mdp_policy_iteration
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                  ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    n = f(h, w, goal[0], goal[1], gamma, grid.ctypes.data, J.ctypes.data, A.ctypes.data,
          res.ctypes.data, chg.ctypes.data, max_rounds)
    return J.reshape(h, w), A.reshape(h, w), n, res[:n // 50].copy(), chg[:n // 50].copy()


def tables(grid, goal):
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    h, w = grid.shape
    tp = np.zeros(h * w * 81, np.float32)
    sc = np.zeros(h * w * 9, np.float32)
    lib().oracle_mdp_generate_model(h, w, goal[0], goal[1], grid.ctypes.data,
                                    tp.ctypes.data, sc.ctypes.data)
    return tp.reshape(h * w, 9, 9), sc.reshape(h * w, 9)


def plan(belief, action):
    b = np.ascontiguousarray(belief, dtype=np.float32).reshape(-1)
    a = np.ascontiguousarray(action, dtype=np.uint8).reshape(-1)
    return lib().oracle_mdp_plan(b.size, b.ctypes.data, a.ctypes.data)


def waypoints(action, start, max_len=None):
    a = np.ascontiguousarray(action, dtype=np.uint8)
    h, w = a.shape
    max_len = max_len or h * w
    out = np.zeros(max_len, np.uint32)
    n = lib().oracle_mdp_waypoints(h, w, a.ctypes.data, start[0], start[1],
                                   out.ctypes.data, max_len)
    return out[:n].copy()
