"""ctypes access to the TEST-ONLY POMDP oracle (oracle/liboracle_pomdp.so)
and, on a GPU box, to the reference POMDP kernels (oracle/_ref)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_vp, _u32, _i32, _f, _u8 = (ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32,
                            ctypes.c_float, ctypes.c_uint8)
_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_pomdp.so"))
        L.oracle_pomdp_generate_model.argtypes = [_u32, _u32, _i32, _i32, _vp, _vp, _vp, _vp]
        L.oracle_pomdp_bayes_update.argtypes = [_u32, _u32, _vp, _vp, _vp, _u8, _u8, _vp]
        L.oracle_pomdp_normalize.restype = _f
        L.oracle_pomdp_normalize.argtypes = [ctypes.c_uint64, _vp]
        L.oracle_pomdp_evaluate_fib.argtypes = [ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp]
        L.oracle_pomdp_evaluate_pbvi.argtypes = [ctypes.c_uint64, _vp, _vp, _u32, _vp, _vp, _vp]
        L.oracle_pomdp_fib_solve.restype = ctypes.c_int
        L.oracle_pomdp_fib_solve.argtypes = [_u32, _u32, _f, _vp, _vp, _vp, _vp, ctypes.c_int]
        L.oracle_pomdp_blind_policy.argtypes = [_u32, _u32, _f, _vp, _vp, _u8, ctypes.c_int, _vp]
        L.oracle_glibc_rand_fill.argtypes = [_u32, _u32, _vp]
        L.oracle_pomdp_tree_create.restype = _vp
        L.oracle_pomdp_tree_create.argtypes = [_u32, _u32, _f, _vp, _vp, _vp, _vp, _vp,
                                               _vp, _vp, _u32, _vp, _u32, _vp]
        L.oracle_pomdp_tree_destroy.argtypes = [_vp]
        L.oracle_pomdp_tree_depth.restype = _u32
        L.oracle_pomdp_tree_depth.argtypes = [_vp]
        L.oracle_pomdp_tree_expand.restype = ctypes.c_int
        L.oracle_pomdp_tree_expand.argtypes = [_vp]
        L.oracle_pomdp_tree_best_action.argtypes = [_vp, _vp, _vp]
        L.oracle_pomdp_tree_update.restype = ctypes.c_int
        L.oracle_pomdp_tree_update.argtypes = [_vp, _u8, _u8]
        L.oracle_pomdp_plan.restype = ctypes.c_int
        L.oracle_pomdp_plan.argtypes = [_vp, _u32, _u32, _vp, _vp, _vp]
        L.oracle_pomdp_tree_root_bounds.argtypes = [_vp, _vp, _vp]
        L.oracle_pomdp_tree_root_q.restype = ctypes.c_int
        L.oracle_pomdp_tree_root_q.argtypes = [_vp, _vp, _vp, _vp]
        L.oracle_pomdp_forward_sampling.restype = ctypes.c_int
        L.oracle_pomdp_tree_dump.restype = ctypes.c_int64
        L.oracle_pomdp_tree_dump.argtypes = [_vp, _vp, ctypes.c_uint64]
        _lib = L
    return _lib


class Model:
    """B1 tables of the oracle for one map."""

    def __init__(self, grid, goal):
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.h, self.w = self.grid.shape
        self.hw = self.h * self.w
        self.goal = goal
        self.tp = np.zeros(self.hw * 81, np.float32)
        self.mp = np.zeros(self.hw * 16, np.float32)
        self.sr = np.zeros(self.hw * 9, np.float32)
        lib().oracle_pomdp_generate_model(self.h, self.w, goal[0], goal[1],
                                          self.grid.ctypes.data, self.tp.ctypes.data,
                                          self.mp.ctypes.data, self.sr.ctypes.data)

    def bayes(self, belief, u, z, normalize=False):
        b = np.ascontiguousarray(belief, dtype=np.float32).reshape(-1)
        out = np.zeros(self.hw, np.float32)
        lib().oracle_pomdp_bayes_update(self.h, self.w, self.tp.ctypes.data,
                                        self.mp.ctypes.data, b.ctypes.data, u, z,
                                        out.ctypes.data)
        s = None
        if normalize:
            s = lib().oracle_pomdp_normalize(self.hw, out.ctypes.data)
        return out, s

    def fib(self, gamma, max_sweeps=0):
        al = np.zeros(self.hw * 9, np.float32)
        n = lib().oracle_pomdp_fib_solve(self.h, self.w, gamma, self.tp.ctypes.data,
                                         self.mp.ctypes.data, self.sr.ctypes.data,
                                         al.ctypes.data, max_sweeps)
        return al.reshape(self.hw, 9), n

    def blind(self, gamma, a, sweeps=200):
        out = np.zeros(self.hw, np.float32)
        lib().oracle_pomdp_blind_policy(self.h, self.w, gamma, self.tp.ctypes.data,
                                        self.sr.ctypes.data, a, sweeps, out.ctypes.data)
        return out


def evaluate(belief, fib_alphas, pbvi_alphas, fib_actions=None, pbvi_actions=None):
    b = np.ascontiguousarray(belief, dtype=np.float32).reshape(-1)
    fa = np.ascontiguousarray(fib_alphas, dtype=np.float32)
    pa = np.ascontiguousarray(pbvi_alphas, dtype=np.float32)
    up, lo = _f(), _f()
    ua, la = _u8(), _u8()
    fac = fib_actions.ctypes.data if fib_actions is not None else None
    pac = pbvi_actions.ctypes.data if pbvi_actions is not None else None
    lib().oracle_pomdp_evaluate_fib(b.size, b.ctypes.data, fa.ctypes.data, fac,
                                    ctypes.addressof(up), ctypes.addressof(ua))
    lib().oracle_pomdp_evaluate_pbvi(b.size, b.ctypes.data, pa.ctypes.data, pa.shape[0],
                                     pac, ctypes.addressof(lo), ctypes.addressof(la))
    return up.value, ua.value, lo.value, la.value


class Tree:
    def __init__(self, model, gamma, fib_alphas, pbvi_alphas, uniforms, belief,
                 fib_actions=None, pbvi_actions=None, seed=1):
        self.keep = (model, np.ascontiguousarray(fib_alphas, np.float32),
                     np.ascontiguousarray(pbvi_alphas, np.float32),
                     np.ascontiguousarray(uniforms, np.float32),
                     np.ascontiguousarray(belief, np.float32).reshape(-1),
                     fib_actions, pbvi_actions)
        m, fa, pa, un, b, fac, pac = self.keep
        self.t = lib().oracle_pomdp_tree_create(
            m.h, m.w, gamma, m.tp.ctypes.data, m.mp.ctypes.data, m.sr.ctypes.data,
            fa.ctypes.data, fac.ctypes.data if fac is not None else None,
            pa.ctypes.data, pac.ctypes.data if pac is not None else None,
            pa.shape[0], un.ctypes.data, seed, b.ctypes.data)

    def close(self):
        if self.t:
            lib().oracle_pomdp_tree_destroy(self.t)
            self.t = None

    def __del__(self):
        self.close()

    def plan(self, max_depth=50, max_iter=15):
        a, r = _u8(), _f()
        stats = np.zeros(5, np.uint64)
        rc = lib().oracle_pomdp_plan(self.t, max_depth, max_iter, ctypes.addressof(a),
                                     ctypes.addressof(r), stats.ctypes.data)
        return a.value, r.value, stats, rc

    def expand(self):
        return lib().oracle_pomdp_tree_expand(self.t)

    def update(self, a, z):
        return lib().oracle_pomdp_tree_update(self.t, a, z)

    def best(self):
        a, r = _u8(), _f()
        lib().oracle_pomdp_tree_best_action(self.t, ctypes.addressof(a), ctypes.addressof(r))
        return a.value, r.value

    @property
    def depth(self):
        return lib().oracle_pomdp_tree_depth(self.t)

    def root_bounds(self):
        u, l = _f(), _f()
        lib().oracle_pomdp_tree_root_bounds(self.t, ctypes.addressof(u), ctypes.addressof(l))
        return u.value, l.value

    def dump(self):
        n = lib().oracle_pomdp_tree_dump(self.t, None, 0)
        out = np.zeros((n, 9), np.float32)
        lib().oracle_pomdp_tree_dump(self.t, out.ctypes.data, n)
        return out

    def root_q(self):
        u = np.zeros(9, np.float32); l = np.zeros(9, np.float32); r = np.zeros(9, np.float32)
        n = lib().oracle_pomdp_tree_root_q(self.t, u.ctypes.data, l.ctypes.data, r.ctypes.data)
        return u[:n], l[:n], r[:n]


def glibc_rand(seed, n):
    out = np.zeros(n, np.uint32)
    lib().oracle_glibc_rand_fill(seed, n, out.ctypes.data)
    return out


# ---- reference POMDP kernels (GPU box only) --------------------------------
def ref():
    global _ref
    if _ref is None:
        R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpp2d_ref_pomdp.so"))
        R.ref_pomdp_model.argtypes = [_u32, _u32, _vp, _i32, _i32, _vp, _vp, _vp]
        R.ref_pomdp_bayes.argtypes = [_u32, _u32, _vp, _i32, _i32, _vp, _u32, _vp, _vp, _vp]
        R.ref_pomdp_fib.restype = ctypes.c_int
        R.ref_pomdp_fib.argtypes = [_u32, _u32, _vp, _i32, _i32, _f, _vp, ctypes.c_int]
        R.ref_pomdp_uniforms.argtypes = [_u32, _vp]
        R.ref_pomdp_forward_sampling.argtypes = [_u32, _u32, _vp, _i32, _i32, _u32, _vp, _u8, _vp]
        _ref = R
    return _ref


class RefFull:
    """The reference's complete POMDP stack (oracle/_ref/libpp2d_ref_pomdp_full.so:
    the four unmodified translation units + oracle/ref_pomdp_full_driver.cu).
    The reference keeps its state in process globals: ONE instance per process
    (tests run it in a subprocess).  GPU box only."""

    def __init__(self, grid, goal, gamma, n_pbvi):
        R = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpp2d_ref_pomdp_full.so"))
        R.ref_full_init.argtypes = [_u32, _u32, _vp, _i32, _i32, _f, _u32]
        R.ref_full_model.argtypes = [_vp, _vp, _vp]
        R.ref_full_set_alphas.argtypes = [_vp, _vp, _vp, _vp]
        R.ref_full_solve_fib.argtypes = [_vp, _vp]
        R.ref_full_solve_pbvi.argtypes = [_vp, _u32, _vp, _vp, _vp]
        R.ref_full_belief_set.argtypes = [_vp, _u32, _vp]
        R.ref_full_backup.argtypes = [_vp, _vp, _vp]
        R.ref_full_evaluate.argtypes = [_vp, _vp, _vp, _vp, _vp]
        R.ref_full_tree_create.argtypes = [_vp, _u32]
        R.ref_full_tree_depth.restype = _u32
        R.ref_full_tree_best_action.argtypes = [_vp, _vp]
        R.ref_full_tree_update.argtypes = [_u8, _u8]
        R.ref_full_tree_root_bounds.argtypes = [_vp, _vp]
        R.ref_full_tree_plan.argtypes = [_u32, _u32, _vp, _vp, _vp]
        R.ref_full_tree_dump.restype = ctypes.c_int64
        R.ref_full_tree_dump.argtypes = [_vp, ctypes.c_uint64]
        R.ref_full_tree_root_belief.argtypes = [_vp]
        R.ref_full_save_data.argtypes = [ctypes.c_char_p]
        R.ref_full_load_data.argtypes = [ctypes.c_char_p]
        R.ref_full_get_alphas.argtypes = [_vp, _vp, _vp, _vp]
        self.R = R
        self.grid = np.ascontiguousarray(grid, np.uint8)
        self.h, self.w = self.grid.shape
        self.hw = self.h * self.w
        self.n_pbvi = n_pbvi
        rc = R.ref_full_init(self.h, self.w, self.grid.ctypes.data, goal[0], goal[1],
                             gamma, n_pbvi)
        assert rc == 0, "the reference planner is one-per-process"

    def close(self):
        self.R.ref_full_shutdown()

    def model(self):
        tp = np.zeros((self.hw, 9, 9), np.float32)
        mp = np.zeros((self.hw, 16), np.float32)
        sr = np.zeros((self.hw, 9), np.float32)
        self.R.ref_full_model(tp.ctypes.data, mp.ctypes.data, sr.ctypes.data)
        return tp, mp, sr

    def set_alphas(self, fib, pbvi, fa=None, pa=None):
        fib = np.ascontiguousarray(fib, np.float32)
        pbvi = np.ascontiguousarray(pbvi, np.float32)
        assert fib.shape == (self.hw, 9) and pbvi.shape == (self.n_pbvi, self.hw)
        fa = None if fa is None else np.ascontiguousarray(fa, np.uint8)
        pa = None if pa is None else np.ascontiguousarray(pa, np.uint8)
        self.R.ref_full_set_alphas(fib.ctypes.data, fa.ctypes.data if fa is not None else None,
                                   pbvi.ctypes.data, pa.ctypes.data if pa is not None else None)

    def save_data(self, directory):
        """saveDataCallback: the reference's seven text files into `directory`."""
        assert self.R.ref_full_save_data(str(directory).encode()) == 0

    def load_data(self, directory):
        """read_data_from_file=true: the reference's own loaders."""
        assert self.R.ref_full_load_data(str(directory).encode()) == 0

    def get_alphas(self):
        fib = np.zeros((self.hw, 9), np.float32)
        pbvi = np.zeros((self.n_pbvi, self.hw), np.float32)
        fa, pa = np.zeros(9, np.uint8), np.zeros(self.n_pbvi, np.uint8)
        self.R.ref_full_get_alphas(fib.ctypes.data, fa.ctypes.data, pbvi.ctypes.data,
                                   pa.ctypes.data)
        return fib, pbvi, fa, pa

    def solve_fib(self):
        al = np.zeros((self.hw, 9), np.float32)
        ac = np.zeros(9, np.uint8)
        self.R.ref_full_solve_fib(al.ctypes.data, ac.ctypes.data)
        return al, ac

    def belief_set(self, b0, seed=1):
        b0 = np.ascontiguousarray(b0, np.float32).reshape(-1)
        out = np.zeros((self.n_pbvi, self.hw), np.float32)
        self.R.ref_full_belief_set(b0.ctypes.data, seed, out.ctypes.data)
        return out

    def backup(self, belief_set):
        bs = np.ascontiguousarray(belief_set, np.float32)
        al = np.zeros((self.n_pbvi, self.hw), np.float32)
        ac = np.zeros(self.n_pbvi, np.uint8)
        self.R.ref_full_backup(bs.ctypes.data, al.ctypes.data, ac.ctypes.data)
        return al, ac

    def solve_pbvi(self, b0, seed=1):
        b0 = np.ascontiguousarray(b0, np.float32).reshape(-1)
        bs = np.zeros((self.n_pbvi, self.hw), np.float32)
        al = np.zeros((self.n_pbvi, self.hw), np.float32)
        ac = np.zeros(self.n_pbvi, np.uint8)
        self.R.ref_full_solve_pbvi(b0.ctypes.data, seed, bs.ctypes.data, al.ctypes.data,
                                   ac.ctypes.data)
        return bs, al, ac

    def evaluate(self, belief):
        b = np.ascontiguousarray(belief, np.float32).reshape(-1)
        up, lo = _f(), _f()
        ua, la = _u8(), _u8()
        self.R.ref_full_evaluate(b.ctypes.data, ctypes.addressof(up), ctypes.addressof(ua),
                                 ctypes.addressof(lo), ctypes.addressof(la))
        return up.value, ua.value, lo.value, la.value

    # SearchTree -- same method names as Tree above
    def create(self, belief, seed=1):
        b = np.ascontiguousarray(belief, np.float32).reshape(-1)
        self.R.ref_full_tree_create(b.ctypes.data, seed)

    def expand(self):
        return self.R.ref_full_tree_expand()

    def update(self, a, z):
        return self.R.ref_full_tree_update(a, z)

    @property
    def depth(self):
        return self.R.ref_full_tree_depth()

    def best(self):
        a, r = _u8(), _f()
        self.R.ref_full_tree_best_action(ctypes.addressof(a), ctypes.addressof(r))
        return a.value, r.value

    def root_bounds(self):
        u, l = _f(), _f()
        self.R.ref_full_tree_root_bounds(ctypes.addressof(u), ctypes.addressof(l))
        return u.value, l.value

    def plan(self, max_depth=50, max_iter=15):
        a, r = _u8(), _f()
        stats = np.zeros(4, np.uint64)
        self.R.ref_full_tree_plan(max_depth, max_iter, ctypes.addressof(a),
                                  ctypes.addressof(r), stats.ctypes.data)
        return a.value, r.value, stats

    def dump(self):
        n = self.R.ref_full_tree_dump(None, 0)
        out = np.zeros((n, 9), np.float32)
        self.R.ref_full_tree_dump(out.ctypes.data, n)
        return out
