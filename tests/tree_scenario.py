"""One QV-tree scenario replayed on three implementations: the reference's own
SearchTree (oracle/_ref/libpp2d_ref_pomdp_full.so, GPU box only), the CPU
oracle (oracle/pomdp_oracle.c) and the product (pp2d_tree_* through the C
ABI).  Every record is compared bit for bit.

Scenario (exercises tree:161-242, 251-286, 311-366, 397-450, 490-524 and both
branches of SearchTree::update, tree:548-626):
  1. SearchTree(belief)                      -> root bounds
  2. n_expand x expand()                     -> depth, best action, its value
  3. dump
  4. update(best action, an observation that HAS a child)   [re-root, kept]
     plan(50, 4), dump
  5. update(best action, an observation WITHOUT a child)    [new root]
     plan(50, 3), dump
"""
import numpy as np

import cases
import pomdp_fixtures as pf

# name -> (map, goal, n_pbvi, belief seeds, expansions)
CASES = {
    "map_3x3": ("map_3x3", (1, 1), 12, (0, 1), 6),
    "map_10x10": ("map_10x10", (8, 7), 20, (0, 1, 3), 8),
    "sparse_map_100x40": ("sparse_map_100x40", (95, 34), 40, (0, 3), 15),
}


def inputs(case):
    name, goal, n_pbvi, seeds, n_expand = CASES[case]
    grid = cases.load_bundled(name)
    m, fib, pbvi, fa, pa = pf.alphas(name, goal, n_pbvi=n_pbvi)
    beliefs = [pf.gaussian_beliefs(grid, 1, seed=s)[0] for s in seeds]
    if case == "map_3x3":                       # a flat and a peaked belief
        free = (grid.reshape(-1) == 0).astype(np.float32)
        beliefs = [free / free.sum(dtype=np.float32), beliefs[1]]
    return grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand


def checksum(*arrays):
    import zlib
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)


def _root_q_children(dump, action):
    """Observations of the V children of the root's Q node for `action`."""
    assert dump[0, 0] == 0
    i, nq = 1, int(dump[0, 7])
    for _ in range(nq):
        assert dump[i, 0] == 1
        a, nv = int(dump[i, 1]), int(dump[i, 7])
        j = i + 1
        obs = []
        for _ in range(nv):
            obs.append(int(dump[j, 1]))
            j = _skip(dump, j)
        if a == action:
            return obs
        i = j
    return []


def _skip(dump, i):
    """Index just after the subtree rooted at row i (pre-order)."""
    n = int(dump[i, 7])
    i += 1
    for _ in range(n):
        i = _skip(dump, i)
    return i


def run(t, belief, n_expand):
    """t: object with create/expand/depth/best/root_bounds/update/plan/dump."""
    rec = {}
    t.create(belief)
    rec["root_bounds"] = np.array(t.root_bounds(), np.float32)
    steps = []
    for _ in range(n_expand):
        if t.expand() != 0:
            break
        a, r = t.best()
        steps.append((t.depth, a, np.float32(r).view(np.uint32)))
    rec["steps"] = np.array(steps, np.uint32).reshape(-1, 3)
    rec["dump1"] = t.dump()
    a, _ = t.best()
    obs = _root_q_children(rec["dump1"], a)
    rec["update1"] = np.array([a, obs[len(obs) // 2]], np.uint8)
    assert t.update(int(a), int(obs[len(obs) // 2])) == 0
    pa, pr = t.plan(50, 4)
    rec["plan2"] = np.array([pa, np.float32(pr).view(np.uint32), t.depth], np.uint32)
    rec["dump2"] = t.dump()
    a, _ = t.best()
    obs = _root_q_children(rec["dump2"], a)
    missing = [z for z in range(16) if z not in obs]
    if missing and rec["dump2"].shape[0] > 1:
        rec["update2"] = np.array([a, missing[0]], np.uint8)
        assert t.update(int(a), int(missing[0])) == 0
        pa, pr = t.plan(50, 3)
        rec["plan3"] = np.array([pa, np.float32(pr).view(np.uint32), t.depth], np.uint32)
        rec["dump3"] = t.dump()
    return rec


class OracleBackend:
    def __init__(self, m, gamma, fib, pbvi, fa, pa):
        import pomdp_oracle_py as po
        self.po, self.args, self.t = po, (m, gamma, fib, pbvi, pf.uniforms()), None
        self.fa, self.pa = fa, pa

    def create(self, belief):
        if self.t is not None:
            self.t.close()
        m, gamma, fib, pbvi, un = self.args
        self.t = self.po.Tree(m, gamma, fib, pbvi, un, belief, self.fa, self.pa)

    def expand(self): return self.t.expand()
    def best(self): return self.t.best()
    def root_bounds(self): return self.t.root_bounds()
    def update(self, a, z): return self.t.update(a, z)
    def dump(self): return self.t.dump()
    @property
    def depth(self): return self.t.depth

    def plan(self, d, n):
        a, r, _, _ = self.t.plan(d, n)
        return a, r


class ProductBackend:
    def __init__(self, planner):
        self.p, self.t = planner, None

    def create(self, belief):
        from path_planning_2d_b200 import SearchTree
        if self.t is not None:
            self.t.close()
        self.t = SearchTree(self.p, belief)

    def expand(self):
        self.t.expand()
        return 0

    def best(self): return self.t.getOptimalAction()
    def root_bounds(self): return self.t.rootBounds()

    def update(self, a, z):
        self.t.update(a, z)
        return 0

    def dump(self): return self.t.dump()
    @property
    def depth(self): return self.t.getDepth()
    def plan(self, d, n): return self.t.plan(d, n)


class RefBackend:
    def __init__(self, ref):
        self.r = ref

    def create(self, belief): self.r.create(belief, seed=1)
    def expand(self): return self.r.expand()
    def best(self): return self.r.best()
    def root_bounds(self): return self.r.root_bounds()
    def update(self, a, z): return self.r.update(a, z)
    def dump(self): return self.r.dump()
    @property
    def depth(self): return self.r.depth

    def plan(self, d, n):
        a, r, _ = self.r.plan(d, n)
        return a, r


def same_record(got, want):
    """Bit-equal records (NaN == NaN); returns the first differing key or None."""
    for k in want:
        if k not in got:
            return k + " (missing)"
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            return f"{k} (shape {a.shape} != {b.shape})"
        if a.dtype == np.float32:
            ok = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        else:
            ok = a == b
        if not np.all(ok):
            idx = np.argwhere(~ok)[0]
            return f"{k} at {tuple(idx)}: {a[tuple(idx)]!r} != {b[tuple(idx)]!r}"
    return None


def run_reference_case(case, data_dir=None):
    """GPU box: the whole case on the reference stack (one process per case).
    data_dir: the reference first writes its seven text files there with its
    own save*DataToFile and reads them back with its own load*DataFromFile
    (the read_data_from_file=true start-up, src/pomdp/path_planning_2d.cu:
    127-143), so that the scenario runs on the "%15.8f"-rounded tables and
    alpha vectors; they are returned with the records."""
    import pomdp_oracle_py as po
    grid, goal, m, fib, pbvi, fa, pa, beliefs, n_expand = inputs(case)
    ref = po.RefFull(grid, goal, cases.GAMMA, pbvi.shape[0])
    ref.set_alphas(fib, pbvi, fa, pa)
    out = {"inputs_crc": checksum(grid, fib, pbvi, fa, pa, *beliefs)}
    if data_dir is not None:
        ref.save_data(data_dir)
        ref.load_data(data_dir)
        out["trans_prob"], out["meas_prob"], out["stage_reward"] = ref.model()
        out["fib"], out["pbvi"], out["fib_actions"], out["pbvi_actions"] = ref.get_alphas()
        # two belief callbacks as PomdpPathPlanning2d makes them
        # (path_planning_2d.cu:199-241): fresh tree, then update(a0, z = 0)
        be = RefBackend(ref)
        be.create(beliefs[0])
        a0, r0 = be.plan(50, n_expand)
        assert be.update(int(a0), 0) == 0
        a1, r1 = be.plan(50, n_expand)
        out["callbacks"] = np.array([a0, np.float32(r0).view(np.uint32), a1,
                                     np.float32(r1).view(np.uint32)], np.uint32)
    tp, mp, sr = ref.model()
    out["model_crc"] = checksum(tp, mp, sr)
    ev = [ref.evaluate(b) for b in beliefs]
    out["evaluate"] = np.array([[np.float32(e[0]).view(np.uint32), e[1],
                                 np.float32(e[2]).view(np.uint32), e[3]] for e in ev], np.uint32)
    for i, b in enumerate(beliefs):
        for k, v in run(RefBackend(ref), b, n_expand).items():
            out[f"b{i}_{k}"] = v
    ref.R.ref_full_tree_destroy()
    ref.close()
    return out
