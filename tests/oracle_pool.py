"""Runs CPU-oracle replicates (TEST INFRASTRUCTURE, oracle/) on all host cores for the distribution tests.

The oracle costs 8-32 ms per reference-sized run, so 5,000-10,000 runs per scenario -- the batch size
BASELINE.json's config 2 asks for -- need the box's cores.  Workers are spawned (never forked from a process that
holds a CUDA context) and import only numpy, the oracle and the host-side engine (parameter arrays, no device).
"""
import multiprocessing as mp
import os

import numpy as np

_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        n = max(1, min(32, (os.cpu_count() or 2) - 1))
        _POOL = mp.get_context("spawn").Pool(n)
    return _POOL


def _engine(name, seed):
    from scenarios import SCENARIOS
    from vgsim_b200._engine import BirthDeathModel as Eng
    (U, K, S), setup = SCENARIOS[name]
    e = Eng(U, K, S, seed, False, False, int(1e6), 0.0)
    setup(e)
    return e


def tree_shape_stats(tree, times):
    """Host restatement of the summary kernel's tree statistics: height, total branch length, cherries, Sackin."""
    n = len(tree)
    nz = tree >= 0
    kids = np.bincount(tree[nz], minlength=n)
    leaf = kids == 0
    leaf_kids = np.bincount(tree[nz & leaf], minlength=n)
    cherries = int((leaf_kids == 2).sum())
    depth = np.zeros(n, np.int64)
    par = tree.tolist()
    d = [0] * n
    for i in range(n - 1, -1, -1):  # parents have larger ids than their children
        p = par[i]
        if p >= 0:
            d[i] = d[p] + 1
    depth[:] = d
    return (times.max() - times.min(), float((times[nz] - times[tree[nz]]).sum()), cherries, int(depth[leaf].sum()))


def _work(task):
    """One chunk of oracle runs; returns a dict of lists (one entry per run that produced a value)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import oracle as O
    kind, name, seeds, kw = task
    out = {}

    def put(k, v):
        out.setdefault(k, []).append(v)

    for seed in seeds:
        e = _engine(name, int(seed))
        if kw.get("state") is not None:
            Sx0, I0 = kw["state"]
            e._susceptible[...] = Sx0
            e._infectious[...] = I0
        om = O.OracleModel.from_engine(e)
        if kind == "curves":
            # total infectious at fixed epidemic times and sampling-time statistics from the exported log, read on the
            # device side's grid (value at a fixed time = the last of T+1 grid points of [0, final time] before it)
            I0_total = int(e._infectious.sum()) or 1          # FirstInfection adds the index case
            om.simulate(10 ** 7, sample_size=10 ** 9, epidemic_time=kw["epidemic_time"])
            chain = om.events()
            if chain.shape[1] <= 100:
                continue
            T, fixed = kw["T"], np.asarray(kw["fixed"])
            ct = om.counters()["time"]
            grid = np.array([i * ct / T for i in range(T + 1)])
            at = grid[np.searchsorted(grid, fixed, side="right") - 1]
            t, ty = chain[0], chain[1].astype(int)
            delta = np.where((ty == 0) | (ty == 5), 1, 0) - np.where((ty == 1) | (ty == 2), 1, 0)
            cum = np.concatenate([[0], np.cumsum(delta)])
            inf = I0_total + cum[np.searchsorted(t, at, side="right")]
            ts = t[ty == 2]
            for j in range(len(fixed)):
                put("inf_%d" % j, int(inf[j]))
            put("samples", len(ts))
            if len(ts):
                put("first_sample", float(ts[0]))
                put("mean_sample", float(ts.mean()))
            continue
        if kind in ("direct", "tree"):
            om.simulate(kw["iterations"], sample_size=kw.get("sample_size"), epidemic_time=kw.get("epidemic_time", -1),
                        attempts=kw.get("attempts", 200))
        elif kind == "tau":
            om.simulate(kw["iterations"], sample_size=10 ** 9, epidemic_time=kw.get("epidemic_time", -1), method="tau",
                        attempts=kw.get("attempts", 1))
        oc = om.counters()
        for k, v in oc.items():
            put(k, v)
        Sx, I = om.get_state()
        put("inf_total", int(I.sum()))
        put("inf_deme0", int(I[0].sum()))
        put("sus_group0", int(Sx[:, 0].sum()))
        if kw.get("lockdowns"):
            st, pop, t = om.lockdowns()
            put("n_lockdowns", len(st))
            put("first_lockdown", float(t[0]) if len(t) else -1.0)
            put("first_lockdown_deme", int(pop[0]) if len(t) else -1)
            put("n_on", int((st == 1).sum()))
        if kind == "tree":
            if oc["sCounter"] < 2:
                put("has_tree", 0)
                continue
            put("has_tree", 1)
            om.genealogy(int(seed) % 100003)
            tree, pop, times = om.tree()
            height, bl, cherries, sackin = tree_shape_stats(tree, times)
            put("samples", (len(tree) + 1) / 2)
            put("height", height)
            put("branch_length", bl)
            put("cherries", cherries)
            put("sackin", sackin)
            put("mutations", len(om.mutations()[0]))
            put("migrations", len(om.migrations()[0]))
            put("root_time", float(times.min()))
    return out


def run(kind, name, seeds, **kw):
    """kind: 'direct' | 'tau' | 'tree' | 'curves'.  Returns {statistic: np.array over runs} gathered from all workers."""
    seeds = list(seeds)
    pool = _pool()
    nchunk = max(1, min(len(seeds), 4 * pool._processes))
    chunks = [seeds[i::nchunk] for i in range(nchunk)]
    res = pool.map(_work, [(kind, name, c, kw) for c in chunks])
    out = {}
    for r in res:
        for k, v in r.items():
            out.setdefault(k, []).extend(v)
    return {k: np.asarray(v) for k, v in out.items()}
