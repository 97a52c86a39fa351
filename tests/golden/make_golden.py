#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference engine (oracle/_ref, built by
oracle/build_ref.py from the sources under /root/reference).  TEST INFRASTRUCTURE.

The reference ships no golden vectors for this path that are usable here: testing/reference_{1..9}.npy
are absent from the mount (.MISSING_LARGE_BLOBS) and tests/test_interface.py pins setters only
(SURVEY.md §8c).  So the fixtures are outputs of the reference itself run in this container:

  direct_<scn>.npz   testing/check_simulator.py scenario <scn>, seed 2020, simulate(100000) direct:
                     first/last HEAD rows of the exported 6xN chain + sha256 of the whole chain,
                     final compartments, then genealogy(seed=7): parent/time arrays (head + sha256),
                     mutation and migration tables (sha256 + counts).
  tau_<scn>.npz      direct warm-up then SimulatePopulation_tau: MULTITYPE row times, final
                     compartments, genealogy over the mixed direct+tau log (pins the restated
                     numpy Poisson / hypergeometric consumption order).
  prop_<scn>.npz     PrintPropensities (src/_BirthDeath.pyx:2615-2649) of a mid-epidemic state, parsed
                     from its repr(float) prints: the P propensities in positional channel order.

  curves_<kind>_<scn>.npz  get_data_infectious / get_data_susceptible (src/_BirthDeath.pyx:1967-2045) of the
                     direct_<scn> / tau_<scn> runs for a few compartments, 37 grid steps.

Run:  python tests/golden/make_golden.py        (needs oracle/_ref; ~1 min)
      python tests/golden/make_golden.py curves (only the curves fixtures)
"""
import contextlib
import hashlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402

HEAD = 400
DIRECT = ["s1", "s2", "s3", "s4", "s5", "s6", "s7", "s8", "s9"]
SEED = 2020
GEN_SEED = 7
# (scenario, forward seed, warm-up end time, tau end time, tau iterations)
# (s9 is absent on purpose: its tau-log genealogy drives the reference into undefined behaviour -- the
# BIRTH branch hands random_hypergeometric a negative `bad` count, src/_BirthDeath.pyx:885 -- and aborts.)
TAU = [("example", 1234, 60.0, 75.0, 400), ("t3small", 5, 50.0, 70.0, 400), ("s5", 2020, 3.0, 5.0, 400),
       ("s1", 2020, 3.0, 5.0, 400), ("s8", 2020, 3.0, 5.0, 400)]
PROP = [("s9", 2020, 4.0), ("example", 1234, 70.0), ("t3small", 5, 60.0), ("t3", 11, 70.0), ("table3_k10", 3, 60.0)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def chain_of(ref):
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "chain")
        ref.export_chain_events(fn)
        return np.load(fn + ".npy")


WARM_ROWS = 2000000


def used_rows(chain):
    """export_chain_events saves the whole ALLOCATION (events.size rows, src/_BirthDeath.pyx:1849-1851);
    rows past events.ptr are zero-filled and every real row has time > 0."""
    return int(np.count_nonzero(chain[0]))


def tree_of(ref):
    tree, times, mut, _pops = _quiet(ref.output_tree_mutations)
    with tempfile.TemporaryDirectory() as d:
        ref.export_migrations("mig", d)
        rows = [l.split("\t") for l in open(os.path.join(d, "mig.tsv")).read().splitlines()[1:]]
    mig = np.array([[float(x) for x in r] for r in rows], dtype=np.float64).reshape(-1, 4)
    mut = np.array(mut, dtype=np.float64).T.reshape(-1, 5)  # nodeId, AS, site, DS, time
    return np.asarray(tree, np.int64).copy(), np.asarray(times, np.float64).copy(), mut, mig


def _quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def make_ref(name, seed):
    (U, K, S), setup = SCENARIOS[name]
    ref = O.make_reference(U, K, S, seed)
    setup(ref)
    return ref


def pack_tree(out, tree, times, mut, mig):
    out.update(tree_n=len(tree), tree_head=tree[:HEAD], tree_tail=tree[-HEAD:], tree_sha=sha(tree),
               times_head=times[:HEAD], times_sha=sha(times), mut_n=len(mut), mut_head=mut[:HEAD], mut_sha=sha(mut),
               mig_n=len(mig), mig_head=mig[:HEAD], mig_sha=sha(mig))


def gen_direct(name):
    ref = make_ref(name, SEED)
    _quiet(ref.SimulatePopulation, 100000, 100000, -1, 200)
    chain = chain_of(ref)
    out = dict(chain_n=chain.shape[1], chain_head=chain[:, :HEAD], chain_tail=chain[:, -HEAD:], chain_sha=sha(chain),
               Sx_end=np.asarray(ref.susceptible).copy(), I_end=np.asarray(ref.infectious).copy())
    _quiet(ref.GetGenealogy, GEN_SEED)
    pack_tree(out, *tree_of(ref))
    np.savez_compressed(os.path.join(HERE, "direct_%s.npz" % name), **out)
    print("direct", name, chain.shape, out["chain_sha"][:16], "tree", out["tree_n"], out["tree_sha"][:16],
          "mut", out["mut_n"], "mig", out["mig_n"])


def gen_tau(name, seed, t_warm, t_end, iters):
    # pass 1 finds how many rows the warm-up to t_warm takes; pass 2 repeats it with iterations == that
    # count, so events.size == events.ptr when the tau call starts and its CreateEvents(iterations) grows
    # the log by exactly `iters` rows: the leap loop then cannot outrun the multiEvents allocation
    # (reference quirk Q3: out-of-bounds writes otherwise).
    ref = make_ref(name, seed)
    _quiet(ref.SimulatePopulation, WARM_ROWS, 10 ** 9, t_warm, 200)
    n_direct = used_rows(chain_of(ref))
    assert 100 < n_direct < WARM_ROWS
    ref = make_ref(name, seed)
    _quiet(ref.SimulatePopulation, n_direct, 10 ** 9, t_warm, 200)
    assert used_rows(chain_of(ref)) == n_direct
    _quiet(ref.SimulatePopulation_tau, iters, 10 ** 9, t_end, 200)
    chain = chain_of(ref)
    chain = chain[:, :used_rows(chain)]
    multi = chain[:, n_direct:]
    assert multi.shape[1] > 0 and np.all(multi[1] == 6), "tau phase produced no leaps"
    out = dict(n_direct=n_direct, direct_sha=sha(chain[:, :n_direct]), leaps=multi.shape[1], leap_times=multi[0].copy(),
               leap_first=multi[2].copy(), leap_last=multi[3].copy(), Sx_end=np.asarray(ref.susceptible).copy(),
               I_end=np.asarray(ref.infectious).copy(), t_warm=t_warm, t_end=t_end, iters=iters, seed=seed)
    _quiet(ref.GetGenealogy, GEN_SEED)
    pack_tree(out, *tree_of(ref))
    np.savez_compressed(os.path.join(HERE, "tau_%s.npz" % name), **out)
    print("tau", name, "direct rows", n_direct, "leaps", out["leaps"], "tree", out["tree_n"], out["tree_sha"][:16],
          "mut", out["mut_n"], "mig", out["mig_n"])


def gen_prop(name, seed, t_warm):
    ref = make_ref(name, seed)
    _quiet(ref.SimulatePopulation, WARM_ROWS, 10 ** 9, t_warm, 200)
    assert used_rows(chain_of(ref)) < WARM_ROWS
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.PrintPropensities()
    vals = []
    for line in buf.getvalue().splitlines():
        parts = line.split()
        if len(parts) >= 2 and parts[0] not in ("Migrations", "Susceptibility"):
            vals.append(float(parts[-1]))
    (U, K, S), _ = SCENARIOS[name]
    H = 4 ** U
    P = K * ((K - 1) * H * S + S * (S - 1) + H * (2 + 3 * U + S))
    assert len(vals) == P, (len(vals), P)
    out = dict(prop=np.array(vals), Sx=np.asarray(ref.susceptible).copy(), I=np.asarray(ref.infectious).copy(),
               cd=np.asarray(ref.contact_density).copy(), seed=seed, t_warm=t_warm)  # cd: live value (lockdowns)
    np.savez_compressed(os.path.join(HERE, "prop_%s.npz" % name), **out)
    print("prop", name, "P", P, "nonzero", int((out["prop"] != 0).sum()))


CURVE_STEPS = 37
# (kind, scenario): the direct run of gen_direct / the mixed run of gen_tau (same seeds and arguments), then
# get_data_infectious / get_data_susceptible of the reference for a handful of compartments
CURVES = [("direct", "s4"), ("direct", "s7"), ("direct", "s9"), ("tau", "example"), ("tau", "t3small"), ("tau", "s5")]


def gen_curves(kind, name):
    (U, K, S), _ = SCENARIOS[name]
    H = 4 ** U
    if kind == "direct":
        ref = make_ref(name, SEED)
        _quiet(ref.SimulatePopulation, 100000, 100000, -1, 200)
    else:
        _, seed, t_warm, t_end, iters = [a for a in TAU if a[0] == name][0]
        ref = make_ref(name, seed)
        _quiet(ref.SimulatePopulation, WARM_ROWS, 10 ** 9, t_warm, 200)
        n_direct = used_rows(chain_of(ref))
        ref = make_ref(name, seed)
        _quiet(ref.SimulatePopulation, n_direct, 10 ** 9, t_warm, 200)
        _quiet(ref.SimulatePopulation_tau, iters, 10 ** 9, t_end, 200)
    I_end = np.asarray(ref.infectious)
    cells = sorted({(p, h) for p in (0, K - 1) for h in (0, int(np.argmax(I_end.sum(axis=0))), H - 1)})
    groups = sorted({(p, s) for p in (0, K - 1) for s in (0, S - 1)})
    out = dict(cells=np.array(cells), groups=np.array(groups), steps=CURVE_STEPS)
    for k, (p, h) in enumerate(cells):
        Data, Sample, tp, _ld = ref.get_data_infectious(p, h, CURVE_STEPS)
        out["inf_%d" % k] = np.asarray(Data, np.float64)
        out["smp_%d" % k] = np.asarray(Sample, np.float64)
        out["tp"] = np.asarray(tp, np.float64)
    for k, (p, s) in enumerate(groups):
        Data, tp, _ld = ref.get_data_susceptible(p, s, CURVE_STEPS)
        out["sus_%d" % k] = np.asarray(Data, np.float64)
    np.savez_compressed(os.path.join(HERE, "curves_%s_%s.npz" % (kind, name)), **out)
    print("curves", kind, name, "cells", cells, "groups", groups, "final Data", [float(out["inf_%d" % k][-1]) for k in range(len(cells))])


def main():
    if not O.reference_available():
        sys.exit("oracle/_ref is not built: run python oracle/build_ref.py (needs /root/reference)")
    if sys.argv[1:] == ["curves"]:
        for args in CURVES:
            gen_curves(*args)
        return
    for name in DIRECT:
        gen_direct(name)
    for args in TAU:
        gen_tau(*args)
    for args in PROP:
        gen_prop(*args)
    for args in CURVES:
        gen_curves(*args)


if __name__ == "__main__":
    main()
