"""Pins the CPU oracle (oracle/vgsim_oracle.cpp) against outputs of the UNMODIFIED reference engine.

tests/golden/*.npz were produced by tests/golden/make_golden.py from the out-of-tree reference build
(oracle/_ref): direct chains of the nine testing/check_simulator.py scenarios, genealogies over them,
mixed direct+tau logs with their genealogies, and PrintPropensities taps.  Integer / index outputs and
event times must match BIT-EXACTLY (sha256 of the raw arrays); propensities to 1e-12 relative
(in practice they are bit-equal too, both builds use -ffp-contract=off).
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GEN_SEED = 7


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def names(prefix):
    return sorted(os.path.basename(f)[len(prefix) + 1:-4] for f in glob.glob(os.path.join(GOLD, prefix + "_*.npz")))


def load(prefix, name):
    return np.load(os.path.join(GOLD, "%s_%s.npz" % (prefix, name)))


def make_oracle(name, seed, live_cd=None):
    """Parameters go through the product's host-side setters (vgsim_b200._engine, no GPU needed), so the
    host logic (wildcards, allele deletion, migration diagonal) is pinned by the same fixtures."""
    (U, K, S), setup = SCENARIOS[name]
    e = Eng(U, K, S, seed, False, False, int(1e6), 0.0)
    setup(e)
    if live_cd is not None:  # contact density as left by the lockdowns of the warm-up (not a setter call)
        e.contactDensity[...] = live_cd
    return O.OracleModel.from_engine(e)


def check_tree(g, om):
    tree, pop, times = om.tree()
    assert len(tree) == int(g["tree_n"])
    assert sha(tree) == str(g["tree_sha"])
    assert sha(times) == str(g["times_sha"])
    node, AS, DS, site, t = om.mutations()
    mut = np.stack([node, AS, site, DS, t], axis=1).astype(np.float64).reshape(-1, 5)
    assert len(mut) == int(g["mut_n"])
    assert sha(mut) == str(g["mut_sha"])
    node, t, oldp, newp = om.migrations()
    mig = np.stack([node, t, oldp, newp], axis=1).astype(np.float64).reshape(-1, 4)
    assert len(mig) == int(g["mig_n"])
    assert sha(mig) == str(g["mig_sha"])


def test_fixtures_present():
    assert len(names("direct")) == 9 and len(names("tau")) >= 4 and len(names("prop")) >= 4


@pytest.mark.parametrize("name", names("direct"))
def test_direct_chain_and_genealogy(name):
    g = load("direct", name)
    om = make_oracle(name, 2020)
    om.simulate(100000)  # sample_size defaults to iterations, like Simulator.simulate (src/_interface.py:816-817)
    chain = om.events()
    assert chain.shape == (6, int(g["chain_n"]))
    np.testing.assert_array_equal(chain[:, :400], g["chain_head"])
    assert sha(chain) == str(g["chain_sha"])
    Sx, I = om.get_state()
    np.testing.assert_array_equal(Sx, g["Sx_end"])
    np.testing.assert_array_equal(I, g["I_end"])
    om.genealogy(GEN_SEED)
    check_tree(g, om)


@pytest.mark.parametrize("name", names("tau"))
def test_tau_log_and_genealogy(name):
    g = load("tau", name)
    om = make_oracle(name, int(g["seed"]))
    n_direct = int(g["n_direct"])
    om.simulate(n_direct, sample_size=10 ** 9, epidemic_time=float(g["t_warm"]))
    assert om.events().shape[1] == n_direct
    assert sha(om.events()) == str(g["direct_sha"])
    om.simulate(int(g["iters"]), sample_size=10 ** 9, epidemic_time=float(g["t_end"]), method="tau")
    chain = om.events()
    multi = chain[:, n_direct:]
    assert multi.shape[1] == int(g["leaps"])
    np.testing.assert_array_equal(multi[0], g["leap_times"])
    np.testing.assert_array_equal(multi[2], g["leap_first"])
    np.testing.assert_array_equal(multi[3], g["leap_last"])
    Sx, I = om.get_state()
    np.testing.assert_array_equal(Sx, g["Sx_end"])
    np.testing.assert_array_equal(I, g["I_end"])
    om.genealogy(GEN_SEED)
    assert om.clamped() == 0
    check_tree(g, om)


@pytest.mark.parametrize("name", names("prop"))
def test_propensities(name):
    g = load("prop", name)
    om = make_oracle(name, int(g["seed"]), live_cd=g["cd"])
    om.set_state(g["Sx"], g["I"])
    prop, dI, dS, tau = om.propensities()
    want = g["prop"]
    assert prop.shape == want.shape
    np.testing.assert_array_equal(prop == 0, want == 0)
    nz = want != 0
    assert (np.abs(prop[nz] - want[nz]) / np.abs(want[nz])).max() < 1e-12


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_direct_matches_oracle():
    """Where the reference build is present, one fresh seed beyond the committed fixtures."""
    import tempfile
    name, seed = "s9", 31337
    (U, K, S), setup = SCENARIOS[name]
    ref = O.make_reference(U, K, S, seed)
    setup(ref)
    with O.quiet():
        ref.SimulatePopulation(30000, 30000, -1, 200)
    with tempfile.TemporaryDirectory() as d:
        ref.export_chain_events(os.path.join(d, "c"))
        want = np.load(os.path.join(d, "c.npy"))
    om = make_oracle(name, seed)
    om.simulate(30000)
    got = om.events()  # the reference exports its whole allocation: rows past events.ptr are zero
    n = got.shape[1]
    assert n > 100 and int(np.count_nonzero(want[0])) == n
    np.testing.assert_array_equal(got, want[:, :n])
