"""GPU parity tests of the genealogy kernel: given an event log and a random stream, parent arrays,
node demes/times, mutation and migration tables must be bit-identical to the oracle's (north_star)."""
import numpy as np
import pytest
from numpy.random import PCG64, Generator, SeedSequence

from oracle import oracle as O
from scenarios import SCENARIOS, example_phase2
from vgsim_b200 import _capi
from test_gpu_tau import make_engine, warm_state

pytestmark = pytest.mark.gpu


def pcg(seed, num=0):
    return Generator(PCG64(SeedSequence(seed, spawn_key=(num,))))


def assert_same_genealogy(h, r, om):
    tree, pop, times = h.get_tree(r)
    t2, p2, tm2 = om.tree()
    assert np.array_equal(tree, t2)
    assert np.array_equal(pop, p2)
    assert np.array_equal(times, tm2)
    for a, b in zip(h.get_mutations(r), om.mutations()):
        assert np.array_equal(a, b)
    for a, b in zip(h.get_migrations(r), om.migrations()):
        assert np.array_equal(a, b)


def test_device_hypergeometric_matches_numpy():
    rs = np.random.RandomState(3)
    n = 4000
    good = rs.randint(1, 3000, n)
    bad = rs.randint(1, 3000, n)
    sample = np.array([rs.randint(1, g + b + 1) for g, b in zip(good, bad)])
    small = rs.rand(n) < 0.3
    sample[small] = np.minimum(sample[small], rs.randint(1, 12, small.sum()))
    g = pcg(11, 0)
    state = g.bit_generator.state
    want = np.array([g.hypergeometric(a, b, c) for a, b, c in zip(good, bad, sample)])
    g2 = Generator(PCG64())
    g2.bit_generator.state = state
    words = g2.bit_generator.random_raw(200000)
    got, used = _capi.test_hypergeometric(good, bad, sample, words)
    assert used > 0
    assert np.array_equal(want, got)


@pytest.mark.parametrize("name,iters,gseed", [("s1", 20000, 7), ("s4", 100000, 3), ("s7", 30000, 5), ("s8", 20000, 9),
                                              ("s9", 40000, 2020), ("example", 60000, 1)])
def test_direct_log_genealogy_bit_exact(name, iters, gseed):
    """Oracle forward run -> its event chain is loaded into the device (set_event_log) -> both replay it
    with the SAME uniform stream (the reference's genealogy(seed) stream, PCG64(SeedSequence(seed,(0,))))."""
    om = O.OracleModel.from_engine(make_engine(name, 2020))
    om.simulate(iters)
    chain = om.events()
    Sx_end, I_end = om.get_state()
    assert om.counters()["sCounter"] >= 2
    e = make_engine(name, 2020, replicates=2)
    h = e._sync_params()
    for r in range(2):
        h.set_event_log(r, chain, I_end)
    u = pcg(gseed).random(chain.shape[1] * 4 + 16)
    h.genealogy(uniform_stream=[u, u])
    om.genealogy(gseed)
    for r in range(2):
        assert_same_genealogy(h, r, om)
    # the replay rewinds the infectious counts to the initial state (reference quirk Q9)
    _, I_dev = h.get_state()
    assert np.array_equal(I_dev[0], om.get_state()[1])


@pytest.mark.parametrize("name,seed,t0,leaps", [("t3small", 5, 60.0, 80), ("s9", 2020, 4.0, 90), ("example", 1234, 60.0, 60)])
def test_tau_log_genealogy_bit_exact(name, seed, t0, leaps):
    """Device direct + tau run -> the same log is handed to the oracle -> both replay it with the same raw
    PCG64 word stream (uniforms and numpy-compatible hypergeometric draws)."""
    e = make_engine(name, seed, replicates=3)
    e.SimulatePopulation(10**6, 10**9, t0, 200)
    e.SimulatePopulation_tau(leaps, 10**9, -1, 1)
    h = e._handle
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    words = [pcg(100 + r).bit_generator.random_raw(400000) for r in range(3)]
    chains = [h.get_event_log(r) for r in range(3)]
    logs = [h.get_tau_log(r) for r in range(3)]
    h.genealogy(uniform_stream=words, raw_words=True)
    for r in range(3):
        if c["sCounter"][r] < 2:
            continue
        e1 = make_engine(name, seed)
        om = O.OracleModel.from_engine(e1)
        direct_rows = chains[r][:, chains[r][1] < 6]
        om.set_events(direct_rows)
        om.set_state(Sx_f[r], I_f[r])
        counts, tt = logs[r]
        om.append_tau_log(counts, tt[:, 0])
        assert om.counters()["sCounter"] == c["sCounter"][r]
        # oracle consumes the same PCG64 stream natively
        om.genealogy(100 + r)
        assert_same_genealogy(h, r, om)


def test_philox_genealogy_is_a_valid_tree():
    e = make_engine("s9", 42, replicates=64)
    e.SimulatePopulation(20000, 20000, -1, 200)
    e.GetGenealogy(None)
    h = e._handle
    c = h.get_counters()
    for r in range(0, 64, 7):
        tree, pop, times = h.get_tree(r)
        n = int(c["sCounter"][r])
        assert len(tree) == 2 * n - 1
        roots = np.where(tree == -1)[0]
        assert len(roots) == 1 and roots[0] == 2 * n - 2
        kids = np.bincount(tree[tree >= 0], minlength=len(tree))
        assert set(np.unique(kids)) <= {0, 2} and (kids == 0).sum() == n
        nz = tree >= 0
        assert np.all(times[nz] >= times[tree[nz]])  # children are younger (later) than parents
    s = h.summaries()
    assert np.all(s[:, 16] == 1) and np.all(s[:, 13] == 2 * c["sCounter"] - 1)


def tree_shape_stats(tree, times):
    """Host restatement of the summary kernel's tree statistics (numpy): height, total branch length,
    cherries, Sackin index."""
    n = len(tree)
    nz = tree >= 0
    kids = np.bincount(tree[nz], minlength=n)
    leaf = kids == 0
    leaf_kids = np.bincount(tree[nz & leaf], minlength=n)
    cherries = int((leaf_kids == 2).sum())
    depth = np.zeros(n, np.int64)
    for i in range(n - 1, -1, -1):  # parents have larger ids than their children
        if tree[i] >= 0:
            depth[i] = depth[tree[i]] + 1
    return (times.max() - times.min(), float((times[nz] - times[tree[nz]]).sum()), cherries, int(depth[leaf].sum()))


def test_summary_tree_statistics_match_host():
    e = make_engine("s9", 42, replicates=40)
    e.SimulatePopulation(20000, 20000, -1, 200)
    e.GetGenealogy(None)
    h = e._handle
    s = h.summaries()
    for r in range(0, 40, 3):
        tree, pop, times = h.get_tree(r)
        height, bl, cherries, sackin = tree_shape_stats(tree, times)
        assert s[r, 20] == cherries and s[r, 21] == sackin
        assert s[r, 14] == height
        assert abs(s[r, 15] - bl) <= 1e-9 * bl


@pytest.mark.parametrize("name", ["s1", "s5", "s9"])
def test_tree_statistics_distribution_matches_oracle(name):
    """Config 2 of BASELINE.json on the genealogy side: forward direct run + genealogy per replicate, device
    (Philox) vs oracle == reference algorithm (PCG64); KS at alpha = 0.01 (Bonferroni) on tree statistics."""
    from test_gpu_direct import _ks_all
    R, N = 600, 2500
    e = make_engine(name, 7000, replicates=R)
    e.SimulatePopulation(N, N, -1, 200)
    e.GetGenealogy(None)
    s = e._handle.summaries()
    has_tree = s[:, 13] > 0          # replicates with fewer than two samples have no genealogy
    assert has_tree.mean() > 0.9 and np.all(s[has_tree, 16] == 1)
    s = s[has_tree]
    keys = ["samples", "height", "branch_length", "cherries", "sackin", "mutations", "migrations", "root_time"]
    dev = {"samples": (s[:, 13] + 1) / 2, "height": s[:, 14], "branch_length": s[:, 15], "cherries": s[:, 20],
           "sackin": s[:, 21], "mutations": s[:, 17], "migrations": s[:, 18], "root_time": s[:, 19]}
    ora = {k: [] for k in keys}
    for r in range(R):
        om = O.OracleModel.from_engine(make_engine(name, 3000 + r))
        om.simulate(N)
        if om.counters()["sCounter"] < 2:
            continue
        om.genealogy(r)
        tree, pop, times = om.tree()
        height, bl, cherries, sackin = tree_shape_stats(tree, times)
        ora["samples"].append((len(tree) + 1) / 2)
        ora["height"].append(height)
        ora["branch_length"].append(bl)
        ora["cherries"].append(cherries)
        ora["sackin"].append(sackin)
        ora["mutations"].append(len(om.mutations()[0]))
        ora["migrations"].append(len(om.migrations()[0]))
        ora["root_time"].append(times.min())
    bad = _ks_all(dev, ora, keys)
    assert not bad, bad


def test_chain_export_import_roundtrip(tmp_path):
    """export_chain_events -> set_chain_events on a fresh engine -> genealogy over the imported log with the same
    uniform stream gives the same tree; output_epidemiology_timelines of the exported log equals
    the literal restatement of the reference loop."""
    from test_gpu_tau import make_engine
    a = make_engine("s9", 2020)
    a.SimulatePopulation(20000, 20000, -1, 200)
    fn = str(tmp_path / "chain")
    a.export_chain_events(fn)
    chain = a.get_chain_events(0)
    log = a.output_epidemiology_timelines(15, False)
    want = O.ref_epidemiology_timelines(chain, list(a.sizes), a.popNum, a.susNum, a.hapNum, float(a.counters()["time"][0]), 15)
    assert log["time"] == want[0] and log["P0"]["H0"] == list(want[2][:, 0, 0]) and log["P2"]["S1"] == list(want[1][:, 2, 1])
    u = O.OracleRng(7, 0).doubles(4 * chain.shape[1] + 16)
    a.GetGenealogy(None, uniform_stream=[u])
    tree_a, times_a = a.get_tree(0)
    b = make_engine("s9", 2020)
    b.set_chain_events(fn)
    assert np.array_equal(b.get_chain_events(0), chain)
    b.GetGenealogy(None, uniform_stream=[u])
    tree_b, times_b = b.get_tree(0)
    assert len(tree_a) > 100 and np.array_equal(tree_a, tree_b) and np.array_equal(times_a, times_b)
