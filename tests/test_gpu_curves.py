"""GPU parity tests of the epidemic-curves kernel (csrc/curves_kernel.cu, through the C ABI and the engine's
get_data_infectious / get_data_susceptible wrappers).

Oracle: oracle.ref_data_infectious / ref_data_susceptible, the literal restatement of the reference's
src/_BirthDeath.pyx:1967-2045 that tests/test_curves_oracle.py pins against the reference's own output.  Both
sides read the SAME log (the device log exported in the reference's layouts), so every comparison is exact."""
import numpy as np
import pytest

from oracle import oracle as O
from test_gpu_tau import make_engine
import oracle_pool

pytestmark = pytest.mark.gpu
STEPS = 29


def first_infection(Sx0, I0):
    Sx, I = Sx0.copy(), I0.copy()
    if I.sum() == 0:
        I[0, 0] += 1
        Sx[0, np.nonzero(Sx0[0])[0][0]] -= 1
    return Sx, I


def run(name, seed, R, direct_t, tau_leaps=0, tau_t=None):
    e = make_engine(name, seed, replicates=R)
    Sx0, I0 = first_infection(e._susceptible, e._infectious)
    e.SimulatePopulation(10 ** 6, 10 ** 9, direct_t, 200)
    if tau_leaps:
        e.SimulatePopulation_tau(tau_leaps, 10 ** 9, tau_t, 200)
    return e, Sx0, I0


def check_true_counts(e, R):
    """Size-independent properties: the last grid point is the final state, the first the initial one, cumulative
    tallies are monotone and equal the counters, and infectious + susceptible is conserved per deme."""
    c = e.epidemic_curves(STEPS)
    Sx_f, I_f = e._handle.get_state()
    cnt = e.counters()
    assert c["infectious"].shape == (R, STEPS + 1, e.popNum, e.hapNum)
    np.testing.assert_array_equal(c["infectious"][:, -1], I_f)
    np.testing.assert_array_equal(c["susceptible"][:, -1], Sx_f)
    assert c["infectious"].min() >= 0 and c["susceptible"].min() >= 0
    assert np.all(np.diff(c["removed"], axis=1) >= 0) and np.all(np.diff(c["sampled"], axis=1) >= 0)
    np.testing.assert_array_equal(c["sampled"][:, -1].sum(axis=(1, 2)), cnt["sCounter"])
    np.testing.assert_array_equal(c["removed"][:, -1].sum(axis=(1, 2)), cnt["sCounter"] + cnt["dCounter"])
    tot = c["infectious"].sum(axis=3) + c["susceptible"].sum(axis=3)  # [R, T+1, K]: a deme never changes size
    assert np.all(tot == tot[:, :1])
    np.testing.assert_array_equal(c["time_points"][:, -1], (STEPS * cnt["time"]) / STEPS)  # the reference's i*T/n (:1968)
    return c


@pytest.mark.parametrize("name,seed,direct_t,tau_leaps,tau_t", [("s9", 11, 5.0, 0, None), ("s4", 3, 8.0, 0, None),
                                                               ("example", 1234, 45.0, 50, 55.0),
                                                               ("t3small", 5, 50.0, 60, 62.0), ("s5", 7, 3.0, 40, 4.5)])
def test_curves_equal_restated_reference(name, seed, direct_t, tau_leaps, tau_t):
    R = 5
    e, Sx0, I0 = run(name, seed, R, direct_t, tau_leaps, tau_t)
    c = check_true_counts(e, R)
    K, H, S = e.popNum, e.hapNum, e.susNum
    cnt = e.counters()
    for r in (0, R - 1):
        chain = e.get_chain_events(r)
        multi = e.get_multievents(r) if tau_leaps else None
        if tau_leaps:
            assert (chain[1] == 6).sum() == cnt["leaps"][r] > 0
        I_end = c["infectious"][r, -1]
        cells = sorted({(p, h) for p in (0, K - 1) for h in (0, int(np.argmax(I_end.sum(axis=0))), H - 1)})
        for p, h in cells:
            want = O.ref_data_infectious(chain, multi, I0[p, h], float(cnt["time"][r]), p, h, STEPS)
            Data, Sample, tp, ld = e.get_data_infectious(p, h, STEPS, replicate=r)
            np.testing.assert_array_equal(np.asarray(tp), np.asarray(want[2]))
            np.testing.assert_array_equal(Data, want[0])
            np.testing.assert_array_equal(Sample, want[1])
        for p, s in sorted({(p, s) for p in (0, K - 1) for s in (0, S - 1)}):
            want = O.ref_data_susceptible(chain, multi, Sx0[p, s], float(cnt["time"][r]), p, s, STEPS)
            Data, tp, ld = e.get_data_susceptible(p, s, STEPS, replicate=r)
            np.testing.assert_array_equal(Data, want[0])
        st, lp, lt = e._handle.get_lockdowns(r)
        assert e.get_data_infectious(0, 0, STEPS, replicate=r)[3] == [[int(st[i]), float(lt[i])] for i in range(len(st)) if lp[i] == 0]


def test_curves_at_bench_shape():
    """T3 (10 demes x 64 haplotypes x 3 groups) with a tau phase, 64 replicates, replicate sub-range calls."""
    R = 64
    e, Sx0, I0 = run("t3", 21, R, 40.0, 24, 1e9)
    c = check_true_counts(e, R)
    part = e.epidemic_curves(STEPS, rep_first=17, rep_count=9)
    for k in ("infectious", "susceptible", "removed", "sampled", "time_points", "last_point"):
        np.testing.assert_array_equal(part[k], c[k][17:26])
    assert np.all(c["last_point"] == STEPS)


def test_curves_argument_errors():
    from vgsim_b200._capi import VgsimError
    e, _, _ = run("s1", 1, 2, 2.0)
    with pytest.raises(VgsimError):
        e.epidemic_curves(0)
    with pytest.raises(VgsimError):
        e.epidemic_curves(10, rep_first=1, rep_count=2)


def _oracle_curves(chain, I0_total, grid):
    """Total infectious at the grid times and sampling-time statistics from an oracle (== reference algorithm) log."""
    t, ty = chain[0], chain[1].astype(int)
    delta = np.where((ty == 0) | (ty == 5), 1, 0) - np.where((ty == 1) | (ty == 2), 1, 0)
    cum = np.concatenate([[0], np.cumsum(delta)])
    inf = I0_total + cum[np.searchsorted(t, grid, side="right")]
    ts = t[ty == 2]
    return inf, ts


@pytest.mark.parametrize("name,t_end", [("s9", 5.0), ("s4", 8.0), ("s7", 7.0)])
def test_curve_and_sample_time_distributions_match_oracle(name, t_end):
    """BASELINE north_star: 'two-sample KS at alpha = 0.01 on ... epidemic curves, sample times'.  Device replicates
    (Philox, curves from the curves kernel) vs oracle runs (PCG64, curves from the exported log), both stopped at the
    same epidemic time; KS (Bonferroni) on the total infectious count at 8 fixed times, the number of sampled cases,
    and the mean / first sampling time."""
    from scipy import stats
    from test_gpu_tau import _ks_all
    R, RO, T = 1000, 400, 64
    e = make_engine(name, 9100, replicates=R)
    e.SimulatePopulation(10 ** 7, 10 ** 9, t_end, 200)
    c = e.epidemic_curves(T, want=("infectious", "sampled"))
    cnt = e.counters()
    alive = cnt["events"] > 100                      # the reference restarts runs that die within 100 events
    fixed = np.linspace(0.1, 0.95, 8) * t_end
    tp = c["time_points"]                            # [R, T+1]; every replicate's grid ends at its own last event
    tot = c["infectious"].sum(axis=(2, 3))
    smp = c["sampled"].sum(axis=(2, 3))
    dev = {}
    idx = np.stack([np.searchsorted(tp[r], fixed, side="right") - 1 for r in range(R)])   # last grid point <= fixed time
    for j in range(len(fixed)):
        dev["inf_%d" % j] = tot[np.arange(R), idx[:, j]][alive]
    dev["samples"] = cnt["sCounter"][alive]
    keys = list(dev)
    # 400 oracle runs (the configuration this test was verified with in round 1) on the host cores of the box
    ora = oracle_pool.run("curves", name, range(40000, 40000 + RO), epidemic_time=t_end, T=T, fixed=fixed.tolist())
    ora_first, ora_mean, dev_first, dev_mean = ora["first_sample"], ora["mean_sample"], [], []
    assert len(ora["samples"]) > 0.5 * RO
    bad = _ks_all(dev, ora, keys)
    assert not bad, bad
    # sampling times on the device side, from the cumulative `sampled` curve: first grid time with a sample, and the
    # mean sampling time up to the grid resolution (same reduction applied to the oracle's exact times)
    for r in np.flatnonzero(alive):
        if smp[r, -1] == 0:
            continue
        inc = np.diff(np.concatenate([[0], smp[r]]))
        dev_first.append(tp[r][np.flatnonzero(inc)[0]])
        dev_mean.append((inc * tp[r]).sum() / inc.sum())
    step = t_end / T                                  # grid resolution: the device times are rounded UP to a grid point
    p1 = stats.ks_2samp(np.asarray(dev_first) - step / 2, ora_first).pvalue
    p2 = stats.ks_2samp(np.asarray(dev_mean) - step / 2, ora_mean).pvalue
    assert p1 > 0.005 and p2 > 0.005, (p1, p2)
