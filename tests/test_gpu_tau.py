"""GPU parity tests of the tau-leap path (through the C ABI) against the CPU oracle.

Deterministic pieces (propensities, tau) must agree to 1e-12 relative (north_star); stochastic output
is compared distributionally (two-sample KS, alpha = 0.01 Bonferroni-corrected) because the device
draws from Philox while the reference/oracle draws from a sequential PCG64 stream.
"""
import numpy as np
import pytest
from scipy import stats

from oracle import oracle as O
from scenarios import SCENARIOS
from vgsim_b200 import _capi
from vgsim_b200._engine import BirthDeathModel as Eng

pytestmark = pytest.mark.gpu


def make_engine(name, seed=1, replicates=1):
    (U, K, S), setup = SCENARIOS[name]
    e = Eng(U, K, S, seed, False, False, int(1e6), 0.0, replicates=replicates)
    setup(e)
    return e


def warm_state(name, seed, t_end):
    """A mid-epidemic state produced by the oracle's direct method (deterministic given the seed)."""
    e = make_engine(name, seed)
    om = O.OracleModel.from_engine(e)
    om.simulate(10**7, sample_size=10**9, epidemic_time=t_end)
    return om.get_state()


@pytest.mark.parametrize("name,seed,t_end", [("s9", 2020, 4.0), ("example", 1234, 70.0), ("t3small", 5, 60.0),
                                             ("t3", 11, 70.0), ("table3_k10", 3, 60.0)])
def test_propensities_match_oracle(name, seed, t_end):
    Sx, I = warm_state(name, seed, t_end)
    assert I.sum() > 0
    e = make_engine(name, seed)
    e._susceptible[...] = Sx
    e._infectious[...] = I
    prop, dI, dS, tau = e.propensities()
    om = O.OracleModel.from_engine(e)
    p2, dI2, dS2, tau2 = om.propensities()
    assert prop.shape == p2.shape
    nz = p2 != 0
    assert np.array_equal(prop == 0, p2 == 0)
    rel = np.abs(prop[nz] - p2[nz]) / np.abs(p2[nz])
    assert rel.max() < 1e-12, rel.max()
    # drifts are signed sums with cancellation: compare against the magnitude of what is summed
    scale = max(np.abs(p2).sum(), 1.0)
    assert np.abs(dI - dI2).max() / scale < 1e-13
    assert np.abs(dS - dS2).max() / scale < 1e-13
    # tau = max(1, eps*count/2) / |drift| of the binding compartment (ChooseTau, :2432-2450), so its relative error IS the
    # relative error of that one drift -- a signed sum whose error is bounded against the magnitude of its terms, not
    # against its own (possibly cancelled) value.  Hence two exact statements instead of one loose tolerance:
    #   (a) the device applies the reference's formula: on the device's own drifts the host restatement gives the device's
    #       tau to the last bits;  (b) against the oracle, tau moves by no more than the drift error allows.
    tau_h, _ = O.choose_tau(dI, dS, I, Sx)
    assert abs(tau - tau_h) <= 4e-16 * tau_h, (tau, tau_h)
    tau_o, bind = O.choose_tau(dI2, dS2, I, Sx)
    assert abs(tau_o - tau2) <= 4e-16 * tau2, (tau_o, tau2)            # the restatement is the oracle's formula
    derr = max(np.abs(dI - dI2).max(), np.abs(dS - dS2).max())
    assert abs(tau - tau2) / tau2 <= 1e-12 + (4.0 * derr / bind if bind > 0 else 0.0), (tau, tau2, derr, bind)


def test_default_model_propensities():
    # sites = 0 (one haplotype), one deme, one group: P = 3 channels, unaligned log row
    e = Eng(0, 1, 1, 7, False, False, int(1e6), 0.0)
    e._infectious[0, 0] = 1234
    e._susceptible[0, 0] -= 1234
    prop, dI, dS, tau = e.propensities()
    p2, dI2, dS2, tau2 = O.OracleModel.from_engine(e).propensities()
    assert prop.shape == (3,)
    np.testing.assert_allclose(prop, p2, rtol=1e-13)
    np.testing.assert_allclose(tau, tau2, rtol=1e-12)


def test_device_poisson_sampler():
    """chi-square of the device sampler (inversion / PTRS over Philox) against the Poisson pmf."""
    n = 200000
    for lam in [1e-7, 1e-3, 0.3, 0.999, 1.0, 3.7, 9.99, 10.0, 14.2, 87.5, 1234.5, 2.5e5]:
        x = _capi.test_poisson(np.full(n, lam), seed=int(lam * 1000) + 17)
        assert x.min() >= 0
        if lam < 1e-5:
            # P(n >= 1) = lam: the count of non-zeros is Binomial(n, ~lam)
            assert (x > 0).sum() <= 5
            continue
        lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam))
        if hi - lo > 60:  # coarse bins for large lambda
            edges = np.unique(stats.poisson.ppf(np.linspace(0, 1, 41)[1:-1], lam).astype(int))
        else:
            edges = np.arange(lo, hi + 1)
        cdf = stats.poisson.cdf(edges, lam)
        probs = np.diff(np.concatenate([[0.0], cdf, [1.0]]))
        obs = np.bincount(np.searchsorted(edges, x, side="left"), minlength=len(probs)).astype(float)
        keep = probs * n >= 5
        obs_k = np.append(obs[keep], obs[~keep].sum())
        exp_k = np.append(probs[keep] * n, probs[~keep].sum() * n)
        if exp_k[-1] < 5:
            obs_k[-2] += obs_k[-1]
            exp_k[-2] += exp_k[-1]
            obs_k, exp_k = obs_k[:-1], exp_k[:-1]
        chi2 = ((obs_k - exp_k) ** 2 / exp_k).sum()
        p = stats.chi2.sf(chi2, len(exp_k) - 1)
        assert p > 1e-4, (lam, chi2, p)
        assert abs(x.mean() - lam) < 6 * np.sqrt(lam / n) + 1e-12, (lam, x.mean())


def _replay_dense_log(e, h, r, Sx0, I0):
    """Apply replicate r's dense tau log to (Sx0, I0) on the host; returns the final state and per-type totals."""
    counts, tt = h.get_tau_log(r)
    me = h.get_multievents(r)
    K, H, S = e.popNum, e.hapNum, e.susNum
    Sx, I = Sx0.copy(), I0.copy()
    num, typ, hap, pop, nhap, npop = (me[k] for k in ("num", "type", "hap", "pop", "nhap", "npop"))
    assert num.sum() == counts.sum()
    tot = {t: int(num[typ == t].sum()) for t in range(6)}
    for t, sgnI, sgnS in ((0, +1, -1),):  # BIRTH: I[pop,hap]++, Sx[pop,nhap]--
        sel = typ == t
        np.add.at(I, (pop[sel], hap[sel]), num[sel])
        np.add.at(Sx, (pop[sel], nhap[sel]), -num[sel])
    for t in (1, 2):  # DEATH / SAMPLING: I[pop,hap]--, Sx[pop,nhap]++
        sel = typ == t
        np.add.at(I, (pop[sel], hap[sel]), -num[sel])
        np.add.at(Sx, (pop[sel], nhap[sel]), num[sel])
    sel = typ == 3  # MUTATION: I[pop,hap]--, I[pop,nhap]++
    np.add.at(I, (pop[sel], hap[sel]), -num[sel])
    np.add.at(I, (pop[sel], nhap[sel]), num[sel])
    sel = typ == 4  # SUSCCHANGE: Sx[pop,hap(src group)]--, Sx[pop,nhap]++
    np.add.at(Sx, (pop[sel], hap[sel]), -num[sel])
    np.add.at(Sx, (pop[sel], nhap[sel]), num[sel])
    sel = typ == 5  # MIGRATION: I[npop,hap]++, Sx[npop,nhap(group)]--
    np.add.at(I, (npop[sel], hap[sel]), num[sel])
    np.add.at(Sx, (npop[sel], nhap[sel]), -num[sel])
    return Sx, I, tot, tt


@pytest.mark.parametrize("name,seed,t0,t1", [("t3small", 5, 60.0, 75.0), ("s9", 2020, 4.0, 6.0)])
def test_tau_log_replays_to_final_state(name, seed, t0, t1):
    """Size-independent property: the dense log is a complete record — replaying it from the initial
    state gives exactly the device's final state, and its per-type sums are the counters."""
    Sx0, I0 = warm_state(name, seed, t0)
    R = 8
    e = make_engine(name, seed, replicates=R)
    e._susceptible[...] = Sx0
    e._infectious[...] = I0
    h = e._sync_params()
    h.simulate_tau(100, -1, t1 - t0, 1)  # iterations <= 100 keeps the extinction-restart rule out of the way
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    for r in range(R):
        assert c["leaps"][r] > 3
        Sx, I, tot, tt = _replay_dense_log(e, h, r, Sx0, I0)
        assert np.array_equal(Sx, Sx_f[r]) and np.array_equal(I, I_f[r])
        assert tot[0] == c["bCounter"][r] and tot[1] == c["dCounter"][r] and tot[2] == c["sCounter"][r]
        assert tot[3] == c["mCounter"][r] and tot[4] == c["iCounter"][r] and tot[5] == c["migPlus"][r]
        assert np.all(np.diff(tt[:, 0]) > 0) and np.all(tt[:, 1] > 0) and np.all(tt[:, 1] <= 1.0)
        assert abs(tt[-1, 0] - c["time"][r]) == 0
        ev = h.get_event_log(r)
        assert ev.shape[1] == c["events"][r] and np.all(ev[1] == 6)
        assert np.array_equal(ev[0], tt[:, 0])
        # population is conserved
        assert Sx.sum() + I.sum() == Sx0.sum() + I0.sum()
    # replicates use different streams
    assert len({int(x) for x in c["bCounter"]}) > 1


def _ks_all(dev, ora, names, alpha=0.01):
    bad = []
    for k in names:
        a, b = np.asarray(dev[k], float), np.asarray(ora[k], float)
        if a.std() == 0 and b.std() == 0 and a[0] == b[0]:
            continue
        p = stats.ks_2samp(a, b).pvalue
        if p < alpha / len(names):
            bad.append((k, p, a.mean(), b.mean()))
    return bad


@pytest.mark.parametrize("name,seed,t0,dt,R", [("t3small", 5, 60.0, 12.0, 1500), ("s9", 2020, 4.0, 1.5, 1500)])
def test_tau_distribution_matches_oracle(name, seed, t0, dt, R):
    """Same start state, same stop rule: R device replicates (Philox) vs R oracle runs (PCG64, seeds r)."""
    Sx0, I0 = warm_state(name, seed, t0)
    e = make_engine(name, seed, replicates=R)
    e._susceptible[...] = Sx0
    e._infectious[...] = I0
    h = e._sync_params()
    h.simulate_tau(100, -1, dt, 1)
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    keys = ["bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "leaps", "time", "inf_total", "inf_deme0"]
    dev = {k: c[k] for k in keys if k in c}
    dev["inf_total"] = I_f.sum(axis=(1, 2))
    dev["inf_deme0"] = I_f[:, 0, :].sum(axis=1)
    ora = {k: [] for k in keys}
    for r in range(R):
        e1 = make_engine(name, 1000 + r)
        e1._susceptible[...] = Sx0
        e1._infectious[...] = I0
        om = O.OracleModel.from_engine(e1)
        # a FRESH tau log in the reference gets events.size = 2 * iterations (CreateEvents is called twice,
        # SURVEY quirk Q3); the device gives `iterations` leaps of capacity as documented -> 50 here == 100 there
        om.simulate(50, sample_size=10**9, epidemic_time=dt, method="tau", attempts=1)
        oc = om.counters()
        _, If = om.get_state()
        for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "time"):
            ora[k].append(oc[k])
        ora["leaps"].append(oc["events"])
        ora["inf_total"].append(If.sum())
        ora["inf_deme0"].append(If[0].sum())
    bad = _ks_all(dev, ora, keys)
    assert not bad, bad


def _first_leap_counts(name, seed, t0, R, variant, seed0):
    """R replicates of ONE leap from the same mid-epidemic state: counts[R, P], tau, propensities."""
    Sx0, I0 = warm_state(name, seed, t0)
    e = make_engine(name, seed0, replicates=R)
    e._susceptible[...] = Sx0
    e._infectious[...] = I0
    prop, _dI, _dS, tau = e.propensities()
    h = e._sync_params()
    h.set_tau_variant(variant)
    h.simulate_tau(1, -1, -1.0, 1)
    c = h.get_counters()
    assert np.all(c["leaps"] == 1)
    counts = np.stack([h.get_tau_log(r)[0][0] for r in range(R)])
    tt = h.get_tau_log(0)[1]
    return counts, float(tt[0, 1]), prop, tau, c


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name,seed,t0", [("t3small", 5, 75.0), ("s9", 2020, 5.0), ("t3", 11, 80.0)])
def test_first_leap_counts_are_independent_poisson_per_channel(name, seed, t0, variant):
    """The reference draws an independent Poisson(prop_c * tau) for every channel (src/_BirthDeath.pyx:2531-2532).
    Variant 1 of the kernel does exactly that (one draw per channel); variant 0 (product) draws ONE Poisson for
    the total of a cell's mutation channels / out-migration channels when that total is small and splits it
    multinomially, which has the same joint distribution.  Check both against theory on R replicates of one
    leap from the same state: per-channel means (chi-square against R*lambda_c), per-channel dispersion
    (variance/mean = 1), and zero counts wherever the propensity is zero."""
    R = 3000
    counts, tau_used, prop, tau, c = _first_leap_counts(name, seed, t0, R, variant, seed0=90 + variant)
    assert tau_used == tau  # no halving happened in replicate 0 (the state is far from the bounds)
    lam = prop * tau
    assert counts.shape == (R, len(prop)) and counts.min() >= 0
    assert np.all(counts[:, lam == 0] == 0)
    O_c = counts.sum(axis=0).astype(float)
    E_c = R * lam
    big = E_c >= 8
    assert big.sum() > 5
    chi2 = ((O_c[big] - E_c[big]) ** 2 / E_c[big]).sum()
    df = int(big.sum())
    small = (~big) & (lam > 0)
    if E_c[small].sum() >= 8:
        chi2 += (O_c[small].sum() - E_c[small].sum()) ** 2 / E_c[small].sum()
        df += 1
    pval = stats.chi2.sf(chi2, df)
    assert pval > 1e-4, (chi2, df, pval)
    # dispersion: var/mean = 1 for a Poisson sample; Var(s^2/m) = (2 + 1/m)/R (Poisson fourth moment m + 3m^2)
    sel = np.where(E_c >= 50)[0]
    m = counts[:, sel].mean(axis=0)
    v = counts[:, sel].var(axis=0, ddof=1)
    z = (v / m - 1.0) / np.sqrt((2.0 + 1.0 / lam[sel]) / R)
    assert np.abs(z).max() < 5.0, (z.min(), z.max())
    # independence between a cell's aggregated channels and its other channels: correlation of the
    # per-replicate type totals is ~0 (mutation vs recovery counts)
    if c["mCounter"].std() > 0:
        rho = np.corrcoef(c["mCounter"], c["dCounter"])[0, 1]
        assert abs(rho) < 5.0 / np.sqrt(R), rho


@pytest.mark.parametrize("name,seed,t0,dt", [("t3small", 5, 60.0, 12.0), ("s7", 2020, 6.0, 3.0)])
def test_aggregated_and_per_channel_variants_agree(name, seed, t0, dt):
    """Multi-leap runs of the two variants from the same state: KS on counters, time, leaps, infectious totals."""
    Sx0, I0 = warm_state(name, seed, t0)
    R = 2000
    res = []
    for variant in (0, 1):
        e = make_engine(name, 300 + variant, replicates=R)
        e._susceptible[...] = Sx0
        e._infectious[...] = I0
        h = e._sync_params()
        h.set_tau_variant(variant)
        h.simulate_tau(100, -1, dt, 1)
        c = h.get_counters()
        _, I_f = h.get_state()
        d = {k: c[k] for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "leaps", "time")}
        d["inf_total"] = I_f.sum(axis=(1, 2))
        res.append(d)
    bad = _ks_all(res[0], res[1], list(res[0]))
    assert not bad, bad


def _run_kernel(name, Sx0, I0, R, variant, leaps, dt, seed0, two_points=False):
    """R replicates from one state through vgsim_simulate_tau with the given variant; returns everything
    the kernel leaves behind for replicate-by-replicate comparison."""
    (U, K, S), setup = SCENARIOS[name]
    e = Eng(U, K, S, seed0, False, False, int(1e6), 0.0)
    setup(e)
    h = _capi.Handle(U, K, S, R, 2 if two_points else 1, None)
    h.set_seeds((np.uint64(seed0) + np.arange(R, dtype=np.uint64)).astype(np.uint64))
    h.upload_params(0, e.param_arrays())
    if two_points:
        e2 = Eng(U, K, S, seed0, False, False, int(1e6), 0.0)
        setup(e2)
        e2.set_transmission_rate(0.4, None)
        e2.set_recovery_rate(0.12, None)
        h.upload_params(1, e2.param_arrays())
        h.set_replicate_params((np.arange(R) % 3 == 1).astype(np.int32))
    h.set_state(np.ascontiguousarray(np.broadcast_to(Sx0, (R,) + Sx0.shape)),
                np.ascontiguousarray(np.broadcast_to(I0, (R,) + I0.shape)))
    h.set_tau_variant(variant)
    h.simulate_tau(leaps, -1, dt, 1)
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    logs = [h.get_tau_log(r) for r in range(0, R, max(1, R // 16))]
    return c, Sx_f, I_f, logs, h.error_flags() if hasattr(h, "error_flags") else None


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name,seed,t0,dt,two", [("t3small", 5, 60.0, 6.0, False), ("t3", 11, 70.0, 2.0, False),
                                                 ("s9", 2020, 4.0, 2.0, False), ("s7", 2020, 6.0, 3.0, False),
                                                 ("table3_k10", 3, 60.0, 3.0, False), ("t3small", 5, 60.0, 6.0, True),
                                                 ("w", 4, 80.0, 0.5, False)])
def test_warp_kernel_reproduces_team_kernel(name, seed, t0, dt, two, variant):
    """The warp-per-replicate kernel (default for large batches) and the team kernel (default for few replicates) share the drift / propensity
    expressions, the summation orders, the Philox addressing and the samplers, so from the same state and seeds
    they must leave the same log: identical leap counts, per-type counters, final compartments and dense rows.
    Covers the mask path (H <= 64, K <= 32), the generic path (w: K = 100), lockdown flips (s7, s9, table3) and
    parameter points staged per warp (two points mixed over the replicates)."""
    Sx0, I0 = warm_state(name, seed, t0)
    assert I0.sum() > 0
    R = 24 if name == "w" else 150
    a = _run_kernel(name, Sx0, I0, R, variant | 8, 40, dt, 700 + variant, two)   # bit 3: warp kernel even for few replicates
    b = _run_kernel(name, Sx0, I0, R, variant | 4, 40, dt, 700 + variant, two)   # bit 2: team kernel
    ca, cb = a[0], b[0]
    assert ca["leaps"].min() >= 1
    for k in ("leaps", "bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "swapLockdown"):
        if k in ca:
            assert np.array_equal(ca[k], cb[k]), (k, np.flatnonzero(ca[k] != cb[k])[:10])
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    np.testing.assert_allclose(ca["time"], cb["time"], rtol=1e-13, atol=0)
    for (cnt_a, tt_a), (cnt_b, tt_b) in zip(a[3], b[3]):
        assert np.array_equal(cnt_a, cnt_b)
        np.testing.assert_allclose(tt_a, tt_b, rtol=1e-13, atol=0)
    exact = all(np.array_equal(x[1], y[1]) for x, y in zip(a[3], b[3]))
    assert exact, "tau/time streams agree to 1e-13 but not bit for bit"


def test_schedule_does_not_change_results():
    """Lockstep generations and the size-sorted replicate schedule (more replicates than warps in flight) are
    scheduling only: the free-running, unsorted kernel must leave exactly the same counters, states and logs."""
    name, seed, t0 = "t3small", 5, 60.0
    Sx0, I0 = warm_state(name, seed, t0)
    R = 2300  # > 148 SMs x 14 warps, so the sorted boustrophedon walk has a second visit
    # unequal replicates: a third of them start from a thinned-out state
    a = _run_kernel(name, Sx0, I0, R, 16, 12, 3.0, 4100)   # variant bit 4: free-running warps
    b = _run_kernel(name, Sx0, I0, R, 0, 12, 3.0, 4100)
    c = _run_kernel(name, Sx0, I0, R, 32, 12, 3.0, 4100)   # variant bit 5: lockstep, unsorted schedule
    for other in (b, c):
        for k in ("leaps", "bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "time"):
            assert np.array_equal(a[0][k], other[0][k]), k
        assert np.array_equal(a[1], other[1]) and np.array_equal(a[2], other[2])
        for (cnt_a, tt_a), (cnt_b, tt_b) in zip(a[3], other[3]):
            assert np.array_equal(cnt_a, cnt_b) and np.array_equal(tt_a, tt_b)
