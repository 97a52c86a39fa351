"""GPU parity tests of the batched direct-method kernel (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
from scipy import stats

from oracle import oracle as O
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
from test_gpu_tau import make_engine, warm_state, _ks_all

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,seed,t_end", [("s9", 2020, 4.0), ("example", 1234, 70.0), ("t3small", 5, 60.0),
                                             ("s7", 2020, 6.0), ("table3_k10", 3, 60.0)])
def test_rate_hierarchy_matches_oracle(name, seed, t_end):
    """UpdateAllRates (reference :279-351) from a given state: deterministic, 1e-12 relative."""
    Sx, I = warm_state(name, seed, t_end)
    e = make_engine(name, seed)
    e._susceptible[...] = Sx
    e._infectious[...] = I
    dev = e.rates()
    ora = O.OracleModel.from_engine(e).rates()
    for k in ("A", "eff", "maxEBM", "ev", "hp", "popRate", "migPop", "totals"):
        a, b = dev[k], ora[k]
        assert np.array_equal(a == 0, b == 0), k
        nz = b != 0
        if nz.any():
            assert (np.abs(a[nz] - b[nz]) / np.abs(b[nz])).max() < 1e-12, k


def _replay_direct(e, chain, Sx0, I0):
    Sx, I = Sx0.copy(), I0.copy()
    H = e.hapNum
    t, ty, hap, pop, nhap, npop = chain
    ty, hap, pop, nhap, npop = (x.astype(np.int64) for x in (ty, hap, pop, nhap, npop))
    for i in range(chain.shape[1]):
        k = ty[i]
        if k == 0:
            assert npop[i] == H  # sentinel of non-recombinant births (reference :600)
            I[pop[i], hap[i]] += 1; Sx[pop[i], nhap[i]] -= 1
        elif k in (1, 2):
            I[pop[i], hap[i]] -= 1; Sx[pop[i], nhap[i]] += 1
        elif k == 3:
            I[pop[i], hap[i]] -= 1; I[pop[i], nhap[i]] += 1
        elif k == 4:
            Sx[pop[i], hap[i]] -= 1; Sx[pop[i], nhap[i]] += 1
        elif k == 5:
            I[npop[i], hap[i]] += 1; Sx[npop[i], nhap[i]] -= 1
        assert Sx.min() >= 0 and I.min() >= 0
    return Sx, I


@pytest.mark.parametrize("name", ["s9", "s4", "s7"])
def test_direct_log_replays_to_final_state(name):
    R = 6
    e = make_engine(name, 77, replicates=R)
    h = e._sync_params()
    h.simulate_direct(3000, -1, -1.0, 200)
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    Sx0 = e._susceptible.copy()
    I0 = e._infectious.copy()
    for r in range(R):
        chain = h.get_event_log(r)
        assert chain.shape[1] == c["events"][r] <= 3000
        if chain.shape[1] < 3000:  # stopped early only by extinction (after > 100 events, else it restarts)
            assert c["globalInfectious"][r] == 0 and chain.shape[1] > 100
        assert np.all(np.diff(chain[0]) >= 0)
        # first infection (reference FirstInfection): one individual, deme 0, haplotype 0, first non-empty group
        I1 = I0.copy(); Sx1 = Sx0.copy(); I1[0, 0] += 1; Sx1[0, np.nonzero(Sx0[0])[0][0]] -= 1
        Sx, I = _replay_direct(e, chain, Sx1, I1)
        assert np.array_equal(Sx, Sx_f[r]) and np.array_equal(I, I_f[r])
        ty = chain[1].astype(int)
        for code, key in enumerate(["bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"]):
            assert (ty == code).sum() == c[key][r], key
        assert c["good_attempt"][r] >= 1 and c["time"][r] == chain[0, -1]


@pytest.mark.parametrize("name", ["s1", "s2", "s3", "s4", "s5", "s6", "s7", "s8", "s9"])
def test_direct_distribution_matches_oracle(name):
    """The nine scenarios of the reference's testing/check_simulator.py as replicate batches:
    device (Philox, seeds 5000+r) vs oracle == reference algorithm (PCG64, seeds 1000+r)."""
    R, N = 1200, 2500
    e = make_engine(name, 5000, replicates=R)
    h = e._sync_params()
    h.simulate_direct(N, N, -1.0, 200)
    c = h.get_counters()
    Sx_f, I_f = h.get_state()
    keys = ["bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "migNonPlus", "time", "good_attempt",
            "inf_total", "events"]
    dev = {k: c[k] for k in keys if k in c}
    dev["inf_total"] = I_f.sum(axis=(1, 2))
    ora = {k: [] for k in keys}
    for r in range(R):
        om = O.OracleModel.from_engine(make_engine(name, 1000 + r))
        om.simulate(N)
        oc = om.counters()
        for k in keys:
            if k in oc:
                ora[k].append(oc[k])
        ora["inf_total"].append(om.get_state()[1].sum())
    bad = _ks_all(dev, ora, keys)
    assert not bad, bad


def test_direct_then_tau_continues_the_log():
    """testing/example.py pattern: direct warm-up, parameter change, then tau-leaping on the same log."""
    from scenarios import example_phase2
    R = 4
    e = make_engine("example", 1234, replicates=R)
    e.SimulatePopulation(10**6, 10**6, 40.0, 200)
    c1 = e.counters()
    example_phase2(e)
    e.SimulatePopulation_tau(60, 10**9, 55.0, 200)
    c2 = e.counters()
    assert np.all(c2["events"] == c1["events"] + c2["leaps"]) and np.all(c2["leaps"] > 0)
    assert np.all(c2["time"] >= 55.0 - 1e-9) or np.all(c2["leaps"] == 60)
    chain = e.get_chain_events(0)
    n1 = int(c1["events"][0])
    assert np.all(chain[1, :n1] < 6) and np.all(chain[1, n1:] == 6)
    P = e._handle.P
    assert np.array_equal(chain[2, n1:], np.arange(c2["leaps"][0]) * P)
    assert np.array_equal(chain[3, n1:], (np.arange(c2["leaps"][0]) + 1) * P)
    # contact density set between the calls reached the device
    _, _, cd, _ = e._handle.get_state(full=True)
    assert np.all(cd[:, 0] == 0.7) or np.any(e._handle.get_lockdowns(0)[0] == 1)


@pytest.mark.parametrize("name,N", [("s5", 1000), ("s6", 1000), ("s9", 1000), ("t3small", 1000), ("table3_k10", 1000), ("s7", 20000)])
def test_incremental_totals_equal_fresh_recompute(name, N):
    """UpdateRates (src/_BirthDeath.pyx:516-546) keeps totalRate and totalMigrationRate by increments; after N events the
    kernel's running totals must equal a fresh recompute of the whole hierarchy from the final state (vgsim_rates).
    N = 1000 stays below the kernel's own 1,024-iteration resynchronisation, so the increments alone are checked;
    the 20,000-event case crosses it (and lockdown flips) as well."""
    R = 16
    e = make_engine(name, 4242, replicates=R)
    h = e._sync_params()
    h.simulate_direct(N, -1, -1.0, 200)
    s = h.summaries()
    c = h.get_counters()
    checked = 0
    for r in range(R):
        if c["globalInfectious"][r] == 0:
            continue
        fresh = h.rates(r)["totals"]
        assert abs(s[r, 22] - fresh[0]) <= 1e-11 * fresh[0], (r, s[r, 22], fresh[0])
        assert abs(s[r, 23] - fresh[1]) <= 1e-9 * max(fresh[0], fresh[1]), (r, s[r, 23], fresh[1])
        checked += 1
    assert checked >= R // 2


def test_contact_density_property_follows_lockdowns():
    """The reference's CheckLockdown overwrites contactDensity (src/_BirthDeath.pyx:698-710), so the `contact_density`
    property shows the lockdown value afterwards; a later simulate call must not put the pre-lockdown value back."""
    e = make_engine("example", 1234)
    before = e.contact_density.copy()
    e.SimulatePopulation(10 ** 7, 10 ** 9, 90.0, 200)
    h = e._handle
    _, _, cd, lock = h.get_state(full=True)
    assert lock[0].sum() > 0, "the example model locks demes down by t = 90"
    assert np.array_equal(e.contact_density, cd[0]) and not np.array_equal(e.contact_density, before)
    locked = lock[0] == 1
    assert np.allclose(e.contact_density[locked], e.contactDensityAfterLockdown[locked])
    e.SimulatePopulation(1000, 10 ** 9, -1, 200)      # continues from the live densities
    _, _, cd2, lock2 = h.get_state(full=True)
    same = lock2[0] == lock[0]
    assert np.array_equal(cd2[0][same], cd[0][same])
