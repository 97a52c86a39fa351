"""Parameter sweeps (BASELINE.json configs[4] at reduced size): every group of replicates has its own parameter
point inside one handle; results must not depend on how the point list is sharded over ranks."""
import numpy as np
import pytest

from oracle import oracle as O
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
from vgsim_b200.sweep import Sweep

pytestmark = pytest.mark.gpu
NAME = "t3small"
B = np.linspace(1.2, 4.0, 6) * 0.1          # transmission rate = R0 * (d + s)
MIG = np.logspace(-4, -1, 4)                # total migration probability


def points():
    def mk(b, m):
        def f(e):
            e.set_transmission_rate(float(b), None)
            e.set_total_migration_probability(float(m))
        return f
    return [mk(b, m) for b in B for m in MIG]


def test_each_replicate_uses_its_own_point():
    dims, setup = SCENARIOS[NAME]
    sw = Sweep(dims, setup, points(), replicates_per_point=3, seed=77)
    sw.simulate(2500, 10 ** 9, -1, "direct")
    Sx, I = sw.h.get_state()
    for r in (0, 17, 40, sw.R - 1):
        if I[r].sum() == 0:
            continue
        b, m = B[(r // 3) // len(MIG)], MIG[(r // 3) % len(MIG)]
        e = Eng(*dims, 1, False, False, int(1e6), 0.0)
        setup(e)
        e.set_transmission_rate(float(b), None)
        e.set_total_migration_probability(float(m))
        e._susceptible[...] = Sx[r]
        e._infectious[...] = I[r]
        want = O.OracleModel.from_engine(e).propensities()[0]
        got = sw.h.propensities(r)[0]
        assert np.array_equal(got == 0, want == 0)
        nz = want != 0
        assert (np.abs(got[nz] - want[nz]) / np.abs(want[nz])).max() < 1e-12
    # faster transmission reaches 2,500 events sooner: the last row of the R0 grid beats the first
    t = sw.counters()["time"].reshape(len(B), -1)
    assert np.median(t[-1]) < np.median(t[0])


def test_sharding_does_not_change_results():
    dims, setup = SCENARIOS[NAME]

    def run(rank, world):
        sw = Sweep(dims, setup, points(), replicates_per_point=2, seed=5, rank=rank, world=world)
        sw.simulate(1500, 10 ** 9, -1, "direct")
        sw.simulate(12, 10 ** 9, -1, "tau")
        sw.genealogy(seed=np.arange(sw.lo * 2, sw.hi * 2, dtype=np.uint64) + np.uint64(900))
        return sw.summaries()

    whole = run(0, 1)
    halves = np.concatenate([run(0, 2), run(1, 2)], axis=0)
    thirds = np.concatenate([run(g, 3) for g in range(3)], axis=0)
    assert whole.shape == (len(B) * len(MIG), 2, whole.shape[2])
    np.testing.assert_array_equal(whole, halves)
    np.testing.assert_array_equal(whole, thirds)
    assert (whole[:, :, 10] == 12).any() and (whole[:, :, 13] > 1).any()   # leaps were taken, trees were built


def test_contact_density_follows_the_parameter_point():
    """A sweep over contact density: every replicate's LIVE contact density is its own point's, in either order of
    vgsim_upload_params / vgsim_set_replicate_params (round-1 advisor finding: the upload seeded only replicates already
    mapped to the point, so points 1.. never reached the device when the map came last)."""
    from vgsim_b200 import _capi
    dims, setup = SCENARIOS[NAME]
    cds = [0.4, 0.9, 1.7, 2.5]

    def mk(cd):
        return lambda e: e.set_contact_density(float(cd), None)
    sw = Sweep(dims, setup, [mk(c) for c in cds], replicates_per_point=5, seed=3)
    _, _, cd, lock = sw.h.get_state(full=True)
    assert np.array_equal(cd, np.repeat(np.asarray(cds), 5)[:, None] * np.ones((1, dims[1])))
    # the other order, straight through the C ABI: uploads first, map last
    U, K, S = dims
    e = Eng(U, K, S, 1, False, False, int(1e6), 0.0)
    setup(e)
    h = _capi.Handle(U, K, S, 8, 2, None)
    for pp, c in enumerate((0.3, 2.2)):
        e.set_contact_density(c, None)
        h.upload_params(pp, e.param_arrays())
    h.set_replicate_params(np.array([0, 1, 1, 0, 1, 0, 0, 1], np.int32))
    _, _, cd, _ = h.get_state(full=True)
    assert np.array_equal(cd[:, 0], np.where(np.array([0, 1, 1, 0, 1, 0, 0, 1]) == 1, 2.2, 0.3))
    # and the rates the kernels use follow: the direct-method birth rate scales with the contact density
    r0, r1 = h.rates(0)["ev"][0, 0, 0], h.rates(1)["ev"][0, 0, 0]
    assert abs(r1 / r0 - 2.2 / 0.3) < 1e-9
