"""Pins the oracle's restated third-party random layer (oracle/np_random.h) draw-for-draw against
numpy itself: PCG64 + SeedSequence(entropy, spawn_key=(num,)) seeding (the mc_lib shim rule),
next_double, random_poisson and random_hypergeometric (reference src/_BirthDeath.pyx:21,74,403,2532,885)."""
import numpy as np
import pytest
from numpy.random import PCG64, Generator, SeedSequence

from oracle import oracle as O


def _gen(entropy, num):
    return Generator(PCG64(SeedSequence(entropy, spawn_key=(num,))))


@pytest.mark.parametrize("entropy,num", [(0, 0), (2020, 0), (2020, 2), (1234, 199), (2**40 + 17, 3), (2**63 - 1, 7)])
def test_stream_matches_numpy(entropy, num):
    g = _gen(entropy, num)
    r = O.OracleRng(entropy, num)
    assert np.array_equal(r.doubles(1000), g.random(1000))
    # raw 64-bit words continue the same stream
    assert np.array_equal(r.raw(100), g.bit_generator.random_raw(100))


def test_survey_cross_check():
    # first uniform of (2020, attempt 2) and scenario 1's first event time
    u = O.OracleRng(2020, 2).doubles(1)[0]
    assert u == 0.44016044650747377


def test_poisson_matches_numpy():
    rs = np.random.RandomState(1)
    lam = np.concatenate([rs.uniform(0, 10, 20000), rs.uniform(10, 200, 20000), 10 ** rs.uniform(-9, 6, 20000),
                          np.zeros(100), np.full(100, 10.0), np.full(100, 9.999999)])
    rs.shuffle(lam)
    g = _gen(99, 1)
    want = np.array([g.poisson(l) for l in lam])
    got = O.OracleRng(99, 1).poisson(lam)
    assert np.array_equal(want, got)


def test_hypergeometric_matches_numpy():
    rs = np.random.RandomState(2)
    n = 30000
    good = rs.randint(0, 3000, n)
    bad = rs.randint(0, 3000, n)
    # mix of small samples (urn path) and large samples (HRUA path), and the > total/2 branches
    sample = np.array([rs.randint(0, g + b + 1) for g, b in zip(good, bad)])
    small = rs.rand(n) < 0.3
    sample[small] = np.minimum(sample[small], rs.randint(0, 12, small.sum()))
    big_good = rs.randint(10**6, 10**9 - 1, 2000)
    big_bad = rs.randint(10**6, 10**9 - 1, 2000)
    big_sample = rs.randint(1, 10**5, 2000)
    good = np.concatenate([good, big_good])
    bad = np.concatenate([bad, big_bad])
    sample = np.concatenate([sample, big_sample])
    g = _gen(5, 0)
    want = np.array([g.hypergeometric(a, b, c) if a + b > 0 else 0 for a, b, c in zip(good, bad, sample)])
    ok = (good + bad) > 0
    got = O.OracleRng(5, 0).hypergeometric(good[ok], bad[ok], sample[ok])
    assert np.array_equal(want[ok], got)
    # interleaving with doubles keeps the 32-bit half-word buffer semantics
    g = _gen(6, 0)
    r = O.OracleRng(6, 0)
    for i in range(200):
        a, b, c = int(good[i]) + 1, int(bad[i]) + 1, int(min(sample[i], 5))
        assert g.hypergeometric(a, b, c) == r.hypergeometric([a], [b], [c])[0]
        assert g.random() == r.doubles(1)[0]
