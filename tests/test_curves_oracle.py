"""Pins the restatement of get_data_infectious / get_data_susceptible (oracle/oracle.py: ref_data_*) against the
UNMODIFIED reference (tests/golden/curves_*.npz, produced by make_golden.py from oracle/_ref): the C++ oracle
replays the same runs bit-exactly (test_oracle_golden.py), its log goes through the restatement, and the
curves must equal the reference's own output EXACTLY -- including the operator-precedence quirk that makes
Data negative.  The GPU curves kernel is then checked against this restatement in test_gpu_curves.py."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
from test_oracle_golden import GOLD, load, make_oracle

WARM_ROWS = 2000000


def curve_names():
    return sorted(tuple(os.path.basename(f)[7:-4].split("_", 1)) for f in glob.glob(os.path.join(GOLD, "curves_*.npz")))


def initial_state(om):
    """State after PrepareParameters: FirstInfection (src/_BirthDeath.pyx:234-242) seeds one case if there is none."""
    Sx, I = om.get_state()
    Sx, I = Sx.copy(), I.copy()
    if I.sum() == 0:
        sn = int(np.nonzero(Sx[0])[0][0])
        Sx[0, sn] -= 1
        I[0, 0] += 1
    return Sx, I


def replay(kind, name):
    if kind == "direct":
        om = make_oracle(name, 2020)
        Sx0, I0 = initial_state(om)
        om.simulate(100000)
        return om, Sx0, I0
    g = load("tau", name)
    om = make_oracle(name, int(g["seed"]))
    Sx0, I0 = initial_state(om)
    om.simulate(int(g["n_direct"]), sample_size=10 ** 9, epidemic_time=float(g["t_warm"]))
    om.simulate(int(g["iters"]), sample_size=10 ** 9, epidemic_time=float(g["t_end"]), method="tau")
    return om, Sx0, I0


def test_curve_fixtures_present():
    assert len(curve_names()) >= 6


@pytest.mark.parametrize("kind,name", curve_names())
def test_restated_curves_equal_reference(kind, name):
    g = load("curves_" + kind, name)
    om, Sx0, I0 = replay(kind, name)
    chain, multi = om.events(), om.multievents()
    ct = om.counters()["time"]
    steps = int(g["steps"])
    for k, (p, h) in enumerate(g["cells"]):
        Data, Sample, tp = O.ref_data_infectious(chain, multi, I0[p, h], ct, int(p), int(h), steps)
        np.testing.assert_array_equal(np.asarray(tp), g["tp"])
        np.testing.assert_array_equal(Data, g["inf_%d" % k])
        np.testing.assert_array_equal(Sample, g["smp_%d" % k])
    for k, (p, s) in enumerate(g["groups"]):
        Data, tp = O.ref_data_susceptible(chain, multi, Sx0[p, s], ct, int(p), int(s), steps)
        np.testing.assert_array_equal(Data, g["sus_%d" % k])


@pytest.mark.parametrize("name,steps", [("s9", 23), ("s7", 1000), ("s4", 7)])
def test_vectorised_timelines_equal_the_literal_loop(name, steps):
    """output_epidemiology_timelines: the product's numpy implementation (vgsim_b200/io.py) against the literal
    restatement of the reference loop, on a reference-exact direct chain (upstream the method itself raises
    AttributeError, so its loop is the specification)."""
    from scenarios import SCENARIOS
    from vgsim_b200 import io as vio
    (U, K, S), _ = SCENARIOS[name]
    H = 4 ** U
    om = make_oracle(name, 2020)
    om.simulate(100000)
    chain, ct = om.events(), om.counters()["time"]
    sizes = [1000000] * K if name != "s7" else None
    from test_oracle_golden import Eng
    (_, _, _), setup = SCENARIOS[name]
    e = Eng(U, K, S, 1, False, False, int(1e6), 0.0)
    setup(e)
    sizes = list(e.sizes)
    want = O.ref_epidemiology_timelines(chain, sizes, K, S, H, ct, steps)
    got = vio.epidemiology_timelines(chain, sizes, K, S, H, ct, steps)
    assert len(got[0]) == len(want[0]) > 0 and got[0] == want[0]
    np.testing.assert_array_equal(got[1], want[1])
    np.testing.assert_array_equal(got[2], want[2])
    log = vio.timelines_as_dict(*got)
    assert log["time"] == want[0] and log["P0"]["H0"] == list(want[2][:, 0, 0]) and len(log["P%d" % (K - 1)]) == S + H


def test_timeline_log_files(tmp_path):
    from vgsim_b200 import io as vio
    times = [0.0, 0.5]
    sus = np.array([[[9, 1]], [[8, 2]]], np.int64)          # [pts][K=1][S=2]
    inf = np.array([[[1, 0, 0, 0]], [[2, 0, 1, 0]]], np.int64)
    vio.write_timelines(times, sus, inf, directory=str(tmp_path / "logs"))
    lines = open(tmp_path / "logs" / "PID0.log").read().splitlines()
    assert lines == ["time S0 S1 H0 H1 H2 H3", "0.0 9 1 1 0 0 0 ", "0.5 8 2 2 0 1 0 "]
