"""Sparse archive of the dense tau log (vgsim_archive_tau_log, vgsim_simulate_tau_blocks).

The archive is an alternative storage of the same record (SURVEY 8(d) allows a sparse log beside the dense one): every
reader of the log -- genealogy replay, epidemic curves, the exporters -- must give bit-identical results whether a leap
still has its dense row or has been archived, and a run in leap blocks must leave the log a single call leaves.
"""
import numpy as np
import pytest

from test_gpu_tau import make_engine, warm_state

pytestmark = pytest.mark.gpu


def _engine_at(name, seed, t0, R):
    Sx0, I0 = warm_state(name, seed, t0)
    e = make_engine(name, seed, replicates=R)
    e._susceptible[...] = Sx0
    e._infectious[...] = I0
    return e, e._sync_params()


def _snapshot(e, h, gseed):
    """Everything a reader of the log can see."""
    c = h.get_counters()
    out = {"counters": {k: np.array(v) for k, v in c.items()}}
    out["state"] = h.get_state()
    R = h.R
    out["tau_log"] = [h.get_tau_log(r) for r in range(R)]
    out["multi"] = h.get_multievents(0)
    out["events"] = [h.get_event_log(r) for r in range(R)]
    out["curves"] = h.epidemic_curves(12)
    h.genealogy(gseed)
    out["trees"] = [h.get_tree(r) for r in range(R)]
    out["mut"] = [h.get_mutations(r) for r in range(R)]
    out["mig"] = [h.get_migrations(r) for r in range(R)]
    out["summaries"] = h.summaries()
    return out


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + "[%d]" % i)
    else:
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), path


def _run(name, seed, t0, R, calls, archive_after=()):
    """A fresh handle taken through `calls` = [(leaps, sample_size), ...]; archive after the calls listed."""
    e, h = _engine_at(name, seed, t0, R)
    for i, (leaps, sample) in enumerate(calls):
        h.simulate_tau(leaps, sample, -1.0, 1)
        if i in archive_after:
            h.archive_tau_log()
    return e, h


@pytest.mark.parametrize("name,seed,t0,sample", [("t3small", 5, 60.0, 400), ("s9", 2020, 4.0, 300), ("example", 1234, 60.0, 300)])
def test_archived_leaps_read_like_dense_rows(name, seed, t0, sample):
    # (the genealogy rewinds the infectious counts in place, quirk Q9: every snapshot is taken on a fresh twin)
    R = 6
    dense = _snapshot(*_run(name, seed, t0, R, [(40, sample)]), 17)
    assert dense["counters"]["leaps"].min() > 3
    e, h = _run(name, seed, t0, R, [(40, sample)], archive_after=(0,))
    assert h.archive_stats()["leaps_archived"] == dense["counters"]["leaps"].sum()
    _same(dense, _snapshot(e, h, 17), "archived")
    # more leaps on top of the archive (mixed log: archived + dense rows) against a twin that never archived
    calls = [(40, sample), (30, 2 * sample)]
    mixed = _snapshot(*_run(name, seed, t0, R, calls), 23)
    _same(mixed, _snapshot(*_run(name, seed, t0, R, calls, archive_after=(0,)), 23), "mixed")
    _same(mixed, _snapshot(*_run(name, seed, t0, R, calls, archive_after=(0, 1)), 23), "archived twice")


@pytest.mark.parametrize("variant", [8, 4], ids=["warp_kernel", "team_kernel"])
@pytest.mark.parametrize("name,seed,t0", [("t3small", 5, 60.0), ("table3_k10", 7, 60.0)])
def test_blocks_leave_the_log_of_a_single_call(name, seed, t0, variant):
    """vgsim_simulate_tau_blocks(N, block) == vgsim_simulate_tau(N): same Philox addressing (absolute leap index, the
    epoch carried over the blocks), so counters, state, log, curves and trees are bit-identical."""
    R = 5
    e1, h1 = _engine_at(name, seed, t0, R)
    e2, h2 = _engine_at(name, seed, t0, R)
    h1.set_tau_variant(variant)
    h2.set_tau_variant(variant)
    h1.simulate_tau(260, 2500, -1.0, 3)
    h2.simulate_tau_blocks(260, 2500, -1.0, 3, leap_block=101)
    a, b = _snapshot(e1, h1, 5), _snapshot(e2, h2, 5)
    assert a["counters"]["leaps"].max() > 101  # at least one block boundary was crossed
    _same(a, b, "blocks")


def test_reset_and_recycle_drop_the_archive():
    e, h = _engine_at("t3small", 5, 60.0, 4)
    h.simulate_tau(20, -1, -1.0, 1)
    h.archive_tau_log()
    h.recycle_log()
    h.simulate_tau(10, -1, -1.0, 1)
    c = h.get_counters()
    assert np.all(c["leaps"] == 10)
    counts, tt = h.get_tau_log(0)
    assert counts.shape[0] == 10 and counts.sum() > 0
    cv = h.epidemic_curves(4)  # the recycled log starts mid-run: only shape and finiteness are checked here
    assert cv["infectious"].shape[1] == 5
