"""Deterministic parity of the device cumulative search (vgsim_b200/csrc/choose.cuh, used by the direct kernel for every
categorical level) with the reference's fastChoose / fastChoose_skip (src/fast_choose.pxi:18-52), entry by entry through
the vgsim_test_choose tap: picked index, the reference's residual, zero weights, the last-index catch-all and `skip`."""
import numpy as np
import pytest

from vgsim_b200 import _capi

pytestmark = pytest.mark.gpu


def fast_choose(w, tw, rn):
    """src/fast_choose.pxi:18-31, literally (returns None where the reference would print '0-weight sampled' and exit)."""
    rn = tw * rn
    i = 0
    total = w[0]
    while total < rn and i < len(w) - 1:
        i += 1
        total += w[i]
    if w[i] == 0.0:
        return None
    return i, (rn - (total - w[i])) / w[i]


def fast_choose_skip(w, tw, rn, skip):
    """src/fast_choose.pxi:36-52, literally."""
    rn = tw * rn
    i = 0
    if skip == 0:
        i += 1
    total = w[i]
    while total < rn and i < len(w) - 1:
        i += 1
        if i != skip:
            total += w[i]
    if w[i] == 0.0:
        return None
    return i, (rn - (total - w[i])) / w[i]


def _check(w, rns, skip, small):
    w = np.asarray(w, float)
    tw = float(np.sum(np.delete(w, skip))) if skip >= 0 else float(w.sum())
    x = rns * tw
    idx, before, wsel, resid = _capi.test_choose(w, x, skip=skip, small=small)
    cum = np.cumsum(np.where(np.arange(len(w)) == skip, 0.0, w))
    n_cmp = 0
    for q, rn in enumerate(rns):
        ref = fast_choose_skip(w, tw, rn, skip) if skip >= 0 else fast_choose(w, tw, rn)
        if ref is None:
            continue
        i_ref, r_ref = ref
        if i_ref == skip:
            continue   # the reference's walk can stop ON the skipped index (its loop bound); the device never returns it
        # a target within rounding of a cell boundary may legitimately fall on either side (the device's scan adds in
        # a different order): compare away from the boundaries
        if np.min(np.abs(cum - x[q])) <= 4e-16 * max(tw, 1.0) * len(w):
            continue
        assert idx[q] == i_ref, (q, rn, idx[q], i_ref)
        assert wsel[q] == w[i_ref]
        # residual (x - before) / w: the rounding of `before` is relative to the running total, not to w
        tol = 8e-16 * tw / w[i_ref] * np.log2(len(w) + 1) + 1e-15
        assert abs(resid[q] - r_ref) <= tol, (q, resid[q], r_ref, tol)
        assert 0.0 <= resid[q] < 1.0
        n_cmp += 1
    return n_cmp


@pytest.mark.parametrize("small", [False, True], ids=["warp_scan", "sequential"])
@pytest.mark.parametrize("n", [1, 3, 4, 10, 31, 32, 33, 64, 100])
def test_pick_matches_fast_choose(n, small):
    rng = np.random.default_rng(1000 + n)
    rns = rng.random(1500)
    for trial in range(3):
        w = rng.random(n) * 10.0 ** rng.integers(-6, 4, n)        # weights spanning ten decades
        if trial == 1 and n > 3:
            w[rng.integers(0, n, max(1, n // 3))] = 0.0             # zero weights in the middle
            w[0] = 1.0
            w[-1] = 1.0
        if trial == 2:
            w = np.floor(w * 1000) + 1.0                            # integer-valued weights (the npy_int64 instantiation)
        assert _check(w, rns, -1, small) > 1000
        if n > 2:
            for skip in (0, n // 2, n - 1):
                assert _check(w, rns, skip, small) > 800


@pytest.mark.parametrize("small", [False, True])
def test_zero_weights_and_catch_all(small):
    # a zero weight is never picked: the reference would print "0-weight sampled" and exit (:29-30)
    w = np.array([0.0, 0.0, 2.0, 0.0, 3.0, 0.0])
    x = np.array([0.0, 1e-300, 1.999, 2.0, 2.0000001, 4.999, 5.0, 5.0 + 1e-9, 7.0])
    idx, before, wsel, resid = _capi.test_choose(w, x, small=small)
    assert list(idx) == [2, 2, 2, 2, 4, 4, 4, 4, 4]        # first positive weight; cumulative >= x; last positive = catch-all
    assert np.all(wsel == w[idx]) and np.all((resid >= 0) & (resid < 1))
    np.testing.assert_allclose(resid[:4], x[:4] / 2.0, rtol=1e-15, atol=0)
    np.testing.assert_allclose(resid[4:6], (x[4:6] - 2.0) / 3.0, rtol=1e-12)
    assert np.all(resid[6:] == 0.9999999999999999)          # above the total: clamped residual of the catch-all
    # every weight zero: no pick (the kernels raise the sticky "zero weight sampled" error bit)
    idx, *_ = _capi.test_choose(np.zeros(40), np.array([0.0, 0.5]), small=small)
    assert list(idx) == [-1, -1]
    # skip: the skipped weight takes no part even when it is the only positive one before the target
    idx, before, wsel, resid = _capi.test_choose(np.array([5.0, 1.0, 1.0]), np.array([0.0, 0.5, 1.5]), skip=0, small=small)
    assert list(idx) == [1, 1, 2] and list(before) == [0.0, 0.0, 1.0]
