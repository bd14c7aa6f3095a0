"""Track-length estimator (TLEDoseActor) on the CPU oracle: properties that do not need
a GPU.  The reference's own acceptance pattern for this actor is self-consistency between the analog
DoseActor and the TLEDoseActor in the same run (test081_tle_*: PDD section means within 10-17 %)."""
import numpy as np
import pytest

from opengate_b200 import abi
from oracle.oracle import TransportOracle
from tle_setups import SQ, c1_tle, c3_tle, c4_tle, long_thin


def run(S, n, first=0, threads=0):
    o = TransportOracle(S.seed)
    sids, dids, _ = S.apply(o)
    o.begin_run()
    o.run_source(sids[0], first, n, threads)
    return o, dids


@pytest.mark.parametrize("mk,tol", [(c1_tle, 0.05), (c3_tle, 0.03), (c4_tle, 0.03)])
def test_tle_total_agrees_with_analog_kerma(mk, tol):
    S = mk(20000)
    o, d = run(S, 20000)
    ea, et = o.dose_grid(d[0], abi.GRID_EDEP), o.dose_grid(d[-1], abi.GRID_EDEP)
    assert et.sum() == pytest.approx(ea.sum(), rel=tol)
    # the estimator scores in many more voxels than the analog one
    assert (et > 0).sum() > (ea > 0).sum()
    st = o.stats()
    assert st["tle_overflows"] == 0 and st["tle_steps"] > 0 and st["tle_entries"] > 0


def test_depth_dose_profile_agrees_like_reference_test081():
    S = c3_tle(60000)
    o, d = run(S, 60000)
    ea, et = o.dose_grid(d[0], abi.GRID_EDEP), o.dose_grid(d[-1], abi.GRID_EDEP)
    # profile along the beam axis (x), in 8 sections (test081_tle_helpers.py:177-236 compares section means)
    pa, pt = ea.sum(axis=(0, 1)), et.sum(axis=(0, 1))
    sec_a = pa.reshape(8, -1).sum(axis=1)
    sec_t = pt.reshape(8, -1).sum(axis=1)
    big = sec_a > 0.02 * sec_a.sum()
    np.testing.assert_allclose(sec_t[big], sec_a[big], rtol=0.10)


def test_single_history_square_is_square_of_sum():
    # one event: edep_squared must be (sum of the event's deposits in the voxel)^2, also when the
    # history crosses the same voxel on several flights (the interaction voxel always is)
    S = c3_tle(1)
    for h in (0, 3, 11, 12345, 99999):
        o, d = run(S, 1, first=h)
        e, e2 = o.dose_grid(d[-1], abi.GRID_EDEP), o.dose_grid(d[-1], abi.GRID_EDEP_SQ)
        dd, d2 = o.dose_grid(d[-1], abi.GRID_DOSE), o.dose_grid(d[-1], abi.GRID_DOSE_SQ)
        assert np.array_equal(e2, e * e) and np.array_equal(d2, dd * dd)
        assert o.stats()["tle_entries"] == (e != 0).sum()   # one table entry per voxel the event touched


def test_threshold_zero_makes_the_tle_actor_an_analog_dose_actor():
    S = c3_tle(8000, flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ | abi.SCORE_COUNTS,
               hit=abi.HIT_POST, threshold=0.0)
    o, d = run(S, 8000)
    for which in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_DOSE, abi.GRID_DOSE_SQ, abi.GRID_COUNTS):
        np.testing.assert_allclose(o.dose_grid(d[-1], which), o.dose_grid(d[0], which), rtol=1e-12)
    assert o.stats()["tle_steps"] == 0


def test_energy_threshold_splits_the_spectrum():
    # 120 kVp spectrum, threshold 60 keV: photons above score analog, below by track length
    S = c3_tle(20000, threshold=0.06)
    o, d = run(S, 20000)
    ea, et = o.dose_grid(d[0], abi.GRID_EDEP), o.dose_grid(d[-1], abi.GRID_EDEP)
    assert et.sum() == pytest.approx(ea.sum(), rel=0.03)
    S0 = c3_tle(20000)
    o0, d0 = run(S0, 20000)
    assert 0 < o.stats()["tle_steps"] < o0.stats()["tle_steps"]


def test_coarse_unaligned_grid_conserves_the_total():
    Sa, Sc = c3_tle(6000), c3_tle(6000, coarse=True)
    oa, da = run(Sa, 6000)
    oc, dc = run(Sc, 6000)
    ta, tc = oa.dose_grid(da[-1], abi.GRID_EDEP).sum(), oc.dose_grid(dc[-1], abi.GRID_EDEP).sum()
    # the same estimator on a coarser, unaligned grid: the totals agree statistically (the two runs share only the
    # first flight of each history -- the coarse grid draws a random point per chord, which moves every later
    # random number; batches of 6000 histories scatter by 1 %)
    assert tc == pytest.approx(ta, rel=3e-2)


def test_table_overflow_is_counted_and_conserves_sums():
    S = long_thin(100)
    o, d = run(S, 100)
    st = o.stats()
    assert st["tle_overflows"] > 0
    e = o.dose_grid(d[-1], abi.GRID_EDEP)
    assert (e > 0).sum() > 3072
    # water voxels (every 7th) get ~800x the air deposit: density and mu_en per material are applied
    water, air = e[0, 0, ::7].mean(), e[0, 0, 1::7].mean()
    assert water > 300 * air


def test_thread_count_does_not_change_the_result():
    S = c3_tle(3000)
    o1, d1 = run(S, 3000, threads=1)
    o4, d4 = run(S, 3000, threads=4)
    for which in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_DOSE):
        np.testing.assert_allclose(o1.dose_grid(d1[-1], which), o4.dose_grid(d4[-1], which), rtol=1e-12)
    assert o1.stats()["tle_steps"] == o4.stats()["tle_steps"]


@pytest.mark.parametrize("mk,n,cap", [(lambda: c3_tle(6000), 6000, None), (lambda: c3_tle(6000, threshold=0.06), 6000, None),
                                      (lambda: c4_tle(6000), 6000, None), (lambda: long_thin(200), 200, None),
                                      (lambda: c3_tle(6000), 6000, 2), (lambda: c3_tle(6000, threshold=0.06), 6000, 1)])
def test_record_fold_statement_equals_the_table(mk, n, cap):
    """The device scores ONE TLE actor with squares on the image grid without a per-history table: the sum a voxel
    holds is re-evaluated from the records of the history's flights (closed form of a voxel chord) and point deposits.
    The oracle carries that statement behind a switch; it must give the table's sums bit for bit (event sums are
    float on both sides) and the same squares to the rounding of (p + e)^2 - p^2, overflow paths included."""
    res = []
    for fold in (False, True):
        S = mk()
        o = TransportOracle(S.seed)
        sids, dids, _ = S.apply(o)
        if cap is not None:
            o.set_tle_rec_cap(cap)
        if fold:
            o.set_tle_fold(True)
        o.begin_run()
        o.run_source(sids[0], 0, n, 0)
        res.append((o, dids))
    (a, d), (b, _) = res
    sa, sb = a.stats(), b.stats()
    for k in ("tle_steps", "tle_entries", "tle_overflows", "deposits", "tentative_collisions"):
        assert sa[k] == sb[k], k
    if cap is not None or n == 200:
        assert sa["tle_overflows"] > 0
    for which in (abi.GRID_EDEP, abi.GRID_DOSE):
        assert np.array_equal(a.dose_grid(d[-1], which), b.dose_grid(d[-1], which))
    for which in (abi.GRID_EDEP_SQ, abi.GRID_DOSE_SQ):
        x, y = a.dose_grid(d[-1], which), b.dose_grid(d[-1], which)
        assert np.array_equal(x != 0, y != 0)
        np.testing.assert_allclose(x, y, rtol=1e-14, atol=0)


def test_record_fold_needs_one_aligned_tle_actor_with_squares():
    S = c3_tle(10, coarse=True)
    o = TransportOracle(S.seed)
    S.apply(o)
    with pytest.raises(ValueError):
        o.set_tle_fold(True)
