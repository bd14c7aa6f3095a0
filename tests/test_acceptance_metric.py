"""The BASELINE.json acceptance metric (opengate_b200/acceptance.py) exercised on the CPU oracle: two
independent runs of the same configuration pass; a run with a 2 % dose bias, or with a shifted beam, fails."""
import numpy as np

from opengate_b200 import abi, acceptance
from opengate_b200.actors import uncertainty_from_squares
from oracle.oracle import TransportOracle
from small_setups import c3


def run(seed, n, scale=1.0, tilt=0.0):
    """edep per primary and its relative uncertainty from one oracle run of the small CT configuration;
    scale: global bias of the result; tilt: +-tilt on the two halves of the grid (the total stays)."""
    S = c3(n)
    o = TransportOracle(seed)
    so, do, _ = S.apply(o)
    o.begin_run()
    o.run_source(so[0], 0, n, 0)
    # edep, not dose: the total of BASELINE.json is the total edep (a dose total is dominated by the rare
    # deposits in air voxels, 1 / rho = 800)
    d = o.dose_grid(do[0], abi.GRID_EDEP) * scale
    d2 = o.dose_grid(do[0], abi.GRID_EDEP_SQ) * scale * scale
    if tilt:
        f = np.ones_like(d)
        f[:, : d.shape[1] // 2, :] += tilt
        f[:, d.shape[1] // 2:, :] -= tilt
        d, d2 = d * f, d2 * f * f
    return d / n, uncertainty_from_squares(d, d2, n)


def test_independent_runs_are_statistically_consistent_and_biased_ones_are_not():
    n = 1500000
    a, ua = run(11, n)
    b, ub = run(22, n)
    r = acceptance.compare_dose(a, ua, b, ub)
    assert r["voxels_compared"] > 500
    assert r["ok"], r
    assert 0.93 <= r["fraction_within_n_sigma"] <= 0.99       # 95.4 % for Gaussian pulls
    assert abs(r["mean_pull"]) < 0.1 and 0.9 < r["rms_pull"] < 1.1
    # a 2 % bias of the dose is caught by the total (0.5 %) although most voxels still agree within 2 sigma
    c, uc = run(33, n, scale=1.02)
    r = acceptance.compare_dose(c, uc, b, ub)
    assert not r["total_ok"] and not r["ok"]
    # +-15 % on the two halves of the grid keeps the total within 0.5 % but fails per voxel
    d, ud = run(44, n, tilt=0.15)
    r = acceptance.compare_dose(d, ud, b, ub)
    assert not r["voxel_ok"] and not r["ok"], r


def test_label_comparison_is_integer_exact():
    a = np.arange(24, dtype=np.uint16).reshape(2, 3, 4)
    assert acceptance.compare_labels(a, a.copy())["ok"]
    b = a.copy()
    b[1, 2, 3] += 1
    r = acceptance.compare_labels(a, b)
    assert not r["ok"] and r["different_voxels"] == 1


def test_no_comparable_voxel_is_a_failure_not_a_vacuous_pass():
    from opengate_b200 import acceptance
    a = np.ones((4, 4, 4))
    r = acceptance.compare_dose(a, np.zeros_like(a), a, np.zeros_like(a))
    assert r["voxels_compared"] == 0 and not r["voxel_ok"] and not r["ok"]
