"""The acceptance test of BASELINE.json for this path, as a function: a dose run of the B200 engine
against a reference run (Geant4-MT where it exists) on identical inputs —

* total edep within 0.5 %,
* per-voxel dose within 2 sigma of the COMBINED statistical uncertainty in the voxels above 10 % of the
  maximum dose (for independent, statistically consistent runs ~95.4 % of those voxels pass),
* identical label / material images.

Inputs are the arrays the DoseActor outputs hold (`dose.edep.image.array`, `dose.edep_uncertainty.image.array`
= relative uncertainty of the mean, `ImageVolume.label_image.array`).  The reference compares images with
`opengate/tests/utility.py:assert_images` (:261-382: sum tolerance + mean per-voxel difference); this is the
sharper statement `north_star` asks for.
"""

from __future__ import annotations

import numpy as np


def compare_dose(value_a, rel_unc_a, value_b, rel_unc_b, total_tolerance=0.005, n_sigma=2.0,
                 threshold_fraction=0.10, min_pass_fraction=0.93):
    """Returns a dict with the figures and `ok`.  value_*: per-voxel dose or edep of two runs normalised to
    the same number of primaries; rel_unc_*: their relative standard uncertainties (DoseActor
    `*_uncertainty` images).  `min_pass_fraction`: the expected 95.4 % minus a margin for the finite number
    of voxels and the correlation between neighbouring voxels of one history."""
    a = np.asarray(value_a, dtype=np.float64)
    b = np.asarray(value_b, dtype=np.float64)
    sa = a * np.asarray(rel_unc_a, dtype=np.float64)
    sb = b * np.asarray(rel_unc_b, dtype=np.float64)
    total_ratio = float(a.sum() / b.sum()) if b.sum() != 0 else float("inf")
    # the voxels are selected on the MEAN of the two runs: selecting on one of them alone picks its upward
    # fluctuations and biases the pulls (regression to the mean, visible as soon as the per-voxel statistics are poor)
    m = 0.5 * (a + b)
    sel = m > threshold_fraction * m.max()
    sigma = np.sqrt(sa[sel] ** 2 + sb[sel] ** 2)
    ok_sigma = sigma > 0
    z = np.abs(a[sel] - b[sel])[ok_sigma] / sigma[ok_sigma]
    # no voxel with a usable uncertainty (uncertainty images missing or all zero): nothing was compared, which
    # is a FAILED acceptance, never a vacuous pass
    frac = float(np.mean(z <= n_sigma)) if z.size else 0.0
    out = {
        "total_ratio": total_ratio,
        "total_ok": abs(total_ratio - 1.0) <= total_tolerance,
        "voxels_compared": int(z.size),
        "fraction_within_n_sigma": frac,
        "mean_pull": float(np.mean((a[sel] - b[sel])[ok_sigma] / sigma[ok_sigma])) if z.size else 0.0,
        "rms_pull": float(np.sqrt(np.mean(z ** 2))) if z.size else 0.0,
        "voxel_ok": bool(z.size > 0 and frac >= min_pass_fraction),
    }
    out["ok"] = bool(out["total_ok"] and out["voxel_ok"])
    return out


def compare_labels(labels_a, labels_b):
    """Integer-exact material / HU assignment."""
    a, b = np.asarray(labels_a), np.asarray(labels_b)
    same = a.shape == b.shape and bool(np.array_equal(a, b))
    return {"ok": same, "different_voxels": int((a != b).sum()) if a.shape == b.shape else -1}
