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
                 threshold_fraction=0.10, min_pass_fraction=0.93, total_over_selected=False, mask=None):
    """Returns a dict with the figures and `ok`.  value_*: per-voxel dose or edep of two runs normalised to
    the same number of primaries; rel_unc_*: their relative standard uncertainties (DoseActor
    `*_uncertainty` images).  `min_pass_fraction`: the expected 95.4 % minus a margin for the finite number
    of voxels and the correlation between neighbouring voxels of one history.  `total_over_selected`: take the
    total over the compared voxels only -- for DOSE images, whose plain sum is dominated by the rare deposits
    in air voxels (dose = edep / 1.2e-3 g/cm3) and is not the "total edep" of the acceptance statement.
    `mask`: boolean array of the voxels that may be compared at all (a DOSE image of a CT is compared inside the
    body: the maximum of the whole image is a rare event in an air voxel, not the dose maximum)."""
    a = np.asarray(value_a, dtype=np.float64)
    b = np.asarray(value_b, dtype=np.float64)
    sa = a * np.asarray(rel_unc_a, dtype=np.float64)
    sb = b * np.asarray(rel_unc_b, dtype=np.float64)
    # the voxels are selected on the MEAN of the two runs: selecting on one of them alone picks its upward
    # fluctuations and biases the pulls (regression to the mean, visible as soon as the per-voxel statistics are poor)
    m = 0.5 * (a + b)
    if mask is not None:
        m = np.where(np.asarray(mask, dtype=bool), m, 0.0)
    sel = m > threshold_fraction * m.max()
    ta, tb = (a[sel].sum(), b[sel].sum()) if total_over_selected else (a.sum(), b.sum())
    total_ratio = float(ta / tb) if tb != 0 else float("inf")
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


def depth_dose_sections(value_a, value_b, sections, tolerance, axis=0, half_width=0):
    """Section means of the central-axis depth-dose profile of two runs, the reference's own acceptance pattern
    for dose engines (opengate/tests/src/actors/test081_tle_helpers.py:177-236: `mean_rel_diff` per material
    section of the profile `arr[:, ny/2, nx/2]`).  `sections`: [(i1, i2)] index ranges along `axis`;
    `half_width`: also average over +-half_width voxels around the axis (0 = the reference's single column).
    Returns {"ok", "sections": [(i1, i2, mean_a, mean_b, relative difference)]}."""
    a = np.moveaxis(np.asarray(value_a, dtype=np.float64), axis, 0)
    b = np.moveaxis(np.asarray(value_b, dtype=np.float64), axis, 0)
    c1, c2 = a.shape[1] // 2, a.shape[2] // 2
    sl = (slice(None), slice(c1 - half_width, c1 + half_width + 1), slice(c2 - half_width, c2 + half_width + 1))
    pa, pb = a[sl].mean(axis=(1, 2)), b[sl].mean(axis=(1, 2))
    rows, ok = [], True
    for i1, i2 in sections:
        v1, v2 = float(pa[i1:i2].mean()), float(pb[i1:i2].mean())
        r = abs(v1 - v2) / v1 if v1 != 0 else float("inf")
        rows.append((int(i1), int(i2), v1, v2, r))
        ok = ok and r < tolerance
    return {"ok": bool(ok), "sections": rows}
