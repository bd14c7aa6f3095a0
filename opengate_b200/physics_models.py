"""Per-element photon interaction models used to split NIST total attenuation into processes.

Why this file exists
--------------------
In the reference all photon physics lives in Geant4 (not in the tree) and no
per-process cross-section is exported to Python (F7).  The production path for the tables
consumed by the GPU engine is an exporter built on `G4EmCalculator` /
`G4VEmModel::CrossSectionPerVolume` (template: reference
`core/opengate_core/opengate_lib/GateMaterialMuHandler.cpp:296-549`); that exporter can only
be compiled where Geant4 exists.  Until then the tables are produced
here, from data that IS pinned in the reference tree plus textbook atomic physics:

* total mu/rho per element: NIST XAAMDI as shipped in the reference
  (`opengate/data/PhotonAttenuation.py`), resampled by `tools/gen_nist_tables.py`;
* Compton: Klein-Nishina differential cross-section times the incoherent scattering function
  S(q,Z) (the Livermore / Penelope style binding correction, no Doppler broadening);
* Rayleigh: Thomson cross-section times the squared atomic form factor F(q,Z)^2;
* F and S from an independent-electron model with Slater-type orbitals and Slater's (1930)
  screening rules (closed-form Fourier transforms, no data tables needed);
* pair (+triplet): the Geant4 Bethe-Heitler parameterised total cross-section
  (Geant4 Physics Reference Manual, "Gamma conversion", the a/b/c polynomial in ln(E/mc2));
* photoelectric = NIST total - Compton - Rayleigh - pair (so that the SUM of the processes
  reproduces the NIST totals exactly on the grid); where the photoelectric share drops below
  2 % it is continued by a power law and the (sub-percent) residual is absorbed by Compton.

Everything here is host-side numpy in float64; it runs once per simulation.
"""

from __future__ import annotations

import functools
import os

import numpy as np

ELECTRON_MASS_MEV = 0.51099891
R_E2_BARN = 0.07940787  # classical electron radius squared, barn
N_AVOGADRO = 6.02214076e23
# photon wave number in 1/bohr per MeV:  k = E / (hbar c) ; hbar c = 197.3269804 MeV fm ;
# bohr = 52917.721 fm
K_PER_MEV = 52917.72109 / 197.3269804

_DATA = os.path.join(os.path.dirname(__file__), "data", "nist_photon_totals.npz")


@functools.lru_cache(maxsize=1)
def nist():
    """NIST totals on the master grid (see tools/gen_nist_tables.py)."""
    d = np.load(_DATA)
    return {k: d[k] for k in d.files}


def element_symbol_to_z(symbol: str) -> int:
    sy = list(nist()["symbol"])
    return sy.index(symbol) + 1


def atomic_mass(z: int) -> float:
    """A in g/mol from the NIST Z/A column."""
    return z / float(nist()["z_over_a"][z - 1])


# --------------------------------------------------------------------------------------
# Slater-type-orbital independent electron model
# --------------------------------------------------------------------------------------

# Madelung filling order: (n, l, capacity)
_FILL = [
    (1, 0, 2), (2, 0, 2), (2, 1, 6), (3, 0, 2), (3, 1, 6), (4, 0, 2), (3, 2, 10), (4, 1, 6),
    (5, 0, 2), (4, 2, 10), (5, 1, 6), (6, 0, 2), (4, 3, 14), (5, 2, 10), (6, 1, 6), (7, 0, 2),
    (5, 3, 14), (6, 2, 10), (7, 1, 6),
]
_NSTAR = {1: 1.0, 2: 2.0, 3: 3.0, 4: 3.7, 5: 4.0, 6: 4.2, 7: 4.4}


def electron_configuration(z: int):
    """Ground-state configuration by the Madelung rule: list of (n, l, occupancy)."""
    left = z
    conf = []
    for n, l, cap in _FILL:
        if left <= 0:
            break
        occ = min(cap, left)
        conf.append((n, l, occ))
        left -= occ
    return conf


def slater_groups(z: int):
    """Slater groups [(n_star, zeta, n_electrons)], zeta = (Z - s)/n* in 1/bohr."""
    conf = electron_configuration(z)
    # group key: (n, 'sp') or (n, 'd') or (n, 'f')
    groups = {}
    for n, l, occ in conf:
        key = (n, "sp" if l <= 1 else ("d" if l == 2 else "f"))
        groups[key] = groups.get(key, 0) + occ
    out = []
    for (n, kind), occ in groups.items():
        s = 0.0
        for (n2, kind2), occ2 in groups.items():
            if (n2, kind2) == (n, kind):
                s += (0.30 if n == 1 else 0.35) * (occ - 1)
            elif kind == "sp":
                if n2 == n - 1:
                    s += 0.85 * occ2
                elif n2 < n - 1:
                    s += 1.00 * occ2
                # same n, d/f groups lie outside: no screening
            else:  # d or f group: everything listed before it screens fully
                if n2 < n or (n2 == n and kind2 == "sp") or (n2 == n and kind == "f" and kind2 == "d"):
                    s += 1.00 * occ2
        nstar = _NSTAR[n]
        zeta = max(z - s, 0.3) / nstar
        out.append((nstar, zeta, occ))
    return out


def _sto_form_factor(q, nstar, zeta):
    """Fourier transform of a normalised Slater radial density r^(2n*) exp(-2 zeta r).

    f(q) = (2z)^(2n+1) sin(2n atan(q/2z)) / (2n q (4z^2+q^2)^n),  f(0) = 1.
    """
    q = np.asarray(q, dtype=np.float64)
    a = 2.0 * zeta
    t = np.maximum(q, 1e-12) / a
    two_n = 2.0 * nstar
    val = np.sin(two_n * np.arctan(t)) / (two_n * t * (1.0 + t * t) ** nstar)
    return np.where(q < 1e-9, 1.0, val)


def form_factor(z: int, q):
    """Atomic form factor F(q,Z); q = momentum transfer in 1/bohr. F(0)=Z."""
    tot = 0.0
    for nstar, zeta, occ in slater_groups(z):
        tot = tot + occ * _sto_form_factor(q, nstar, zeta)
    return tot


def incoherent_function(z: int, q):
    """Incoherent scattering function S(q,Z) = sum_e (1 - f_e(q)^2) (Waller-Hartree without
    exchange terms). S(0)=0, S(inf)=Z."""
    tot = 0.0
    for nstar, zeta, occ in slater_groups(z):
        f = _sto_form_factor(q, nstar, zeta)
        tot = tot + occ * (1.0 - f * f)
    return tot


# --------------------------------------------------------------------------------------
# Differential cross sections and their integrals (barn/atom)
# --------------------------------------------------------------------------------------

def _mu_nodes(n=400):
    """Gauss-Legendre nodes in s = sin^2(theta/2) mapped logarithmically: forward peaked
    integrands (F^2) need resolution at small angles."""
    # integrate over u = ln(1 - cos(theta) + eps)
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


def rayleigh_xs_barn(z: int, e_mev):
    """sigma_R(E) = pi r_e^2 int_{-1}^{1} (1+c^2) F(q)^2 dc ,  q = k sqrt(2(1-c))."""
    e = np.atleast_1d(np.asarray(e_mev, dtype=np.float64))
    x, w = _mu_nodes()
    eps = 1e-10
    lo, hi = np.log(eps), np.log(2.0 + eps)
    u = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
    omc = np.exp(u) - eps  # 1 - cos(theta)
    jac = 0.5 * (hi - lo) * np.exp(u)
    c = 1.0 - omc
    out = np.zeros_like(e)
    for i, ei in enumerate(e):
        k = K_PER_MEV * ei
        q = k * np.sqrt(2.0 * np.maximum(omc, 0.0))
        f = form_factor(z, q)
        out[i] = np.pi * R_E2_BARN * np.sum(w * jac * (1.0 + c * c) * f * f)
    return out


def klein_nishina_dsdc(e_mev, c):
    """d sigma_KN / d cos(theta) per electron in barn."""
    k = e_mev / ELECTRON_MASS_MEV
    r = 1.0 / (1.0 + k * (1.0 - c))  # E'/E
    return np.pi * R_E2_BARN * r * r * (r + 1.0 / r - (1.0 - c * c))


def compton_xs_barn(z: int, e_mev):
    """sigma_C(E) = int KN(E,c) S(q,Z) dc, q = k sqrt(2(1-c)) (momentum transfer of the
    elastic kinematics, as in the Livermore model). Also returns the mean fraction of the
    photon energy transferred to the electron (for kerma checks)."""
    e = np.atleast_1d(np.asarray(e_mev, dtype=np.float64))
    x, w = _mu_nodes()
    eps = 1e-10
    lo, hi = np.log(eps), np.log(2.0 + eps)
    u = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
    omc = np.exp(u) - eps
    jac = 0.5 * (hi - lo) * np.exp(u)
    c = 1.0 - omc
    sig = np.zeros_like(e)
    ftr = np.zeros_like(e)
    for i, ei in enumerate(e):
        k = K_PER_MEV * ei
        q = k * np.sqrt(2.0 * np.maximum(omc, 0.0))
        s = incoherent_function(z, q)
        d = klein_nishina_dsdc(ei, c) * s
        sig[i] = np.sum(w * jac * d)
        kk = ei / ELECTRON_MASS_MEV
        r = 1.0 / (1.0 + kk * omc)
        ftr[i] = np.sum(w * jac * d * (1.0 - r)) / max(sig[i], 1e-300)
    return sig, ftr


def pair_xs_barn(z: int, e_mev):
    """Bethe-Heitler total conversion cross-section (nuclear + electron field), Geant4
    parameterisation: sigma = (Z+1) Z [F1(X) + F2(X) Z + F3(X)/Z] microbarn, X = ln(E/mc2),
    with the (E-2mc2)^2 threshold behaviour below 1.5 MeV."""
    e = np.atleast_1d(np.asarray(e_mev, dtype=np.float64))
    a = [8.7842e2, -1.9625e3, 1.2949e3, -2.0028e2, 1.2575e1, -2.8333e-1]
    b = [-1.0342e1, 1.7692e1, -8.2381, 1.3063, -9.0815e-2, 2.3586e-3]
    c = [-4.5263e2, 1.1161e3, -8.6749e2, 2.1773e2, -2.0467e1, 6.5372e-1]
    e_lim = 1.5
    ee = np.maximum(e, e_lim)
    x = np.log(ee / ELECTRON_MASS_MEV)
    f1 = sum(a[i] * x**i for i in range(6))
    f2 = sum(b[i] * x**i for i in range(6))
    f3 = sum(c[i] * x**i for i in range(6))
    xs = (z + 1.0) * z * (f1 + f2 * z + f3 / z) * 1e-6  # barn
    thr = 2.0 * ELECTRON_MASS_MEV
    low = e < e_lim
    xx = np.where(low, (e - thr) / (e_lim - thr), 1.0)
    xs = xs * np.where(low, xx * xx, 1.0)
    xs = np.where(e <= thr, 0.0, xs)
    return np.maximum(xs, 0.0)


# --------------------------------------------------------------------------------------
# The per-element split on an arbitrary energy grid
# --------------------------------------------------------------------------------------

def barn_to_cm2_g(z: int) -> float:
    return 1e-24 * N_AVOGADRO / atomic_mass(z)


def nist_total_cm2_g(z: int, e_mev):
    """log-log linear interpolation of the master-grid NIST mu/rho."""
    d = nist()
    return np.exp(np.interp(np.log(e_mev), np.log(d["energy_MeV"]),
                            np.log(d["mu_over_rho_cm2_g"][z - 1].astype(np.float64))))


def nist_mu_en_cm2_g(z: int, e_mev):
    d = nist()
    return np.exp(np.interp(np.log(e_mev), np.log(d["energy_MeV"]),
                            np.log(np.maximum(d["mu_en_over_rho_cm2_g"][z - 1].astype(np.float64), 1e-300))))


_COARSE_N = 160


def _compute_element_models_coarse(z: int):
    """Compton / Rayleigh / pair on a coarse log grid (they are smooth in E)."""
    e = np.exp(np.linspace(np.log(1e-3), np.log(20.0), _COARSE_N))
    sc, ftr = compton_xs_barn(z, e)
    sr = rayleigh_xs_barn(z, e)
    sp = pair_xs_barn(z, e)
    return e, sc, ftr, sr, sp


@functools.lru_cache(maxsize=1)
def _element_models_file():
    """data/element_models_coarse.npz (tools/gen_element_models.py): the function above tabulated for Z = 1..100"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "element_models_coarse.npz")
    if not os.path.exists(path) or os.environ.get("GATE_B200_NO_MODEL_CACHE"):
        return None
    d = np.load(path)
    return {k: d[k] for k in d.files}


@functools.lru_cache(maxsize=128)
def _element_models_coarse(z: int):
    d = _element_models_file()
    if d is not None and f"sc_{z}" in d and len(d["energy_MeV"]) == _COARSE_N:
        return d["energy_MeV"], d[f"sc_{z}"], d[f"ftr_{z}"], d[f"sr_{z}"], d[f"sp_{z}"]
    return _compute_element_models_coarse(z)


def element_process_xs(z: int, e_mev):
    """Return dict of per-process mass cross-sections (cm2/g) at energies e_mev (MeV):
    'phot', 'compt', 'rayl', 'conv', 'total' (== NIST), 'compt_ftr' (mean energy-transfer
    fraction of Compton). The split is computed once per element on the master grid and
    interpolated linearly in ln E, so it does not depend on the requested grid."""
    m = _element_split_master(z)
    le = np.log(np.asarray(e_mev, dtype=np.float64))
    lm = np.log(nist()["energy_MeV"])
    out = {k: np.interp(le, lm, m[k]) for k in ("phot", "compt", "rayl", "conv", "total", "compt_ftr")}
    out["switch_energy"] = m["switch_energy"]
    return out


@functools.lru_cache(maxsize=128)
def _element_split_master(z: int):
    e_mev = np.asarray(nist()["energy_MeV"], dtype=np.float64)
    ec, sc, ftr, sr, sp = _element_models_coarse(z)
    le = np.log(e_mev)
    lec = np.log(ec)
    b2g = barn_to_cm2_g(z)

    def loglog(y):
        return np.exp(np.interp(le, lec, np.log(np.maximum(y, 1e-300))))

    compt = loglog(sc) * b2g
    rayl = loglog(sr) * b2g
    conv = pair_xs_barn(z, e_mev) * b2g
    f_tr = np.interp(le, lec, ftr)
    total = nist_total_cm2_g(z, e_mev)
    if z <= 10:
        # Light elements have no fluorescence or bremsstrahlung losses worth mentioning below
        # ~0.3 MeV, so NIST's second pinned table gives the photoelectric part directly:
        #   mu_en = phot + compt * <T/E>   =>   phot = mu_en - compt*f_tr ,
        # and Rayleigh takes what is left of the total.  Used only where it is well
        # conditioned (photoelectric share >= 2 %, Rayleigh within a factor 2 of the model).
        mu_en = nist_mu_en_cm2_g(z, e_mev)
        phot_c = mu_en - compt * f_tr
        rayl_c = total - compt - conv - phot_c
        good = (e_mev < 0.3) & (phot_c >= 0.02 * total) & (rayl_c > 0.5 * rayl) & (rayl_c < 2.0 * rayl)
        # keep one contiguous low-energy range so the tables stay smooth
        if good.any():
            last = np.where(good)[0].max()
            first_bad = np.where(~good[: last + 1])[0]
            lo = (first_bad.max() + 1) if first_bad.size else 0
            sel = np.zeros_like(good)
            sel[lo: last + 1] = True
            # above the range keep the model's shape but continue from the constrained value
            ratio_hi = rayl_c[last] / rayl[last]
            ratio_lo = rayl_c[lo] / rayl[lo]
            new_rayl = rayl.copy()
            new_rayl[sel] = rayl_c[sel]
            new_rayl[last + 1:] = rayl[last + 1:] * ratio_hi
            new_rayl[:lo] = rayl[:lo] * ratio_lo
            rayl = new_rayl
    resid = total - compt - rayl - conv
    share = resid / total
    # highest energy up to which the photoelectric residual is a solid (>=2%) share
    ok = share >= 0.02
    if not ok.any():
        ix = 0
    else:
        bad = np.where(~ok)[0]
        # first failure above the last absorption edge region: take first bad index after
        # the maximum of the share
        start = int(np.argmax(share))
        bad = bad[bad > start]
        ix = (bad[0] - 1) if bad.size else len(e_mev) - 1
    phot = resid.copy()
    if ix < len(e_mev) - 1:
        # local power-law slope estimated over the preceding decade/3
        j0 = max(0, np.searchsorted(e_mev, e_mev[ix] / 1.6))
        if j0 < ix and resid[j0] > 0 and resid[ix] > 0:
            p = np.log(resid[j0] / resid[ix]) / np.log(e_mev[ix] / e_mev[j0])
        else:
            p = 3.0
        p = float(np.clip(p, 1.0, 3.5))
        # the photoelectric slope flattens towards ~E^-1 at MeV energies
        tail = e_mev[ix + 1:]
        p_eff = np.maximum(1.0, p - 0.6 * np.log(tail / e_mev[ix]) / np.log(10.0))
        phot[ix + 1:] = max(resid[ix], 0.0) * (e_mev[ix] / tail) ** p_eff
        phot[: ix + 1] = np.maximum(phot[: ix + 1], 0.0)
        compt2 = total - rayl - conv - phot
        # the sub-percent model residual goes to Compton above the switch energy
        compt = np.where(np.arange(len(e_mev)) > ix, compt2, compt)
    phot = np.maximum(phot, 0.0)
    compt = np.maximum(compt, 0.0)
    # below the switch a negative residual (model overshoot) is clipped: renormalise the
    # scattering parts so the sum still equals the NIST total
    s = phot + compt + rayl + conv
    scale = np.where(s > 0, total / s, 1.0)
    return {
        "phot": phot * scale,
        "compt": compt * scale,
        "rayl": rayl * scale,
        "conv": conv * scale,
        "total": total,
        "compt_ftr": f_tr,
        "switch_energy": float(e_mev[ix]),
    }
