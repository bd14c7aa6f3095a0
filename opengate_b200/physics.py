"""Per-material photon tables for the device (the "table export" of BASELINE.json's north_star).

Reference seam: `opengate/physics.py` builds the Geant4 physics list
(`create_reference_physics_list_class`, physics.py:300-337) and Geant4 owns the cross-sections.
For the GPU engine the same information must exist as energy look-up tables per material.
`build_photon_tables` produces them; its source of per-element cross-sections is pluggable:

* default: `physics_models.element_process_xs` (NIST totals pinned in the reference tree +
  analytic split, see that module);
* where Geant4 exists: an exporter following `GateMaterialMuHandler::SimulateMaterialTable`
  (GateMaterialMuHandler.cpp:296-549) that fills the same `PhotonTables`.

Mixture rule = mass-fraction weighting of mu/rho, as `GateMaterialMuHandler::ConstructMaterial`
(GateMaterialMuHandler.cpp:159-263).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import physics_models as pm


@dataclass
class MaterialSpec:
    """A material as the engine needs it: density (g/cm3) and element mass fractions."""

    name: str
    density_g_cm3: float
    fractions: dict  # Z (int) -> mass fraction

    def normalised(self):
        s = sum(self.fractions.values())
        return {int(z): float(w) / s for z, w in self.fractions.items() if w > 0}


@dataclass
class PhotonTables:
    """Host copy of everything `gate_b200_set_tables` uploads (float32, C-contiguous)."""

    names: list
    density: np.ndarray          # [M]
    e_min: float
    e_max: float
    e_cut: float
    n_energy: int
    mu_total: np.ndarray         # [M][NE] 1/mm
    cum_phot: np.ndarray
    cum_compt: np.ndarray
    cum_rayl: np.ndarray
    n_q: int
    q2_min: float
    q2_max: float
    rayl_cum_ff2: np.ndarray     # [M][NQ]
    compt_sf: np.ndarray         # [M][NQ]
    pair_mode: int = 0
    # float64 extras for tests / TLE-style checks (not uploaded)
    energy: np.ndarray = field(default=None, repr=False)
    mu_en_over_rho: np.ndarray = field(default=None, repr=False)   # [M][NE] cm2/g (NIST)
    process_mu: dict = field(default=None, repr=False)             # name -> [M][NE] 1/mm

    def material_index(self, name):
        return self.names.index(name)

    def mu_per_mm(self, material, energy_mev):
        """Total linear attenuation with the same interpolation as the device."""
        m = material if isinstance(material, int) else self.material_index(material)
        t = (np.log(energy_mev) - np.log(self.e_min)) / (np.log(self.e_max) - np.log(self.e_min)) * (self.n_energy - 1)
        i = np.clip(np.floor(t).astype(int), 0, self.n_energy - 2)
        f = t - i
        row = self.mu_total[m].astype(np.float64)
        return row[i] + f * (row[i + 1] - row[i])


def energy_grid(e_min, e_max, n):
    g = np.exp(np.linspace(np.log(e_min), np.log(e_max), n))
    g[0], g[-1] = e_min, e_max
    return g


def q2_grid(q2_min, q2_max, n):
    return np.exp(np.log(q2_min) + (np.log(q2_max) - np.log(q2_min)) * np.arange(n) / (n - 1))


def build_photon_tables(materials, e_max_mev, e_cut_mev=1.0e-3, n_energy=1024, n_q=256,
                        pair_mode=0, element_xs=None):
    """Build the device tables for a list of MaterialSpec (index in the list = label).

    e_max_mev: highest photon energy that can occur (source maximum); the grid spans
    [e_cut, e_max*1.0001].  NIST data cover 1 keV .. 20 MeV.
    """
    if element_xs is None:
        element_xs = pm.element_process_xs
    e_cut = max(float(e_cut_mev), 1.0e-3)
    e_hi = min(float(e_max_mev) * 1.0001, 20.0)
    if e_hi <= e_cut * 1.01:
        e_hi = e_cut * 1.01
    e = energy_grid(e_cut, e_hi, n_energy)
    q2_min = 1.0e-4
    q2_max = max((2.0 * pm.K_PER_MEV * e_hi) ** 2 * 1.001, 1.0)
    q2 = q2_grid(q2_min, q2_max, n_q)

    zs = sorted({int(z) for m in materials for z in m.fractions})
    xs = {z: element_xs(z, e) for z in zs}
    mu_en_el = {z: pm.nist_mu_en_cm2_g(z, e) if z <= 92 else np.zeros_like(e) for z in zs}

    # squared form factor / scattering function on a refined q grid (8 sub-steps per bin)
    sub = 8
    lq = np.log(q2)
    fine = np.concatenate([np.linspace(lq[j], lq[j + 1], sub, endpoint=False) for j in range(n_q - 1)] + [lq[-1:]])
    q2f = np.exp(fine)
    qf = np.sqrt(q2f)
    ff2_el = {z: pm.form_factor(z, qf) ** 2 for z in zs}
    sf_el = {z: pm.incoherent_function(z, qf) for z in zs}
    ff2_0 = {z: pm.form_factor(z, np.sqrt(np.linspace(0, q2_min, 16))) ** 2 for z in zs}

    M = len(materials)
    dens = np.zeros(M, np.float32)
    mu = np.zeros((M, n_energy))
    proc = {k: np.zeros((M, n_energy)) for k in ("phot", "compt", "rayl", "conv")}
    muen = np.zeros((M, n_energy))
    ray = np.zeros((M, n_q))
    sfm = np.zeros((M, n_q))
    for i, m in enumerate(materials):
        fr = m.normalised()
        rho = float(m.density_g_cm3)
        dens[i] = rho
        to_mm = rho * 0.1  # cm2/g * g/cm3 = 1/cm -> 1/mm
        ff2 = np.zeros_like(q2f)
        sfv = np.zeros_like(q2f)
        ff0 = np.zeros(16)
        sinf = 0.0
        for z, w in fr.items():
            for k in proc:
                proc[k][i] += w * xs[z][k] * to_mm
            muen[i] += w * mu_en_el[z]
            per_g = w / pm.atomic_mass(z)
            ff2 += per_g * ff2_el[z]
            ff0 += per_g * ff2_0[z]
            sfv += per_g * sf_el[z]
            sinf += per_g * z
        mu[i] = proc["phot"][i] + proc["compt"][i] + proc["rayl"][i] + proc["conv"][i]
        # cumulative int_0^{q2} F^2 d(q2): first segment [0, q2_min] then trapezoids
        a0 = np.trapezoid(ff0, np.linspace(0, q2_min, 16))
        seg = 0.5 * (ff2[1:] + ff2[:-1]) * np.diff(q2f)
        cum = a0 + np.concatenate([[0.0], np.cumsum(seg)])
        ray[i] = cum[::sub][:n_q]
        sfm[i] = np.clip(sfv[::sub][:n_q] / sinf, 0.0, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        p_ph = np.where(mu > 0, proc["phot"] / mu, 1.0)
        p_co = np.where(mu > 0, proc["compt"] / mu, 0.0)
        p_ra = np.where(mu > 0, proc["rayl"] / mu, 0.0)
    c1 = p_ph
    c2 = np.clip(p_ph + p_co, 0.0, 1.0)
    c3 = np.clip(p_ph + p_co + p_ra, 0.0, 1.0)
    # where there is no conversion the third cumulative must be exactly 1 in float32
    no_conv = proc["conv"] <= 0.0
    c3 = np.where(no_conv, 1.0, c3)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    return PhotonTables(
        names=[m.name for m in materials], density=dens, e_min=float(e[0]), e_max=float(e[-1]),
        e_cut=e_cut, n_energy=n_energy, mu_total=f32(mu), cum_phot=f32(c1), cum_compt=f32(c2),
        cum_rayl=f32(c3), n_q=n_q, q2_min=q2_min, q2_max=float(q2_max), rayl_cum_ff2=f32(ray),
        compt_sf=f32(sfm), pair_mode=int(pair_mode), energy=e, mu_en_over_rho=muen,
        process_mu=proc,
    )
