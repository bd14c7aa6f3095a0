"""mu/rho and mu_en/rho per material from the element database — the tables the TLEDoseActor uses.

Restates, in numpy/float64, the in-tree reference code of the attenuation database:

* `GateMaterialMuHandler::InitElementTable` (GateMaterialMuHandler.cpp:265-295): one raw
  (energy, mu/rho, mu_en/rho) table per element Z from `GateMuDatabase.h` ("NIST" or "EPDL");
* `GateMaterialMuHandler::ConstructMaterial` (GateMaterialMuHandler.cpp:159-263): union of the element
  energy grids (an energy doubled inside ONE element table is an absorption edge and is kept twice,
  the same energy coming from two tables is kept once), mass-fraction mix with log-log interpolation
  (`interpolation`, :108-111) where an element has no node, storage as natural logarithms;
* `GateMuTable::GetMuOverRho / GetMuEnOverRho` (GateMuTables.cpp:47-107): binary search on ln E,
  linear interpolation of the stored logarithms, value of the lower node outside the open interval.

The element data live in data/mu_database.npz (tools/gen_mu_database.py).  The restatement is checked
against the reference's own GateMuTables.cpp compiled into oracle/_ref (tests/test_mu_handler.py).
"""

from __future__ import annotations

import math
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mu_database.npz")
_cache = {}


def element_tables(database="EPDL"):
    """dict Z -> (energy[MeV], mu/rho, mu_en/rho) raw arrays (InitElementTable)."""
    if database not in ("NIST", "EPDL"):
        raise ValueError(f"GateMaterialMuHandler -- mu/mu_en database option '{database}' doesn't exist. "
                         "Available database are 'NIST' and 'EPDL'")
    if database not in _cache:
        z = np.load(_DATA)
        counts, data = z[database + "_counts"], z[database + "_data"]
        out, k = {}, 0
        for zel, n in enumerate(counts):
            out[zel] = (data[k:k + n, 0].copy(), data[k:k + n, 1].copy(), data[k:k + n, 2].copy())
            k += n
        _cache[database] = out
    return _cache[database]


def _interpolation(xa, xb, ya, yb, x):
    # GateMaterialMuHandler.cpp:108-111
    return math.exp(math.log(ya) + math.log(yb / ya) / math.log(xb / xa) * math.log(x / xa))


class MuTable:
    """GateMuTable: ln(E), ln(mu/rho), ln(mu_en/rho) nodes + the two look-ups."""

    def __init__(self, ln_e, ln_mu, ln_mu_en, density_g_cm3=-1.0):
        self.ln_e = np.asarray(ln_e, float)
        self.ln_mu = np.asarray(ln_mu, float)
        self.ln_mu_en = np.asarray(ln_mu_en, float)
        self.density = float(density_g_cm3)

    def _lookup(self, values, energy):
        le = math.log(energy)
        inf, sup = 0, len(self.ln_e) - 1
        while sup - inf > 1:
            mid = (inf + sup) // 2
            if self.ln_e[mid] > le:
                sup = mid
            else:
                inf = mid
        e_inf, e_sup = self.ln_e[inf], self.ln_e[sup]
        if e_inf < le < e_sup:
            return math.exp(values[inf] + (values[sup] - values[inf]) * (le - e_inf) / (e_sup - e_inf))
        return math.exp(values[inf])

    def mu_over_rho(self, energy_mev):
        return self._lookup(self.ln_mu, energy_mev)

    def mu_en_over_rho(self, energy_mev):
        return self._lookup(self.ln_mu_en, energy_mev)

    def mu(self, energy_mev):
        return self.mu_over_rho(energy_mev) * self.density

    def mu_en(self, energy_mev):
        return self.mu_en_over_rho(energy_mev) * self.density


def construct_material(elements, fractions, density_g_cm3=-1.0, database="EPDL"):
    """ConstructMaterial for elements [Z...] with mass fractions [w...] (in the material's order)."""
    tabs = element_tables(database)
    ne = len(elements)
    e_t = [tabs[int(z)][0] for z in elements]
    mu_t = [tabs[int(z)][1] for z in elements]
    en_t = [tabs[int(z)][2] for z in elements]
    size = [len(t) for t in e_t]
    nb_e = sum(size)
    energies = [0.0] * nb_e
    index = [0] * ne
    i = 0
    while i < nb_e:
        mt = 0
        while index[mt] >= size[mt]:
            mt += 1
        for j in range(mt + 1, ne):
            if index[j] < size[j] and e_t[j][index[j]] < e_t[mt][index[mt]]:
                mt = j
        energies[i] = e_t[mt][index[mt]]
        if i > 0 and energies[i] == energies[i - 1]:
            if index[mt] > 0 and e_t[mt][index[mt]] == e_t[mt][index[mt] - 1]:
                pass
            else:
                i -= 1
                nb_e -= 1
        index[mt] += 1
        i += 1
    energies = energies[:nb_e]
    mu = [0.0] * nb_e
    mu_en = [0.0] * nb_e
    index = [0] * ne
    for i in range(nb_e):
        cur = energies[i]
        for j in range(ne):
            if e_t[j][index[j]] < cur:
                index[j] += 1
            k = index[j]
            if e_t[j][k] == cur:
                mu[i] += fractions[j] * mu_t[j][k]
                mu_en[i] += fractions[j] * en_t[j][k]
                if i != nb_e - 1 and k + 1 < size[j] and e_t[j][k] == e_t[j][k + 1]:
                    index[j] += 1
            else:
                mu[i] += fractions[j] * _interpolation(e_t[j][k - 1], e_t[j][k], mu_t[j][k - 1], mu_t[j][k], cur)
                mu_en[i] += fractions[j] * _interpolation(e_t[j][k - 1], e_t[j][k], en_t[j][k - 1], en_t[j][k], cur)
    return MuTable(np.log(energies), np.log(mu), np.log(mu_en), density_g_cm3)


def material_table(material, database="EPDL"):
    """MuTable of a `physics.MaterialSpec` (density g/cm3, fractions Z -> mass fraction)."""
    fr = material.normalised()
    zs = list(fr.keys())
    return construct_material(zs, [fr[z] for z in zs], material.density_g_cm3, database)


def mu_en_table_on_grid(materials, energy_mev, database="EPDL"):
    """[n_materials][n_energy] float32 mu_en/rho (cm2/g): the reference look-up evaluated on the
    device's energy nodes (what `gate_b200_tle_desc.mu_en_over_rho` uploads)."""
    out = np.zeros((len(materials), len(energy_mev)), np.float32)
    for i, m in enumerate(materials):
        t = material_table(m, database)
        out[i] = [t.mu_en_over_rho(float(e)) for e in energy_mev]
    return out
