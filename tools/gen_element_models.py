#!/usr/bin/env python
"""Generate opengate_b200/data/element_models_coarse.npz: the per-element Compton / Rayleigh / pair cross
sections of physics_models.py (_element_models_coarse: integrals of Klein-Nishina x S(q,Z), F(q,Z)^2 and the
pair model on a 160-node log grid, 1 keV - 20 MeV) for Z = 1..100, so that building the photon tables of a
simulation does not start with 0.7 s of numerical integration.  Pure function of Z: the test
tests/test_physics_tables.py recomputes a few elements and compares."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from opengate_b200 import physics_models as pm  # noqa: E402


def main():
    out = {}
    grid = None
    for z in range(1, 101):
        e, sc, ftr, sr, sp = pm._compute_element_models_coarse(z)
        grid = e
        out[f"sc_{z}"], out[f"ftr_{z}"], out[f"sr_{z}"], out[f"sp_{z}"] = sc, ftr, sr, sp
    path = os.path.join(os.path.dirname(os.path.abspath(pm.__file__)), "data", "element_models_coarse.npz")
    np.savez_compressed(path, energy_MeV=grid, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
