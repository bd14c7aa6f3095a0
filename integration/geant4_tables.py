"""Adaptor between the Geant4 exporter (GateB200TableExporter.cpp, built into opengate_core) and the table
builder of the B200 engine.  Needs `opengate_core` in production; the CPU test tier runs it against the exporter compiled with the Geant4
shim headers of integration/g4shim (a fake G4EmCalculator that answers from registered tables).

    from opengate_b200 import physics
    from physics_b200 import geant4_element_xs
    tables = physics.build_photon_tables(materials, e_max_mev, element_xs=geant4_element_xs)

`build_photon_tables` calls `element_xs(z, e_mev)` once per element present and mixes by mass fraction; the
keys are those of `opengate_b200.physics_models.element_process_xs`.  The form factors / scattering functions
(Rayleigh and Compton angular sampling) still come from `physics_models.form_factor` /
`incoherent_function`; replacing them by the G4EMLOW tabulations (`$G4LEDATA/livermore/rayl`, `.../comp`) is
the remaining step of INTEGRATION.md section 3.
"""
import numpy as np


def geant4_element_xs(z, e_mev, g4=None):
    """`element_xs` callback of build_photon_tables fed by the exporter.  g4: the module that carries
    GateB200ExportElementPhotonCrossSections -- `opengate_core` in a GATE installation (the default); the CPU test
    tier passes the exporter compiled against integration/g4shim (tests/test_g4_exporter_shim.py)."""
    if g4 is None:
        import opengate_core as g4  # the reference's pybind11 module, with init_GateB200TableExporter registered
    from opengate_b200 import physics_models as pm

    e = np.asarray(e_mev, dtype=np.float64)
    d = g4.GateB200ExportElementPhotonCrossSections(int(z), [float(x) for x in e])
    out = {k: np.asarray(d[k], dtype=np.float64) for k in ("phot", "compt", "rayl", "conv", "total")}
    # mean energy-transfer fraction of Compton: a property of the Klein-Nishina (+ binding) final state the
    # device samples itself; only reported by the table builder (mu_en of the analog path)
    own = pm.element_process_xs(int(z), e)
    out["compt_ftr"] = own["compt_ftr"]
    out["switch_energy"] = own["switch_energy"]
    return out
