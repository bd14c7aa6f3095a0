"""The Geant4 table exporter (integration/GateB200TableExporter.cpp) compiled and run against the
stand-in Geant4 headers of integration/g4shim: the exporter -> adaptor (integration/geant4_tables.py) ->
`build_photon_tables(element_xs=...)` seam end to end.  The fake G4EmCalculator answers from the in-tree
per-process numbers, so the tables that come out through the exporter must be the ones the native path builds;
a second set of numbers (one process scaled) must change exactly the tables it should."""
import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INTEG = os.path.join(ROOT, "integration")
PROC = {"phot": "phot", "compt": "compt", "rayl": "Rayl", "conv": "conv"}    # our key -> Geant4 process name


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    pybind11 = pytest.importorskip("pybind11")
    d = tmp_path_factory.mktemp("g4shim")
    out = d / ("opengate_core_g4shim" + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(INTEG, "g4shim"), "-I", pybind11.get_include(),
           "-I", sysconfig.get_paths()["include"], os.path.join(INTEG, "g4shim", "shim_module.cpp"),
           os.path.join(INTEG, "GateB200TableExporter.cpp"), "-o", str(out)]
    subprocess.check_call(cmd)
    spec = importlib.util.spec_from_file_location("opengate_core_g4shim", out)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _adaptor():
    spec = importlib.util.spec_from_file_location("geant4_tables", os.path.join(INTEG, "geant4_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _register(shim, zs, grid, scale=None):
    from opengate_b200 import physics_models as pm
    for z in zs:
        xs = pm.element_process_xs(z, grid)
        table = {g4: [float(v) * (scale or {}).get(k, 1.0) for v in xs[k]] for k, g4 in PROC.items()}
        shim.ShimRegisterElement(int(z), 1.0 + 0.1 * z, [float(x) for x in grid], table)


def test_exporter_returns_the_registered_cross_sections(shim):
    from opengate_b200 import physics_models as pm
    grid = np.exp(np.linspace(np.log(1e-3), np.log(10.0), 400))
    _register(shim, [1, 8, 20], grid)
    d = shim.GateB200ExportElementPhotonCrossSections(8, [float(x) for x in grid[::7]])
    want = pm.element_process_xs(8, grid[::7])
    for k in ("phot", "compt", "rayl", "conv"):
        np.testing.assert_allclose(np.asarray(d[k]), want[k], rtol=1e-12, atol=0)
    np.testing.assert_allclose(np.asarray(d["total"]), want["phot"] + want["compt"] + want["rayl"] + want["conv"], rtol=1e-12)
    with pytest.raises(RuntimeError, match="unknown element"):
        shim.GateB200ExportElementPhotonCrossSections(92, [0.1])


def test_tables_through_the_exporter_equal_the_native_tables(shim):
    from opengate_b200 import physics, setups
    mats, _ = setups._materials_from_db(["G4_WATER", "G4_AIR", "G4_BONE_COMPACT_ICRU"])
    native = physics.build_photon_tables(mats, e_max_mev=0.5)
    zs = sorted({int(z) for m in mats for z in m.fractions})
    _register(shim, zs, native.energy)                     # the table's own nodes: the shim answers them exactly
    g4t = _adaptor()
    via = physics.build_photon_tables(mats, e_max_mev=0.5, element_xs=lambda z, e: g4t.geant4_element_xs(z, e, g4=shim))
    for k in ("mu_total", "cum_phot", "cum_compt", "cum_rayl", "rayl_cum_ff2", "compt_sf", "density"):
        a, b = getattr(native, k), getattr(via, k)
        np.testing.assert_allclose(b, a, rtol=2e-6, atol=0, err_msg=k)     # float32 tables of double sums in another order
    # sum of the processes = NIST total (what pins the native split) survives the trip
    from opengate_b200 import physics_models as pm
    e = native.energy
    water = mats[0].normalised()
    nist = sum(w * pm.nist_total_cm2_g(z, e) for z, w in water.items()) * float(mats[0].density_g_cm3) * 0.1
    np.testing.assert_allclose(via.mu_total[0], nist, rtol=2e-3)


def test_a_different_physics_list_changes_the_tables_it_should(shim):
    """Photoelectric cross-sections 20 % higher (a stand-in for another EM constructor, physics.py:300-337): total
    attenuation and the process probabilities follow; the angular tables do not."""
    from opengate_b200 import physics, setups
    mats, _ = setups._materials_from_db(["G4_WATER"])
    native = physics.build_photon_tables(mats, e_max_mev=0.2)
    _register(shim, [1, 8], native.energy, scale={"phot": 1.2})
    g4t = _adaptor()
    via = physics.build_photon_tables(mats, e_max_mev=0.2, element_xs=lambda z, e: g4t.geant4_element_xs(z, e, g4=shim))
    ph_n = native.process_mu["phot"][0]
    np.testing.assert_allclose(via.process_mu["phot"][0], 1.2 * ph_n, rtol=1e-9)
    np.testing.assert_allclose(via.mu_total[0], native.mu_total[0].astype(np.float64) + 0.2 * ph_n, rtol=1e-5)
    assert np.all(via.cum_phot[0] >= native.cum_phot[0])
    assert np.array_equal(via.rayl_cum_ff2, native.rayl_cum_ff2) and np.array_equal(via.compt_sf, native.compt_sf)
