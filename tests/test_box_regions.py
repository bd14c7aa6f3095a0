"""Homogeneous Box regions of their own material in the world (a couch, a detector slab): known answers on the CPU
oracle; the GPU parity cases are in test_gpu_parity.py."""
import numpy as np
import pytest

import opengate_b200 as gate
from opengate_b200 import abi
from oracle.oracle import TransportOracle
from region_setups import slab_in_front

mm, m = gate.g4_units.mm, gate.g4_units.m


def run(S, n):
    o = TransportOracle(S.seed)
    ids = S.apply(o)
    o.begin_run()
    o.run_source(ids[0][0], 0, n, 0)
    return o, ids


@pytest.mark.parametrize("thickness,tilt", [(40.0, 0.0), (25.0, 30.0)])
def test_unscattered_fraction_behind_a_water_slab_is_the_attenuation_law(thickness, tilt):
    n, E = 40000, 0.2
    S = slab_in_front(n, thickness, tilt, energy=E)
    o, (sids, dids, pids) = run(S, n)
    before, behind = o.phsp(pids[0]), o.phsp(pids[1])
    assert o.stats()["phsp_dropped"] == 0
    n_in = (np.isclose(before["KineticEnergy"], E, rtol=1e-6) & (before["Direction"][:, 2] > 0.999999)).sum()
    assert n_in > 0.99 * n
    unscattered = np.isclose(behind["KineticEnergy"], E, rtol=1e-6) & (np.abs(behind["Direction"][:, 2] - 1.0) < 1e-6)
    T = S.tables
    ie = np.searchsorted(T.energy, E) - 1
    f = (np.log(E) - np.log(T.energy[ie])) / (np.log(T.energy[ie + 1]) - np.log(T.energy[ie]))
    mu_w = T.mu_total[0, ie] + f * (T.mu_total[0, ie + 1] - T.mu_total[0, ie])
    mu_a = T.mu_total[1, ie] + f * (T.mu_total[1, ie + 1] - T.mu_total[1, ie])
    path_w = thickness / np.cos(np.radians(tilt))
    expect = np.exp(-mu_w * path_w - mu_a * (190.0 - path_w))     # plane at -390 -> plane at -200
    got = unscattered.sum() / n_in
    assert got == pytest.approx(expect, abs=4 * np.sqrt(expect * (1 - expect) / n))
    # what interacts in the slab deposits energy that no dose actor sees
    st = o.stats()
    assert st["edep_total_MeV"] > o.dose_grid(dids[0], abi.GRID_EDEP).sum() * (1 + 1e-6)


def test_a_slab_of_world_material_changes_nothing_but_the_step_count():
    n = 5000
    a, _ = run(slab_in_front(n, material=1, planes=False), n)
    S = slab_in_front(n, material=1, planes=False)
    S.regions = []
    b, _ = run(S, n)
    sa, sb = a.stats(), b.stats()
    # the same physics, but a flight that crosses a face is cut in two (a fresh block of random numbers for the
    # second part): the histories differ, the expectation does not
    assert sa["steps"] > sb["steps"]
    assert sa["edep_total_MeV"] == pytest.approx(sb["edep_total_MeV"], rel=0.03)


def test_script_level_couch_under_the_ct(tmp_path):
    """a BoxVolume of its own material besides the Image volume is a region, not an error; overlaps are refused"""
    from test_host_api import ct_sim
    from oracle_backend import OracleBackend
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=3000)
    couch = sim.add_volume("Box", "couch")
    couch.size = [400 * mm, 400 * mm, 30 * mm]
    couch.translation = [0, 0, -250 * mm]          # between the source (z = -400 mm) and the CT
    couch.material = "G4_WATER"
    out = sim.run(_backend=OracleBackend)
    with_couch = dose.edep.image.array.sum()
    sim2, ct2, src2, dose2, _ = ct_sim(tmp_path / "b", n_primaries=3000)
    sim2.run(_backend=OracleBackend)
    assert 0.3 * dose2.edep.image.array.sum() < with_couch < 0.9 * dose2.edep.image.array.sum()   # 30 mm of water at ~70 keV
    assert out.device_stats[0]["histories"] == 3000
    couch.translation = [0, 0, 0]                  # now inside the CT
    with pytest.raises(Exception, match="overlaps"):
        sim.run(_backend=OracleBackend)


def test_detector_slab_with_phase_space_actor(tmp_path):
    """a Box of its own material that also carries a PhaseSpaceActor: scoring plane AND material region (the crossing
    is reported as the plane's event; the medium follows from the position)"""
    from test_host_api import ct_sim
    from oracle_backend import OracleBackend
    res = {}
    for mat in ("G4_AIR", "G4_WATER"):
        sim, ct, src, dose, hu = ct_sim(tmp_path / mat, n_primaries=4000)
        slab = sim.add_volume("Box", "detector")
        slab.size = [400 * mm, 400 * mm, 40 * mm]
        slab.translation = [0, 0, -250 * mm]
        slab.material = mat
        ph = sim.add_actor("PhaseSpaceActor", "phsp")
        ph.attached_to = slab
        ph.attributes = ["KineticEnergy", "PrePosition", "PreDirection", "EventID"]
        ph.steps_to_store = "entering exiting"
        ph.max_rows = 40000                      # entering + exiting + what the CT scatters back
        ph.output_filename = "det.npz"
        sim.run(_backend=OracleBackend)
        res[mat] = (ph, dose.edep.image.array.sum())
    (pa, ea), (pw, ew) = res["G4_AIR"], res["G4_WATER"]
    # every primary enters the slab from below in both cases; fewer leave it unscattered when it is made of water
    for ph in (pa, pw):
        z = ph.data["PrePosition_Z"]
        first = (np.abs(z + 270.0) < 0.1) & (ph.data["PreDirection_Z"] > 0.999999)
        assert 3975 <= first.sum() <= 4000        # a few interact in the 130 mm of air on the way
    def straight_out(ph):
        z = ph.data["PrePosition_Z"]
        return ((np.abs(z + 230.0) < 0.1) & (ph.data["PreDirection_Z"] > 0.999999)).sum()
    assert straight_out(pa) > 3980
    assert 0.3 * 4000 < straight_out(pw) < 0.6 * 4000      # 40 mm of water at 40 ... 120 keV
    assert ew < 0.75 * ea


def test_couch_touching_the_phantom():
    """a slab that lies flat on the face of the phantom (its exit face IS the entry face of the image volume)"""
    n = 20000
    S = slab_in_front(n, 30.0, 0.0, energy=0.1, touching=True)
    o, (sids, dids, pids) = run(S, n)
    R = slab_in_front(n, 30.0, 0.0, energy=0.1, touching=True)
    R.regions = []
    r, (_, rd, _) = run(R, n)
    a, b = o.dose_grid(dids[0], abi.GRID_EDEP), r.dose_grid(rd[0], abi.GRID_EDEP)
    T = S.tables
    ie = np.searchsorted(T.energy, 0.1) - 1
    # the primary fluence that reaches the water box is attenuated by 30 mm of water; scatter from the couch adds some
    att = np.exp(-T.mu_total[0, ie] * 30.0)
    assert att * 0.98 < a.sum() / b.sum() < att * 1.35
    assert o.stats()["histories"] == n and o.stats()["escaped"] > 0
