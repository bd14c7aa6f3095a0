"""The drop-in layer: a GATE-10-style script (Simulation / add_volume('Image') / GenericSource /
VoxelSource / DoseActor / PhaseSpaceActor / SimulationStatisticsActor) run through
GpuSimulationEngine.  The host logic is exercised on CPU by injecting the oracle as the
transport back end (tests/oracle_backend.py); the same scripts run on the CUDA engine in
tests/test_gpu_parity.py::test_script_level_*."""
import json
import os

import numpy as np
import pytest

import opengate_b200 as gate
from opengate_b200 import image as gimage
from opengate_b200.base import GateDeprecationError
from oracle_backend import OracleBackend

mm, cm, m = gate.g4_units.mm, gate.g4_units.cm, gate.g4_units.m
keV, MeV, Bq, sec, deg = gate.g4_units.keV, gate.g4_units.MeV, gate.g4_units.Bq, gate.g4_units.s, gate.g4_units.deg


def make_ct(tmp_path, n=(24, 20, 16), spacing=(10.0, 10.0, 12.0)):
    nx, ny, nz = n
    hu = np.full((nz, ny, nx), -1000, dtype=np.int16)
    hu[2:-2, 3:-3, 3:-3] = 40
    hu[4:-4, 6:-6, 6:10] = -750
    hu[4:-4, 8:12, 14:17] = 900
    img = gimage.Image(hu, spacing, origin=(-50.0, 20.0, 7.0))
    p = tmp_path / "ct.mhd"
    gimage.write_image(img, str(p))
    return str(p), hu


def ct_sim(tmp_path, n_primaries=4000, **kw):
    path, hu = make_ct(tmp_path)
    sim = gate.Simulation()
    sim.random_seed = 987654321
    sim.output_dir = str(tmp_path / "out")
    sim.world.size = [2 * m, 2 * m, 2 * m]
    ct = sim.add_volume("Image", "ct")
    ct.image = path
    ct.material = "G4_AIR"
    tol = 0.05 * gate.g4_units.g_cm3
    ct.voxel_materials, _ = gate.geometry.materials.HounsfieldUnit_to_material(sim, tol)
    ct.translation = [5 * mm, -10 * mm, 0]
    sim.physics_manager.physics_list_name = "G4EmStandardPhysics_option4"
    sim.physics_manager.set_production_cut("world", "all", 1 * m)
    sim.physics_manager.set_production_cut("ct", "all", 1 * m)
    src = sim.add_source("GenericSource", "xray")
    src.particle = "gamma"
    src.energy.type = "spectrum_histogram"
    src.energy.spectrum_energy_bin_edges = [40 * keV, 60 * keV, 80 * keV, 120 * keV]
    src.energy.spectrum_weights = [1.0, 2.0, 0.5]
    src.position.type = "disc"
    src.position.radius = 30 * mm
    src.position.translation = [0, 0, -400 * mm]
    src.direction.type = "momentum"
    src.direction.momentum = [0, 0, 1]
    src.number_of_primaries = n_primaries
    dose = sim.add_actor("DoseActor", "dose")
    dose.attached_to = "ct"
    dose.size = "like_image_volume"
    dose.spacing = "like_image_volume"
    dose.hit_type = "random"
    dose.output_filename = "ct.mhd"
    return sim, ct, src, dose, hu


def rotated_ct_sim(tmp_path, n_primaries=4000):
    """CT volume rotated and shifted in the world, source attached to it (position and direction
    follow the volume), coarse dose grid with its own size / spacing / translation."""
    from scipy.spatial.transform import Rotation
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=n_primaries)
    ct.rotation = Rotation.from_euler("zx", [30, 10], degrees=True).as_matrix()
    ct.translation = [40 * mm, -25 * mm, 15 * mm]
    src.attached_to = "ct"
    src.direction_relative_to_attached_volume = True
    src.position.translation = [10 * mm, 0, -140 * mm]
    src.position.rotation = Rotation.from_euler("y", 5, degrees=True).as_matrix()
    dose.size = [12, 10, 8]
    dose.spacing = [20 * mm, 20 * mm, 24 * mm]
    dose.translation = [10 * mm, -5 * mm, 12 * mm]
    dose.hit_type = "random"
    return sim, ct, src, dose, hu


def test_rigid_motion_of_the_whole_setup_leaves_the_dose_unchanged(tmp_path):
    """Moving phantom + attached source + dose grid together must not change the physics: same
    histories up to float rounding of the frame changes."""
    from scipy.spatial.transform import Rotation
    imgs = []
    for k, (rot, tr) in enumerate(((np.eye(3), [0, 0, 0]),
                                   (Rotation.from_euler("zx", [30, 10], degrees=True).as_matrix(), [40.0, -25.0, 15.0]))):
        sim, ct, src, dose, hu = rotated_ct_sim(tmp_path / f"r{k}", n_primaries=6000)
        ct.rotation, ct.translation = rot, tr
        sim.run(_backend=OracleBackend)
        imgs.append(dose.edep.image.array)
    a, b = imgs
    assert a.sum() > 0 and b.sum() == pytest.approx(a.sum(), rel=2e-3)
    assert np.corrcoef(a.ravel(), b.ravel())[0, 1] > 0.999


def test_ct_dose_script_end_to_end(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    dose.dose.active = True
    dose.dose_uncertainty.active = True
    dose.edep_uncertainty.active = True
    dose.density.active = True
    dose.counts.active = True
    stats = sim.add_actor("SimulationStatisticsActor", "stats")
    out = sim.run(_backend=OracleBackend)
    assert out.current_random_seed == 987654321 and out.expected_number_of_events == 4000
    assert stats.counts["events"] == 4000 and stats.counts["runs"] == 1 and stats.pps > 0
    e = dose.edep.image
    assert list(e.size) == [24, 20, 16] and np.allclose(e.spacing, [10, 10, 12])
    # local coordinate system: centred on the volume (doseactors.py:170-232)
    assert np.allclose(e.origin, [-115.0, -95.0, -90.0])
    assert e.array.dtype == np.float64 and e.array.sum() > 0
    # edep only where the beam passes; nothing scored outside the grid
    st = out.device_stats[0]
    assert e.array.sum() <= st["edep_total_MeV"] * (1 + 1e-9)
    # dose [Gy] = sum(edep_i / rho_i) / (voxel volume) in Gy: check one homogeneous voxel by hand
    d = dose.dose.image.array
    dens = dose.density.image.array
    vox_mm3 = 10 * 10 * 12
    MeV_per_g_to_Gy = 1.602176634e-10
    body = (hu == 40) & (e.array > 0)
    k = np.argwhere(body)[0]
    want = e.array[tuple(k)] / (dens[tuple(k)] * vox_mm3 * 1e-3) * MeV_per_g_to_Gy
    assert d[tuple(k)] == pytest.approx(want, rel=1e-6)
    # uncertainty images follow the reference algebra on (sum, sum of squares, N)
    sq = dose.user_output["edep_with_uncertainty"].get_data("merged", 1).array
    n = 4000
    var = np.clip((sq / n - (e.array / n) ** 2) / (n - 1), 0, None)
    want_u = np.divide(np.sqrt(var), e.array / n, out=np.zeros_like(var), where=e.array != 0)
    np.testing.assert_allclose(dose.edep_uncertainty.image.array, want_u, rtol=1e-12)
    assert np.all(dose.counts.image.array[e.array > 0] >= 1)
    # files: names get the item suffix; squared items that were only activated for the uncertainty
    # are not written (doseactors.py:575-589); sidecar json carries number_of_samples
    files = sorted(os.listdir(tmp_path / "out"))
    for f in ("ct_edep.mhd", "ct_edep_uncertainty.mhd", "ct_dose.mhd", "ct_dose_uncertainty.mhd", "ct_density.mhd", "ct_counts.mhd"):
        assert f in files, (f, files)
    assert "ct_edep_squared.mhd" not in files and "ct_dose_squared.mhd" not in files
    assert json.load(open(tmp_path / "out" / "ct_edep.json")) == {"number_of_samples": 4000}
    back = gimage.read_image(str(dose.edep.get_output_path()))
    np.testing.assert_array_equal(back.array, e.array)
    assert str(dose.edep.get_output_path()).endswith("out/ct_edep.mhd")


def test_label_volume_is_integer_exact_and_placed_like_the_reference(tmp_path):
    from opengate_b200.engines import GpuSimulationEngine
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    sim.resolve_and_validate_config()
    S = GpuSimulationEngine(sim).build_setup()
    # label rule: np.digitize over the Schneider intervals, label 0 = volume material, in order of HU
    lut = ct.material_to_label_lut
    names = list(lut)
    assert names[0] == "G4_AIR" and lut["Air_0"] == 1
    vm = sorted(ct.voxel_materials, key=lambda r: r[0])
    want = np.zeros_like(hu, dtype=np.uint16)
    for lo, hi, name in vm:
        want[(hu >= lo) & (hu < hi)] = lut[name]
    # the engine hands the HU image + the digitize table to the device; the host restatement of the same rule
    assert S.labels is None and S.hu is not None and S.label_bytes == 1
    lab = S.label_array()
    assert np.array_equal(lab.astype(np.uint16), want)
    assert lab.dtype == np.uint8 and len(S.tables.names) == 75
    # the input image origin is NOT used for placement: the box is centred on `translation`
    assert np.allclose(S.world_to_volume[:, 3], [-5, 10, 0]) and np.allclose(S.world_to_volume[:, :3], np.eye(3))
    # output_coordinate_system = attached_to_image uses the image's own origin
    dose.output_coordinate_system = "attached_to_image"
    o, r = dose.output_origin_rotation()
    assert np.allclose(o, [-50.0, 20.0, 7.0])


def test_two_runs_and_run_merging(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    sim.run_timing_intervals = [[0, 1 * sec], [1 * sec, 2 * sec]]
    src.number_of_primaries = [1500, 2500]
    dose.edep_squared.active = True
    dose.edep.keep_data_per_run = True
    out = sim.run(_backend=OracleBackend)
    assert [s["histories"] for s in out.device_stats] == [1500, 2500]
    o = dose.user_output["edep_with_uncertainty"]
    assert o.merged_samples == 4000
    r0, r1 = dose.edep.get_data(0).array, dose.edep.get_data(1).array
    np.testing.assert_allclose(dose.edep.image.array, r0 + r1, rtol=1e-14)
    # the same 4000 histories in ONE run give the same total (history ids continue across runs)
    sim2, ct2, src2, dose2, _ = ct_sim(tmp_path)
    sim2.output_dir = str(tmp_path / "out2")
    sim2.run(_backend=OracleBackend)
    np.testing.assert_allclose(dose2.edep.image.array, dose.edep.image.array, rtol=1e-12)


def test_voxel_source_and_phase_space_plane(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    sim.source_manager.sources.clear()
    act = np.zeros((8, 10, 12), dtype=np.float32)
    act[3:5, 4:6, 5:8] = 1.0
    gimage.write_image(gimage.Image(act, (20.0, 20.0, 24.0)), str(tmp_path / "act.mhd"))
    vs = sim.add_source("VoxelSource", "tc99m")
    vs.attached_to = "ct"
    vs.image = str(tmp_path / "act.mhd")
    vs.particle = "gamma"
    vs.energy.mono = 140.511 * keV
    vs.direction.type = "iso"
    vs.number_of_primaries = 6000
    plane = sim.add_volume("Box", "plane")
    plane.size = [600 * mm, 2 * mm, 600 * mm]
    plane.translation = [0, 300 * mm, 0]
    plane.material = "G4_AIR"
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "PrePosition", "PreDirection", "Weight", "EventID", "TimeFromBeginOfEvent"]
    ph.steps_to_store = "entering"
    ph.store_absorbed_event = True
    ph.output_filename = "plane.root"
    sim.run(_backend=OracleBackend)
    n = ph.total_number_of_entries
    assert 0 < n < 6000 and ph.number_of_absorbed_events == 6000 - len(np.unique(ph.data["EventID"]))
    assert set(ph.data) == {"KineticEnergy", "PrePosition_X", "PrePosition_Y", "PrePosition_Z", "PreDirection_X",
                            "PreDirection_Y", "PreDirection_Z", "Weight", "EventID", "TimeFromBeginOfEvent"}
    assert ph.data["EventID"].dtype == np.int32 and ph.data["KineticEnergy"].dtype == np.float64
    # entering the plane box from below: on its -y face, moving towards +y; primaries keep 140.5 keV
    y = ph.data["PrePosition_Y"]
    on_face = np.abs(y - 299.0) < 2e-2
    assert on_face.mean() > 0.98 and np.all((y > 298.99) & (y < 301.01))
    assert np.all(ph.data["PreDirection_Y"][on_face] > 0)
    assert ph.data["KineticEnergy"].max() == pytest.approx(0.140511, rel=1e-6)
    assert np.all(ph.data["KineticEnergy"] > 0.001) and np.all(ph.data["EventID"] < 6000)
    # the ROOT file: a flat TTree named after the actor with the same columns, same types (rootio.py)
    from opengate_b200 import rootio
    assert rootio.list_trees(tmp_path / "out" / "plane.root") == ["phsp"]
    cols, avail = rootio.read_tree(tmp_path / "out" / "plane.root")
    assert avail == list(ph.data)
    for k, v in ph.data.items():
        assert cols[k].dtype == v.dtype and np.array_equal(cols[k], v), k
    # emission positions: all inside the hot block of the activity image, in the CT frame (+ translation)
    d = dose.edep.image.array
    assert d.sum() > 0


def test_activity_source_draws_poisson_counts(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    src.number_of_primaries = 0
    src.activity = 3000 * Bq
    sim.run_timing_intervals = [[0, 1 * sec]]
    out = sim.run(_backend=OracleBackend)
    n = out.device_stats[0]["histories"]
    # spectrum weights sum (x bin width) scales the activity (GateGenericSource.cpp:622-626)
    scale = 1.0 * 0.02 + 2.0 * 0.02 + 0.5 * 0.04
    assert abs(n - 3000 * scale) < 6 * np.sqrt(3000 * scale)


def test_unsupported_configurations_are_fatal(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    with pytest.raises(Exception, match="gammas only"):
        src.particle = "e-"
    with pytest.raises(Exception, match="not an allowed value"):
        dose.hit_type = "somewhere"
    with pytest.raises(GateDeprecationError):
        dose.dose = True
    with pytest.raises(GateDeprecationError):
        src.n = 10
    with pytest.raises(Exception, match="not supported"):
        sim.add_actor("LETActor", "let")
    with pytest.raises(Exception, match="not supported"):
        sim.add_volume("Sphere", "s")
    sim.engine = "geant4"
    with pytest.raises(Exception, match="reference's own SimulationEngine"):
        sim.run(_backend=OracleBackend)
    sim.engine = "b200"
    dose.score_in = "G4_WATER"
    with pytest.raises(Exception, match="score_in"):
        sim.run(_backend=OracleBackend)
    dose.score_in = "material"
    src.direction.type = "histogram"            # supported, but needs its histograms (generic.py:194-205)
    with pytest.raises(Exception, match="histogram_theta_weights"):
        sim.run(_backend=OracleBackend)
    src.direction.type = "sideways"
    with pytest.raises(Exception, match="direction type"):
        sim.run(_backend=OracleBackend)


def test_uncertainty_goal_stops_the_run_early(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=200000)
    dose.uncertainty_goal = 0.25
    dose.uncertainty_first_check_after_n_events = 2000
    dose.uncertainty_top_voxels_count = 5
    dose.uncertainty_voxel_edep_threshold = 0.7

    class B(OracleBackend):
        def mean_uncertainty(self, actor, n_events, top, thr):
            from oracle.oracle import mean_uncertainty
            from opengate_b200 import abi
            return mean_uncertainty(self.dose_grid(actor, abi.GRID_EDEP), self.dose_grid(actor, abi.GRID_EDEP_SQ), n_events, top, thr)

    out = sim.run(_backend=B)
    n = out.device_stats[0]["histories"]
    assert 2000 <= n < 200000
    from opengate_b200.engines import host_mean_uncertainty
    o = dose.user_output["edep_with_uncertainty"]
    u = host_mean_uncertainty(o.get_data("merged", 0).array, o.get_data("merged", 1).array, n, 5, 0.7)
    assert u <= 0.25


def test_tle_dose_actor_next_to_analog_dose_actor(tmp_path):
    """The reference's own TLE check (test081_tle_2_vox.py:80-98): a DoseActor and a TLEDoseActor on the
    same voxelised volume in one run; the two estimators must agree, the TLE one with far less noise."""
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=20000)
    dose.edep_uncertainty.active = True
    tle = sim.add_actor("TLEDoseActor", "tle")
    tle.attached_to = "ct"
    tle.size = "like_image_volume"
    tle.spacing = "like_image_volume"
    tle.output_filename = "tle.mhd"
    tle.dose.active = True
    tle.edep_uncertainty.active = True
    tle.database = "NIST"
    out = sim.run(_backend=OracleBackend)
    ea, et = dose.edep.image.array, tle.edep.image.array
    assert et.sum() == pytest.approx(ea.sum(), rel=0.03)
    ua, ut = dose.edep_uncertainty.image.array, tle.edep_uncertainty.image.array
    body = (hu == 40) & (ea > 0)
    assert np.median(ut[body]) < 0.5 * np.median(ua[body])
    assert out.device_stats[0]["tle_steps"] > 0 and out.device_stats[0]["tle_overflows"] == 0
    files = sorted(os.listdir(tmp_path / "out"))
    for f in ("tle_edep.mhd", "tle_edep_uncertainty.mhd", "tle_dose.mhd", "ct_edep.mhd"):
        assert f in files, (f, files)
    assert json.load(open(tmp_path / "out" / "tle_edep.json")) == {"number_of_samples": 20000}


def test_tle_dose_actor_user_info(tmp_path):
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=2000)
    tle = sim.add_actor("TLEDoseActor", "tle")
    tle.attached_to = "ct"
    tle.size = "like_image_volume"
    tle.spacing = "like_image_volume"
    assert tle.database == "EPDL" and tle.tle_threshold_type == "None" and tle.energy_min == 0.0
    with pytest.raises(Exception):
        tle.database = "simulated"
    # energy threshold: gammas above it score like a DoseActor
    tle.tle_threshold_type = "energy"
    tle.tle_threshold = 70 * keV
    out = sim.run(_backend=OracleBackend)
    assert 0 < out.device_stats[0]["tle_steps"]
    assert tle.edep.image.array.sum() == pytest.approx(dose.edep.image.array.sum(), rel=0.1)
    for bad, msg in (("max range", "CSDA"), ("bogus", "correct threshold type")):
        sim2, ct2, src2, dose2, _ = ct_sim(tmp_path, n_primaries=10)
        t2 = sim2.add_actor("TLEDoseActor", "tle")
        t2.attached_to = "ct"
        t2.tle_threshold_type = bad
        with pytest.raises(Exception, match=msg):
            sim2.run(_backend=OracleBackend)
    sim3, ct3, src3, dose3, _ = ct_sim(tmp_path, n_primaries=10)
    t3 = sim3.add_actor("TLEDoseActor", "tle")
    t3.attached_to = "ct"
    t3.score_in = "G4_WATER"
    with pytest.raises(Exception, match="TLEDoseActor cannot score in G4_WATER"):
        sim3.run(_backend=OracleBackend)


def test_user_facing_api_surface_checklist():
    """every attribute of the reference's user interface on this path exists on the mirror."""
    sim = gate.Simulation()
    missing = []

    def need(obj, prefix, names, key=False):
        for n in names:
            ok = (n in obj) if key else True
            if not key:
                try:
                    getattr(obj, n)
                except Exception:
                    ok = False
            if not ok:
                missing.append(f"{prefix}.{n}")

    need(sim, "sim", ["output_dir", "random_seed", "number_of_threads", "run_timing_intervals", "progress_bar", "run"])
    need(sim.world, "world", ["size", "material"])
    need(sim.volume_manager, "volume_manager", ["add_material_database"])
    need(sim.volume_manager.material_database, "material_database", ["add_material_weights"])
    need(sim.physics_manager, "physics_manager", ["physics_list_name", "global_production_cuts", "set_production_cut",
                                                  "apply_cuts", "enable_decay"])
    need(sim.physics_manager.global_production_cuts, "global_production_cuts",
         ["all", "gamma", "electron", "positron", "proton"], key=True)
    v = sim.add_volume("Image", "ct")
    need(v, "image", ["image", "voxel_materials", "material", "mother", "translation", "rotation", "dump_label_image",
                      "set_materials_from_voxelisation"])
    for kind in ("GenericSource", "VoxelSource"):
        s = sim.add_source(kind, kind)
        need(s, kind, ["particle", "number_of_primaries", "activity", "half_life", "attached_to", "weight"])
        need(s.position, kind + ".position", ["type", "radius", "size", "translation", "rotation", "sigma_x", "sigma_y"], key=True)
        need(s.direction, kind + ".direction", ["type", "momentum", "focus_point", "theta", "phi", "sigma"], key=True)
        need(s.energy, kind + ".energy", ["type", "mono", "sigma_gauss", "spectrum_weights", "spectrum_energies",
                                          "spectrum_energy_bin_edges", "spectrum_histogram_interpolation"], key=True)
    need(sim.source_manager.sources["VoxelSource"], "VoxelSource", ["image"])
    ps = sim.add_source("PhaseSpaceSource", "ps")
    need(ps, "PhaseSpaceSource", ["phsp_file", "entry_start", "particle", "global_flag", "isotropic_direction", "translate_position",
                                  "rotate_direction", "batch_size", "position_key", "direction_key", "energy_key", "weight_key",
                                  "PDGCode_key", "generate_until_next_primary", "cycle_count"])
    for kind in ("DoseActor", "TLEDoseActor"):
        d = sim.add_actor(kind, kind)
        need(d, kind, ["attached_to", "size", "spacing", "translation", "rotation", "hit_type", "output_coordinate_system",
                       "output_filename", "score_in", "uncertainty_goal", "edep", "edep_squared", "edep_uncertainty", "dose",
                       "dose_squared", "dose_uncertainty", "density", "counts"])
        need(d.edep, kind + ".edep", ["active", "write_to_disk", "output_filename", "keep_data_per_run", "get_output_path",
                                      "get_data", "image"])
    need(sim.actor_manager.actors["TLEDoseActor"], "TLEDoseActor", ["energy_min", "database"])
    p = sim.add_actor("PhaseSpaceActor", "p")
    need(p, "PhaseSpaceActor", ["attached_to", "attributes", "steps_to_store", "store_absorbed_event", "output_filename",
                                "get_output_path", "total_number_of_entries", "number_of_absorbed_events"])
    st = sim.add_actor("SimulationStatisticsActor", "st")
    need(st, "SimulationStatisticsActor", ["counts", "pps", "track_types_flag"])
    assert hasattr(gate.geometry.materials, "HounsfieldUnit_to_material")
    assert missing == []


def test_output_filename_semantics_follow_the_reference(tmp_path):
    """actoroutput.py:996-1058 + actors/base.py:449-457: the actor-level name fans out with the item suffix,
    a name set on ONE output is kept verbatim, 'auto' gets the suffix when the name is generated."""
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    assert dose.edep.output_filename == "ct_edep.mhd"
    assert dose.edep_uncertainty.output_filename == "ct_edep_uncertainty.mhd"
    assert dose.dose.output_filename == "ct_dose.mhd"
    dose.edep.output_filename = "x.mhd"
    assert dose.edep.output_filename == "x.mhd"                       # verbatim, no suffix
    assert dose.edep_uncertainty.output_filename == "ct_edep_uncertainty.mhd"
    assert str(dose.edep.get_output_path()).endswith("/x.mhd")
    dose2 = sim.add_actor("DoseActor", "d2")
    dose2.attached_to = "ct"
    assert dose2.edep.output_filename == "edep_with_uncertainty_from_doseactor_d2_edep.mhd"   # actoroutput.py:488-490
    dose.dose.output_filename = None
    assert dose.dose.output_filename is None


def test_lean_outputs_keep_the_squares_on_the_engine_side(tmp_path):
    """sim.lean_outputs: the squares scored only for an uncertainty image are not brought back in a single-run
    simulation; the uncertainty image is evaluated by the engine and equals the host algebra of the default path."""
    imgs = []
    for lean in (False, True):
        (tmp_path / str(lean)).mkdir()
        sim, ct, src, dose, hu = ct_sim(tmp_path / str(lean), n_primaries=3000)
        sim.lean_outputs = lean
        dose.dose.active = True
        dose.dose_uncertainty.active = True
        dose.edep_uncertainty.active = True
        sim.run(_backend=OracleBackend)
        imgs.append((dose.edep.image.array, dose.edep_uncertainty.image.array, dose.dose.image.array, dose.dose_uncertainty.image.array,
                     dose.user_output["edep_with_uncertainty"].get_data("merged", 1)))
    a, b = imgs
    assert a[4] is not None and b[4] is None
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1], b[1])                                   # same expressions, same bits
    np.testing.assert_allclose(a[3], b[3], rtol=1e-12, atol=0)          # dose: evaluated before / after the Gy scaling
    assert (b[1] > 0).any()


def test_event_time_stamps_follow_the_source_timing(tmp_path):
    """GlobalTime of a phase-space row = time of the event + time of flight.  Reference: GateVSource.cpp:55-89: an
    `activity` source spreads its events over the run interval as a Poisson process (decaying with a half-life),
    a `number_of_primaries` source emits them at the start of the interval.  Here the time of an event is a pure
    function of (seed, history id): timing.event_times."""
    import opengate_b200 as gate
    from opengate_b200 import timing
    s_, mm, keV, Bq = gate.g4_units.s, gate.g4_units.mm, gate.g4_units.keV, gate.g4_units.Bq
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=0)
    sim.random_seed = 4242
    sim.run_timing_intervals = [[2 * s_, 3 * s_]]
    src.number_of_primaries = 0
    src.energy.type = "mono"
    src.energy.mono = 100 * keV
    src.activity = 20000 * Bq
    src.half_life = 0.4 * s_
    plane = sim.add_volume("Box", "plane")
    plane.size = [900 * mm, 900 * mm, 2 * mm]
    plane.translation = [0, 0, 350 * mm]
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "EventID", "GlobalTime", "TimeFromBeginOfEvent"]
    ph.write_to_disk = False
    out = sim.run(_backend=OracleBackend)
    n = out.device_stats[0]["histories"]
    assert 8500 < n < 10500          # A0 / lambda * (1 - exp(-lambda * 1 s)) = 9500
    g, tof, ev = ph.data["GlobalTime"], ph.data["TimeFromBeginOfEvent"], ph.data["EventID"]
    assert len(g) > 2000
    t_ev = g - tof
    assert t_ev.min() >= 2 * s_ and t_ev.max() <= 3 * s_ and np.all(tof > 0) and np.all(tof < 1000.0)   # ns of flight
    # the same function of the id, evaluated directly
    want = timing.event_times(4242, ev, (2 * s_, 3 * s_), activity=True, half_life=0.4 * s_)
    np.testing.assert_allclose(t_ev, want, rtol=0, atol=1e-3)
    # decaying rate: more events early; mean of exp(-lam t) on one second with T1/2 = 0.4 s
    lam = np.log(2) / 0.4
    mean = 1 / lam - np.exp(-lam) / (1 - np.exp(-lam))
    assert abs((t_ev.mean() - 2 * s_) / s_ - mean) < 0.02
    # a number_of_primaries source: every event at the start of the interval
    sim2, ct2, src2, dose2, hu2 = ct_sim(tmp_path, n_primaries=3000)
    sim2.run_timing_intervals = [[5 * s_, 6 * s_]]
    plane2 = sim2.add_volume("Box", "plane")
    plane2.size = [900 * mm, 900 * mm, 2 * mm]
    plane2.translation = [0, 0, 350 * mm]
    ph2 = sim2.add_actor("PhaseSpaceActor", "phsp")
    ph2.attached_to = plane2
    ph2.attributes = ["GlobalTime", "TimeFromBeginOfEvent"]
    ph2.write_to_disk = False
    sim2.run(_backend=OracleBackend)
    assert len(ph2.data["GlobalTime"]) > 100
    np.testing.assert_allclose(ph2.data["GlobalTime"] - ph2.data["TimeFromBeginOfEvent"], 5 * s_, rtol=0, atol=1e-3)


def test_angular_histogram_and_confine_like_the_reference_tests(tmp_path):
    """direction.type = 'histogram' (reference test010_generic_source_angular_distribution.py: theta 80..100 deg,
    five phi bins) and position.confine (test010_generic_source_confine.py, here to a Box) through the script API,
    primaries traced on the oracle."""
    from oracle.oracle import TransportOracle
    deg = gate.g4_units.deg
    sim, ct, src, dose, hu = ct_sim(tmp_path)
    src.position.type = "point"
    src.direction.type = "histogram"
    src.direction.histogram_theta_angles = [80 * deg, 100 * deg]
    src.direction.histogram_theta_weights = [1]
    src.direction.histogram_phi_angles = [60 * deg, 70 * deg, 80 * deg, 100 * deg, 110 * deg, 120 * deg]
    src.direction.histogram_phi_weights = [0.3, 0.5, 1, 0.5, 0.3]
    stuff = sim.add_volume("Box", "stuff")
    stuff.size = [20 * mm, 10 * mm, 40 * mm]
    stuff.translation = [-50 * mm, 5 * mm, -300 * mm]
    stuff.material = "G4_AIR"
    src2 = sim.add_source("GenericSource", "confined")
    src2.particle = "gamma"
    src2.energy.mono = 0.1 * gate.g4_units.MeV
    src2.position.type = "box"
    src2.position.size = [100 * mm, 100 * mm, 100 * mm]
    src2.position.translation = [-50 * mm, 0, -300 * mm]
    src2.position.confine = "stuff"
    src2.direction.type = "momentum"
    src2.direction.momentum = [0, 0, 1]
    src2.number_of_primaries = 100
    for s in (src, src2):
        s.resolve_and_validate_config(sim.run_timing_intervals)
    from small_setups import c1
    S = c1(10)
    S.sources = [src.to_abi(None), src2.to_abi(None)]
    orc = TransportOracle(3)
    so, _, _ = S.apply(orc)
    d = np.array([orc.trace(so[0], h).d0[:] for h in range(20000)])
    theta, phi = np.arccos(-d[:, 2]), np.arctan2(-d[:, 1], -d[:, 0]) % (2 * np.pi)
    assert theta.min() >= 80 * deg - 1e-5 and theta.max() <= 100 * deg + 1e-5
    assert np.allclose(np.histogram(theta, bins=4, range=(80 * deg, 100 * deg))[0] / len(theta), 0.25, atol=0.012)
    h = np.histogram(phi, bins=np.array([60, 70, 80, 100, 110, 120]) * deg)[0] / len(phi)
    assert np.allclose(h, np.array([0.3, 0.5, 1, 0.5, 0.3]) / 2.6, atol=0.012)
    p = np.array([orc.trace(so[1], h).p0[:] for h in range(5000)])
    q = p - np.array([-50.0, 5.0, -300.0])
    assert np.all(np.abs(q) <= np.array([10.0, 5.0, 20.0]) + 1e-4)
    assert np.allclose(q.std(axis=0), np.array([10.0, 5.0, 20.0]) / np.sqrt(3), rtol=0.05)
    # validation as the reference: n weights for n + 1 angles; a confine volume that does not exist
    src.direction.histogram_phi_weights = [1, 2]
    with pytest.raises(Exception, match="histogram_phi_weights"):
        src.resolve_and_validate_config(sim.run_timing_intervals)
    src2.position.confine = "nowhere"
    with pytest.raises(Exception, match="confine"):
        src2.resolve_and_validate_config(sim.run_timing_intervals)


def test_compressed_metaimage_round_trip(tmp_path):
    """CT images often come as .mhd + .zraw (ITK MetaImageIO: the element data as one zlib stream): read and written"""
    a = np.random.default_rng(3).integers(-1000, 2000, (6, 8, 10)).astype(np.int16)
    img = gimage.Image(a, (0.5, 0.6, 0.7), (1.0, 2.0, 3.0))
    for name in ("ct.mhd", "ct.mha"):
        gimage.write_image(img, tmp_path / name, compressed=True)
        assert "CompressedData = True" in open(tmp_path / name, "rb").read(400).decode("latin-1")
        r = gimage.read_image(tmp_path / name)
        assert r.array.dtype == a.dtype and np.array_equal(r.array, a)
        assert np.allclose(r.spacing, img.spacing) and np.allclose(r.origin, img.origin)
    assert (tmp_path / "ct.zraw").exists()
    (tmp_path / "ct.zraw").write_bytes(b"not zlib")
    with pytest.raises(Exception, match="decompress"):
        gimage.read_image(tmp_path / "ct.mhd")


def test_particle_name_column_reaches_the_root_file(tmp_path):
    """ParticleName (a string attribute, `S` column of the reference's digi collections) is written as a TLeafC branch"""
    from opengate_b200 import rootio
    sim, ct, src, dose, hu = ct_sim(tmp_path, n_primaries=500)
    plane = sim.add_volume("Box", "plane")
    plane.size = [400 * mm, 400 * mm, 2 * mm]
    plane.translation = [0, 0, -300 * mm]
    plane.material = "G4_AIR"
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "ParticleName", "PDGCode", "EventID"]
    ph.output_filename = "names.root"
    sim.run(_backend=OracleBackend)
    cols, avail = rootio.read_tree(tmp_path / "out" / "names.root")
    assert avail == ["KineticEnergy", "ParticleName", "PDGCode", "EventID"]
    assert len(cols["ParticleName"]) == ph.total_number_of_entries >= 490
    assert set(cols["ParticleName"].tolist()) == {"gamma"} and set(cols["PDGCode"].tolist()) == {22}
