"""Phase-space replay source (reference: sources/phspsources.py, GatePhaseSpaceSource.cpp:92-178).

CPU tier: the oracle's restatement against the reference's rules (sequential entries wrapping at the end
of the file, local -> world frame, normalised direction, per-entry weight) and the script-level chain
PhaseSpaceActor file -> PhaseSpaceSource.  GPU tier: the CUDA engine against the oracle, bit for bit."""
import numpy as np
import pytest

import opengate_b200 as gate
from opengate_b200 import abi

from small_setups import ALL_FLAGS, SMALL  # noqa: E402


def phsp_columns(n, seed=5, with_dir=True):
    rng = np.random.default_rng(seed)
    c = {"pos_x": rng.uniform(-20, 20, n), "pos_y": rng.uniform(-20, 20, n), "pos_z": rng.uniform(-5, 5, n),
         "energy": rng.uniform(0.03, 0.4, n), "weight": rng.uniform(0.25, 2.0, n)}
    if with_dir:
        d = rng.normal(size=(n, 3)) * 0.15 + np.array([0.0, 0.0, 1.0])
        d *= rng.uniform(0.5, 3.0, n)[:, None]          # not normalised in the file
        c.update(dir_x=d[:, 0], dir_y=d[:, 1], dir_z=d[:, 2])
    return {k: v.astype(np.float32) for k, v in c.items()}


ROT = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])


def phsp_setup(n_entries=700, entry_start=650, local=True, iso=False, n=10):
    S = SMALL["c1"](n)
    cols = phsp_columns(n_entries, with_dir=not iso)
    kw = dict(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, dir_type=abi.DIR_ISO if iso else abi.DIR_PHSP,
              phsp=cols, phsp_entry_start=entry_start, weight=1.0)
    if local:
        kw.update(pos_rotation=ROT, pos_translation=(-220.0, 3.0, -4.0), dir_relative=True, dir_rotation=ROT)
    else:
        kw.update(pos_translation=(3.0, -4.0, -200.0))
    S.sources = [abi.make_source(**kw)]
    return S, cols


def test_oracle_replays_entries_in_order_and_wraps():
    from oracle.oracle import TransportOracle
    S, c = phsp_setup(n_entries=700, entry_start=650, local=True)
    orc = TransportOracle(S.seed)
    so, _, _ = S.apply(orc)
    t = np.array([-220.0, 3.0, -4.0])
    for hist in (0, 1, 49, 50, 51, 700, 1399):       # 650 + 50 = 700 -> wraps to entry 0
        e = (650 + hist) % 700
        tr = orc.trace(so[0], hist)
        p = ROT @ np.array([c["pos_x"][e], c["pos_y"][e], c["pos_z"][e]], dtype=np.float64) + t
        d = np.array([c["dir_x"][e], c["dir_y"][e], c["dir_z"][e]], dtype=np.float64)
        d = ROT @ (d / np.linalg.norm(d))
        assert np.allclose(np.array(tr.p0[:]), p, rtol=0, atol=1e-4)
        assert np.allclose(np.array(tr.d0[:]), d, rtol=0, atol=1e-6)
        assert tr.E0 == c["energy"][e]


def test_oracle_weights_follow_the_entries():
    """Everything a history deposits carries the weight of ITS entry: with photons that cannot leave the
    world unseen, sum(edep) + escaped energy is bounded by sum(w E) and edep of a 1-entry file scales with w."""
    from oracle.oracle import TransportOracle
    out = []
    for w in (1.0, 2.5):
        S = SMALL["c1"](10)
        cols = {"pos_x": np.zeros(1, np.float32), "pos_y": np.zeros(1, np.float32), "pos_z": np.full(1, -100.0, np.float32),
                "dir_x": np.zeros(1, np.float32), "dir_y": np.zeros(1, np.float32), "dir_z": np.full(1, 2.0, np.float32),
                "energy": np.full(1, 0.06, np.float32), "weight": np.full(1, w, np.float32)}
        S.sources = [abi.make_source(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, dir_type=abi.DIR_PHSP, phsp=cols)]
        orc = TransportOracle(S.seed)
        so, do, _ = S.apply(orc)
        orc.begin_run()
        orc.run_source(so[0], 0, 3000, 0)
        out.append((orc.stats()["edep_total_MeV"], orc.dose_grid(do[0], abi.GRID_EDEP).sum(),
                    orc.dose_grid(do[0], abi.GRID_EDEP_SQ).sum()))
    (e1, g1, q1), (e2, g2, q2) = out
    assert e1 > 0 and g1 > 0
    assert e2 == pytest.approx(2.5 * e1, rel=1e-5)       # same histories (same seed), weight 2.5
    assert g2 == pytest.approx(2.5 * g1, rel=1e-5)
    assert q2 == pytest.approx(6.25 * q1, rel=1e-5)


def test_source_abi_validation_messages():
    from oracle.oracle import TransportOracle  # noqa: F401  (host-only checks below)
    with pytest.raises(ValueError):
        c = phsp_columns(10)
        c["energy"] = c["energy"][:5]
        abi.make_source(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, dir_type=abi.DIR_PHSP, phsp=c)


def _plane_sim(tmp_path, backend_n=3000):
    """Point source -> water box with a dose actor, and a scoring plane behind it writing a phase space."""
    m, mm, keV = gate.g4_units.m, gate.g4_units.mm, gate.g4_units.keV
    sim = gate.Simulation()
    sim.random_seed = 1234
    sim.output_dir = str(tmp_path)
    sim.world.size = [2 * m, 2 * m, 2 * m]
    box = sim.add_volume("Box", "water")
    box.size = [200 * mm, 200 * mm, 100 * mm]
    box.material = "G4_WATER"
    sim.physics_manager.set_production_cut("world", "all", 1 * m)
    dose = sim.add_actor("DoseActor", "dose")
    dose.attached_to = "water"
    dose.size = [20, 20, 10]
    dose.spacing = [10 * mm, 10 * mm, 10 * mm]
    dose.write_to_disk = False
    return sim, box, dose


def _load_columns(path):
    if str(path).endswith(".root"):
        from opengate_b200 import rootio
        return rootio.read_tree(path)[0]
    return np.load(path)


@pytest.mark.parametrize("plane_file", ["plane.npz", "plane.root"])
def test_script_level_phase_space_actor_file_feeds_phase_space_source(tmp_path, plane_file):
    """The wire format either side of the path: a PhaseSpaceActor file (local coordinates of its plane)
    replayed by a PhaseSpaceSource attached to the same plane reproduces the stored particles in the world."""
    from scipy.spatial.transform import Rotation
    from oracle_backend import OracleBackend
    mm, keV = gate.g4_units.mm, gate.g4_units.keV
    sim, box, dose = _plane_sim(tmp_path)
    src = sim.add_source("GenericSource", "beam")
    src.energy.mono = 150 * keV
    src.position.type = "disc"
    src.position.radius = 20 * mm
    src.position.translation = [0, 0, -300 * mm]
    src.direction.type = "iso"
    src.direction.theta = [170 * gate.g4_units.deg, 180 * gate.g4_units.deg]   # iso flux points to -(...): towards +z
    src.number_of_primaries = 3000
    plane = sim.add_volume("Box", "plane")
    plane.size = [400 * mm, 400 * mm, 2 * mm]
    plane.translation = [10 * mm, -5 * mm, 120 * mm]
    plane.rotation = Rotation.from_euler("y", 12, degrees=True).as_matrix()
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "PrePosition", "PreDirection", "PrePositionLocal", "PreDirectionLocal",
                     "Weight", "PDGCode", "EventID"]
    ph.output_filename = plane_file
    sim.run(_backend=OracleBackend)
    n_rows = ph.total_number_of_entries
    assert n_rows > 100
    f = _load_columns(ph.get_output_path())
    # local <-> world relation of the stored columns
    R, t = plane.world_transform()
    loc = np.column_stack([f[f"PrePositionLocal_{a}"] for a in "XYZ"])
    wor = np.column_stack([f[f"PrePosition_{a}"] for a in "XYZ"])
    assert np.allclose(loc @ R.T + t, wor, atol=1e-9)
    assert np.all(np.abs(np.abs(loc[:, 2]) - 1.0) < 1e-2) or np.all(np.abs(loc) <= 200.0 + 1e-2)

    # replay
    sim2, box2, dose2 = _plane_sim(tmp_path / "replay")
    plane2 = sim2.add_volume("Box", "plane")
    plane2.size, plane2.translation, plane2.rotation = plane.size, plane.translation, plane.rotation
    s2 = sim2.add_source("PhaseSpaceSource", "replay")
    s2.attached_to = "plane"
    s2.phsp_file = str(ph.get_output_path())
    s2.number_of_primaries = n_rows + 7              # passes the end of the file once
    s2.entry_start = 3
    sim2.run(_backend=OracleBackend)
    assert s2.num_entries == n_rows
    assert s2.cycle_count == 1
    from oracle.oracle import TransportOracle
    st, _keep = s2.to_abi(plane2)
    assert st.pos_type == abi.POS_PHSP and st.phsp_n_entries == n_rows and st.phsp_entry_start == 3
    S = SMALL["c1"](10)
    S.sources = [(st, _keep)]
    orc = TransportOracle(1)
    so, _, _ = S.apply(orc)
    for hist in (0, 5, n_rows - 4, n_rows - 3, n_rows + 2):
        e = (3 + hist) % n_rows
        tr = orc.trace(so[0], hist)
        assert np.allclose(np.array(tr.p0[:]), wor[e], atol=2e-3)
        d = np.array([f[f"PreDirection_{a}"][e] for a in "XYZ"])
        assert np.allclose(np.array(tr.d0[:]), d, atol=1e-5)
        assert tr.E0 == pytest.approx(f["KineticEnergy"][e], rel=1e-6)


def test_phase_space_source_user_info_rules(tmp_path):
    sim, box, dose = _plane_sim(tmp_path)
    c = phsp_columns(50)
    np.savez(tmp_path / "in.npz", X=c["pos_x"], Y=c["pos_y"], Z=c["pos_z"], dX=c["dir_x"], dY=c["dir_y"], dZ=c["dir_z"],
             Ekine=c["energy"], PDGCode=np.full(50, 22, np.int32))
    s = sim.add_source("PhaseSpaceSource", "p")
    s.phsp_file = str(tmp_path / "in.npz")
    s.number_of_primaries = 10
    s.global_flag = True
    s.position_key_x, s.position_key_y, s.position_key_z = "X", "Y", "Z"
    s.direction_key_x, s.direction_key_y, s.direction_key_z = "dX", "dY", "dZ"
    s.energy_key = "Ekine"
    s.translate_position = True
    s.position.translation = [1.0, 2.0, -300.0]
    s.rotate_direction = True
    s.position.rotation = ROT
    s.resolve_and_validate_config([[0, 1e9]])
    cols = s.load_columns()
    assert np.array_equal(cols["pos_z"], c["pos_z"] + np.float32(-300.0))          # float32 add, as the reference
    d = np.column_stack((c["dir_x"], c["dir_y"], c["dir_z"])).astype(np.float64) @ ROT.T
    assert np.array_equal(cols["dir_x"], d[:, 0].astype(np.float32))
    assert np.array_equal(cols["weight"], np.ones(50, np.float32))                 # no Weight column: ones
    st, keep = s.to_abi(None)
    assert st.dir_relative == 0 and st.phsp_entry_start == 0
    # missing key
    s2 = sim.add_source("PhaseSpaceSource", "q")
    s2.phsp_file = str(tmp_path / "in.npz")
    s2.number_of_primaries = 1
    s2.resolve_and_validate_config([[0, 1e9]])
    with pytest.raises(Exception, match="Key 'PrePositionLocal_X' not found"):
        s2.load_columns()
    # non-gamma content
    np.savez(tmp_path / "e.npz", PrePositionLocal_X=c["pos_x"], PrePositionLocal_Y=c["pos_y"], PrePositionLocal_Z=c["pos_z"],
             PreDirectionLocal_X=c["dir_x"], PreDirectionLocal_Y=c["dir_y"], PreDirectionLocal_Z=c["dir_z"],
             KineticEnergy=c["energy"], PDGCode=np.full(50, 11, np.int32))
    s3 = sim.add_source("PhaseSpaceSource", "r")
    s3.phsp_file = str(tmp_path / "e.npz")
    s3.number_of_primaries = 1
    s3.resolve_and_validate_config([[0, 1e9]])
    with pytest.raises(Exception, match="gammas only"):
        s3.load_columns()
    s3.particle = "gamma"                               # the user says what the particles are: PDGCode ignored
    s3._columns = None
    assert len(s3.load_columns()["energy"]) == 50
    with pytest.raises(Exception, match="gammas only"):
        s4 = sim.add_source("PhaseSpaceSource", "s")
        s4.phsp_file = str(tmp_path / "e.npz")
        s4.number_of_primaries = 1
        s4.particle = "e-"
        s4.resolve_and_validate_config([[0, 1e9]])


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("local,iso,n_entries,start", [(True, False, 700, 650), (False, False, 5000, 0), (True, True, 333, 7)])
def test_gpu_phase_space_source_matches_oracle(cuda_required, local, iso, n_entries, start):
    from test_gpu_parity import compare, run_both
    S, _ = phsp_setup(n_entries=n_entries, entry_start=start, local=local, iso=iso)
    eng, orc, ids = run_both(S, 6000, first=123)
    compare(eng, orc, S, ids)
    assert eng.stats()["histories"] == 6000
    eng.close()


@pytest.mark.gpu
def test_gpu_phase_space_source_weights_reach_phase_space_rows_and_tle(cuda_required):
    """Per-entry weights in every scorer: phase-space rows (C4 plane), analog grids and a TLE actor."""
    from test_gpu_parity import compare, run_both
    S = SMALL["c4"](10)
    cols = phsp_columns(900, seed=11)
    cols["pos_z"] = cols["pos_z"] * 0.0
    cols["energy"] = (cols["energy"] * 0.3 + 0.05).astype(np.float32)
    S.sources = [abi.make_source(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, dir_type=abi.DIR_ISO, phsp=cols)]
    eng, orc, ids = run_both(S, 5000)
    compare(eng, orc, S, ids)
    rows = eng.phsp(ids[2][0])
    assert len(rows["EventID"]) > 50
    assert np.array_equal(rows["Weight"], cols["weight"][rows["EventID"] % 900])
    eng.close()
    import tle_setups
    T = tle_setups.c3_tle(10)
    tc = phsp_columns(400, seed=3)
    tc["energy"] = (tc["energy"] * 0.2 + 0.03).astype(np.float32)      # inside the 120 kVp tables of the C3 setup
    T.sources = [abi.make_source(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, dir_type=abi.DIR_PHSP, phsp=tc,
                                 pos_translation=(0.0, 0.0, -200.0))]
    from test_gpu_tle import compare_all
    eng, orc, ids = run_both(T, 1500)
    compare_all(eng, orc, T, ids)
    assert eng.stats()["tle_steps"] > 0
    eng.close()


@pytest.mark.gpu
def test_gpu_script_level_phase_space_chain_matches_oracle_backend(cuda_required, tmp_path):
    """Script level, both directions of the wire format: a PhaseSpaceActor file written by a GPU run is
    replayed by a PhaseSpaceSource on the GPU and on the oracle back end: identical dose grids."""
    from scipy.spatial.transform import Rotation
    from oracle_backend import OracleBackend
    mm, keV = gate.g4_units.mm, gate.g4_units.keV
    sim, box, dose = _plane_sim(tmp_path / "a")
    src = sim.add_source("GenericSource", "beam")
    src.energy.mono = 150 * keV
    src.position.type = "disc"
    src.position.radius = 20 * mm
    src.position.translation = [0, 0, 300 * mm]
    src.direction.type = "iso"
    src.direction.theta = [0, 10 * gate.g4_units.deg]          # the iso flux of G4 points towards -z
    src.number_of_primaries = 20000
    plane = sim.add_volume("Box", "plane")
    plane.size = [400 * mm, 400 * mm, 2 * mm]
    plane.translation = [10 * mm, -5 * mm, 120 * mm]
    plane.rotation = Rotation.from_euler("y", 12, degrees=True).as_matrix()
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "PrePositionLocal", "PreDirectionLocal", "Weight", "PDGCode"]
    ph.output_filename = "plane.npz"
    sim.run()                                                    # CUDA engine
    assert ph.total_number_of_entries > 1000
    out = []
    for backend, sub in ((None, "gpu"), (OracleBackend, "cpu")):
        sim2, box2, dose2 = _plane_sim(tmp_path / sub)
        p2 = sim2.add_volume("Box", "plane")
        p2.size, p2.translation, p2.rotation = plane.size, plane.translation, plane.rotation
        s2 = sim2.add_source("PhaseSpaceSource", "replay")
        s2.attached_to = "plane"
        s2.phsp_file = str(ph.get_output_path())
        s2.number_of_primaries = 30000                           # wraps around the file
        dose2.edep_uncertainty.active = True
        sim2.run(_backend=backend)
        out.append((dose2.edep.image.array.copy(), dose2.edep_squared.image.array.copy()))
    (ge, gs), (ce, cs) = out
    assert ge.sum() > 0
    np.testing.assert_allclose(ge, ce, rtol=1e-9, atol=0)
    np.testing.assert_allclose(gs, cs, rtol=1e-9, atol=0)
