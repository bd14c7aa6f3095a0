"""GPU parity tests proper: the CUDA engine, called through the C-ABI, against the CPU oracle
on identical inputs. Integer quantities (event counters, counts grid, phase-space rows) must
be bit-exact; the fp64 grids may differ only by the order of the double additions
(rtol 1e-12)."""
import numpy as np
import pytest

from opengate_b200 import abi

pytestmark = pytest.mark.gpu

from small_setups import ALL_FLAGS, SMALL, c1  # noqa: E402

INT_KEYS = ["histories", "tracks", "steps", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "n_conv",
            "deposits", "flushed_entries", "buffer_overflows", "escaped", "phsp_rows", "phsp_dropped",
            "absorbed_events"]


def run_both(S, n, first=0, tle_rec_cap=None):
    from opengate_b200.core import Engine
    from oracle.oracle import TransportOracle
    eng = Engine(0, S.seed)
    orc = TransportOracle(S.seed)
    se, de, pe = S.apply(eng)
    so, do, po = S.apply(orc)
    if tle_rec_cap is not None:   # the engine reads GATE_B200_TLE_REC_CAP at launch (set by the caller)
        orc.set_tle_rec_cap(tle_rec_cap)
    eng.begin_run()
    orc.begin_run()
    eng.run_source(se[0], first, n)
    eng.sync()
    orc.run_source(so[0], first, n, 0)
    return eng, orc, (se, de, pe)


def compare(eng, orc, S, ids):
    se, de, pe = ids
    a, b = eng.stats(), orc.stats()
    for k in INT_KEYS:
        assert a[k] == b[k], (k, a[k], b[k])
    assert a["edep_total_MeV"] == pytest.approx(b["edep_total_MeV"], rel=1e-6)
    for d in de:
        flags = S.dose[d].score_flags | abi.SCORE_EDEP
        for which, fl in ((abi.GRID_EDEP, abi.SCORE_EDEP), (abi.GRID_EDEP_SQ, abi.SCORE_EDEP_SQ),
                          (abi.GRID_DOSE, abi.SCORE_DOSE), (abi.GRID_DOSE_SQ, abi.SCORE_DOSE_SQ),
                          (abi.GRID_COUNTS, abi.SCORE_COUNTS)):
            if not flags & fl:
                continue
            g, o = eng.dose_grid(d, which), orc.dose_grid(d, which)
            assert np.array_equal(g != 0, o != 0), f"support differs grid {which}"
            if which == abi.GRID_COUNTS:
                assert np.array_equal(g, o)
            else:
                np.testing.assert_allclose(g, o, rtol=1e-12, atol=0)
    for p in pe:
        g, o = eng.phsp(p), orc.phsp(p)
        assert len(g["EventID"]) == len(o["EventID"])
        ig = np.lexsort((g["Time"], g["EventID"]))
        io = np.lexsort((o["Time"], o["EventID"]))
        for k in g:
            assert np.array_equal(g[k][ig], o[k][io]), k


@pytest.mark.parametrize("name", list(SMALL))
def test_engine_matches_oracle_bit_exact(cuda_required, name):
    S = SMALL[name](20000)
    eng, orc, ids = run_both(S, 20000)
    compare(eng, orc, S, ids)
    assert eng.stats()["histories"] == 20000
    eng.close()


@pytest.mark.parametrize("hit", [abi.HIT_RANDOM, abi.HIT_PRE, abi.HIT_MIDDLE])
def test_hit_types_match_oracle(cuda_required, hit):
    S = c1(5000, hit_type=hit)
    eng, orc, ids = run_both(S, 5000)
    compare(eng, orc, S, ids)
    eng.close()


def test_history_offset_and_split_invariance(cuda_required):
    """Histories are keyed by their global id: running [0,n) in one launch or in two launches
    (or on two GPUs) gives the same grids (up to fp64 summation order)."""
    from opengate_b200.core import Engine
    S = SMALL["c3"](8000)
    e1 = Engine(0, S.seed)
    s1, d1, _ = S.apply(e1)
    e1.begin_run()
    e1.run_source(s1[0], 0, 8000)
    e1.sync()
    e2 = Engine(0, S.seed)
    s2, d2, _ = S.apply(e2)
    e2.begin_run()
    e2.run_source(s2[0], 0, 3000)
    e2.run_source(s2[0], 3000, 5000)
    e2.sync()
    for which in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_DOSE):
        np.testing.assert_allclose(e1.dose_grid(0, which), e2.dose_grid(0, which), rtol=1e-12)
    a, b = e1.stats(), e2.stats()
    for k in INT_KEYS:
        assert a[k] == b[k]
    e1.close()
    e2.close()


def test_pair_mode_emits_annihilation_photons(cuda_required):
    from opengate_b200 import setups
    S = setups.make_c2(n=48, n_primaries=20000, field_mm=30.0)
    S.tables.pair_mode = 1
    eng, orc, ids = run_both(S, 20000)
    compare(eng, orc, S, ids)
    st = eng.stats()
    assert st["n_conv"] > 0 and st["tracks"] == st["histories"] + 2 * st["n_conv"]
    eng.close()


def test_energy_conservation_full_size_property(cuda_required):
    """Size-independent property at a larger size than the oracle is run at: every primary's
    energy is either deposited or leaves the world; with a closed (huge, dense) phantom the
    deposited total equals N * E."""
    from opengate_b200 import setups
    from opengate_b200.core import Engine
    S = setups.make_c1(n_primaries=200000, grid=(64, 64, 64), spacing_mm=1000.0, energy=0.2)
    S.spacing = (64000.0, 64000.0, 64000.0)   # 64 m of water: nothing escapes
    S.world_half = (40000.0, 40000.0, 40000.0)
    S.sources[0][0].pos_translation[2] = 0.0
    eng = Engine(0, 7)
    s, d, _ = S.apply(eng)
    eng.begin_run()
    eng.run_source(s[0], 0, 200000)
    eng.sync()
    st = eng.stats()
    assert st["escaped"] == 0
    assert st["edep_total_MeV"] == pytest.approx(200000 * 0.2, rel=1e-5)
    assert eng.dose_grid(0, abi.GRID_EDEP).sum() == pytest.approx(200000 * 0.2, rel=1e-5)
    eng.close()


# ------------------------------------------------------------------------------ script level

def test_script_level_ct_dose_matches_oracle(cuda_required, tmp_path):
    """The drop-in layer on the CUDA engine: the same GATE-style script run with the product back end
    and with the oracle back end gives the same DoseActor images (fp64 sums: order of additions)."""
    import os
    from test_host_api import ct_sim
    from oracle_backend import OracleBackend
    res = []
    for backend, sub in ((None, "gpu"), (OracleBackend, "cpu")):
        os.makedirs(tmp_path / sub, exist_ok=True)
        sim, ct, src, dose, hu = ct_sim(tmp_path / sub, n_primaries=30000)
        dose.dose.active = True
        dose.dose_uncertainty.active = True
        dose.edep_uncertainty.active = True
        dose.counts.active = True
        stats = sim.add_actor("SimulationStatisticsActor", "stats")
        out = sim.run(_backend=backend)
        res.append((dose, stats, out))
    (dg, sg, og), (dc, sc, oc) = res
    for k in ("histories", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "deposits", "flushed_entries", "escaped"):
        assert og.device_stats[0][k] == oc.device_stats[0][k], k
    assert sg.counts["events"] == sc.counts["events"] == 30000 and sg.counts["steps"] == sc.counts["steps"]
    for name in ("edep", "edep_uncertainty", "dose", "dose_uncertainty", "counts"):
        a, b = getattr(dg, name).image.array, getattr(dc, name).image.array
        assert np.array_equal(a != 0, b != 0), name
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=0, err_msg=name)
    assert os.path.exists(dg.edep.get_output_path()) and os.path.exists(dg.dose_uncertainty.get_output_path())


def test_script_level_rotated_volume_attached_source_coarse_dose_grid(cuda_required, tmp_path):
    """Non-trivial frames: rotated + translated CT, source attached to it, dose grid with its own
    spacing and translation, hit_type random: CUDA engine == oracle."""
    from test_host_api import rotated_ct_sim
    from oracle_backend import OracleBackend
    res = []
    for backend, sub in ((None, "gpu"), (OracleBackend, "cpu")):
        (tmp_path / sub).mkdir(exist_ok=True)
        sim, ct, src, dose, hu = rotated_ct_sim(tmp_path / sub, n_primaries=30000)
        dose.dose.active = True
        dose.edep_squared.active = True
        dose.counts.active = True
        out = sim.run(_backend=backend)
        res.append((dose, out))
    (dg, og), (dc, oc) = res
    for k in ("histories", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "deposits", "flushed_entries", "escaped", "steps"):
        assert og.device_stats[0][k] == oc.device_stats[0][k], k
    for name in ("edep", "edep_squared", "dose", "counts"):
        a, b = getattr(dg, name).image.array, getattr(dc, name).image.array
        assert a.sum() > 0
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=0, err_msg=name)


def test_script_level_voxel_source_phase_space(cuda_required, tmp_path):
    import opengate_b200 as gate
    from opengate_b200 import image as gimage
    from test_host_api import ct_sim
    from oracle_backend import OracleBackend
    keV, mm = gate.g4_units.keV, gate.g4_units.mm
    rows = []
    for backend, sub in ((None, "gpu"), (OracleBackend, "cpu")):
        (tmp_path / sub).mkdir(exist_ok=True)
        sim, ct, src, dose, hu = ct_sim(tmp_path / sub)
        sim.source_manager.sources.clear()
        act = np.zeros((8, 10, 12), dtype=np.float32)
        act[3:5, 4:6, 5:8] = 1.0
        vs = sim.add_source("VoxelSource", "tc99m")
        vs.attached_to = "ct"
        vs.image = gimage.Image(act, (20.0, 20.0, 24.0))
        vs.energy.mono = 140.511 * keV
        vs.direction.type = "iso"
        vs.number_of_primaries = 20000
        plane = sim.add_volume("Box", "plane")
        plane.size = [600 * mm, 2 * mm, 600 * mm]
        plane.translation = [0, 300 * mm, 0]
        ph = sim.add_actor("PhaseSpaceActor", "phsp")
        ph.attached_to = plane
        ph.attributes = ["KineticEnergy", "PrePosition", "PreDirection", "Weight", "EventID", "TimeFromBeginOfEvent"]
        ph.store_absorbed_event = True
        ph.write_to_disk = False
        sim.run(_backend=backend)
        rows.append(ph)
    g, c = rows
    assert g.total_number_of_entries == c.total_number_of_entries > 0
    assert g.number_of_absorbed_events == c.number_of_absorbed_events
    for k in g.data:
        assert np.array_equal(g.data[k], c.data[k]), k


# ------------------------------------------------------------------------------ BASELINE.json sizes

@pytest.fixture(scope="module")
def c3_full(cuda_required):
    from opengate_b200 import setups
    return setups.make_c3(n_primaries=4_000_000, flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_COUNTS)


def test_c3_full_size_properties(c3_full):
    """512x512x300 CT, 75 materials (BASELINE.json config 3) beyond what the bit-exact comparison of tests/test_fullsize_parity.py already covers:
    size-independent properties of the scoring.  (i) everything deposited inside the volume is in
    the edep grid; (ii) dose = sum(e / rho(label at the deposit)): with dose grid == label grid,
    dose * rho(label) == edep per voxel; (iii) counts are integers, and squares obey
    sum(e^2) <= (sum e)^2 and >= (sum e)^2 / counts per voxel (Cauchy-Schwarz over the events)."""
    from opengate_b200.core import Engine
    S = c3_full
    n = 4_000_000
    eng = Engine(0, S.seed)
    s, d, _ = S.apply(eng)
    eng.begin_run()
    eng.run_source(s[0], 0, n)
    eng.sync()
    st = eng.stats()
    assert st["histories"] == n and st["buffer_overflows"] == 0
    e = eng.dose_grid(0, abi.GRID_EDEP)
    e2 = eng.dose_grid(0, abi.GRID_EDEP_SQ)
    dose = eng.dose_grid(0, abi.GRID_DOSE)
    cnt = eng.dose_grid(0, abi.GRID_COUNTS)
    assert e.shape == (300, 512, 512)
    # (i) the world is air: a little energy is deposited outside the volume
    assert e.sum() <= st["edep_total_MeV"] * (1 + 1e-9) and e.sum() > 0.98 * st["edep_total_MeV"]
    # (ii)
    rho = S.tables.density.astype(np.float64)[S.labels]
    m = e > 0
    # (a deposit within one float ulp of a voxel face can be indexed into the neighbour voxel by the
    # dose grid's own affine transform, as in the reference where both indices are computed
    # independently: allow a 1e-5 fraction of such voxels)
    bad = np.abs((dose * rho)[m] - e[m]) > 2e-6 * e[m]
    assert bad.mean() < 1e-5, bad.sum()
    # (iii)
    assert np.array_equal(cnt, np.round(cnt)) and cnt.sum() == st["deposits"] - _deposits_outside(st, e, cnt)
    assert np.all(e2[m] <= e[m] ** 2 * (1 + 1e-9))
    assert np.all(e2[m] * cnt[m] >= e[m] ** 2 * (1 - 1e-9))
    assert np.array_equal(m, cnt > 0) and np.array_equal(m, e2 > 0)
    eng.close()


def _deposits_outside(st, e, cnt):
    # deposits in the world material (outside the image volume) are counted in the statistics only
    return st["deposits"] - int(cnt.sum())


def test_c3_full_size_linearity_and_sharding(c3_full):
    """Two launches over disjoint history ranges add up to the launch over the union (what the
    multi-GPU sharding relies on), at the full CT size."""
    from opengate_b200.core import Engine
    S = c3_full
    a = Engine(0, S.seed)
    sa, _, _ = S.apply(a)
    a.begin_run()
    a.run_source(sa[0], 0, 1_000_000)
    a.sync()
    whole = {w: a.dose_grid(0, w) for w in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_DOSE, abi.GRID_COUNTS)}
    st_whole = a.stats()
    a.begin_run()
    a.run_source(sa[0], 0, 400_000)
    a.sync()
    p1 = {w: a.dose_grid(0, w) for w in whole}
    a.begin_run()
    a.run_source(sa[0], 400_000, 600_000)
    a.sync()
    st2 = a.stats()
    for w in whole:
        np.testing.assert_allclose(p1[w] + a.dose_grid(0, w), whole[w], rtol=1e-11, atol=0)
    assert st2["histories"] == 600_000
    a.close()
    assert st_whole["histories"] == 1_000_000


def test_c1_pencil_beam_same_voxel_aggregation(cuda_required):
    """Pencil beam: every history deposits in the same voxel column, the warp-aggregated atomics
    combine lanes that hit the same voxel; squares must still be per history."""
    from opengate_b200 import setups
    from opengate_b200.core import Engine
    from oracle.oracle import TransportOracle
    S = setups.make_c1(n_primaries=30000, grid=(3, 3, 2), spacing_mm=100.0, flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_COUNTS)
    eng, orc = Engine(0, S.seed), TransportOracle(S.seed)
    se, _, _ = S.apply(eng)
    so, _, _ = S.apply(orc)
    eng.begin_run(); orc.begin_run()
    eng.run_source(se[0], 0, 30000); eng.sync()
    orc.run_source(so[0], 0, 30000, 0)
    for w in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_COUNTS):
        np.testing.assert_allclose(eng.dose_grid(0, w), orc.dose_grid(0, w), rtol=1e-11)
    eng.close()


# ------------------------------------------------------------------------------ edge cases

def _both(S, n, first=0, source=0):
    from opengate_b200.core import Engine
    from oracle.oracle import TransportOracle
    eng, orc = Engine(0, S.seed), TransportOracle(S.seed)
    ie, io = S.apply(eng), S.apply(orc)
    eng.begin_run(); orc.begin_run()
    return eng, orc, ie, io


def test_zero_and_tiny_history_counts(cuda_required):
    """Empty and ragged inputs: 0 histories is a no-op, 1 and 33 histories (less than a warp, one
    more than a warp) run and match."""
    S = SMALL["c3"](100)
    for n in (0, 1, 33):
        eng, orc, (se, de, pe), (so, do, po) = _both(S, n)
        eng.run_source(se[0], 7, n); eng.sync()
        if n:
            orc.run_source(so[0], 7, n, 0)
        a, b = eng.stats(), orc.stats()
        assert a["histories"] == n == b["histories"]
        for k in INT_KEYS:
            assert a[k] == b[k], (n, k)
        np.testing.assert_allclose(eng.dose_grid(0, abi.GRID_EDEP), orc.dose_grid(0, abi.GRID_EDEP), rtol=1e-12)
        eng.close()


def test_uint16_label_volume(cuda_required):
    """The reference's label type is unsigned short (GateImageNestedParameterisation.h:24); the
    engine takes uint16 volumes as they are (uint8 is only a packing used when <= 256 materials)."""
    S = SMALL["c3"](8000)
    S.labels = S.labels.astype(np.uint16)
    eng, orc, ids = run_both(S, 8000)
    compare(eng, orc, S, ids)
    eng.close()


def test_two_dose_actors_two_sources_and_exiting_phase_space(cuda_required):
    """Several actors and sources in one engine: a fine and a coarse dose grid (different hit types),
    a second source, phase-space rows on entering AND exiting."""
    from opengate_b200 import setups
    S = SMALL["c4"](6000)
    nx, ny, nz = S.labels.shape[::-1]
    A2 = setups.dose_world_to_index((5, 4, 3), (100.0, 120.0, 90.0), translation=(10.0, -20.0, 5.0))
    S.dose.append(abi.make_dose_desc((5, 4, 3), A2, ALL_FLAGS, abi.HIT_RANDOM))
    s2, keep2 = abi.make_source(pos_type=abi.POS_SPHERE, radius=30.0, pos_translation=(0, -300.0, 0), dir_type=abi.DIR_FOCUSED,
                                focus_point=(0, 0, 0), ene_type=abi.ENE_GAUSS, mono=0.08, sigma=0.005)
    S.sources.append((s2, keep2))
    S.phsp[0].steps_to_store = abi.PHSP_ENTERING | abi.PHSP_EXITING
    from opengate_b200.core import Engine
    from oracle.oracle import TransportOracle
    eng, orc = Engine(0, S.seed), TransportOracle(S.seed)
    se, de, pe = S.apply(eng)
    so, do, po = S.apply(orc)
    eng.begin_run(); orc.begin_run()
    eng.run_source(se[0], 0, 6000); eng.run_source(se[1], 6000, 4000); eng.sync()
    orc.run_source(so[0], 0, 6000, 0); orc.run_source(so[1], 6000, 4000, 0)
    compare(eng, orc, S, (se, de, pe))
    rows = eng.phsp(pe[0])
    assert set(np.unique(rows["Flags"])) == {abi.PHSP_ENTERING, abi.PHSP_EXITING}
    assert rows["EventID"].max() >= 6000          # rows of the second source carry its history ids
    eng.close()


def test_phase_space_capacity_overflow_is_counted_not_written(cuda_required):
    S = SMALL["c4"](20000)
    S.phsp[0].capacity = 50
    eng, orc, ids = run_both(S, 20000)
    a, b = eng.stats(), orc.stats()
    assert a["phsp_rows"] == 50 == b["phsp_rows"] and a["phsp_dropped"] == b["phsp_dropped"] > 0
    assert len(eng.phsp(ids[2][0])["EventID"]) == 50
    eng.close()


def test_energy_source_types_and_positions(cuda_required):
    """Every position / direction / energy sampler of the source ABI against the oracle."""
    base = SMALL["c1"](10)
    edges = np.array([0.02, 0.04, 0.07, 0.1, 0.15])
    w = np.array([1.0, 3.0, 2.0, 0.5])
    from opengate_b200.setups import histogram_cdf
    cdf_lin = np.cumsum(w + 0.1)
    variants = [
        dict(pos_type=abi.POS_BOX, half_size=(20, 10, 5), ene_type=abi.ENE_RANGE, e_min=0.05, e_max=0.3),
        dict(pos_type=abi.POS_SPHERE, radius=25.0, ene_type=abi.ENE_GAUSS, mono=0.3, sigma=0.02),
        dict(pos_type=abi.POS_DISC, radius=15.0, sigma_x=2.0, sigma_y=3.0, ene_type=abi.ENE_SPECTRUM_DISCRETE,
             spectrum_energy=np.array([0.08, 0.14, 0.36]), spectrum_cdf=np.array([0.2, 0.9, 1.0])),
        dict(pos_type=abi.POS_CYLINDER, radius=10.0, half_size=(0, 0, 30.0), ene_type=abi.ENE_SPECTRUM_HISTOGRAM,
             spectrum_energy=edges, spectrum_cdf=histogram_cdf(edges, w)),
        dict(pos_type=abi.POS_POINT, ene_type=abi.ENE_SPECTRUM_HISTOGRAM_LINEAR, spectrum_energy=edges,
             spectrum_cdf=cdf_lin / cdf_lin[-1]),
        dict(pos_type=abi.POS_POINT, ene_type=abi.ENE_CDF, spectrum_energy=np.array([0.05, 0.1, 0.2, 0.4]),
             spectrum_cdf=np.array([0.0, 0.3, 0.8, 1.0])),
        dict(pos_type=abi.POS_SURFACE_SPHERE, radius=40.0, ene_type=abi.ENE_MONO, mono=0.2),
        dict(pos_type=abi.POS_DISC, radius=5.0, ene_type=abi.ENE_MONO, mono=0.15, dir_type=abi.DIR_COS,
             theta=(0.0, 1.2), phi=(0.3, 5.0), dir_relative=True,
             dir_rotation=np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])),
        dict(pos_type=abi.POS_POINT, ene_type=abi.ENE_MONO, mono=0.3, dir_type=abi.DIR_BEAM2D, dir_sigma=(0.05, 0.2),
             dir_relative=True, dir_rotation=np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])),
        # user angular histograms (a zero-weight bin in theta), pointing at the phantom: -z of the flux frame turned to +z
        dict(pos_type=abi.POS_POINT, ene_type=abi.ENE_MONO, mono=0.2, dir_type=abi.DIR_HISTOGRAM,
             theta_hist=([0.0, 0.1, 0.2, 0.5], [1.0, 0.0, 2.0]), phi_hist=([0.5, 1.0, 3.0, 6.0], [0.3, 1.0, 0.2]),
             dir_relative=True, dir_rotation=np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])),
        # a box source confined to a rotated box that it only partly overlaps (rejection loop in the position sampler)
        dict(pos_type=abi.POS_BOX, half_size=(30, 30, 10), ene_type=abi.ENE_MONO, mono=0.25, dir_type=abi.DIR_MOMENTUM,
             momentum=(0.0, 0.1, 1.0), confine_half_size=(12.0, 6.0, 40.0),
             world_to_confine=np.array([[0.8, 0.6, 0.0, -10.0], [-0.6, 0.8, 0.0, 4.0], [0.0, 0.0, 1.0, 195.0]])),
    ]
    rot = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    dirs = [dict(dir_type=abi.DIR_ISO), dict(dir_type=abi.DIR_MOMENTUM, momentum=(0.1, 0.2, 1.0)),
            dict(dir_type=abi.DIR_FOCUSED, focus_point=(5.0, -5.0, 100.0)),
            dict(dir_type=abi.DIR_ISO, theta=(0.1, 0.6), phi=(0.5, 2.0), dir_relative=True, dir_rotation=rot)]
    for i, v in enumerate(variants):
        kw = dict(v)
        if "dir_type" not in kw:
            kw.update(dirs[i % len(dirs)])
        kw.setdefault("pos_translation", (3.0, -4.0, -200.0))
        s, keep = abi.make_source(**kw)
        S = SMALL["c1"](10)
        S.tables = base.tables
        S.sources = [(s, keep)]
        eng, orc, ids = run_both(S, 4000)
        compare(eng, orc, S, ids)
        assert eng.stats()["histories"] == 4000
        eng.close()


def test_pool_reuse_and_release_of_cached_device_memory(cuda_required):
    """Engines allocate from the device's stream-ordered pool: a second engine of the process reuses the
    blocks of a destroyed one; gate_b200_release_cached_memory hands them back to the driver."""
    import torch
    from opengate_b200 import core
    S = SMALL["c3"](1000)
    core.release_cached_memory(0)              # what earlier tests of this process left in the pool
    free0, _ = torch.cuda.mem_get_info(0)
    for _ in range(2):
        eng = core.Engine(0, S.seed)
        se, de, pe = S.apply(eng)
        eng.begin_run()
        eng.run_source(se[0], 0, 1000)
        eng.sync()
        assert eng.stats()["histories"] == 1000
        eng.close()
    cached, _ = torch.cuda.mem_get_info(0)
    assert cached < free0                      # the pool keeps the engine's blocks
    core.release_cached_memory(0)
    after, _ = torch.cuda.mem_get_info(0)
    assert after > cached



@pytest.mark.parametrize("flags", [abi.SCORE_EDEP, abi.SCORE_EDEP | abi.SCORE_DOSE, abi.SCORE_EDEP | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ,
                                   ALL_FLAGS])
def test_voxel_record_buffers_match_the_dense_grids(cuda_required, flags):
    """gate_b200_dose_buffers exposes the record array the multi-GPU reduce works on; its layout
    (multigpu.record_layout) must agree with what get_dose_grid extracts, and set_dose_grid must invert it."""
    import torch
    from opengate_b200 import multigpu
    from opengate_b200.core import Engine
    S = SMALL["c3"](3000)
    S.dose[0].score_flags = flags
    eng = Engine(0, S.seed)
    se, de, pe = S.apply(eng)
    eng.begin_run()
    eng.run_source(se[0], 0, 3000)
    eng.sync()
    bufs = multigpu.buffer_tensors(eng, de[0])
    stride, offset = multigpu.record_layout(flags)
    nx, ny, nz = [int(x) for x in S.dose[0].size]
    assert bufs[0].numel() == stride * nx * ny * nz
    rec = bufs[0].view(nz, ny, nx, stride).cpu().numpy()
    for w in multigpu.active_grids(flags):
        g = eng.dose_grid(de[0], w)
        if w == abi.GRID_COUNTS:
            assert np.array_equal(bufs[1].cpu().numpy().reshape(nz, ny, nx), g)
        else:
            assert np.array_equal(rec[..., offset[w]], g)
    assert eng.dose_grid(de[0], abi.GRID_EDEP).sum() > 0
    # host -> device overwrite of one component leaves the others alone
    e0 = eng.dose_grid(de[0], abi.GRID_EDEP)
    eng.set_dose_grid(de[0], abi.GRID_EDEP, 2.0 * e0)
    assert np.array_equal(eng.dose_grid(de[0], abi.GRID_EDEP), 2.0 * e0)
    for w in multigpu.active_grids(flags):
        if w not in (abi.GRID_EDEP, abi.GRID_COUNTS):
            assert np.array_equal(eng.dose_grid(de[0], w), rec[..., offset[w]])
    eng.close()


@pytest.mark.parametrize("hu_dtype,label_bytes", [(np.int16, 1), (np.float32, 1), (np.float64, 2), (np.int16, 2)])
def test_hu_to_label_on_the_device_is_integer_exact(cuda_required, hu_dtype, label_bytes):
    """gate_b200_set_image_volume_hu: the label volume made on the device equals np.digitize over the interval
    table of ImageVolume.create_label_image (volumes.py:1200-1243) voxel for voxel: Schneider intervals, a gap
    (label 0), unsorted input rows, values on the bin edges, below the first and above the last bin."""
    from opengate_b200 import materials as gmat
    from opengate_b200.core import Engine
    S = SMALL["c3"](2000)
    nmat = len(S.tables.names)
    rng = np.random.default_rng(17)
    # intervals [lo, hi) -> label; a gap between 300 and 350; rows deliberately out of order
    edges = np.array([-1050.0, -950.0, -700.5, -100.0, 0.0, 15.25, 120.0, 300.0])
    rows = [[edges[i], edges[i + 1], f"m{i}"] for i in range(len(edges) - 1)] + [[350.0, 3000.0, "metal"]]
    rows = [rows[i] for i in rng.permutation(len(rows))]
    lut = {"default": 0}
    for i, r in enumerate(sorted(rows, key=lambda x: x[0])):
        lut[r[2]] = 1 + (i * 7) % (nmat - 1)                     # any labels below n_materials
    bins, bin_labels = gmat.label_bins(rows, lut)
    shape = (37, 41, 53)                                          # odd sizes: the scan's vector tail
    hu = rng.uniform(-1200, 3500, shape)
    hu.flat[: len(bins)] = bins                                   # exactly on the edges
    hu.flat[100:110] = [-5000, 5000, 299.999, 300.0, 349.999, 350.0, 0.0, -0.0, 15.25, 15.2499]
    hu = np.round(hu).astype(hu_dtype) if hu_dtype == np.int16 else hu.astype(hu_dtype)
    if hu_dtype != np.int16:
        hu.flat[200] = np.nan
    want = np.array(bin_labels, dtype=np.uint16)[np.digitize(hu, bins=bins)]
    eng = Engine(0, S.seed)
    eng.set_tables(S.tables)
    eng.set_world(S.world_half, S.world_material)
    got = eng.set_image_volume_hu(hu, (2.0, 2.0, 2.5), S.world_to_volume, bins, bin_labels, label_bytes, want_labels=True)
    assert got.dtype == (np.uint8 if label_bytes == 1 else np.uint16)
    assert np.array_equal(got.astype(np.uint16), want)
    # and the engine runs on it exactly as on an uploaded label volume
    src, keep = S.sources[0]
    sid = eng.add_source(src, keep)
    eng.begin_run()
    eng.run_source(sid, 0, 2000)
    eng.sync()
    a = eng.stats()
    eng.close()
    eng2 = Engine(0, S.seed)
    eng2.set_tables(S.tables)
    eng2.set_world(S.world_half, S.world_material)
    eng2.set_image_volume(want.astype(got.dtype), (2.0, 2.0, 2.5), S.world_to_volume)
    sid = eng2.add_source(src, keep)
    eng2.begin_run()
    eng2.run_source(sid, 0, 2000)
    eng2.sync()
    b = eng2.stats()
    eng2.close()
    for k in INT_KEYS:
        assert a[k] == b[k], k


def test_hu_path_rejects_labels_without_material(cuda_required):
    from opengate_b200.core import Engine, GateB200Error
    S = SMALL["c1"](10)
    eng = Engine(0, S.seed)
    eng.set_tables(S.tables)
    eng.set_world(S.world_half, S.world_material)
    hu = np.zeros((4, 4, 4), np.int16)
    hu[1, 2, 3] = 500
    with pytest.raises(GateB200Error, match="label 77 at voxel 27 has no material"):
        eng.set_image_volume_hu(hu, (1.0, 1.0, 1.0), S.world_to_volume, [100.0], [0, 77], 1)
    eng.close()


def test_ten_billion_histories_launch_splitting_and_64_bit_counters(cuda_required):
    """Config 5 size on one GPU: one run_source call of 1e10 histories is split into launches of 2^31, the
    run statistics are 64-bit end to end, energy is conserved to float32 rounding."""
    from opengate_b200 import setups
    from opengate_b200.core import Engine
    n = 10_000_000_000
    S = setups.make_c2(n=64, n_primaries=n, flags=abi.SCORE_EDEP, field_mm=40.0)
    eng = Engine(0, S.seed)
    se, de, pe = S.apply(eng)
    eng.begin_run()
    eng.run_source(se[0], 0, n)
    eng.sync()
    st = eng.stats()
    assert st["histories"] == n
    assert st["tentative_collisions"] > 2 ** 32 and st["steps"] > 2 ** 32
    assert st["tracks"] >= n
    g = eng.dose_grid(de[0], abi.GRID_EDEP)
    assert g.sum() <= st["edep_total_MeV"] * (1 + 1e-9)          # the grid misses what is deposited outside it
    assert g.sum() > 0.5 * st["edep_total_MeV"]
    eng.close()


@pytest.mark.parametrize("kw", [dict(thickness=40.0, tilt_deg=0.0), dict(thickness=25.0, tilt_deg=30.0),
                                dict(thickness=30.0, tilt_deg=-15.0, energy=0.08),
                                dict(thickness=25.0, tilt_deg=20.0, energy=0.1, slab_is_plane=True),
                                dict(thickness=30.0, energy=0.1, touching=True)])
def test_box_regions_match_oracle(cuda_required, kw):
    """a homogeneous Box of its own material in the world (couch / detector slab), scoring planes before and behind:
    counters, phase-space rows and grids as the oracle (general kernel variant)"""
    from region_setups import slab_in_front
    n = 6000
    S = slab_in_front(n, **kw)
    eng, orc, ids = run_both(S, n)
    compare(eng, orc, S, ids)
    assert eng.kernel_name().endswith(", 0>") and eng.stats()["phsp_dropped"] == 0
    eng.close()


@pytest.mark.parametrize("hit,variant", [(abi.HIT_RANDOM, 2), (abi.HIT_MIDDLE, 2), (abi.HIT_PRE, 3)])
def test_hit_type_on_the_image_grid(cuda_required, hit, variant):
    """On a dose grid that is the image grid a point between the pre- and the post-step point of a Geant4 step lies
    in the voxel of the interaction: random and middle are scored like post (same kernel variant, identical grids),
    pre keeps the hit-point emulation."""
    n = 8000
    S = SMALL["c3"](n)
    S.dose[0].hit_type = hit
    eng, orc, ids = run_both(S, n)
    compare(eng, orc, S, ids)
    assert eng.kernel_name().endswith(f", {variant}>")
    if variant == 2:
        P = SMALL["c3"](n)
        P.dose[0].hit_type = abi.HIT_POST
        e2, o2, i2 = run_both(P, n)
        for which in (abi.GRID_EDEP, abi.GRID_DOSE, abi.GRID_COUNTS):
            assert np.array_equal(orc.dose_grid(ids[1][0], which), o2.dose_grid(i2[1][0], which))
        e2.close()
    eng.close()
