"""GPU parity tests proper: the CUDA engine, called through the C-ABI, against the CPU oracle
on identical inputs. Integer quantities (event counters, counts grid, phase-space rows) must
be bit-exact; the fp64 grids may differ only by the order of the double additions
(rtol 1e-12)."""
import numpy as np
import pytest

from opengate_b200 import abi

pytestmark = pytest.mark.gpu

from small_setups import SMALL, c1  # noqa: E402

INT_KEYS = ["histories", "tracks", "steps", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "n_conv",
            "deposits", "flushed_entries", "buffer_overflows", "escaped", "phsp_rows", "phsp_dropped",
            "absorbed_events"]


def run_both(S, n, first=0):
    from opengate_b200.core import Engine
    from oracle.oracle import TransportOracle
    eng = Engine(0, S.seed)
    orc = TransportOracle(S.seed)
    se, de, pe = S.apply(eng)
    so, do, po = S.apply(orc)
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
