"""GPU tier: the uncertainty-goal stop rule (reference GateDoseActor::ComputeMeanUncertainty +
EndOfEventAction, GateDoseActor.cpp:245-330, 401-430) on the CUDA back end: the device reductions of
csrc/uncertainty.cuh against the numpy restatement in the oracle, and Simulation.run() with
`uncertainty_goal` on the product path."""
import numpy as np
import pytest

from opengate_b200 import abi

pytestmark = pytest.mark.gpu

from small_setups import SMALL  # noqa: E402


@pytest.mark.parametrize("name,flags", [("c3", abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE),     # stride-4 records
                                        ("c1", abi.SCORE_EDEP | abi.SCORE_EDEP_SQ),                      # stride-2 records
                                        ("c2", abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ | abi.SCORE_COUNTS)])
def test_device_mean_uncertainty_matches_the_oracle(cuda_required, name, flags):
    from opengate_b200.core import Engine
    from oracle.oracle import mean_uncertainty
    n = 30000
    S = SMALL[name](n)
    S.dose[0].score_flags = flags
    eng = Engine(0, S.seed)
    se, de, _ = S.apply(eng)
    eng.begin_run()
    eng.run_source(se[0], 0, n)
    eng.sync()
    e, e2 = eng.dose_grid(de[0], abi.GRID_EDEP), eng.dose_grid(de[0], abi.GRID_EDEP_SQ)
    assert (e > 0).sum() > 100
    for top, thr in ((5, 0.7), (1, 0.0), (50, 0.2), (1000, 0.05), (10 ** 9, 0.5)):
        want = mean_uncertainty(e, e2, n, top, thr)
        got = eng.mean_uncertainty(de[0], n, min(top, 2 ** 31 - 1), thr)
        assert got == pytest.approx(want, rel=1e-12), (top, thr)
    # ties in the top-N (pencil beam on a coarse grid gives few distinct voxels): still the mean of the N highest
    eng.set_dose_grid(de[0], abi.GRID_EDEP, np.where(e > 0, 3.0, 0.0))
    eng.set_dose_grid(de[0], abi.GRID_EDEP_SQ, np.where(e > 0, 9.0 / 2, 0.0))
    want = mean_uncertainty(np.where(e > 0, 3.0, 0.0), np.where(e > 0, 4.5, 0.0), n, 7, 0.9)
    assert eng.mean_uncertainty(de[0], n, 7, 0.9) == pytest.approx(want, rel=1e-12)
    # an empty grid: nothing above the threshold -> 1.0 (GateDoseActor.cpp:321-326)
    eng.begin_run()
    assert eng.mean_uncertainty(de[0], 10, 5, 0.7) == 1.0
    eng.close()


def test_mean_uncertainty_needs_squares(cuda_required):
    from opengate_b200.core import Engine, GateB200Error
    S = SMALL["c3"](100)
    S.dose[0].score_flags = abi.SCORE_EDEP
    eng = Engine(0, S.seed)
    se, de, _ = S.apply(eng)
    with pytest.raises(GateB200Error, match="edep_squared"):
        eng.mean_uncertainty(de[0], 100, 5, 0.7)
    eng.close()


def test_script_level_uncertainty_goal_on_the_cuda_engine(cuda_required, tmp_path):
    """`dose.uncertainty_goal` through Simulation.run() on the product back end stops at the same event count as
    the oracle back end (same histories, same check schedule) and below the goal."""
    from oracle.oracle import mean_uncertainty
    from oracle_backend import OracleBackend
    from opengate_b200.engines import host_mean_uncertainty
    from test_host_api import ct_sim

    class B(OracleBackend):
        def mean_uncertainty(self, actor, n_events, top, thr):
            return mean_uncertainty(self.dose_grid(actor, abi.GRID_EDEP), self.dose_grid(actor, abi.GRID_EDEP_SQ), n_events, top, thr)

    res = []
    for backend, sub in ((None, "gpu"), (B, "cpu")):
        (tmp_path / sub).mkdir()
        sim, ct, src, dose, hu = ct_sim(tmp_path / sub, n_primaries=200000)
        dose.uncertainty_goal = 0.25
        dose.uncertainty_first_check_after_n_events = 2000
        dose.uncertainty_top_voxels_count = 5
        dose.uncertainty_voxel_edep_threshold = 0.7
        out = sim.run(_backend=backend)
        res.append((out.device_stats[0]["histories"], dose.edep.image.array.copy(), dose.user_output["edep_with_uncertainty"]))
    (ng, eg, og), (nc, ec, oc) = res
    assert ng == nc and 2000 <= ng < 200000
    np.testing.assert_allclose(eg, ec, rtol=1e-10, atol=0)
    u = host_mean_uncertainty(og.get_data("merged", 0).array, og.get_data("merged", 1).array, ng, 5, 0.7)
    assert u <= 0.25


def test_device_side_unit_conversion_and_uncertainty_images(cuda_required):
    """gate_b200_get_dose_grid_scaled / gate_b200_get_uncertainty: the end-of-run algebra of the DoseActor
    (doseactors.py:685-696, utility.py:551-560, dataitems.py:1956-1993) evaluated on the device gives the bits the
    numpy expressions give on the host; a page-locked destination (gate_b200_host_register) changes nothing."""
    from opengate_b200.actors import uncertainty_from_squares
    from opengate_b200.core import Engine
    n = 20000
    S = SMALL["c3"](n)
    S.dose[0].score_flags = abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ
    eng = Engine(0, S.seed)
    se, de, _ = S.apply(eng)
    eng.begin_run()
    eng.run_source(se[0], 0, n)
    eng.sync()
    g = {w: eng.dose_grid(de[0], w) for w in (abi.GRID_EDEP, abi.GRID_EDEP_SQ, abi.GRID_DOSE, abi.GRID_DOSE_SQ)}
    s = 1.0 / (6.241509074460763e+18 * 1.0e-12 * 7.8 * 7.8 * 7.5)
    assert np.array_equal(eng.dose_grid(de[0], abi.GRID_DOSE, scale=s), g[abi.GRID_DOSE] * s)
    assert np.array_equal(eng.dose_grid(de[0], abi.GRID_DOSE_SQ, scale=s * s), g[abi.GRID_DOSE_SQ] * (s * s))
    for v in (abi.GRID_EDEP, abi.GRID_DOSE):
        want = uncertainty_from_squares(g[v], g[v + 1], n)
        assert np.array_equal(eng.uncertainty_grid(de[0], v, n), want)
    out = np.empty(eng.dose_grid_shape(de[0]), dtype=np.float64)
    assert eng.pin(out)
    eng.dose_grid(de[0], abi.GRID_EDEP, out=out)
    eng.unpin(out)
    assert np.array_equal(out, g[abi.GRID_EDEP])
    eng.close()
