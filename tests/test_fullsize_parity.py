"""GPU parity at the sizes BASELINE.json states, bit for bit against the CPU oracle: the same comparison as
tests/test_gpu_parity.py (integer counters and phase-space rows exact, fp64 grids to the order of the
additions) on the full phantoms -- C1 150^3 @ 2 mm / 1e6 histories / seed 12345678, C2 256^3, C3 512x512x300
with 75 Schneider materials, C4 256x256x150 with the 128^3 activity image and the phase-space plane.  The
multithreaded oracle runs a few million full-size histories in about a second (bench.py's cpu_baseline arm
measures 7e6 histories/s on C3)."""
import numpy as np
import pytest

from opengate_b200 import abi, setups

pytestmark = pytest.mark.gpu

from test_gpu_parity import compare, run_both  # noqa: E402


def _check(S, n, first=0):
    eng, orc, ids = run_both(S, n, first)
    try:
        compare(eng, orc, S, ids)
        st = eng.stats()
        assert st["histories"] == n and st["buffer_overflows"] == 0
        return st
    finally:
        eng.close()
        orc.close()


def test_c1_full_size_bit_exact(cuda_required):
    """BASELINE config 1 as stated: 1 MeV pencil beam, 30 cm water box, DoseActor 150^3 @ 2 mm, 1e6 primaries,
    seed 12345678 (the seed of the reference's test008_dose_actor.py:22)."""
    S = setups.make_c1()
    assert tuple(S.dose[0].size) == (150, 150, 150) and S.seed == 12345678 and S.n_primaries == 1_000_000
    st = _check(S, 1_000_000)
    assert st["n_compt"] > 1_000_000          # a 1 MeV photon scatters more than once on average in 30 cm of water


def test_c2_full_size_bit_exact(cuda_required):
    """BASELINE config 2: 6 MV spectrum into the 256^3 water/lung/bone slab phantom, dose + uncertainty (4 sums)."""
    S = setups.make_c2()
    assert S.label_array().shape == (256, 256, 256)
    _check(S, 4_000_000)


def test_c3_full_size_bit_exact(cuda_required):
    """BASELINE config 3 (the bench workload): 120 kVp into the 512x512x300 synthetic CT, 75 materials,
    dose + edep + edep^2; a second history range far from 0 checks the id -> random stream mapping there too."""
    S = setups.make_c3()
    assert S.label_array().shape == (300, 512, 512) and len(S.tables.names) >= 74
    _check(S, 4_000_000)
    _check(S, 500_000, first=3_000_000_000_123)


def test_c3_full_size_default_hit_type_bit_exact(cuda_required):
    """The reference's default hit_type ('random', doseactors.py:84-91) on the full CT: kernel variant 3."""
    S = setups.make_c3()
    S.dose[0].hit_type = abi.HIT_RANDOM
    _check(S, 2_000_000)


def test_c4_full_size_bit_exact(cuda_required):
    """BASELINE config 4: Tc-99m voxelised activity (128^3 @ 4 mm) in the 256x256x150 CT, phase-space plane
    540x400x1 mm: rows bit-exact (after sorting by event and time), dose grid, counters."""
    S = setups.make_c4(n_primaries=2_000_000)
    assert S.label_array().shape == (150, 256, 256)
    st = _check(S, 2_000_000)
    assert st["phsp_rows"] > 30_000 and st["phsp_dropped"] == 0
