"""The tabulated per-element models (data/element_models_coarse.npz, tools/gen_element_models.py) are the numbers
the code computes: recompute a few elements, and build the photon tables of a material set with and without the
file."""
import numpy as np


def test_tabulated_element_models_equal_the_computed_ones():
    from opengate_b200 import physics_models as pm
    assert pm._element_models_file() is not None
    for z in (1, 6, 8, 20, 53, 82):
        a, b = pm._element_models_coarse(z), pm._compute_element_models_coarse(z)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), z


def test_photon_tables_do_not_depend_on_the_model_file(monkeypatch):
    from opengate_b200 import physics, physics_models as pm, setups
    mats, _ = setups._materials_from_db(["G4_WATER", "G4_AIR", "G4_BONE_COMPACT_ICRU"])
    a = physics.build_photon_tables(mats, e_max_mev=0.2)
    monkeypatch.setenv("GATE_B200_NO_MODEL_CACHE", "1")
    for f in (pm._element_models_file, pm._element_models_coarse, pm._element_split_master):
        f.cache_clear()
    b = physics.build_photon_tables(mats, e_max_mev=0.2)
    for f in (pm._element_models_file, pm._element_models_coarse, pm._element_split_master):
        f.cache_clear()
    for k in ("mu_total", "cum_phot", "cum_compt", "cum_rayl", "rayl_cum_ff2", "compt_sf", "density"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
