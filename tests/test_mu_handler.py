"""mu/rho, mu_en/rho database of the TLEDoseActor: the numpy restatement
(opengate_b200/mu_handler.py) against the REFERENCE's own code — GateMuTables.cpp + GateMuDatabase.h
compiled unmodified into oracle/_ref/libref_mu.so (oracle/Makefile `ref`) — and against known NIST
values.  The look-ups must agree to the last bit but for libm's exp/log (rtol 1e-14)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from opengate_b200 import materials as pmat
from opengate_b200 import mu_handler as mh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_mu.so")
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_LIB) and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libref_mu.so not built (reference tree absent)")
    return C.CDLL(REF_LIB)


def _ref_nodes(L, db, z):
    n = L.ref_mu_element_size(db, z)
    e, mu, en = np.zeros(n), np.zeros(n), np.zeros(n)
    L.ref_mu_element_nodes(db, z, e.ctypes.data_as(DP), mu.ctypes.data_as(DP), en.ctypes.data_as(DP))
    return e, mu, en


def _ref_lookup(L, t, energies):
    e = np.ascontiguousarray(energies, float)
    out = [np.zeros(len(e)) for _ in range(4)]
    L.ref_mu_table_lookup(len(t.ln_e), t.ln_e.ctypes.data_as(DP), t.ln_mu.ctypes.data_as(DP),
                          t.ln_mu_en.ctypes.data_as(DP), C.c_double(t.density), len(e), e.ctypes.data_as(DP),
                          *[o.ctypes.data_as(DP) for o in out])
    return out


@pytest.mark.parametrize("db,name", [(0, "NIST"), (1, "EPDL")])
def test_element_tables_identical_to_reference_database(ref, db, name):
    T = mh.element_tables(name)
    nel = 92 if name == "NIST" else 100
    assert ref.ref_mu_element_size(db, nel + 1) == -1
    for z in range(0, nel + 1):
        e, mu, en = _ref_nodes(ref, db, z)
        assert np.array_equal(e, T[z][0]) and np.array_equal(mu, T[z][1]) and np.array_equal(en, T[z][2]), z


@pytest.mark.parametrize("name", ["NIST", "EPDL"])
def test_lookup_matches_reference_GateMuTable(ref, name):
    db = pmat.MaterialDatabase()
    db.read_from_file(pmat.DEFAULT_DB)
    rng = np.random.default_rng(5)
    for mat in ("G4_WATER", "G4_AIR", "RibBone", "Lung", "G4_LEAD_OXIDE"):
        try:
            spec = db.find_or_build_material(mat)
        except Exception:
            continue
        t = mh.material_table(spec, name)
        nodes = np.exp(t.ln_e)
        # random energies, every node exactly, below the first and above the last node
        e = np.concatenate([np.exp(rng.uniform(np.log(5e-4), np.log(3.0), 400)), nodes, [1e-4, 50.0]])
        mu_r, en_r, mul_r, enl_r = _ref_lookup(ref, t, e)
        mine = np.array([[t.mu_over_rho(x), t.mu_en_over_rho(x), t.mu(x), t.mu_en(x)] for x in e])
        np.testing.assert_allclose(mine[:, 0], mu_r, rtol=1e-14)
        np.testing.assert_allclose(mine[:, 1], en_r, rtol=1e-14)
        np.testing.assert_allclose(mine[:, 2], mul_r, rtol=1e-14)
        np.testing.assert_allclose(mine[:, 3], enl_r, rtol=1e-14)


def test_single_element_material_reproduces_element_nodes():
    for name in ("NIST", "EPDL"):
        T = mh.element_tables(name)
        for z in (1, 6, 13, 26, 82):
            t = mh.construct_material([z], [1.0], 1.0, name)
            e, mu, en = T[z]
            assert len(t.ln_e) == len(e)            # edges (doubled energies) are kept twice
            np.testing.assert_allclose(np.exp(t.ln_mu), mu, rtol=1e-14)
            np.testing.assert_allclose(np.exp(t.ln_mu_en), en, rtol=1e-14)


def test_mixture_rule_union_grid_and_edges():
    # water: union of the H and O grids, cross-table duplicates once; values = mass-fraction mix
    T = mh.element_tables("NIST")
    w = {1: 0.111894, 8: 0.888106}
    t = mh.construct_material([1, 8], [w[1], w[8]], 1.0, "NIST")
    e = np.exp(t.ln_e)
    assert np.all(np.diff(e) >= 0)
    union = sorted(set(T[1][0]) | set(T[8][0]))
    dup_o = len(T[8][0]) - len(set(T[8][0]))      # oxygen K edge below 1 keV -> none in range
    assert len(e) == len(union) + dup_o
    # at a node shared by both elements the mix is exact
    k = int(np.argmin(np.abs(e - 0.1)))
    i1 = int(np.argmin(np.abs(T[1][0] - 0.1))); i8 = int(np.argmin(np.abs(T[8][0] - 0.1)))
    assert np.exp(t.ln_mu[k]) == pytest.approx(w[1] * T[1][1][i1] + w[8] * T[8][1][i8], rel=1e-14)
    # NIST XAAMDI water: mu/rho(100 keV) = 0.1707, mu_en/rho = 0.02546 cm2/g
    assert t.mu_over_rho(0.1) == pytest.approx(0.1707, rel=2e-3)
    assert t.mu_en_over_rho(0.1) == pytest.approx(0.02546, rel=2e-3)
    # a material with an absorption edge inside the range keeps the doubled node and jumps there
    pb = mh.construct_material([82, 8], [0.9283, 0.0717], 9.53, "NIST")
    le = pb.ln_e
    dbl = np.where(np.diff(le) == 0)[0]
    assert len(dbl) >= 3                                # Pb M, L, K edges
    kedge = dbl[-1]
    assert np.exp(pb.ln_mu[kedge + 1]) > 2.0 * np.exp(pb.ln_mu[kedge])   # K-edge jump at 88 keV


def test_lookup_outside_and_on_nodes():
    t = mh.construct_material([8], [1.0], 1.0, "NIST")
    first, last = np.exp(t.ln_e[0]), np.exp(t.ln_e[-1])
    assert t.mu_over_rho(first * 0.5) == pytest.approx(np.exp(t.ln_mu[0]))
    # above the last node the binary search ends on (n-2, n-1) and returns the LOWER node's value
    # (GateMuTables.cpp:66-70: the value is only interpolated strictly inside the interval)
    assert t.mu_en_over_rho(last * 2.0) == pytest.approx(np.exp(t.ln_mu_en[-2]))
    assert t.mu_over_rho(float(np.exp(t.ln_e[5]))) == pytest.approx(np.exp(t.ln_mu[5]), rel=1e-12)


def test_bad_database_name():
    with pytest.raises(ValueError):
        mh.element_tables("simulated")
