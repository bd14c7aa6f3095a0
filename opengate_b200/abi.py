"""ctypes mirror of include/gate_b200.h (the C-ABI structs and enums).

Kept separate from the library loader so that test infrastructure (the CPU oracle, which
takes the same descriptor structs) can share the definitions without loading CUDA.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 5

# enums --------------------------------------------------------------------------------
POS_POINT, POS_BOX, POS_SPHERE, POS_DISC, POS_CYLINDER, POS_VOXELS, POS_PHSP, POS_SURFACE_SPHERE = range(8)
DIR_ISO, DIR_MOMENTUM, DIR_FOCUSED, DIR_PHSP, DIR_COS, DIR_BEAM2D, DIR_HISTOGRAM = range(7)
(ENE_MONO, ENE_GAUSS, ENE_RANGE, ENE_SPECTRUM_DISCRETE, ENE_SPECTRUM_HISTOGRAM,
 ENE_SPECTRUM_HISTOGRAM_LINEAR, ENE_CDF, ENE_PHSP) = range(8)
HIT_RANDOM, HIT_PRE, HIT_POST, HIT_MIDDLE = range(4)
SCORE_EDEP, SCORE_EDEP_SQ, SCORE_DOSE, SCORE_DOSE_SQ, SCORE_COUNTS = 1, 2, 4, 8, 16
GRID_EDEP, GRID_EDEP_SQ, GRID_DOSE, GRID_DOSE_SQ, GRID_COUNTS = range(5)
PHSP_ENTERING, PHSP_EXITING = 1, 2

HIT_TYPES = {"random": HIT_RANDOM, "pre": HIT_PRE, "post": HIT_POST, "middle": HIT_MIDDLE}

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)


class Tables(C.Structure):
    _fields_ = [
        ("n_materials", C.c_int32), ("n_energy", C.c_int32),
        ("e_min_MeV", C.c_double), ("e_max_MeV", C.c_double), ("e_cut_MeV", C.c_double),
        ("density_g_cm3", _fp), ("mu_total_per_mm", _fp), ("cum_phot", _fp),
        ("cum_compt", _fp), ("cum_rayl", _fp),
        ("n_q", C.c_int32), ("q2_min", C.c_double), ("q2_max", C.c_double),
        ("rayl_cum_ff2", _fp), ("compt_sf", _fp), ("pair_mode", C.c_int32),
    ]


class Source(C.Structure):
    _fields_ = [
        ("pos_type", C.c_int32), ("half_size", C.c_double * 3),
        ("radius", C.c_double), ("sigma_x", C.c_double), ("sigma_y", C.c_double),
        ("pos_rotation", C.c_double * 9), ("pos_translation", C.c_double * 3),
        ("dir_type", C.c_int32), ("theta", C.c_double * 2), ("phi", C.c_double * 2),
        ("momentum", C.c_double * 3), ("focus_point", C.c_double * 3),
        ("dir_relative", C.c_int32), ("dir_rotation", C.c_double * 9),
        ("ene_type", C.c_int32), ("mono_MeV", C.c_double), ("sigma_MeV", C.c_double),
        ("e_min_MeV", C.c_double), ("e_max_MeV", C.c_double),
        ("n_spectrum", C.c_int32), ("spectrum_energy", _dp), ("spectrum_cdf", _dp),
        ("weight", C.c_double),
        ("vox_size", C.c_int32 * 3), ("vox_spacing", C.c_double * 3), ("vox_origin", C.c_double * 3),
        ("cdf_z", _fp), ("cdf_y", _fp), ("cdf_x", _fp),
        ("phsp_n_entries", C.c_uint64), ("phsp_entry_start", C.c_uint64),
        ("phsp_pos_x", _fp), ("phsp_pos_y", _fp), ("phsp_pos_z", _fp),
        ("phsp_dir_x", _fp), ("phsp_dir_y", _fp), ("phsp_dir_z", _fp),
        ("phsp_energy", _fp), ("phsp_weight", _fp),
        ("dir_sigma", C.c_double * 2),
        ("n_dir_theta", C.c_int32), ("n_dir_phi", C.c_int32),
        ("dir_theta_edges", _dp), ("dir_theta_cdf", _dp), ("dir_phi_edges", _dp), ("dir_phi_cdf", _dp),
        ("confine", C.c_int32), ("confine_half_size", C.c_double * 3), ("world_to_confine", C.c_double * 12),
    ]


class DoseDesc(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("world_to_index", C.c_double * 12),
                ("score_flags", C.c_int32), ("hit_type", C.c_int32)]


class TleDesc(C.Structure):
    _fields_ = [("energy_min_MeV", C.c_double), ("energy_threshold_MeV", C.c_double),
                ("mu_en_over_rho", _fp)]


class PhspDesc(C.Structure):
    _fields_ = [("half_size", C.c_double * 3), ("world_to_box", C.c_double * 12),
                ("steps_to_store", C.c_int32), ("capacity", C.c_uint64)]


class RegionDesc(C.Structure):
    _fields_ = [("half_size", C.c_double * 3), ("world_to_box", C.c_double * 12), ("material", C.c_int32)]


def make_region_desc(half_size, world_to_box12, material):
    d = RegionDesc()
    d.half_size = (C.c_double * 3)(*[float(x) for x in half_size])
    d.world_to_box = (C.c_double * 12)(*np.asarray(world_to_box12, float).reshape(-1))
    d.material = int(material)
    return d


class PhspColumns(C.Structure):
    _fields_ = [("kinetic_energy", _fp), ("position", _fp), ("direction", _fp), ("weight", _fp),
                ("time_ns", _fp), ("event_id", C.POINTER(C.c_int64)), ("flags", C.POINTER(C.c_int32))]


class Stats(C.Structure):
    _fields_ = [
        ("histories", C.c_uint64), ("tracks", C.c_uint64), ("steps", C.c_uint64),
        ("tentative_collisions", C.c_uint64),
        ("n_phot", C.c_uint64), ("n_compt", C.c_uint64), ("n_rayl", C.c_uint64), ("n_conv", C.c_uint64),
        ("deposits", C.c_uint64), ("flushed_entries", C.c_uint64), ("buffer_overflows", C.c_uint64),
        ("escaped", C.c_uint64), ("phsp_rows", C.c_uint64), ("phsp_dropped", C.c_uint64),
        ("absorbed_events", C.c_uint64),
        ("edep_total_MeV", C.c_double), ("kernel_ms", C.c_double),
        ("tle_steps", C.c_uint64), ("tle_entries", C.c_uint64), ("tle_overflows", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# helpers -------------------------------------------------------------------------------

def fptr(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def dptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def affine12(rotation, translation):
    """Row-major 3x4 [R|t] as a ctypes double[12]."""
    r = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
    t = np.asarray(translation, dtype=np.float64).reshape(3)
    m = np.concatenate([r, t[:, None]], axis=1).reshape(-1)
    return (C.c_double * 12)(*m)


def make_tables(t):
    """gate_b200_tables from a physics.PhotonTables; returns (struct, keepalive)."""
    s = Tables()
    s.n_materials = len(t.names)
    s.n_energy = t.n_energy
    s.e_min_MeV, s.e_max_MeV, s.e_cut_MeV = t.e_min, t.e_max, t.e_cut
    s.density_g_cm3 = fptr(t.density)
    s.mu_total_per_mm = fptr(t.mu_total)
    s.cum_phot, s.cum_compt, s.cum_rayl = fptr(t.cum_phot), fptr(t.cum_compt), fptr(t.cum_rayl)
    s.n_q, s.q2_min, s.q2_max = t.n_q, t.q2_min, t.q2_max
    s.rayl_cum_ff2, s.compt_sf = fptr(t.rayl_cum_ff2), fptr(t.compt_sf)
    s.pair_mode = t.pair_mode
    return s, t


def make_source(pos_type=POS_POINT, half_size=(0, 0, 0), radius=0.0, sigma_x=0.0, sigma_y=0.0,
                pos_rotation=None, pos_translation=(0, 0, 0), dir_type=DIR_MOMENTUM,
                theta=(0.0, np.pi), phi=(0.0, 2 * np.pi), momentum=(0, 0, 1),
                focus_point=(0, 0, 0), dir_relative=False, dir_rotation=None,
                ene_type=ENE_MONO, mono=0.0, sigma=0.0, e_min=0.0, e_max=0.0,
                spectrum_energy=None, spectrum_cdf=None, weight=1.0,
                vox_size=(0, 0, 0), vox_spacing=(1, 1, 1), vox_origin=(0, 0, 0),
                cdf_z=None, cdf_y=None, cdf_x=None, phsp=None, phsp_entry_start=0, dir_sigma=(0.0, 0.0),
                theta_hist=None, phi_hist=None, confine_half_size=None, world_to_confine=None):
    """Fill a gate_b200_source; returns (struct, keepalive list).
    phsp: dict of float32 arrays pos_x/y/z, energy and optionally dir_x/y/z, weight (phase-space replay).
    theta_hist / phi_hist (dir_type HISTOGRAM): (edges [n+1] rad, weights [n]) of the user angular histograms.
    confine_half_size / world_to_confine (3x4): box the position is confined to."""
    s = Source()
    keep = []
    s.pos_type = pos_type
    s.half_size = (C.c_double * 3)(*[float(x) for x in half_size])
    s.radius, s.sigma_x, s.sigma_y = float(radius), float(sigma_x), float(sigma_y)
    eye = np.eye(3)
    s.pos_rotation = (C.c_double * 9)(*np.asarray(eye if pos_rotation is None else pos_rotation, float).reshape(-1))
    s.pos_translation = (C.c_double * 3)(*[float(x) for x in pos_translation])
    s.dir_type = dir_type
    s.theta = (C.c_double * 2)(*[float(x) for x in theta])
    s.phi = (C.c_double * 2)(*[float(x) for x in phi])
    mom = np.asarray(momentum, float)
    n = np.linalg.norm(mom)
    if n > 0:
        mom = mom / n
    s.momentum = (C.c_double * 3)(*mom)
    s.focus_point = (C.c_double * 3)(*[float(x) for x in focus_point])
    s.dir_relative = 1 if dir_relative else 0
    s.dir_sigma = (C.c_double * 2)(*[float(x) for x in dir_sigma])
    s.dir_rotation = (C.c_double * 9)(*np.asarray(eye if dir_rotation is None else dir_rotation, float).reshape(-1))
    s.ene_type = ene_type
    s.mono_MeV, s.sigma_MeV, s.e_min_MeV, s.e_max_MeV = float(mono), float(sigma), float(e_min), float(e_max)
    if spectrum_cdf is not None:
        se = np.ascontiguousarray(spectrum_energy, dtype=np.float64)
        sc = np.ascontiguousarray(spectrum_cdf, dtype=np.float64)
        keep += [se, sc]
        s.n_spectrum = len(sc)
        s.spectrum_energy, s.spectrum_cdf = dptr(se), dptr(sc)
    s.weight = float(weight)
    s.vox_size = (C.c_int32 * 3)(*[int(x) for x in vox_size])
    s.vox_spacing = (C.c_double * 3)(*[float(x) for x in vox_spacing])
    s.vox_origin = (C.c_double * 3)(*[float(x) for x in vox_origin])
    if cdf_x is not None:
        cz = np.ascontiguousarray(cdf_z, dtype=np.float32)
        cy = np.ascontiguousarray(cdf_y, dtype=np.float32)
        cx = np.ascontiguousarray(cdf_x, dtype=np.float32)
        keep += [cz, cy, cx]
        s.cdf_z, s.cdf_y, s.cdf_x = fptr(cz), fptr(cy), fptr(cx)
    for name, hist in (("theta", theta_hist), ("phi", phi_hist)):
        if hist is None:
            continue
        edges = np.ascontiguousarray(hist[0], dtype=np.float64)
        w = np.asarray(hist[1], dtype=np.float64)
        if len(edges) != len(w) + 1 or len(w) < 1 or np.any(w < 0) or w.sum() <= 0 or np.any(np.diff(edges) < 0):
            raise ValueError(f"{name} histogram: n weights >= 0 (not all zero) and n + 1 non-decreasing angles are needed")
        cdf = np.ascontiguousarray(np.concatenate([[0.0], np.cumsum(w) / w.sum()]))
        cdf[-1] = 1.0
        keep += [edges, cdf]
        setattr(s, f"n_dir_{name}", len(w))
        setattr(s, f"dir_{name}_edges", dptr(edges))
        setattr(s, f"dir_{name}_cdf", dptr(cdf))
    if confine_half_size is not None:
        s.confine = 1
        s.confine_half_size = (C.c_double * 3)(*[float(x) for x in confine_half_size])
        s.world_to_confine = (C.c_double * 12)(*np.asarray(world_to_confine, float).reshape(-1))
    if phsp is not None:
        cols = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in phsp.items() if v is not None}
        n = len(cols["pos_x"])
        for k, v in cols.items():
            if len(v) != n:
                raise ValueError(f"phase-space column '{k}' has {len(v)} entries, expected {n}")
            keep.append(v)
            setattr(s, "phsp_" + k, fptr(v))
        s.phsp_n_entries = n
        s.phsp_entry_start = int(phsp_entry_start)
    return s, keep


def make_dose_desc(size, world_to_index12, score_flags, hit_type):
    d = DoseDesc()
    d.size = (C.c_int32 * 3)(*[int(x) for x in size])
    d.world_to_index = (C.c_double * 12)(*np.asarray(world_to_index12, float).reshape(-1))
    d.score_flags = int(score_flags)
    d.hit_type = int(hit_type)
    return d


def make_tle_desc(mu_en_over_rho, energy_threshold=float("inf"), energy_min=0.0):
    """gate_b200_tle_desc; mu_en_over_rho: float32 [n_materials][n_energy] on the table's energy nodes.
    Returns (struct, keepalive)."""
    t = TleDesc()
    a = np.ascontiguousarray(mu_en_over_rho, dtype=np.float32)
    t.energy_min_MeV = float(energy_min)
    t.energy_threshold_MeV = float(energy_threshold)
    t.mu_en_over_rho = fptr(a)
    return t, a


def make_phsp_desc(half_size, world_to_box12, steps_to_store, capacity):
    d = PhspDesc()
    d.half_size = (C.c_double * 3)(*[float(x) for x in half_size])
    d.world_to_box = (C.c_double * 12)(*np.asarray(world_to_box12, float).reshape(-1))
    d.steps_to_store = int(steps_to_store)
    d.capacity = int(capacity)
    return d
