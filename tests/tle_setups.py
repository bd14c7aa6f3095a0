"""Small TLE configurations shared by the CPU and GPU tests (test infrastructure)."""
import numpy as np

from opengate_b200 import abi, physics, setups
from small_setups import ALL_FLAGS, c1, c3, c4

SQ = abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ


def _w2i(S, k=0):
    return np.array(S.dose[k].world_to_index[:]).reshape(3, 4)


def c1_tle(n=5000, hit=abi.HIT_RANDOM, flags=SQ, **kw):
    """homogeneous Box (one-voxel label volume) + 6 mm TLE grid: chord = whole flight, random point"""
    S = c1(n)
    S.add_tle_actor(tuple(S.dose[0].size), _w2i(S), flags, hit_type=hit, database="NIST", **kw)
    return S


def c3_tle(n=5000, flags=SQ, hit=abi.HIT_RANDOM, coarse=False, **kw):
    """CT phantom; TLE grid = image grid (aligned) or a coarser grid (unaligned, 2x2x2 voxels)"""
    S = c3(n)
    if coarse:
        size = tuple(int(x) // 2 for x in S.dose[0].size)
        sp = tuple(2.0 * s for s in S.spacing)
        A = setups.dose_world_to_index(size, sp)
    else:
        size, A = tuple(S.dose[0].size), _w2i(S)
    S.add_tle_actor(size, A, flags, hit_type=hit, database="EPDL", **kw)
    return S


def c4_tle(n=5000, flags=SQ):
    S = c4(n)
    S.add_tle_actor(tuple(S.dose[0].size), _w2i(S), flags, database="EPDL")
    return S


def long_thin(n=200, nvox=4000):
    """4000 x 1 x 1 voxels of 0.1 mm air, photons along +x: one flight crosses more voxels than the
    per-history table holds (3072) -> overflow path"""
    from opengate_b200 import materials
    db = materials.MaterialDatabase()
    db.read_from_file(materials.DEFAULT_DB)
    mats = [db.find_or_build_material("G4_AIR"), db.find_or_build_material("G4_WATER")]
    T = physics.build_photon_tables(mats, e_max_mev=0.1)
    labels = np.zeros((1, 1, nvox), np.uint8)
    labels[0, 0, ::7] = 1
    sp = (0.1, 5.0, 5.0)
    s, keep = abi.make_source(pos_type=abi.POS_POINT, pos_translation=(-nvox * 0.05 - 1.0, 0.3, -0.2), dir_type=abi.DIR_MOMENTUM,
                              momentum=(1, 0.0005, 0.0003), ene_type=abi.ENE_MONO, mono=0.08)
    A = setups.dose_world_to_index((nvox, 1, 1), sp)
    d = abi.make_dose_desc((nvox, 1, 1), A, ALL_FLAGS, abi.HIT_POST)
    S = setups.LowLevelSetup("long thin air/water stack", 7, T, (400, 400, 400), 0, labels, sp, setups.identity34(),
                             [(s, keep)], [d], [], n, materials=mats)
    S.add_tle_actor((nvox, 1, 1), A, SQ, database="NIST")
    return S
