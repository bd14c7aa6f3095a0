"""Small versions of the BASELINE.json configs that the CPU oracle finishes in seconds."""
from opengate_b200 import abi, setups

ALL_FLAGS = abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ | abi.SCORE_COUNTS


def c1(n=20000, hit_type=abi.HIT_POST, **kw):
    return setups.make_c1(n_primaries=n, grid=(50, 50, 50), spacing_mm=6.0, flags=ALL_FLAGS, hit_type=hit_type, **kw)


def c2(n=20000):
    return setups.make_c2(n=48, n_primaries=n, flags=ALL_FLAGS, field_mm=30.0)


def c3(n=20000):
    return setups.make_c3(nx=64, ny=64, nz=40, spacing=(7.8, 7.8, 7.5), n_primaries=n, flags=ALL_FLAGS)


def c4(n=20000):
    return setups.make_c4(nx=64, ny=64, nz=40, spacing=(8, 8, 7.5), n_act=32, act_spacing=16.0, n_primaries=n)


SMALL = {"c1": c1, "c2": c2, "c3": c3, "c4": c4}
