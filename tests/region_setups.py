"""Set-ups with homogeneous Box regions of their own material in the world (test infrastructure)."""
import numpy as np

from opengate_b200 import abi, setups


def slab_in_front(n=20000, thickness=40.0, tilt_deg=0.0, energy=0.2, material=0, planes=True, slab_is_plane=False, touching=False):
    """c1's water box and pencil beam (source moved back to z = -400 mm); a slab of `material` (0 = water, 1 = air)
    across the beam at z = -300 mm, tilted about x; two thin scoring planes of world material before and behind it."""
    S = setups.make_c1(n_primaries=n, grid=(50, 50, 50), spacing_mm=6.0, energy=energy,
                       flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_COUNTS)
    s, keep = abi.make_source(pos_type=abi.POS_POINT, pos_translation=(0, 0, -400.0), dir_type=abi.DIR_MOMENTUM,
                              momentum=(0, 0, 1), ene_type=abi.ENE_MONO, mono=energy)
    S.sources = [(s, keep)]
    c, sn = np.cos(np.radians(tilt_deg)), np.sin(np.radians(tilt_deg))
    R = np.array([[1.0, 0.0, 0.0], [0.0, c, -sn], [0.0, sn, c]])      # box -> world
    # touching: the slab lies flat on the -z face of the water box (z = -150 mm), like a couch under a phantom
    t = np.array([0.0, 0.0, -150.0 - thickness / 2 if touching else -300.0])
    w2b = np.zeros((3, 4))
    w2b[:, :3] = R.T
    w2b[:, 3] = -R.T @ t
    S.regions = [abi.make_region_desc((150.0, 150.0, thickness / 2), w2b, material)]
    if slab_is_plane:
        # the slab itself carries a phase-space actor (entering + exiting): region and scoring plane at once
        S.phsp.append(abi.make_phsp_desc((150.0, 150.0, thickness / 2), w2b, abi.PHSP_ENTERING | abi.PHSP_EXITING, 8 * n + 16))
    if planes:
        for z in (-390.0, -200.0):
            B = np.zeros((3, 4))
            B[:, :3] = np.eye(3)
            B[:, 3] = (0.0, 0.0, -z)
            S.phsp.append(abi.make_phsp_desc((200.0, 200.0, 0.5), B, abi.PHSP_ENTERING, 4 * n + 16))   # room for the photons the water box scatters back
    return S
