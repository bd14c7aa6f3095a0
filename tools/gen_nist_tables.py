#!/usr/bin/env python
"""Generate opengate_b200/data/nist_photon_totals.npz from the reference's NIST module.

Runs ONLY where /root/reference exists (the authoring container). It imports
`opengate/data/PhotonAttenuation.py` (Hubbell-Seltzer NIST tables, numpy+scipy only;
reference file: opengate/data/PhotonAttenuation.py:17 `energy`, :58 `mac`, :3863 `meac`,
:7365 `edges`, :8556 `PhotonAttenuationEl`) and resamples, with the reference's own
interpolation rule (log-log PCHIP, linear next to absorption edges), the total mass
attenuation coefficient mu/rho (Z=1..100) and the mass energy-absorption coefficient
mu_en/rho (Z=1..92) onto a log-uniform master grid of 2048 nodes from 1 keV to 20 MeV.

The output is physics DATA (NIST XAAMDI, public domain), not reference source code. The
engine needs it at run time on machines where the reference package is not importable.
"""
import sys
import os
import numpy as np

REF = os.environ.get("GATE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "opengate", "data"))
import PhotonAttenuation as P  # noqa: E402

N_MASTER = 2048
E_MIN, E_MAX = 1.0e-3, 20.0  # MeV


def main():
    e = np.exp(np.linspace(np.log(E_MIN), np.log(E_MAX), N_MASTER))
    e[0], e[-1] = E_MIN, E_MAX
    mac = np.zeros((100, N_MASTER))
    meac = np.zeros((100, N_MASTER))
    for z in range(1, 101):
        if z <= 10:
            mac[z - 1] = P.PhotonAttenuationEl(z, e, 1)[0]
            meac[z - 1] = P.PhotonAttenuationEl(z, e, 2)[0]
        else:
            # the reference's edge-aware branch only accepts scalar energies
            for i, ei in enumerate(e):
                mac[z - 1, i] = float(P.PhotonAttenuationEl(z, float(ei), 1)[0])
                if z <= P.meac.shape[0]:
                    meac[z - 1, i] = float(P.PhotonAttenuationEl(z, float(ei), 2)[0])
    z_over_a = np.array([float(r[0]) for r in P.PropsEl[:100]])
    symbols = np.array([str(r[2]) for r in P.PropsEl[:100]])
    names = np.array([str(r[3]) for r in P.PropsEl[:100]])
    # K/L... absorption edges (MeV) per element so that the device grid builder can refine
    edges = np.asarray(P.edges[:, :2], dtype=np.float64)
    out = os.path.join(os.path.dirname(__file__), "..", "opengate_b200", "data",
                       "nist_photon_totals.npz")
    np.savez_compressed(
        out,
        energy_MeV=e,
        mu_over_rho_cm2_g=mac.astype(np.float32),
        mu_en_over_rho_cm2_g=meac.astype(np.float32),
        z_over_a=z_over_a,
        symbol=symbols,
        name=names,
        edges_Z_EMeV=edges,
        nist_energy_MeV=np.asarray(P.energy, dtype=np.float64),
        nist_mac=np.asarray(P.mac, dtype=np.float64),
        nist_meac=np.asarray(P.meac, dtype=np.float64),
    )
    print("wrote", os.path.abspath(out), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
