#!/usr/bin/env python
"""Generate opengate_b200/data/mu_database.npz from the reference's mu / mu_en database.

Runs ONLY where /root/reference exists (the authoring container).  It parses the numeric
arrays of core/opengate_core/opengate_lib/GateMuDatabase.h (NIST: :17-26, EPDL: :2112-2123):
per element Z = 0..N a run of (energy MeV, mu/rho cm2/g, mu_en/rho cm2/g) triples, the run
lengths in *_energyNumber.  These are the tables GateMaterialMuHandler::InitElementTable
(GateMaterialMuHandler.cpp:265-295) loads for `database = "NIST" | "EPDL"` of the TLEDoseActor.

The output is physics DATA (NIST XAAMDI / LLNL EPDL97, public domain), stored as flat float64
arrays + offsets; no reference source code is copied.
"""
import os
import re
import numpy as np

REF = os.environ.get("GATE_REFERENCE", "/root/reference")
HDR = os.path.join(REF, "core", "opengate_core", "opengate_lib", "GateMuDatabase.h")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "opengate_b200", "data", "mu_database.npz")


def array_of(text, name):
    m = re.search(r"static const (?:int|double) " + name + r"\[\d+\]\s*=\s*\{(.*?)\};", text, re.S)
    return m.group(1)


def main():
    text = open(HDR).read()
    out = {}
    for db in ("NIST", "EPDL"):
        counts = np.array([int(v) for v in array_of(text, db + "_mu_muen_data_energyNumber").replace("\n", " ").split(",") if v.strip()])
        data = np.array([float(v) for v in array_of(text, db + "_mu_muen_data").replace("\n", " ").split(",") if v.strip()])
        n_el = int(re.search(db + r"_mu_muen_data_elementNumber = (\d+);", text).group(1))
        assert len(counts) == n_el + 1 and len(data) == 3 * counts.sum(), (db, len(counts), len(data), counts.sum())
        out[db + "_counts"] = counts.astype(np.int32)
        out[db + "_data"] = data.reshape(-1, 3)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
