#!/usr/bin/env python
"""Small runs of every kernel variant for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck|initcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from opengate_b200 import abi  # noqa: E402
from opengate_b200.core import Engine  # noqa: E402
import small_setups  # noqa: E402
import tle_setups  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cases = {"c1": small_setups.c1(n), "c2": small_setups.c2(n), "c3": small_setups.c3(n), "c4": small_setups.c4(n),
         "c3_tle_sq": tle_setups.c3_tle(n), "c3_tle_nosq": tle_setups.c3_tle(n, flags=abi.SCORE_EDEP | abi.SCORE_DOSE),
         "c4_tle": tle_setups.c4_tle(n)}
S = small_setups.c3(n)
S.labels = S.labels.astype(np.uint16)
cases["c3_u16"] = S
S = small_setups.c3(n)
S.dose[0].hit_type = abi.HIT_RANDOM          # on the image's own grid: scored like post, kernel variant 2
cases["c3_hit_random"] = S
cases["c1_hit_random"] = small_setups.c1(n, hit_type=abi.HIT_RANDOM)   # a grid of its own: hit-point emulation, variant 3
S = small_setups.c3(n)
S.dose[0].hit_type = abi.HIT_PRE             # pre-step points keep the emulation on any grid: variant 3
cases["c3_hit_pre"] = S
S = small_setups.c3(n)
S.dose[0].hit_type = abi.HIT_POST
S.dose[0].score_flags = abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE
cases["c3_bench_shape"] = S
import region_setups  # noqa: E402
cases["box_region"] = region_setups.slab_in_front(n, 25.0, 30.0)
cases["c3_tle_sq_table"] = tle_setups.c3_tle(n)      # the same actor forced through the table variant (1) of the kernel
cases["c3_tle_sq_hybrid_cap2"] = tle_setups.c3_tle(n, threshold=0.06)   # record variant: point deposits, record overflows
for name, S in cases.items():
    os.environ.pop("GATE_B200_TLE_TABLE", None)
    os.environ.pop("GATE_B200_TLE_REC_CAP", None)
    if name.endswith("_table"):
        os.environ["GATE_B200_TLE_TABLE"] = "1"
    if name.endswith("_cap2"):
        os.environ["GATE_B200_TLE_REC_CAP"] = "2"
    eng = Engine(0, S.seed)
    se, de, pe = S.apply(eng)
    eng.begin_run()
    eng.run_source(se[0], 0, n)
    eng.sync()
    st = eng.stats()
    for d in de:
        eng.dose_grid(d, abi.GRID_EDEP)
        eng.dose_grid(d, abi.GRID_EDEP, scale=0.5)
        nx, ny, nz = eng._dose_sizes[d]
        out = np.zeros((nz, ny, nx))
        eng.dose_grid_range(d, abi.GRID_EDEP, 7, nx * ny * nz - 11, out)
    if name == "c3_bench_shape":
        eng.uncertainty_grid(de[0], abi.GRID_EDEP, n)
        eng.mean_uncertainty(de[0], n, 5, 0.7)
    for p in pe:
        eng.phsp(p)
    print(name, eng.kernel_name(), st["histories"], st["tentative_collisions"], st["tle_steps"], st["tle_overflows"], flush=True)
    eng.close()
