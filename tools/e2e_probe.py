import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ["GATE_B200_TRACE"] = "1"
from bench import build_setup
from opengate_b200.core import Engine
S = build_setup("c3", 1000000)
for i in range(3):
    print("--- engine", i, file=sys.stderr)
    t0 = time.perf_counter(); eng = Engine(0, S.seed); t1 = time.perf_counter()
    sids, dids, pids = S.apply(eng); t2 = time.perf_counter()
    eng.begin_run(); eng.run_source(sids[0], 0, 1000000); eng.sync(); t3 = time.perf_counter()
    eng.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} apply {1e3*(t2-t1):.2f} run {1e3*(t3-t2):.2f} close {1e3*(t4-t3):.2f}", file=sys.stderr)
