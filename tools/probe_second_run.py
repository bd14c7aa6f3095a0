"""Two Simulation.run() calls of the bench's script-level workload in one process, the second one profiled
(development aid: why is the second run slower than the first?)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
t0 = time.perf_counter(); r1 = bench._script_level_run(n); print("run 1", r1["wall_s"], flush=True)
pr = cProfile.Profile(); pr.enable()
r2 = bench._script_level_run(n)
pr.disable(); print("run 2", r2["wall_s"], flush=True)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
