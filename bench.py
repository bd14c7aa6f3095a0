#!/usr/bin/env python
"""bench.py — photon primaries/s of the GPU dose engine on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c1|c4] [--histories H]
    python bench.py --impl reference ...      # the CPU arm (oracle port on all host cores)

A "step" is one batch of H photon histories through the hot path on an already configured
engine (label volume, tables and accumulators resident in HBM).  N > 1 (torchrun): one process
per GPU, histories sharded by global history id (weak scaling: H per GPU per step), one NCCL
reduce of every accumulator grid at the end of the run, inside the timed region.
Prints ONE JSON line (see the field list in DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "photon primaries/sec"
UNIT = "primaries/s"


def build_setup(name, histories):
    from opengate_b200 import setups
    if name == "c3":
        return setups.make_c3(n_primaries=histories)
    if name == "c2":
        return setups.make_c2(n_primaries=histories)
    if name == "c1":
        return setups.make_c1(n_primaries=histories)
    if name == "c4":
        return setups.make_c4(n_primaries=histories)
    if name in ("c3tle", "c3tle_nosq"):
        # config 3 scored by the track-length estimator (TLEDoseActor, EPDL database) instead of the
        # analog DoseActor: dose + edep with event-wise squares (c3tle) or plain sums (c3tle_nosq)
        from opengate_b200 import abi
        S = setups.make_c3(n_primaries=histories)
        d0 = S.dose[0]
        A = np.array(d0.world_to_index[:]).reshape(3, 4)
        flags = abi.SCORE_EDEP | abi.SCORE_DOSE | (abi.SCORE_EDEP_SQ if name == "c3tle" else 0)
        S.dose = []
        S.add_tle_actor(tuple(d0.size), A, flags, database="EPDL")
        S.name = S.name.replace("DoseActor dose+edep2", "TLEDoseActor dose+edep" + ("+edep2" if name == "c3tle" else ""))
        return S
    if name == "c3_hit_random":
        # config 3 with the reference's DEFAULT hit_type ("random": a point between the pre- and post-step
        # points of the Geant4 step).  A Geant4 step never leaves its voxel, so on a dose grid equal to the image
        # grid that point lies in the voxel of the interaction: engine and oracle score it like "post" (no
        # emulation, no random number; the kernel variant of the main line); grids of their own keep the
        # hit-point emulation (kernel variant 3).  This variant shows that the default costs nothing here.
        from opengate_b200 import abi
        S = setups.make_c3(n_primaries=histories)
        S.dose[0].hit_type = abi.HIT_RANDOM
        S.name += " hit_type=random"
        return S
    if name == "c5ct":
        # config 5 on the CT: the 6 MV spectrum of config 2 into the 512x512x300 synthetic CT (dose + squares)
        return setups.make_c3(n_primaries=histories, spectrum="6mv")
    if name == "c3small":  # CPU-container debugging only
        return setups.make_c3(nx=128, ny=128, nz=75, spacing=(3.9, 3.9, 4.0), n_primaries=histories)
    raise SystemExit(f"unknown config {name}")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic_per_history(workload_key, kernel_name):
    """DRAM bytes per history of the transport kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic-json).  The figure is used only when the
    capture is of the SAME kernel instantiation as the one just launched (entry["kernel"] == kernel_name);
    otherwise (None, why)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None, "no profiles/traffic.json"
    e = d.get(workload_key)
    if not e:
        return None, f"no ncu capture of workload '{workload_key}'"
    if e.get("kernel") != kernel_name:
        return None, f"capture {e.get('source')} is of '{e.get('kernel')}', launched '{kernel_name}'"
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ncu_summary import kernel_source_sha
    if e.get("kernel_source_sha") != kernel_source_sha():
        return None, f"capture {e.get('source')} is of an older build of '{kernel_name}' (kernel sources changed since)"
    return float(e["dram_bytes_per_history"]), f"{e.get('source')} ({e.get('kernel')}, {e.get('histories_in_capture'):.3g} histories)"


def algorithmic_bytes_per_history(st, label_bytes, n_grids, phsp_row_bytes=0):
    """B_hist = (N_tent + N_chord)*s_label + (N_flush + N_tle)*G*2*8 + f_out*s_rec  (SURVEY.md section 8d;
    a TLE actor adds one label gather per voxel chord and one RMW per grid and (voxel, history) sum)."""
    n = max(st["histories"], 1)
    n_tent = st["tentative_collisions"] / n
    n_dep = (st["flushed_entries"] + st.get("tle_entries", 0)) / n
    n_chord = st.get("tle_steps", 0) / n
    f_out = st["phsp_rows"] / n
    return (n_tent + n_chord) * label_bytes + n_dep * n_grids * 2 * 8 + f_out * phsp_row_bytes, n_tent, n_dep, f_out


def desc0(S):
    """descriptor of the first scoring grid (a DoseActor, else the TLEDoseActor)"""
    return S.dose[0] if S.dose else S.tle[0][0]


def n_active_grids(S):
    from opengate_b200 import abi
    f = desc0(S).score_flags | abi.SCORE_EDEP
    return bin(f & 31).count("1")


def cpu_baseline(S, seconds_target=15.0, threads=None):
    """The oracle port on all host cores over a bounded sample of the same workload."""
    from oracle.oracle import TransportOracle
    threads = threads or os.cpu_count() or 1
    orc = TransportOracle(S.seed)
    sids, _, _ = S.apply(orc)
    orc.begin_run()
    probe = 20000 * threads
    t0 = time.perf_counter()
    orc.run_source(sids[0], 0, probe, threads)
    dt = time.perf_counter() - t0
    rate = probe / dt
    n = int(min(max(rate * seconds_target, probe), 5e8))
    t0 = time.perf_counter()
    orc.run_source(sids[0], probe, n, threads)
    dt = time.perf_counter() - t0
    orc.close()
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} histories of the same workload in {dt:.1f} s (CPU oracle, not Geant4)"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores. The reference's
    own implementation is Geant4-MT, which is not in the reference tree and cannot be built
    here, so this arm times the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import TransportOracle
    S = build_setup(args.config, args.histories)
    threads = os.cpu_count() or 1
    orc = TransportOracle(S.seed)
    sids, _, _ = S.apply(orc)
    orc.begin_run()
    probe = 10000 * threads
    t0 = time.perf_counter()
    orc.run_source(sids[0], 0, probe, threads)
    rate = probe / (time.perf_counter() - t0)
    per_step = int(max(rate * args.ref_seconds_per_step, probe))
    first = probe
    for _ in range(args.warmup):
        orc.run_source(sids[0], first, per_step // 4 + 1, threads)
        first += per_step // 4 + 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.run_source(sids[0], first, per_step, threads)
        first += per_step
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": S.name, "histories_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} histories per step x {args.steps} steps (CPU oracle on {threads} threads; Geant4 not available)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3")
    ap.add_argument("--histories", type=int, default=500_000_000, help="histories per step per GPU")
    ap.add_argument("--total-histories", type=float, default=None,
                    help="STRONG scaling: this many histories in total over all GPUs and timed steps (config 5: 1e10); "
                         "histories per step per GPU = total / (gpus * steps)")
    ap.add_argument("--e2e-histories", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-script-level", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds-per-step", type=float, default=8.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)

    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly ONE line (the JSON): anything libraries print while the bench runs
    # (NCCL's version banner on communicator creation) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from opengate_b200 import abi, multigpu
    from opengate_b200.core import Engine

    H = args.histories
    scaling = "weak"
    if args.total_histories:
        H = int(args.total_histories) // (world * args.steps)
        scaling = "strong"
    S = build_setup(args.config, H)
    eng = Engine(local, S.seed)
    stream = torch.cuda.Stream()          # a real (non-default) stream shared by torch and the engine
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    sids, dids, pids = S.apply(eng)
    sid = sids[0]
    G = n_active_grids(S)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.begin_run()
    next_first = [rank * H]

    def step():
        flush_buf.fill_(1)  # evict L2 between steps (the timing rule); ~0.1 ms
        eng.run_source(sid, next_first[0], H)
        next_first[0] += world * H

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.begin_run()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    reduce_check = None
    if world > 1:
        # correctness of the one exchange step, checked on the numbers themselves: the sum over the voxel records
        # on rank 0 after the reduce must equal the sum over ranks of the per-rank sums taken before it
        bufs = [t for a in dids for t in multigpu.buffer_tensors(eng, a)]
        before = torch.stack([t.sum() for t in bufs])
        multigpu.reduce_dose_grids(eng, dids, S, dst=0)
        after = torch.stack([t.sum() for t in bufs])
    ev1.record(stream)
    barrier()
    if world > 1:
        dist.all_reduce(before)
        nhist = torch.tensor([eng.stats()["histories"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(nhist)
        if rank == 0:
            rel = float(((after - before).abs() / before.abs().clamp_min(1e-300)).max())
            ok = rel < 1e-9 and int(nhist.item()) == world * H * args.steps and float(before.min()) > 0
            reduce_check = "ok" if ok else f"FAILED: rel {rel:.3e}, histories {int(nhist.item())} vs {world * H * args.steps}"
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = eng.stats()
    total_hist = world * H * args.steps
    value = total_hist / (ms * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes / kernel time on this rank
    label_bytes = S.labels.dtype.itemsize
    row_bytes = 76 if S.phsp else 0
    b_hist, n_tent, n_dep, f_out = algorithmic_bytes_per_history(st, label_bytes, G, row_bytes)
    kernel_ms = st["kernel_ms"] / max(args.steps, 1)
    achieved = b_hist * H / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    kernel_name = eng.kernel_name()
    t_hist, t_src = measured_traffic_per_history(args.config, kernel_name)

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": S.name, "histories_per_step_per_gpu": H, "seed": S.seed,
                       "total_histories": (world * H * args.steps if scaling == "strong" else None),
                       "label_dtype": str(S.labels.dtype), "n_materials": len(S.tables.names),
                       "l2": "256 MB L2 flush before every step; accumulator grids "
                             f"({G} x {desc0(S).size[0] * desc0(S).size[1] * desc0(S).size[2] * 8 / 1e6:.0f} MB) exceed L2",
                       "parallelism": f"histories sharded over {world} GPU(s), NCCL reduce of {G} grids at end of run" if world > 1 else "1 GPU"},
            "clocks": clocks,
            "gpu_launches": args.steps,
            "reduce_check": reduce_check,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (t_hist * H if t_hist else None), "traffic_source": t_src,
                         "algorithmic_bytes_per_launch": b_hist * H, "peak_source": peak_src,
                         "kernel": kernel_name, "kernel_ms_per_launch": kernel_ms,
                         # the reference's label type is unsigned short (GateImageNestedParameterisation.h:24);
                         # the engine packs <= 256 materials into one byte: the same run counted with 2-byte labels
                         "frac_with_2_byte_labels": (b_hist + n_tent * (2 - label_bytes)) * H / (kernel_ms * 1e-3) / 1e9 / peak,
                         "bytes_per_history": b_hist, "tentative_collisions_per_history": n_tent,
                         "deposit_entries_per_history": n_dep, "phsp_rows_per_history": f_out,
                         "label_bytes": label_bytes, "grids_per_deposit": G,
                         "tle_chords_per_history": st.get("tle_steps", 0) / max(st["histories"], 1)},
            "stats": {k: st[k] for k in ("histories", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "n_conv",
                                         "flushed_entries", "buffer_overflows", "escaped", "edep_total_MeV",
                                         "tle_steps", "tle_entries", "tle_overflows")},
        }
    eng.close()
    # the same workload with the reference's DEFAULT hit_type ("random", doseactors.py:84-91; the headline uses
    # "post"): a second, shorter measurement with its own kernel variant, reported under config.variants
    if args.config == "c3" and world == 1 and rank == 0:
        Sv = build_setup("c3_hit_random", H)
        ev = Engine(local, Sv.seed)
        ev.set_stream(stream.cuda_stream)
        sv, _, _ = Sv.apply(ev)
        ev.begin_run()
        for k in range(2):
            flush_buf.fill_(1)
            ev.run_source(sv[0], k * H, H)
        ev.sync()
        ev.begin_run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for k in range(3):
            flush_buf.fill_(1)
            ev.run_source(sv[0], (2 + k) * H, H)
        b.record(stream)
        torch.cuda.synchronize()
        line["config"]["variants"] = {"c3_hit_random": {"value": 3 * H / (a.elapsed_time(b) * 1e-3), "unit": UNIT, "steps": 3, "warmup": 2,
                                                          "kernel": ev.kernel_name(), "what": "hit_type='random', the reference's default"}}
        ev.close()
    del flush_buf
    torch.cuda.empty_cache()

    # e2e: the same metric through the C-ABI with HOST buffers: every step uploads tables +
    # label volume, runs, and reads the accumulator grids back to host memory.
    He = args.e2e_histories or H
    e2e = measure_e2e(S, He, local, rank, world, torch, dist, steps=max(2, min(args.steps, 3)))
    if rank == 0:
        line["e2e"] = e2e
        if world == 1 and not args.no_script_level:
            line["e2e_script"] = measure_script_level(args.config, 2 * H, local)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(S, args.cpu_seconds)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_script_level(config, n, local):
    """`e2e_script`: wall clock of Simulation.run() for the bench workload written against the repo's PUBLIC
    script-level API (Simulation / add_volume("Image") with the int16 HU image / HounsfieldUnit_to_material /
    GenericSource / DoseActor dose + edep uncertainty) -- what a GATE user calls.  Everything is inside the timed
    call: HU -> labels, physics tables, H2D, the histories, Gy conversion, D2H of every image the actor keeps.
    Config 3 only (the other configs have no script twin here)."""
    if config != "c3":
        return None
    # two runs, each with a new Simulation: the wall clock of one sim.run() depends on how fast the box page-locks the
    # 1.9 GB of output arrays while the histories run (0.56 ... 1.4 s seen for the same code on different boxes); both
    # are listed, `value` is the faster one (the second run finds the device memory pool of the process warm)
    runs = [_script_level_run(n) for _ in range(2)]
    best = min(runs, key=lambda r: r["wall_s"])
    best["runs_wall_s"] = [r["wall_s"] for r in runs]
    return best


def _script_level_run(n):
    import opengate_b200 as gate
    from opengate_b200 import image as gimage, setups
    mm, m, MeV, deg = gate.g4_units.mm, gate.g4_units.m, gate.g4_units.MeV, gate.g4_units.deg
    hu = setups.synthetic_ct_hu()
    sim = gate.Simulation()
    sim.random_seed = 987654321
    sim.output_dir = "/tmp/gate_b200_bench_script"
    sim.world.size = [2 * m, 2 * m, 2 * m]
    ct = sim.add_volume("Image", "ct")
    ct.image = gimage.Image(hu, (0.9766, 0.9766, 1.0))
    ct.material = "G4_AIR"
    ct.voxel_materials, _ = gate.geometry.materials.HounsfieldUnit_to_material(sim, 0.05 * gate.g4_units.g_cm3)
    sim.physics_manager.set_production_cut("world", "all", 1 * m)
    sim.physics_manager.set_production_cut("ct", "all", 1 * m)
    src = sim.add_source("GenericSource", "xray")
    edges, w = setups.spectrum_120kvp()
    src.energy.type = "spectrum_histogram"
    src.energy.spectrum_energy_bin_edges = list(edges * MeV)
    src.energy.spectrum_weights = list(w)
    src.position.type = "point"
    src.position.translation = [-600 * mm, 0, 0]
    src.direction.type = "iso"
    src.direction.theta = [0, float(np.degrees(np.arctan(0.95 * 250.0 / 600.0))) * deg]
    src.position.rotation = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    src.direction_relative_to_attached_volume = True
    src.number_of_primaries = n
    dose = sim.add_actor("DoseActor", "dose")
    dose.attached_to = "ct"
    dose.size = "like_image_volume"
    dose.spacing = "like_image_volume"
    dose.hit_type = "post"
    dose.dose.active = True
    dose.edep_uncertainty.active = True
    dose.write_to_disk = False
    stats = sim.add_actor("SimulationStatisticsActor", "stats")
    sim.number_of_gpus = 1
    t0 = time.perf_counter()
    sim.run()
    wall = time.perf_counter() - t0
    e = dose.edep.image.array
    assert stats.counts["events"] == n and float(e.sum()) > 0
    return {"value": n / wall, "unit": UNIT, "wall_s": wall, "histories": n,
            "statistics_actor_duration_s": stats.counts.get("duration", 0) / gate.g4_units.s,
            "images_returned": ["edep", "edep_squared", "dose"],
            "what": "wall clock of Simulation.run() of the public script-level API, 1 GPU, CUDA context warm; best of two runs, both listed"}


def measure_e2e(S, H, local, rank, world, torch, dist, steps=2):
    """The same metric through the C-ABI from HOST buffers: every step creates an engine, uploads
    the tables and the label volume from pinned host memory (H2D), runs H histories per GPU,
    reduces the grids onto rank 0 (N > 1) and reads every active grid back into pinned host
    memory (D2H).  Wall-clock around the whole sequence, max over ranks."""
    import copy
    from opengate_b200 import abi, multigpu
    from opengate_b200.core import Engine
    G = n_active_grids(S)
    nx, ny, nz = desc0(S).size[0], desc0(S).size[1], desc0(S).size[2]
    which = multigpu.active_grids(desc0(S).score_flags)
    host_out = [torch.empty((nz, ny, nx), dtype=torch.float64, pin_memory=True).numpy() for _ in range(G)] if (rank == 0 and world == 1) else []
    # N > 1: the result lands in host arrays shared by the ranks of the node; after the all-reduce every rank copies
    # 1/N of every grid, so the device -> host read uses the host links of all GPUs (multigpu.SharedHostGrids)
    shared = multigpu.SharedHostGrids([(nz, ny, nx)] * G) if world > 1 else None
    S = copy.copy(S)
    pinned_labels = torch.empty(S.labels.shape, dtype=torch.uint8 if S.labels.dtype == np.uint8 else torch.int16, pin_memory=True)
    lab = pinned_labels.numpy().view(S.labels.dtype)
    lab[...] = S.labels
    S.labels = lab
    t = S.tables
    h2d = S.labels.nbytes + sum(a.nbytes for a in (t.density, t.mu_total, t.cum_phot, t.cum_compt, t.cum_rayl,
                                                   t.rayl_cum_ff2, t.compt_sf)) + sum(k.nbytes for _, _, k in S.tle)
    d2h = G * nx * ny * nz * 8

    parts = {"create": 0.0, "setup_h2d": 0.0, "run": 0.0, "reduce": 0.0, "d2h": 0.0, "destroy": 0.0}

    def one(first, acc=None):
        # the section clocks only explain the total: the run is launched asynchronously, so the
        # stream is drained at the end of "run" (the D2H copy would wait for it anyway)
        t = [time.perf_counter()]

        def lap(key):
            t.append(time.perf_counter())
            if acc is not None:
                acc[key] += t[-1] - t[-2]

        eng = Engine(local, S.seed)
        lap("create")
        sids, dids, pids = S.apply(eng)       # H2D: tables + label volume (+ source CDFs)
        lap("setup_h2d")
        eng.begin_run()
        eng.run_source(sids[0], first, H)
        eng.sync()
        lap("run")
        if world > 1:
            multigpu.all_reduce_dose_grids(eng, dids)
            torch.cuda.synchronize()
        lap("reduce")
        if world > 1:
            shared.gather(eng, dids[0], which)     # D2H: every rank 1/N of every grid into the shared arrays + barrier
        else:
            for w, out in zip(which, host_out):   # D2H: every active accumulator grid
                eng.dose_grid(dids[0], w, out)
        lap("d2h")
        eng.close()
        lap("destroy")

    one(rank * H)  # warm-up (page-in, allocator)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        one((1 + i) * world * H + rank * H, parts)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    if shared is not None:
        if rank == 0:
            assert float(shared.arrays[0].sum()) > 0.0
        shared.close()
    return {"value": world * H * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "histories_per_step_per_gpu": H, "steps": steps,
            "ms_per_step_by_section_rank0": {k: 1e3 * v / steps for k, v in parts.items()},
            "what": "gate_b200 C-ABI from pinned host buffers: create + set_tables + set_image_volume (H2D) + run"
                    + (" + NCCL all-reduce + every rank copies 1/N of every grid into host arrays shared by the ranks" if world > 1
                       else " + get_dose_grid of every grid (D2H)") + " per step"}


if __name__ == "__main__":
    main()
