"""`GpuSimulationEngine`: the run phase of a `Simulation` on the CUDA engine.

Takes the place of the reference's `SimulationEngine` (opengate/engines.py:1414-1891: initialize
:1721-1863, run_engine :1554-1630, start_and_stop :1632-1667) for the photon-in-image-volume path:
it reads the resolved Python objects (volumes, physics, sources, actors), converts them ONCE into
the C-ABI descriptors (`LowLevelSetup`), drives `libgate_b200.so` run by run the way
`GateSourceManager::StartMasterThread` does (GateSourceManager.cpp:198-244: BeginOfRunAction ->
beamOn -> EndOfRunAction per timing interval) and hands every actor its results in the reference's
end-of-run life cycle.  Unsupported configuration -> `fatal()`.

Multi-GPU: one process per GPU under torchrun (histories split `n//G + (g < n%G)` like the
reference's thread split, sources/base.py:190-197; one sum all-reduce of the accumulator grids at
the end of each run), or `sim.number_of_gpus = N` inside one process (N engines, host-side sum).
"""

from __future__ import annotations

import time

import numpy as np

from . import abi, multigpu, physics
from . import materials as gmat
from .exception import fatal, warning
from .setups import LowLevelSetup, dose_world_to_index
from .units import g4_units
from .volumes import __world_name__


class SimulationOutput:
    """What `Simulation.run()` returns (reference engines.py:1320-1411)."""

    def __init__(self):
        self.actors = {}
        self.sources = {}
        self.current_random_seed = None
        self.expected_number_of_events = 0
        self.warnings = []
        self.log_output = ""
        self.user_hook_log = []
        self.device_stats = []     # gate_b200_stats per run (summed over GPUs)

    def get_actor(self, name):
        if name not in self.actors:
            fatal(f'The actor "{name}" does not exist. Here is the list of actors: {list(self.actors)}')
        return self.actors[name]

    def get_source(self, name):
        if name not in self.sources:
            fatal(f'The source "{name}" does not exist. Here is the list of sources: {list(self.sources)}')
        return self.sources[name]


def invert_rigid(R, t):
    Ri = np.asarray(R, dtype=np.float64).T
    return Ri, -Ri @ np.asarray(t, dtype=np.float64)


def _half_size(v):
    h = v.half_size_mm() if callable(v.half_size_mm) else v.half_size_mm
    return np.asarray(h, dtype=np.float64)


def _boxes_overlap(a, b):
    """separating-axis test of two placed boxes (15 axes): True when they share volume (boxes that merely touch,
    e.g. a couch under the image volume, do not: 1e-6 mm of slack)"""
    Ra, ta = a.world_transform()
    Rb, tb = b.world_transform()
    ha, hb = _half_size(a), _half_size(b)
    R = Ra.T @ Rb
    t = Ra.T @ (tb - ta)
    A = np.abs(R) + 1e-12
    for i in range(3):
        if abs(t[i]) > ha[i] + hb @ A[i] - 1e-6:
            return False
    for j in range(3):
        if abs(t @ R[:, j]) > ha @ A[:, j] + hb[j] - 1e-6:
            return False
    for i in range(3):
        for j in range(3):
            ra = ha[(i + 1) % 3] * A[(i + 2) % 3, j] + ha[(i + 2) % 3] * A[(i + 1) % 3, j]
            rb = hb[(j + 1) % 3] * A[i, (j + 2) % 3] + hb[(j + 2) % 3] * A[i, (j + 1) % 3]
            if abs(t[(i + 2) % 3] * R[(i + 1) % 3, j] - t[(i + 1) % 3] * R[(i + 2) % 3, j]) > ra + rb + 1e-9:
                return False
    return True


def affine34(R, t):
    A = np.zeros((3, 4))
    A[:, :3] = R
    A[:, 3] = t
    return A


class GpuSimulationEngine:
    def __init__(self, simulation, backend=None):
        self.sim = simulation
        self._backend = backend
        self.setup = None
        self.dose_actors = []
        self.phsp_actors = []
        self.stats_actors = []
        self.source_objects = []
        self.engines = []
        self.ids = []
        self.rank, self.world, self.local_rank = 0, 1, 0
        self.dist = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        for e in self.engines:
            try:
                e.close()
            except Exception:
                pass
        self.engines = []

    # ------------------------------------------------------------------------ configuration
    def _phantom_volume(self):
        """The ONE volume photons interact in besides the world: an ImageVolume, or (config 1)
        a homogeneous Box carrying a DoseActor, represented as a 1-voxel label volume."""
        sim = self.sim
        vols = [v for v in sim.volume_manager.volumes.values() if v.name != __world_name__]
        images = [v for v in vols if v.volume_type == "ImageVolume"]
        dose_targets = {a.attached_to for a in sim.actor_manager.actors.values() if a.type_name in ("DoseActor", "TLEDoseActor")}
        phsp_targets = {a.attached_to for a in sim.actor_manager.actors.values() if a.type_name == "PhaseSpaceActor"}
        world_mat = sim.world.material
        phantom = None
        if len(images) > 1:
            fatal("The B200 GPU engine supports one ImageVolume per simulation")
        if images:
            phantom = images[0]
        # Boxes: the one a dose actor is attached to (or, without an Image volume, the first one of another material)
        # is the phantom; every other Box of a material of its own is a homogeneous region of the world (a couch, a
        # detector slab -- photons are attenuated and scatter there, nothing is scored); Boxes of the world
        # material only carry phase-space actors
        self.region_volumes = []
        for v in vols:
            if v is phantom:
                continue
            if v.volume_type != "BoxVolume":
                fatal(f"Volume '{v.name}' of type {v.volume_type} is not supported by the B200 GPU engine")
            if v.name in dose_targets:
                if phantom is not None:
                    fatal(f"DoseActor attached to '{v.name}' besides the phantom '{phantom.name}': on the B200 GPU engine "
                          f"dose actors are attached to ONE phantom volume")
                phantom = v
        for v in vols:
            if v is phantom or v.material == world_mat:
                continue
            if phantom is None and v.name not in phsp_targets:
                phantom = v
                continue
            self.region_volumes.append(v)
        if phantom is None:
            fatal("No phantom volume: add an Image volume (or a Box with a DoseActor attached)")
        if len(self.region_volumes) > 4:
            fatal(f"The B200 GPU engine supports at most 4 Box volumes of their own material besides the phantom "
                  f"({[v.name for v in self.region_volumes]})")
        for v in self.region_volumes:
            m = v.mother_volume
            while m is not None and m.name != __world_name__:
                if m.material != world_mat:
                    fatal(f"Box '{v.name}' must be placed in the world (or in mothers made of the world material) "
                          f"on the B200 GPU engine: nested material regions are not supported")
                m = m.mother_volume
            others = [phantom] + [w for w in self.region_volumes if w is not v]
            for w in others:
                if _boxes_overlap(v, w):
                    fatal(f"Box '{v.name}' overlaps volume '{w.name}': material regions must be disjoint "
                          f"on the B200 GPU engine")
        for t in dose_targets:
            if t != phantom.name:
                fatal(f"DoseActor attached to '{t}': on the B200 GPU engine dose actors must be attached to the "
                      f"phantom volume '{phantom.name}'")
        if phantom.mother != __world_name__:
            # nested placement is fine geometrically (world_transform composes it) as long as the
            # mothers are made of the world material
            m = phantom.mother_volume
            while m is not None and m.name != __world_name__:
                if m.material != world_mat:
                    fatal(f"Mother volume '{m.name}' of the phantom is not made of the world material")
                m = m.mother_volume
        return phantom

    def build_setup(self):
        """Resolved Simulation -> LowLevelSetup (pure host code; no device needed)."""
        sim = self.sim
        pm = sim.physics_manager
        db = sim.volume_manager.material_database
        phantom = self._phantom_volume()
        world = sim.world

        # label volume ------------------------------------------------------------------------
        hu = hu_bins = hu_bin_labels = None
        if phantom.volume_type == "ImageVolume" and not phantom.dump_label_image:
            # the HU image and the np.digitize table go to the engine, which makes the label volume on the
            # device (ImageVolume.create_label_image, volumes.py:1200-1243); no host pass over the voxels
            phantom.material_to_label_lut = phantom.create_material_to_label_lut()
            phantom.load_input_image()
            lut = dict(phantom.material_to_label_lut)
            hu = phantom.itk_image.array
            hu_bins, hu_bin_labels = gmat.label_bins(phantom.voxel_materials, lut)
            labels16 = None
            volume_shape = hu.shape
            spacing = tuple(float(x) for x in phantom.itk_image.spacing)
        elif phantom.volume_type == "ImageVolume":
            label_img = phantom.construct()
            lut = dict(phantom.material_to_label_lut)
            labels16 = label_img.array
            volume_shape = labels16.shape
            spacing = tuple(float(x) for x in label_img.spacing)
        else:
            lut = {phantom.material: 0}
            labels16 = np.zeros((1, 1, 1), dtype=np.uint16)
            volume_shape = labels16.shape
            spacing = tuple(float(x) for x in phantom.size)
        names = list(lut.keys())
        if world.material not in lut:
            names.append(world.material)
        for v in self.region_volumes:
            if v.material not in names:
                names.append(v.material)
        world_label = names.index(world.material)
        label_bytes = 1 if len(names) <= 256 else 2
        labels = None if labels16 is None else labels16.astype(np.uint8 if label_bytes == 1 else np.uint16)
        mats = [db.find_or_build_material(n) for n in names]

        # sources -------------------------------------------------------------------------------
        self.source_objects = list(sim.source_manager.sources.values())
        e_max = 0.0
        src_structs = []
        for s in self.source_objects:
            vol = sim.volume_manager.get_volume(s.attached_to)
            src_structs.append(s.to_abi(None if vol.name == __world_name__ else vol))
            e_max = max(e_max, s.max_energy())
        if pm.energy_range_max is not None:
            e_max = max(e_max, 0.0) if e_max > 0 else float(pm.energy_range_max)
        if not e_max > 0:
            fatal("Cannot determine the maximum photon energy of the sources")
        tables = physics.build_photon_tables(
            mats, e_max_mev=e_max, e_cut_mev=float(pm.photon_energy_cut), n_energy=int(pm.table_energy_nodes),
            pair_mode=1 if pm.pair_production_mode == "annihilation" else 0)
        half_phantom = 0.5 * np.asarray(volume_shape[::-1]) * np.asarray(spacing)
        pm.check_kerma_configuration(float(2 * np.max(half_phantom)))

        # placement -----------------------------------------------------------------------------
        Rv, tv = phantom.world_transform()
        Ri, ti = invert_rigid(Rv, tv)
        w2v = affine34(Ri, ti)
        world_half = tuple(0.5 * float(x) for x in world.size)
        if sim.check_volumes_overlap:
            corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * half_phantom
            cw = corners @ Rv.T + tv
            if np.any(np.abs(cw) > np.asarray(world_half) + 1e-9):
                fatal(f"Volume '{phantom.name}' is not fully inside the world")

        # actors --------------------------------------------------------------------------------
        self.dose_actors, self.phsp_actors, self.stats_actors = [], [], []
        tle_actors, tle_desc = [], []
        dose_desc, phsp_desc = [], []
        n_total = self.expected_number_of_events()
        for a in sim.actor_manager.sorted_actors:
            a.initialize()
            if a.type_name == "DoseActor":
                A = dose_world_to_index(a.size, a.spacing, translation=a.translation, vol_rotation=Rv, vol_translation=tv)
                dose_desc.append(abi.make_dose_desc([int(x) for x in a.size], A, a.score_flags(), abi.HIT_TYPES[a.hit_type]))
                self.dose_actors.append(a)
            elif a.type_name == "TLEDoseActor":
                A = dose_world_to_index(a.size, a.spacing, translation=a.translation, vol_rotation=Rv, vol_translation=tv)
                tle_desc.append(([int(x) for x in a.size], A, a.score_flags(), abi.HIT_TYPES[a.hit_type],
                                 a.tle_energy_threshold(), float(a.energy_min), a.database))
                tle_actors.append(a)
            elif a.type_name == "PhaseSpaceActor":
                box = a.attached_to_volume
                Rb, tb = box.world_transform()
                Rbi, tbi = invert_rigid(Rb, tb)
                cap = a.max_rows if a.max_rows is not None else max(int(n_total) + 1024, 1024)
                phsp_desc.append(abi.make_phsp_desc(box.half_size_mm, affine34(Rbi, tbi), a.steps_mask(), int(cap)))
                self.phsp_actors.append(a)
            elif a.type_name == "SimulationStatisticsActor":
                self.stats_actors.append(a)
            else:
                fatal(f"Actor type {a.type_name} is not supported by the B200 GPU engine")
        self.setup = LowLevelSetup(
            name=f"{sim.name}", seed=int(sim.current_random_seed) & (2 ** 64 - 1), tables=tables,
            world_half=world_half, world_material=world_label,
            labels=None if labels is None else np.ascontiguousarray(labels), spacing=spacing,
            world_to_volume=w2v, sources=src_structs, dose=dose_desc, phsp=phsp_desc, n_primaries=int(n_total),
            materials=mats, hu=hu, hu_bins=hu_bins, hu_bin_labels=hu_bin_labels, label_bytes=label_bytes)
        for v in self.region_volumes:
            Rb, tb = v.world_transform()
            Rbi, tbi = invert_rigid(Rb, tb)
            self.setup.regions.append(abi.make_region_desc(_half_size(v), affine34(Rbi, tbi), names.index(v.material)))
        # TLE dose actors come after the plain ones in the engine's actor id space
        for size, A, flags, hit, thr, emin, database in tle_desc:
            self.setup.add_tle_actor(size, A, flags, hit_type=hit, threshold=thr, energy_min=emin, database=database)
        self.dose_actors += tle_actors
        return self.setup

    def expected_number_of_events(self):
        sim = self.sim
        rng = np.random.default_rng(int(sim.current_random_seed) & 0xFFFFFFFF)
        self._counts = []   # [run][source]
        for r in range(len(sim.run_timing_intervals)):
            self._counts.append([s.primaries_for_run(r, rng) for s in sim.source_manager.sources.values()])
        return int(np.sum(self._counts)) if self._counts else 0

    # ------------------------------------------------------------------------ devices
    def _init_devices(self, seed):
        sim = self.sim
        if self._backend is not None:
            self.engines = [self._backend(0, seed)]
            return
        from .core import Engine
        if self.dist is not None:
            if sim.number_of_gpus not in (None, self.world):
                warning(f"number_of_gpus={sim.number_of_gpus} ignored: running under torchrun with {self.world} ranks")
            self.engines = [Engine(self.local_rank, seed)]
        else:
            n = int(sim.number_of_gpus or 1)
            self.engines = [Engine(g, seed) for g in range(n)]

    @property
    def n_shards(self):
        return self.world if self.dist is not None else len(self.engines)

    # ------------------------------------------------------------------------ run
    def _detect_process_group(self):
        """torch.distributed only matters when the script runs under torchrun and has initialised a process
        group: do not pay the import of torch (4 s) in a plain single-process run."""
        import sys
        if self._backend is not None and not getattr(self._backend, "distributed", False):
            return
        if "torch" in sys.modules:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    self.dist = dist
                    self.rank, self.world = dist.get_rank(), dist.get_world_size()
                    _, _, self.local_rank = multigpu.world_info()
            except ImportError:
                pass

    def _share_random_seed(self):
        """All ranks of a run must use ONE seed: the Philox key of every history (results independent of the
        number of GPUs) and the Poisson draw of activity sources in expected_number_of_events() (all ranks
        must agree on the count they split).  random_seed='auto' draws it per process, so rank 0's is
        broadcast; an explicit integer seed is the same everywhere already and goes through unchanged."""
        if self.dist is None:
            return
        box = [int(self.sim.current_random_seed)]
        self.dist.broadcast_object_list(box, src=0)
        self.sim.current_random_seed = int(box[0])

    def run_engine(self):
        sim = self.sim
        t_init = time.perf_counter()
        self._detect_process_group()
        self._share_random_seed()
        # the CUDA context of a process takes about a second to come up: the devices are opened on a worker thread
        # (ctypes releases the GIL) while this thread turns the simulation into tables and descriptors
        import threading
        seed = int(sim.current_random_seed) & (2 ** 64 - 1)
        opened = {}

        def open_devices():
            try:
                self._init_devices(seed)
            except BaseException as e:   # re-raised on the main thread
                opened["error"] = e

        th = threading.Thread(target=open_devices, daemon=True)
        th.start()
        try:
            self.build_setup()
        finally:
            th.join()
        if "error" in opened:
            raise opened["error"]
        self.ids = [self.setup.apply(e) for e in self.engines]
        out = SimulationOutput()
        out.current_random_seed = sim.current_random_seed
        out.expected_number_of_events = self.setup.n_primaries
        t_start = time.perf_counter()
        base = 0
        totals = {"events": 0, "tracks": 0, "steps": 0}
        for run_index in range(len(sim.run_timing_intervals)):
            st = self._run_one(run_index, base)
            out.device_stats.append(st)
            base += int(np.sum(self._counts[run_index]))
            totals["events"] += st["histories"]
            totals["tracks"] += st["tracks"]
            totals["steps"] += st["steps"]
        t_stop = time.perf_counter()
        for a in self.stats_actors:
            a.store(len(sim.run_timing_intervals), totals["events"], totals["tracks"], totals["steps"],
                    (t_stop - t_start) * g4_units.s, (t_start - t_init) * g4_units.s,
                    extra={"nb_threads": sim.number_of_threads, "nb_gpus": self.n_shards,
                           "tentative_collisions": int(sum(s["tentative_collisions"] for s in out.device_stats))})
        if self.rank == 0:
            for a in sim.actor_manager.sorted_actors:
                a.end_of_simulation()
        out.actors = dict(sim.actor_manager.actors)
        out.sources = dict(sim.source_manager.sources)
        out.warnings = list(sim.warnings)
        return out

    def _launch(self, source_index, first, n):
        """Enqueue n histories [first, first+n) of one source, split over the shards."""
        chunk = int(self.sim.max_histories_per_launch)
        if self.dist is not None:
            off, cnt = multigpu.split_primaries(n, self.world, self.rank)
            parts = [(self.engines[0], self.ids[0], first + off, cnt)]
        else:
            parts = []
            for g, e in enumerate(self.engines):
                off, cnt = multigpu.split_primaries(n, len(self.engines), g)
                parts.append((e, self.ids[g], first + off, cnt))
        for e, ids, f, cnt in parts:
            done = 0
            while done < cnt:
                k = min(chunk, cnt - done)
                e.run_source(ids[0][source_index], f + done, k)
                done += k

    def _sync(self):
        for e in self.engines:
            if hasattr(e, "sync"):
                e.sync()

    # ---- host memory of the outputs, made ready while the device is busy ---------------------------------
    def _prepare_outputs(self, single_run):
        """The histories are in flight and the host has nothing to do: allocate the (z,y,x) float64 arrays the dose
        actors will keep and page-lock them (worker threads: first touch + cudaHostRegister of 0.6 GB per image at
        config-3 size), so that the copies after the run go at the full host-link rate into their final place."""
        self._prepared = {}
        e = self.engines[0] if self.engines else None
        if self.dist is not None or len(self.engines) != 1 or not hasattr(e, "pin"):
            return
        from concurrent.futures import ThreadPoolExecutor
        jobs = [(i, key) for i, a in enumerate(self.dose_actors) for key, _, _, _ in a.fetch_plan(single_run)]
        if not jobs:
            return
        # ONE worker: gate_b200_host_register touches the pages of an array from several threads itself (0.09 s for a
        # 0.6 GB image on the GPU box); three arrays registered side by side contend for the same cores and locks and
        # took 0.35 ... 0.65 s instead of 3 x 0.09 s one after the other (tools/probe_pin.py)
        pool = ThreadPoolExecutor(max_workers=1)

        def make(i):
            arr = np.empty(e.dose_grid_shape(self.ids[0][1][i]), dtype=np.float64)
            return arr, (arr.nbytes >= (1 << 20) and e.pin(arr))

        self._prepared = {job: pool.submit(make, job[0]) for job in jobs}
        pool.shutdown(wait=False)

    def _take_prepared(self, actor_index, key, engine):
        fut = getattr(self, "_prepared", {}).get((actor_index, key))
        if fut is None:
            return None
        arr, pinned = fut.result()
        if pinned:
            self.__dict__.setdefault("_pinned", []).append(arr)
        return arr

    def _release_prepared(self, actor_index, engine):
        for arr in self.__dict__.pop("_pinned", []):
            engine.unpin(arr)

    def _event_time_function(self, run_index, base):
        """global history id -> event time of run `run_index` (timing.event_times): the ids of a run are dealt to
        the sources in order, counts[run][source] each, starting at `base`"""
        from . import timing
        sim = self.sim
        interval = sim.run_timing_intervals[run_index]
        seed = int(sim.current_random_seed) & (2 ** 64 - 1)
        segs, first = [], base
        for s, n in zip(self.source_objects, self._counts[run_index]):
            segs.append((first, first + int(n), s))
            first += int(n)

        def f(ids):
            ids = np.asarray(ids, dtype=np.int64)
            t = np.full(ids.shape, float(interval[0]), dtype=np.float64)
            for lo, hi, s in segs:
                m = (ids >= lo) & (ids < hi)
                if m.any():
                    ui = s.user_info
                    t[m] = timing.event_times(seed, ids[m], interval, start_time=ui.start_time, end_time=ui.end_time,
                                              activity=ui.activity > 0, half_life=ui.half_life)
            return t

        return f

    def _mean_uncertainty(self, actor_index, actor, n_events):
        """GateDoseActor::ComputeMeanUncertainty over the data of ALL shards."""
        if self.n_shards == 1 and hasattr(self.engines[0], "mean_uncertainty"):
            return self.engines[0].mean_uncertainty(self.ids[0][1][actor_index], n_events,
                                                    int(actor.uncertainty_top_voxels_count),
                                                    float(actor.uncertainty_voxel_edep_threshold))
        grids = self._gather_grids(actor_index, (abi.GRID_EDEP, abi.GRID_EDEP_SQ), reduce_in_place=False)
        return host_mean_uncertainty(grids[abi.GRID_EDEP], grids[abi.GRID_EDEP_SQ], n_events,
                                     int(actor.uncertainty_top_voxels_count), float(actor.uncertainty_voxel_edep_threshold))

    def _gather_grids(self, actor_index, which_list, reduce_in_place=True):
        """Sum of the accumulator grids of one dose actor over all shards -> {which: (z,y,x) float64}."""
        out = {}
        if self.dist is not None and not hasattr(self.engines[0], "dose_buffers"):
            # an injected back end without device buffers (tests: the CPU oracle under a gloo group): host reduce
            import torch
            e, ids = self.engines[0], self.ids[0]
            for w in which_list:
                t = torch.from_numpy(np.array(e.dose_grid(ids[1][actor_index], w), dtype=np.float64, copy=True))
                self.dist.all_reduce(t)
                out[w] = t.numpy()
            return out
        if self.dist is not None:
            e, ids = self.engines[0], self.ids[0]
            aid = ids[1][actor_index]
            bufs = multigpu.buffer_tensors(e, aid)     # [records {edep, edep2, dose, dose2} x nvox, (counts)]
            if reduce_in_place:
                for t in bufs:
                    self.dist.all_reduce(t)
                # a (non-async) NCCL all_reduce only orders torch's current stream after the collective; the
                # engine extracts and copies on its OWN non-blocking stream: wait for the reduction on the host
                # before handing the buffers back to it (the engine's earlier work was drained by _sync())
                multigpu.device_synchronize(e.device)
                for w in which_list:
                    out[w] = e.dose_grid(aid, w)
            else:
                import torch
                t = bufs[0].clone()
                self.dist.all_reduce(t)
                nx, ny, nz = [int(x) for x in self.dose_actors[actor_index].size]
                stride, offset = multigpu.record_layout(self.dose_actors[actor_index].score_flags())
                rec = t.view(nz, ny, nx, stride)
                for w in which_list:
                    if w == abi.GRID_COUNTS:
                        c = bufs[1].clone()
                        self.dist.all_reduce(c)
                        out[w] = c.cpu().numpy().reshape(nz, ny, nx)
                    else:
                        out[w] = rec[..., offset[w]].contiguous().cpu().numpy()
                del t, rec
                torch.cuda.empty_cache()
            return out
        for w in which_list:
            acc = None
            for e, ids in zip(self.engines, self.ids):
                g = e.dose_grid(ids[1][actor_index], w)      # a fresh host array per call
                acc = g if acc is None else acc + g
            out[w] = acc
        return out

    def _run_one(self, run_index, base):
        sim = self.sim
        for e in self.engines:
            e.begin_run(run_index)
        counts = self._counts[run_index]
        goal_actors = [(i, a) for i, a in enumerate(self.dose_actors) if a.uncertainty_goal is not None]
        first = base
        n_run = 0
        for si, n in enumerate(counts):
            if n <= 0:
                continue
            if not goal_actors:
                self._launch(si, first, n)
                n_run += n
            else:
                # stop rule of GateDoseActor::EndOfEventAction (GateDoseActor.cpp:245-278): check after
                # N events, then re-estimate N = (u/goal)^2 * n * overshoot
                i, a = goal_actors[0]
                done = 0
                next_check = int(a.uncertainty_first_check_after_n_events)
                while done < n:
                    k = min(max(next_check - (n_run + done), 1), n - done)
                    self._launch(si, first + done, k)
                    done += k
                    self._sync()
                    u = self._mean_uncertainty(i, a, n_run + done)
                    if u <= a.uncertainty_goal:
                        break
                    next_check = int((u / a.uncertainty_goal) ** 2 * (n_run + done) * a.uncertainty_overshoot_factor_N_events)
                n_run += done
            first += n
        self._prepare_outputs(len(sim.run_timing_intervals) == 1)
        self._sync()
        # statistics: summed over shards
        st = None
        for e in self.engines:
            s = e.stats()
            st = dict(s) if st is None else {k: (st[k] + s[k] if k != "kernel_ms" else max(st[k], s[k])) for k in st}
        if self.dist is not None:
            st = multigpu.reduce_counters(st, "cpu" if self._backend is not None else f"cuda:{self.local_rank}")
        n_events = int(st["histories"])
        # dose actors
        single_run = len(sim.run_timing_intervals) == 1
        fast = self.dist is None and len(self.engines) == 1 and n_events > 1
        for i, a in enumerate(self.dose_actors):
            if fast:
                # one device: the unit conversion and the uncertainty are done there, every image comes back once
                # into the array the actor keeps (prepared while the histories were running, see _prepare_outputs)
                e, aid = self.engines[0], self.ids[0][1][i]
                grids = {}
                for key, kind, which, scale in a.fetch_plan(single_run):
                    out = self._take_prepared(i, key, e)
                    if kind == "grid":
                        grids[key] = e.dose_grid(aid, which, out=out, scale=scale)
                    else:
                        grids[key] = e.uncertainty_grid(aid, which, n_events, out=out)
                self._release_prepared(i, e)
                a.store_run(run_index, grids, n_events, scaled=True)
            else:
                which = multigpu.active_grids(a.score_flags())
                grids = self._gather_grids(i, which)
                a.store_run(run_index, grids, n_events)
        # phase space: rows of all shards concatenated (EventID is the global history id of the run)
        for i, a in enumerate(self.phsp_actors):
            rows = None
            for e, ids in zip(self.engines, self.ids):
                r = e.phsp(ids[2][i])
                rows = r if rows is None else {k: np.concatenate([rows[k], r[k]]) for k in rows}
            if self.dist is not None:
                gathered = [None] * self.world
                self.dist.all_gather_object(gathered, rows)
                rows = {k: np.concatenate([g[k] for g in gathered]) for k in rows}
            order = np.lexsort((rows["Time"], rows["EventID"]))
            rows = {k: v[order] for k, v in rows.items()}
            a.store_rows(rows, st["absorbed_events"] if a.store_absorbed_event else 0, event_offset=-base,
                         event_time=self._event_time_function(run_index, base))
            if st["phsp_dropped"]:
                warning(f"PhaseSpaceActor '{a.name}': {st['phsp_dropped']} rows dropped (max_rows too small)")
        return st


def host_mean_uncertainty(edep, edep_sq, n_events, top_voxels, threshold):
    """GateDoseActor::ComputeMeanUncertainty (GateDoseActor.cpp:280-330) on host arrays (multi-shard case)."""
    n = max(float(n_events), 2.0)
    flat = np.asarray(edep).ravel()
    k = min(int(top_voxels), flat.size)
    top = np.partition(flat, flat.size - k)[flat.size - k:] if k > 0 else np.zeros(0)
    thr = (top.mean() if k > 0 else 0.0) * threshold
    sel = flat > thr
    if not sel.any():
        return 1.0
    val = flat[sel] / n
    sq = np.asarray(edep_sq).ravel()[sel] / n
    u = np.sqrt(np.clip((sq - val * val) / (n - 1.0), 0, None)) / val
    u = np.minimum(u, 1.0)
    s = u.sum()
    return float(s / sel.sum()) if s > 0 else 1.0
