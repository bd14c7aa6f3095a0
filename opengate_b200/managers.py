"""`Simulation` and its managers: the user-facing objects of a GATE 10 script, restricted to what the
photon-in-CT dose path needs (SURVEY.md section 3.6 check-list).

Reference: opengate/managers.py (`Simulation` :1805, defaults :1852-2119, `run` :2818-2950,
`_run_simulation_engine` :2803-2816, `PhysicsManager` :695-1079, `VolumeManager`, `SourceManager`,
`ActorManager`).  Managers hold parameters only; the run phase is delegated to an engine chosen by
the new user parameter `Simulation.engine` — the seam the reference lacks:
`"b200"` selects `engines.GpuSimulationEngine`; `"geant4"` is the reference's own engine, which
does not exist in this package (fatal).
"""

from __future__ import annotations

import os
import random
import sys
from pathlib import Path

import numpy as np

from . import actors as gactors
from . import materials as gmat
from . import sources as gsources
from . import volumes as gvolumes
from .base import Box, GateObject
from .exception import fatal, warning
from .units import g4_units

__world_name__ = gvolumes.__world_name__


class VolumeManager:
    def __init__(self, simulation):
        self.simulation = simulation
        self.volumes = {}
        self.material_database = gmat.MaterialDatabase()
        self.material_database.read_from_file(gmat.DEFAULT_DB)
        world = gvolumes.BoxVolume(name=__world_name__, simulation=simulation)
        world.size = [3 * g4_units.m] * 3
        world.material = "G4_AIR"
        world.mother = None
        world.volume_manager = self
        self.volumes[__world_name__] = world

    @property
    def world_volume(self):
        return self.volumes[__world_name__]

    def add_volume(self, volume, name=None):
        if isinstance(volume, str):
            if volume not in gvolumes.volume_types:
                fatal(f"Volume type '{volume}' is not supported by the B200 GPU engine. "
                      f"Supported: {sorted(set(gvolumes.volume_types))}")
            if name is None:
                fatal("You must provide a name for the volume.")
            volume = gvolumes.volume_types[volume](name=name, simulation=self.simulation)
        if volume.name in self.volumes:
            fatal(f"The volume name {volume.name} already exists. Existing volume names are: {list(self.volumes)}")
        volume.volume_manager = self
        volume.simulation = self.simulation
        self.volumes[volume.name] = volume
        return volume

    def get_volume(self, name):
        if hasattr(name, "name"):
            name = name.name
        if name not in self.volumes:
            fatal(f"Cannot find volume {name}. Volumes included in this simulation are: {list(self.volumes)}")
        return self.volumes[name]

    def add_material_database(self, filename):
        if str(filename) in self.material_database.filenames:
            fatal(f'Database "{filename}" already exist.')
        self.material_database.read_from_file(filename)

    def find_or_build_material(self, name):
        return self.material_database.find_or_build_material(name)


class PhysicsManager(GateObject):
    """Physics parameters.  The GPU engine implements ONE physics configuration: photons with
    photoelectric, Compton, Rayleigh (and conversion), secondaries deposited locally — the
    reference's kerma configuration (production cuts larger than the phantom + apply_cuts,
    contrib/dose/doserate.py:102-116).  The parameters below are accepted so that scripts run
    unchanged; those that cannot change the result of this configuration are recorded, those
    that would (e.g. a physics list without EM) are rejected."""

    user_info_defaults = {
        "physics_list_name": ("QGSP_BERT_EMV", {"doc": "Name of the Geant4 physics list (informative on the GPU engine)"}),
        "enable_decay": (False, {"doc": "Radioactive decay (not simulated on the GPU engine)"}),
        "apply_cuts": (True, {"doc": "Secondaries below the production threshold are not created"}),
        "energy_range_min": (None, {"doc": "Lowest energy of the physics tables"}),
        "energy_range_max": (None, {"doc": "Highest energy of the physics tables"}),
        "photon_energy_cut": (1 * g4_units.keV, {"doc": "GPU engine: photons below this energy are absorbed on the spot"}),
        "pair_production_mode": ("local", {"doc": "GPU engine: 'local' deposits the whole photon energy at a "
                                 "conversion (kerma with apply_cuts); 'annihilation' deposits E-2mc2 and emits two 511 keV photons",
                                 "allowed_values": ("local", "annihilation")}),
        "table_energy_nodes": (1024, {"doc": "GPU engine: nodes of the log-uniform energy grid of the device tables"}),
    }

    def __init__(self, simulation):
        super().__init__(name="physics_manager", simulation=simulation)
        self.global_production_cuts = Box({"all": None, "gamma": None, "electron": None, "positron": None, "proton": None})
        self.production_cuts = {}
        self._freeze()

    def set_production_cut(self, volume_name, particle_name, value):
        if hasattr(volume_name, "name"):
            volume_name = volume_name.name
        if particle_name not in ("all", "gamma", "electron", "positron", "proton"):
            fatal(f"Unknown particle '{particle_name}' in set_production_cut")
        if volume_name == __world_name__:
            self.global_production_cuts[particle_name] = value
        else:
            self.production_cuts.setdefault(volume_name, {})[particle_name] = value

    def check_kerma_configuration(self, phantom_extent_mm):
        """The engine always runs the kerma approximation.  Warn when the script's cuts would make
        the reference track secondary electrons over distances that matter (> 1 voxel-ish)."""
        cuts = [self.global_production_cuts.get("electron"), self.global_production_cuts.get("all")]
        for v in self.production_cuts.values():
            cuts += [v.get("electron"), v.get("all")]
        cuts = [c for c in cuts if c is not None]
        if not self.apply_cuts or not cuts or max(cuts) < phantom_extent_mm:
            warning("The B200 GPU engine deposits secondary electrons locally (kerma approximation). For an "
                    "identical reference configuration set production cuts larger than the phantom and "
                    "apply_cuts=True (as opengate/contrib/dose/doserate.py does).")


class SourceManager:
    def __init__(self, simulation):
        self.simulation = simulation
        self.sources = {}

    def add_source(self, source, name=None):
        if isinstance(source, str):
            if source not in gsources.source_types:
                fatal(f"Source type '{source}' is not supported by the B200 GPU engine. "
                      f"Supported: {list(gsources.source_types)}")
            if name is None:
                fatal("You must provide a name for the source.")
            source = gsources.source_types[source](name=name, simulation=self.simulation)
        if source.name in self.sources:
            fatal(f"The source name {source.name} already exists.")
        source.simulation = self.simulation
        self.sources[source.name] = source
        return source

    def get_source(self, name):
        if name not in self.sources:
            fatal(f"Cannot find the source {name} in the list of sources: {list(self.sources)}")
        return self.sources[name]


class ActorManager:
    def __init__(self, simulation):
        self.simulation = simulation
        self.actors = {}

    def add_actor(self, actor, name=None):
        if isinstance(actor, str):
            if actor not in gactors.actor_types:
                fatal(f"Actor type '{actor}' is not supported by the B200 GPU engine. "
                      f"Supported: {list(gactors.actor_types)}")
            if name is None:
                fatal("You must provide a name for the actor.")
            actor = gactors.actor_types[actor](name=name, simulation=self.simulation)
        if actor.name in self.actors:
            fatal(f"The actor name {actor.name} already exists.")
        actor.simulation = self.simulation
        self.actors[actor.name] = actor
        return actor

    def get_actor(self, name):
        if name not in self.actors:
            fatal(f"Cannot find actor {name}. Actors in this simulation are: {list(self.actors)}")
        return self.actors[name]

    @property
    def sorted_actors(self):
        return sorted(self.actors.values(), key=lambda a: a.priority)


class Simulation(GateObject):
    """Main object of a script: holds volumes, sources, actors, physics and `run()`."""

    user_info_defaults = {
        "name": ("simulation", {"doc": "Name of the simulation"}),
        "verbose_level": ("INFO", {"doc": "Gate pre-run verbosity"}),
        "running_verbose_level": (0, {"doc": "Gate verbosity while the simulation is running."}),
        "g4_verbose": (False, {"doc": "unused by the GPU engine"}),
        "g4_verbose_level": (1, {"doc": "unused by the GPU engine"}),
        "visu": (False, {"doc": "Visualisation is not available on the GPU engine"}),
        "visu_type": ("vrml", {"doc": "unused by the GPU engine"}),
        "check_volumes_overlap": (True, {"doc": "Check that daughter boxes fit into the world"}),
        "number_of_threads": (1, {"doc": "Threads of the reference's Geant4 engine; ignored by the GPU engine"}),
        "number_of_gpus": (None, {"doc": "GPU engine: devices to use. None = the torchrun world (one process per "
                                  "GPU) or 1; an int > 1 in a single process uses that many devices in turn"}),
        "random_engine": ("MixMaxRng", {"doc": "Reference RNG name; the GPU engine always uses Philox4x32-10"}),
        "random_seed": ("auto", {"doc": "'auto' or an int (Philox key)"}),
        "run_timing_intervals": ([[0 * g4_units.s, 1 * g4_units.s]], {"doc": "List of [start, end] of each run"}),
        "output_dir": (".", {"doc": "Directory to which outputs are written"}),
        "progress_bar": (False, {"doc": "Display a progress bar"}),
        "store_json_archive": (False, {"doc": "unused"}),
        "engine": ("b200", {"doc": "Run-phase engine: 'b200' (CUDA engine of this package) or 'geant4' (the "
                            "reference's engine, only available in the reference package)",
                            "allowed_values": ("b200", "geant4")}),
        "max_histories_per_launch": (200_000_000, {"doc": "GPU engine: histories per kernel launch"}),
        "lean_outputs": (False, {"doc": "GPU engine: squared-sum images that are scored only because an uncertainty image "
                                        "needs them stay on the device in a single-run simulation and the uncertainty image is "
                                        "evaluated there (same bits as the host algebra); the squared images are then not "
                                        "available on the actor.  Default False: every active item comes back, as in the "
                                        "reference."}),
    }

    def __init__(self, name="simulation", **kwargs):
        super().__init__(name=name, **kwargs)
        self.simulation_dir = str(Path(sys.argv[0]).resolve().parent) if sys.argv and sys.argv[0] else os.getcwd()
        if not os.path.isdir(self.simulation_dir):
            self.simulation_dir = os.getcwd()
        self.volume_manager = VolumeManager(self)
        self.physics_manager = PhysicsManager(self)
        self.source_manager = SourceManager(self)
        self.actor_manager = ActorManager(self)
        self.current_random_seed = None
        self.expected_number_of_events = None
        self.warnings = []
        self.output = None
        self._freeze()

    # -- shortcuts (same names as the reference) ---------------------------------------------
    @property
    def world(self):
        return self.volume_manager.world_volume

    def add_volume(self, volume, name=None):
        return self.volume_manager.add_volume(volume, name)

    def add_source(self, source, name=None):
        return self.source_manager.add_source(source, name)

    def add_actor(self, actor, name=None):
        return self.actor_manager.add_actor(actor, name)

    def get_actor(self, name):
        return self.actor_manager.get_actor(name)

    def get_source(self, name):
        return self.source_manager.get_source(name)

    @property
    def multithreaded(self):
        return self.number_of_threads > 1

    def get_output_path(self, path=None, is_file_or_directory="file", suffix=None):
        root = Path(self.output_dir)
        if not root.is_absolute():
            root = (Path(self.simulation_dir) / root).resolve()
        p = root if path is None else (Path(path) if Path(path).is_absolute() else root / Path(path))
        if suffix is not None:
            p = p.with_name(f"{p.stem}_{suffix}{p.suffix}")
        d = p if (path is None or is_file_or_directory.lower().startswith("d")) else p.parent
        os.makedirs(d, exist_ok=True)
        return p

    # -- run -------------------------------------------------------------------------------------
    def resolve_and_validate_config(self):
        if self.random_seed == "auto":
            self.current_random_seed = random.randrange(sys.maxsize)
        else:
            self.current_random_seed = int(self.random_seed)
        if self.visu:
            fatal("Visualisation is not available with the B200 GPU engine")
        rti = self.run_timing_intervals
        for a, b in rti:
            if b < a:
                fatal(f"Wrong run timing interval [{a}, {b}]")
        for s in self.source_manager.sources.values():
            s.resolve_and_validate_config(rti)
        if not self.source_manager.sources:
            fatal("No source: no particle will be generated")

    def _run_simulation_engine(self, _backend=None):
        """Counterpart of Simulation._run_simulation_engine (managers.py:2803-2816) with the engine
        switch.  `_backend` is a test hook (tests inject the CPU oracle to check the host logic);
        the product default is the CUDA engine and nothing else."""
        self.resolve_and_validate_config()
        if self.engine == "geant4":
            fatal("engine='geant4' selects the reference's own SimulationEngine (opengate/engines.py), which is "
                  "not part of opengate_b200. Use engine='b200'.")
        from .engines import GpuSimulationEngine
        with GpuSimulationEngine(self, backend=_backend) as se:
            return se.run_engine()

    def run(self, start_new_process=False, _backend=None):
        if start_new_process:
            warning("start_new_process=True is not needed with the B200 GPU engine (no per-process Geant4 state); "
                    "running in this process.")
        out = self._run_simulation_engine(_backend=_backend)
        self.output = out
        self.expected_number_of_events = out.expected_number_of_events
        self.warnings = out.warnings
        return out


def create_sim_from_dict(d):  # pragma: no cover - convenience
    sim = Simulation()
    for k, v in d.get("user_info", {}).items():
        setattr(sim, k, v)
    return sim


__all__ = ["Simulation", "VolumeManager", "PhysicsManager", "SourceManager", "ActorManager", "np"]
