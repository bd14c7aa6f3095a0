"""DoseActor, PhaseSpaceActor, SimulationStatisticsActor: host-side mirror of the reference's actor
user interface for the photon-dose path.

Reference: opengate/actors/doseactors.py (VoxelDepositActor :34-288, DoseActor :317-719),
actors/actoroutput.py (interfaces, file names), actors/dataitems.py (variance/uncertainty
:1956-2005, MetaImage + number_of_samples sidecar :1336-1340), actors/digitizers.py
(PhaseSpaceActor :1579-1677), actors/miscactors.py (SimulationStatisticsActor :30-116).
The C++ halves (GateDoseActor, GatePhaseSpaceActor, GateSimulationStatisticsActor) are what the
CUDA engine replaces; these classes only hold parameters and receive the results from
`engines.GpuSimulationEngine` in the same per-run life cycle.
"""

from __future__ import annotations

import json
import os

import numpy as np

from . import image as gimage
from .base import GateObject
from .exception import fatal, warning
from .units import g4_units
from .volumes import __world_name__


# ------------------------------------------------------------------------------ scoring algebra

def calculate_variance(value_array, squared_value_array, number_of_samples):
    """Sample variance of the mean from sum(x) and sum(x^2) over events, clipped at 0
    (reference opengate/utility.py:551-560)."""
    n = number_of_samples
    mean = value_array / n
    return np.clip((squared_value_array / n - mean * mean) / (n - 1), 0, None)


def uncertainty_from_squares(value_array, squared_value_array, number_of_samples, which="uncertainty"):
    """variance | std | relative uncertainty (reference dataitems.py:1956-1993): ones when
    n <= 1, zeros without squares, uncertainty = std / mean where the value is non-zero."""
    value_array = np.asarray(value_array, dtype=np.float64)
    if not number_of_samples > 1:
        warning("You try to compute statistical errors with only one or zero event! "
                "The uncertainty value for all voxels has been fixed at 1")
        return np.ones_like(value_array)
    if squared_value_array is None:
        warning("This data item does not contain squared values so no variance can be calculated. "
                "The variance will be set to 0 everywhere. ")
        return np.zeros_like(value_array)
    out = calculate_variance(value_array, np.asarray(squared_value_array, dtype=np.float64), number_of_samples)
    if which in ("std", "uncertainty"):
        out = np.sqrt(out)
    if which == "uncertainty":
        out = np.divide(out, value_array / number_of_samples, out=np.zeros_like(out), where=value_array != 0)
    return out


def _insert_suffix(path, suffix):
    root, ext = os.path.splitext(str(path))
    return f"{root}_{suffix}{ext}"


# ------------------------------------------------------------------------------ outputs

class ActorOutputImage:
    """One user output made of items (0 = value, 1 = squared value, derived 'uncertainty').
    Holds per-run data and the merged data (sum over runs, samples added), writes MetaImage
    files + the `{"number_of_samples": N}` sidecar like the reference's ItkImageDataItem."""

    def __init__(self, actor, name, items, default_ext="mhd", merge="sum"):
        self.merge = merge      # "sum" (edep, dose, counts) or "mean" (density, ActorOutputSingleMeanImage)
        self.actor = actor
        self.name = name
        # item id -> config
        self.items = {k: {"active": bool(v.get("active", False)), "write_to_disk": v.get("write_to_disk", True),
                          "output_filename": "auto", "suffix": v.get("suffix", None), "keep_data_per_run": False}
                      for k, v in items.items()}
        self.default_ext = default_ext
        self.data_per_run = {}      # run -> {item: Image}
        self.samples_per_run = {}
        self.merged = None
        self.merged_samples = 0
        self.merged_runs = 0
        self.uncertainty_per_run = {}   # run -> Image evaluated by the engine on the device (single-run fast path)

    # configuration ------------------------------------------------------------------------
    def _ids(self, item):
        if item == "all" or item == "any":
            return list(self.items)
        if isinstance(item, (tuple, list)):
            return list(item)
        return [item]

    def get_active(self, item=0):
        ids = self._ids(item)
        flags = [self.items[i]["active"] for i in ids]
        if item == "all":
            return all(flags)
        return any(flags)

    def set_active(self, value, item=0):
        for i in self._ids(item):
            self.items[i]["active"] = bool(value)

    def get_write_to_disk(self, item=0):
        return any(self.items[i]["write_to_disk"] for i in self._ids(item))

    def set_write_to_disk(self, value, item=0):
        for i in self._ids(item):
            self.items[i]["write_to_disk"] = bool(value)

    def set_output_filename(self, value, item="all"):
        """Reference semantics (actoroutput.py:996-1018, actors/base.py:449-457): a name given for ALL items (what
        `actor.output_filename = ...` does) gets the item suffix inserted per item; a name given for ONE item
        (`actor.edep.output_filename = ...`) is stored verbatim."""
        if item == "all":
            for i, cfg in self.items.items():
                if value in (None, "", "auto") or not cfg["suffix"]:
                    cfg["output_filename"] = value
                else:
                    cfg["output_filename"] = _insert_suffix(value, cfg["suffix"])
            return
        for i in self._ids(item):
            self.items[i]["output_filename"] = None if value is None else str(value)

    def get_output_filename(self, item=0):
        f = self.items[item]["output_filename"]
        if f == "auto":
            # _generate_auto_output_filename (actoroutput.py:1041-1052): only the automatic name gets the suffix here
            f = f"{self.name}_from_{self.actor.type_name.lower()}_{self.actor.name}.{self.default_ext}"
            suffix = self.items[item]["suffix"]
            return _insert_suffix(f, suffix) if suffix else f
        if f in (None, ""):
            return None
        return f

    def get_output_path(self, which="merged", item=0):
        f = self.get_output_filename(item)
        if f is None:
            warning(f"No output_filename defined for user output '{self.name}' of {self.actor.type_name} "
                    f"'{self.actor.name}'. Therefore, get_output_path() returns None. ")
            return None
        p = self.actor.simulation.get_output_path(f)
        if which != "merged":
            p = _insert_suffix(p, f"run{int(which):04d}")
        return p

    # data ---------------------------------------------------------------------------------
    def store_data(self, run_index, *images):
        ids = [i for i in self.items if i != "uncertainty"]
        self.data_per_run[run_index] = {i: im for i, im in zip(ids, images)}

    def set_number_of_samples(self, run_index, n):
        self.samples_per_run[run_index] = int(n)

    def store_uncertainty(self, run_index, image):
        """relative uncertainty of the run, already evaluated (on the device) from the sums of this run"""
        self.uncertainty_per_run[run_index] = image

    def end_of_run(self, run_index):
        run = self.data_per_run[run_index]
        keep = any(c["keep_data_per_run"] for c in self.items.values())
        if self.merged is None:
            # the first run IS the merged data; it is copied only when the per-run data stay alive beside it
            self.merged = {i: (None if im is None else (gimage.Image(im.array.copy(), im.spacing, im.origin, im.direction) if keep else im))
                           for i, im in run.items()}
        else:
            for i, im in run.items():
                if im is not None and self.merged.get(i) is not None:
                    if self.merge == "mean":
                        k = len(self.samples_per_run)
                        self.merged[i].array += (im.array - self.merged[i].array) / max(k, 1)
                    else:
                        self.merged[i].array += im.array
        self.merged_samples += self.samples_per_run.get(run_index, 0)
        self.merged_runs += 1
        if not keep:
            # like the reference: per-run data are dropped unless explicitly kept
            self.data_per_run[run_index] = {i: None for i in run}

    def get_data(self, which="merged", item=0):
        src = self.merged if which == "merged" else self.data_per_run.get(int(which))
        n = self.merged_samples if which == "merged" else self.samples_per_run.get(int(which), 0)
        if src is None:
            return None
        if item == "uncertainty" or item in ("std", "variance"):
            v = src.get(0)
            if v is None:
                return None
            if item == "uncertainty":
                # evaluated by the engine from the sums of the run: valid for that run, and for the merged data
                # while they consist of that one run
                runs = list(self.uncertainty_per_run)
                if which != "merged" and int(which) in self.uncertainty_per_run:
                    return self.uncertainty_per_run[int(which)]
                if which == "merged" and self.merged_runs == 1 and len(runs) == 1:
                    return self.uncertainty_per_run[runs[0]]
            sq = src.get(1)
            arr = uncertainty_from_squares(v.array, None if sq is None else sq.array, n, which=item)
            return gimage.Image(arr, v.spacing, v.origin, v.direction)
        return src.get(item)

    def write(self):
        if self.merged is None:
            return
        for item, cfg in self.items.items():
            if not (cfg["active"] and cfg["write_to_disk"]):
                continue
            img = self.get_data("merged", item)
            if img is None:
                continue
            path = self.get_output_path("merged", item)
            if path is None:
                continue
            gimage.write_image(img, path)
            with open(os.path.splitext(str(path))[0] + ".json", "w") as f:
                json.dump({"number_of_samples": self.merged_samples}, f)


class UserInterfaceToActorOutputImage:
    """`actor.edep`, `actor.dose_uncertainty`, ...: the handle users touch
    (actoroutput.py:51-291)."""

    def __init__(self, output, item):
        object.__setattr__(self, "_o", output)
        object.__setattr__(self, "_item", item)

    active = property(lambda s: s._o.get_active(s._item), lambda s, v: s._o.set_active(v, s._item))
    write_to_disk = property(lambda s: s._o.get_write_to_disk(s._item), lambda s, v: s._o.set_write_to_disk(v, s._item))
    output_filename = property(lambda s: s._o.get_output_filename(s._item),
                               lambda s, v: s._o.set_output_filename(v, item=s._item))
    keep_data_per_run = property(lambda s: s._o.items[s._item]["keep_data_per_run"],
                                 lambda s, v: s._o.items[s._item].__setitem__("keep_data_per_run", bool(v)))

    def __setattr__(self, k, v):
        if k not in ("active", "write_to_disk", "output_filename", "keep_data_per_run"):
            fatal(f"Unknown attribute '{k}' of an actor output interface")
        object.__setattr__(self, k, v)

    def get_output_path(self, which="merged", **kw):
        return self._o.get_output_path(which, self._item)

    def get_data(self, which="merged", **kw):
        return self._o.get_data(which, self._item)

    @property
    def image(self):
        return self._o.get_data("merged", self._item)


# ------------------------------------------------------------------------------ actors

class ActorBase(GateObject):
    user_info_defaults = {
        "attached_to": (__world_name__, {"doc": "Name of the volume to which the actor is attached."}),
        "mother": (None, {"deprecated": "The user input parameter 'mother' is deprecated. Use 'attached_to' instead. "}),
        "filters": ([], {"doc": "Filters used by this actor (not supported on the GPU engine)"}),
        "priority": (100, {"doc": "Actors fire in order of increasing priority"}),
        "output_filename": ("auto", {"doc": "Output file name, relative to sim.output_dir; a suffix per item is inserted"}),
        "write_to_disk": (True, {"doc": "Should the output be written to disk"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.user_output = {}

    def __setattr__(self, key, value):
        if key == "attached_to" and hasattr(value, "name"):
            value = value.name
        super().__setattr__(key, value)
        if key == "output_filename" and self.__dict__.get("user_output"):
            for o in self.user_output.values():
                if hasattr(o, "set_output_filename"):
                    o.set_output_filename(value, item="all")
        if key == "write_to_disk" and self.__dict__.get("user_output"):
            for o in self.user_output.values():
                if hasattr(o, "set_write_to_disk"):
                    o.set_write_to_disk(value, item="all")

    @property
    def attached_to_volume(self):
        return self.simulation.volume_manager.get_volume(self.attached_to)

    def initialize(self):
        if self.filters:
            fatal(f"Actor '{self.name}': filters are not supported by the B200 GPU engine")

    def close(self):
        pass


def _hook_uncertainty_goal(self, value):
    if value is not None and (value < 0.0 or value > 1.0):
        fatal(f"Uncertainty goal must be > 0 and < 1, where 1 means 100%. The provided value is: {value}")
    return value


class DoseActor(ActorBase):
    """Deposited energy / absorbed dose in a voxel grid attached to a volume."""

    user_info_defaults = {
        "size": ([10, 10, 10], {"doc": "Size of the dose grid in number of voxels [N_x, N_y, N_z]"}),
        "spacing": ([1 * g4_units.mm] * 3, {"doc": "Voxel spacing along [x, y, z]"}),
        "translation": ([0.0, 0.0, 0.0], {"doc": "Translation of the grid from the centre of the attached volume"}),
        "rotation": (np.eye(3), {"doc": "Rotation of the output image (metadata only, as in the reference)"}),
        "repeated_volume_index": (0, {"doc": "Index of the repeated volume"}),
        "hit_type": ("random", {"doc": "Position to which the deposit is associated",
                                "allowed_values": ("random", "pre", "post", "middle")}),
        "output": (None, {"deprecated": "The output filename is now set via output_filename relative to the output "
                          "directory of the simulation, which can be set via sim.output_dir. "}),
        "img_coord_system": (None, {"deprecated": "Use my_actor.output_coordinate_system='attached_to_image' instead"}),
        "output_coordinate_system": ("local", {"doc": "local | global | attached_to_image | None",
                                               "allowed_values": ("local", "global", "attached_to_image", None)}),
        "square": (True, {"deprecated": "Use: my_actor.edep_squared.active=True/False "}),
        "uncertainty": (True, {"deprecated": "Use: my_actor.dose_uncertainty.active=True/False and "
                               "my_actor.edep_uncertainty.active=True/False "}),
        "dose": (False, {"deprecated": "Use: my_actor.dose.active=True/False "}),
        "to_water": (False, {"deprecated": "Use my_dose_actor.score_in='G4_WATER' instead. "}),
        "score_in": ("material", {"doc": "Only 'material' on the GPU engine"}),
        "uncertainty_goal": (None, {"doc": "Stop once the mean statistical uncertainty is below this value",
                                    "setter_hook": _hook_uncertainty_goal}),
        "uncertainty_first_check_after_n_events": (1e4, {"doc": "Events before the first uncertainty evaluation"}),
        "uncertainty_top_voxels_count": (10, {"doc": "N voxels with the highest edep define the 'mean maximum'"}),
        "uncertainty_voxel_edep_threshold": (0.7, {"doc": "Fraction of the mean maximum above which voxels count"}),
        "uncertainty_overshoot_factor_N_events": (1.05, {"doc": "Factor on the estimated number of events"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        mk = ActorOutputImage
        self.user_output = {
            "edep_with_uncertainty": mk(self, "edep_with_uncertainty",
                                        {0: {"suffix": "edep", "active": True}, 1: {"suffix": "edep_squared"},
                                         "uncertainty": {"suffix": "edep_uncertainty"}}),
            "dose_with_uncertainty": mk(self, "dose_with_uncertainty",
                                        {0: {"suffix": "dose"}, 1: {"suffix": "dose_squared"},
                                         "uncertainty": {"suffix": "dose_uncertainty"}}),
            "density": mk(self, "density", {0: {"suffix": "density"}}, merge="mean"),
            "counts": mk(self, "counts", {0: {"suffix": "counts"}}),
        }
        o = self.user_output
        UI = UserInterfaceToActorOutputImage
        # interfaces are plain instance attributes; `dose` is also the name of a deprecated
        # user parameter, so assignment by the user (`actor.dose = True`) still raises
        for nm, out, item in (("edep", "edep_with_uncertainty", 0), ("edep_squared", "edep_with_uncertainty", 1),
                              ("edep_uncertainty", "edep_with_uncertainty", "uncertainty"),
                              ("dose", "dose_with_uncertainty", 0), ("dose_squared", "dose_with_uncertainty", 1),
                              ("dose_uncertainty", "dose_with_uncertainty", "uncertainty"),
                              ("density", "density", 0), ("counts", "counts", 0)):
            object.__setattr__(self, nm, UI(o[out], item))
        self.NbOfEvent = 0
        self._freeze()

    # reference: DoseActor.initialize (doseactors.py:544-636) -----------------------------------
    def initialize(self):
        super().initialize()
        vol = self.attached_to_volume
        if self.output_coordinate_system == "attached_to_image" and not hasattr(vol, "native_translation"):
            fatal(f"User input 'output_coordinate_system' = {self.output_coordinate_system} of actor {self.name} "
                  f"is not compatible with the volume to which this actor is attached: "
                  f"{self.attached_to} ({vol.volume_type})")
        msg = (f"cannot be used in actor {self.name} because the volume ({self.attached_to}, {vol.type_name}) "
               f"to which the actor is attached does not support it. ")
        if isinstance(self.spacing, str) and self.spacing == "like_image_volume":
            if not hasattr(vol, "size_pix"):
                fatal("spacing = 'like_image_volume' " + msg)
            self.spacing = [float(x) for x in vol.spacing]
        if isinstance(self.size, str) and self.size == "like_image_volume":
            if not hasattr(vol, "size_pix"):
                fatal("size = 'like_image_volume' " + msg)
            self.size = [int(x) for x in vol.size_pix]
        if self.score_in != "material":
            fatal(f"DoseActor '{self.name}': score_in = '{self.score_in}' is not supported by the B200 GPU engine "
                  f"(only 'material')")
        e, d = self.user_output["edep_with_uncertainty"], self.user_output["dose_with_uncertainty"]
        e.set_active(True, item=0)
        # squares the user did not ask for, scored only because an uncertainty needs them
        self.__dict__["_squares_for_uncertainty_only"] = set()
        if e.get_active(item=("uncertainty",)) or self.uncertainty_goal is not None:
            if not e.get_active(item=1):
                e.set_write_to_disk(False, item=1)
                e.set_active(True, item=1)
                self._squares_for_uncertainty_only.add("edep")
        if d.get_active(item=("uncertainty",)):
            d.set_active(True, item=0)
            if not d.get_active(item=1):
                d.set_write_to_disk(False, item=1)
                d.set_active(True, item=1)
                self._squares_for_uncertainty_only.add("dose")
        if d.get_active(item=1):
            d.set_active(True, item=0)
        if self.user_output["density"].get_active() and vol.volume_type != "ImageVolume":
            fatal("The dose actor can only produce a density map if it is attached to an ImageVolume. "
                  f"This actor is attached to a {vol.volume_type} volume. ")
        if self.uncertainty_goal is not None:
            n_vox = int(np.prod(self.size))
            if self.uncertainty_top_voxels_count <= 0 or self.uncertainty_top_voxels_count > n_vox:
                fatal("uncertainty_top_voxels_count must be greater than 0 and cannot be bigger than the number "
                      "of voxels in the image. ")

    def score_flags(self):
        from . import abi
        e, d = self.user_output["edep_with_uncertainty"], self.user_output["dose_with_uncertainty"]
        f = abi.SCORE_EDEP
        if e.get_active(item=1):
            f |= abi.SCORE_EDEP_SQ
        if d.get_active(item=0):
            f |= abi.SCORE_DOSE
        if d.get_active(item=1):
            f |= abi.SCORE_DOSE_SQ
        if self.user_output["counts"].get_active():
            f |= abi.SCORE_COUNTS
        return f

    # output metadata: reference _update_output_coordinate_system (doseactors.py:170-232) ----------
    def output_origin_rotation(self):
        size = np.asarray(self.size, dtype=np.float64)
        spacing = np.asarray(self.spacing, dtype=np.float64)
        origin = -size * spacing / 2.0 + spacing / 2.0
        R = np.asarray(self.rotation, dtype=np.float64)
        origin_local = R @ origin + np.asarray(self.translation, dtype=np.float64)
        ocs = self.output_coordinate_system
        if ocs is None:
            return origin, np.eye(3)
        if ocs == "local":
            return origin_local, R
        vol = self.attached_to_volume
        if ocs == "attached_to_image":
            Rn = vol.native_rotation
            return Rn @ origin_local + vol.native_translation, Rn @ R
        Rw, tw = vol.world_transform()
        return Rw @ origin_local + tw, Rw @ R

    def dose_scale(self):
        """engine dose grid = sum(edep[MeV] / rho[g/cm3]); reference: edep / density (G4 units), then
        / (Gy * voxel volume) (GateDoseActor.cpp:206-217, doseactors.py:685-696)"""
        sp = np.asarray(self.spacing, dtype=np.float64)
        return 1.0 / (g4_units.g_cm3 * g4_units.Gy * sp[0] * sp[1] * sp[2])

    def fetch_plan(self, single_run):
        """What the engine has to bring back for one run: [(key, kind, grid, scale)], kind 'grid' (an accumulator,
        multiplied by `scale` on the device: the Gy conversion of doseactors.py:685-696) or 'uncertainty' (the
        relative-uncertainty image evaluated on the device).  With `sim.lean_outputs`, squares that are scored only
        because an uncertainty image needs them stay on the device when the run is the only one (nothing to merge):
        the uncertainty comes back instead."""
        from . import abi
        e, d = self.user_output["edep_with_uncertainty"], self.user_output["dose_with_uncertainty"]
        s = self.dose_scale()
        lean = set()
        if single_run and self.uncertainty_goal is None and getattr(self.simulation, "lean_outputs", False):
            lean = getattr(self, "_squares_for_uncertainty_only", set())
        plan = [(abi.GRID_EDEP, "grid", abi.GRID_EDEP, 1.0)]
        if e.get_active(item=1):
            plan.append(("edep_unc", "uncertainty", abi.GRID_EDEP, 1.0) if "edep" in lean else (abi.GRID_EDEP_SQ, "grid", abi.GRID_EDEP_SQ, 1.0))
        if d.get_active(item=0):
            plan.append((abi.GRID_DOSE, "grid", abi.GRID_DOSE, s))
        if d.get_active(item=1):
            plan.append(("dose_unc", "uncertainty", abi.GRID_DOSE, 1.0) if "dose" in lean else (abi.GRID_DOSE_SQ, "grid", abi.GRID_DOSE_SQ, s * s))
        if self.user_output["counts"].get_active():
            plan.append((abi.GRID_COUNTS, "grid", abi.GRID_COUNTS, 1.0))
        return plan

    def store_run(self, run_index, grids, n_events, scaled=False):
        """Receive the accumulator grids of one run (dict which -> float64 (z,y,x) array, plus 'edep_unc' /
        'dose_unc' when the engine evaluated the uncertainty) and apply what
        DoseActor.EndOfRunActionMasterThread does (doseactors.py:662-715).  scaled: the dose grids already
        carry the Gy conversion (done on the device)."""
        from . import abi
        self.NbOfEvent = int(n_events)
        origin, rot = self.output_origin_rotation()
        sp = np.asarray(self.spacing, dtype=np.float64)

        def img(a):
            return None if a is None else gimage.Image(a, sp, origin, rot)

        e = self.user_output["edep_with_uncertainty"]
        e.store_data(run_index, img(grids.get(abi.GRID_EDEP)), img(grids.get(abi.GRID_EDEP_SQ)))
        e.set_number_of_samples(run_index, n_events)
        if grids.get("edep_unc") is not None:
            e.store_uncertainty(run_index, img(grids["edep_unc"]))
        d = self.user_output["dose_with_uncertainty"]
        if d.get_active(item="any"):
            scale = 1.0 if scaled else self.dose_scale()
            dose = grids.get(abi.GRID_DOSE)
            dose_sq = grids.get(abi.GRID_DOSE_SQ)
            if scale != 1.0:
                dose = None if dose is None else dose * scale
                dose_sq = None if dose_sq is None else dose_sq * (scale * scale)
            d.store_data(run_index, img(dose), img(dose_sq))
            d.set_number_of_samples(run_index, n_events)
            if grids.get("dose_unc") is not None:
                d.store_uncertainty(run_index, img(grids["dose_unc"]))
        c = self.user_output["counts"]
        if c.get_active():
            c.store_data(run_index, img(grids.get(abi.GRID_COUNTS)))
            c.set_number_of_samples(run_index, n_events)
        den = self.user_output["density"]
        if den.get_active():
            dimg = self.attached_to_volume.create_density_image()
            same = tuple(dimg.size) == tuple(int(x) for x in self.size) and np.allclose(dimg.spacing, sp)
            if not same:
                dimg = _resample_nearest_like(dimg, self.size, sp, np.asarray(self.translation, float))
            den.store_data(run_index, gimage.Image(dimg.array.astype(np.float64), sp, origin, rot))
            den.set_number_of_samples(run_index, n_events)
        for o in self.user_output.values():
            if o.get_active(item="any") and run_index in o.data_per_run:
                o.end_of_run(run_index)

    def end_of_simulation(self):
        for o in self.user_output.values():
            if o.get_active(item="any"):
                o.write()


def _resample_nearest_like(dimg, size, spacing, translation):
    """Density image sampled at the centres of the dose grid (both centred on the volume)."""
    nx, ny, nz = [int(s) for s in size]
    out = np.zeros((nz, ny, nx), dtype=np.float64)
    src_size = dimg.size
    axes = []
    for a, n in enumerate((nx, ny, nz)):
        c = -n * spacing[a] / 2.0 + spacing[a] / 2.0 + translation[a] + np.arange(n) * spacing[a]
        idx = np.floor((c + src_size[a] * dimg.spacing[a] / 2.0) / dimg.spacing[a]).astype(int)
        axes.append(idx)
    ix, iy, iz = axes
    okx, oky, okz = [(i >= 0) & (i < s) for i, s in zip((ix, iy, iz), src_size)]
    sub = dimg.array[np.clip(iz, 0, src_size[2] - 1)][:, np.clip(iy, 0, src_size[1] - 1)][:, :, np.clip(ix, 0, src_size[0] - 1)]
    mask = okz[:, None, None] & oky[None, :, None] & okx[None, None, :]
    out[mask] = sub[mask]
    return gimage.Image(out, spacing, dimg.origin, dimg.direction)


class TLEDoseActor(DoseActor):
    """TLE = Track Length Estimator (reference doseactors.py:722-775, C++ GateTLEDoseActor.cpp): a DoseActor
    whose gammas deposit  w * E * mu_en/rho * L * rho  along every voxel chord of their flight instead of
    at the interaction points; `database` ("EPDL" | "NIST") selects the mu_en/rho tables of
    GateMaterialMuHandler.  On the GPU engine the march over the voxel chords is a phase of the transport
    kernel (`phase_tle`)."""

    user_info_defaults = {
        "energy_min": (0.0, {"doc": "Kill the gamma if below this energy"}),
        "tle_threshold": (np.inf, {"doc": "Define a criterium to enable TLE or not. It can be in terms of gamma energy or "
                                   "in secondary particle range depending on the provided tle_threshold_type"}),
        "tle_threshold_type": ("None", {"doc": "No threshold (None), by gamma energy (energy), or by the range of the "
                                        "secondary electron (max range / average range: need Geant4's CSDA "
                                        "ranges, not available on the GPU engine)"}),
        "database": ("EPDL", {"doc": "which database to use", "allowed_values": ("EPDL", "NIST")}),
    }

    def initialize(self):
        if self.score_in != "material":
            fatal(f"TLEDoseActor cannot score in {self.score_in}, only 'material' is allowed.")
        t = self.tle_threshold_type
        if t not in ("None", "energy", "max range", "average range"):
            fatal("You did not define a correct threshold type")     # GateTLEDoseActor.cpp:47-52
        if t in ("max range", "average range"):
            fatal(f"TLEDoseActor '{self.name}': tle_threshold_type = '{t}' needs the CSDA electron ranges of "
                  f"G4EmCalculator (GateTLEDoseActor.cpp:96-107), which the B200 GPU engine does not have; "
                  f"use 'energy' or 'None'")
        super().initialize()

    def tle_energy_threshold(self):
        """E <= threshold scores by track length; 'None' -> always (GateTLEDoseActor.cpp:224-243)."""
        return float(self.tle_threshold) if self.tle_threshold_type == "energy" else float("inf")


class PhaseSpaceActor(ActorBase):
    """Stores particle attributes when a photon enters / exits the attached Box volume."""

    SUPPORTED = ("KineticEnergy", "Weight", "EventID", "PrePosition", "Position", "PostPosition", "PreDirection",
                 "Direction", "PostDirection", "GlobalTime", "LocalTime", "TimeFromBeginOfEvent", "ParticleName",
                 "PDGCode", "TrackID", "ParentID", "PrePositionLocal", "PostPositionLocal", "PreDirectionLocal")

    user_info_defaults = {
        "attributes": ([], {"doc": "Names of the attributes to store"}),
        "store_absorbed_event": (False, {"doc": "Count the events that never reached the actor"}),
        "debug": (False, {"doc": "print debug info"}),
        "steps_to_store": ("entering", {"doc": "'entering', 'exiting' (or both, separated by a space)"}),
        "max_rows": (None, {"doc": "GPU engine: rows reserved on the device (default: one per primary)"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.total_number_of_entries = 0
        self.number_of_absorbed_events = 0
        self.data = {}
        self._freeze()

    def initialize(self):
        super().initialize()
        steps = self.steps_to_store.split() if isinstance(self.steps_to_store, str) else list(self.steps_to_store)
        allowed = {"entering", "exiting", "first", "all"}
        unknown = [s for s in steps if s not in allowed]
        if unknown:
            fatal(f"Unknown steps_to_store values {unknown} for actor '{self.name}'. Allowed values are {sorted(allowed)}.")
        bad = [s for s in steps if s in ("first", "all")]
        if bad:
            fatal(f"PhaseSpaceActor '{self.name}': steps_to_store {bad} is not supported by the B200 GPU engine")
        self.user_info.steps_to_store = " ".join(steps)
        for a in self.attributes:
            if a not in self.SUPPORTED:
                fatal(f"PhaseSpaceActor '{self.name}': attribute '{a}' is not supported by the B200 GPU engine. "
                      f"Supported: {self.SUPPORTED}")
        if self.attached_to_volume.volume_type != "BoxVolume":
            fatal(f"PhaseSpaceActor '{self.name}' must be attached to a Box volume on the B200 GPU engine")

    def steps_mask(self):
        from . import abi
        m = 0
        if "entering" in self.steps_to_store:
            m |= abi.PHSP_ENTERING
        if "exiting" in self.steps_to_store:
            m |= abi.PHSP_EXITING
        return m

    def get_output_path(self, which="merged"):
        f = self.output_filename
        if f == "auto":
            f = f"phsp_from_{self.type_name.lower()}_{self.name}.root"   # ActorOutputRoot: default suffix root
        return self.simulation.get_output_path(f)

    def store_rows(self, rows, n_absorbed, event_offset=0, event_time=None):
        """rows: dict of engine columns (KineticEnergy, Position, Direction, Weight, Time, EventID, Flags).
        Builds the requested attribute columns with the reference's types
        (D = float64, I = int32, 3 = 3 x float64, GateDigiCollection.cpp:240-271).
        event_time: callable(global history ids) -> time of the event in ns (timing.event_times of the source the
        id belongs to); `GlobalTime` = event time + time of flight (GateDigiAttributeList.cpp: GlobalTime is the
        post-step global time of the track, which starts at the event time)."""
        n = len(rows["EventID"])
        cols = {}
        t_event = event_time(rows["EventID"]) if (event_time is not None and "GlobalTime" in self.attributes) else None
        for a in self.attributes:
            if a == "KineticEnergy":
                cols[a] = rows["KineticEnergy"].astype(np.float64)
            elif a == "Weight":
                cols[a] = rows["Weight"].astype(np.float64)
            elif a == "EventID":
                cols[a] = (rows["EventID"] + event_offset).astype(np.int32)
            elif a in ("PrePosition", "Position", "PostPosition"):
                for i, ax in enumerate("XYZ"):
                    cols[f"{a}_{ax}"] = rows["Position"][:, i].astype(np.float64)
            elif a in ("PreDirection", "Direction", "PostDirection"):
                for i, ax in enumerate("XYZ"):
                    cols[f"{a}_{ax}"] = rows["Direction"][:, i].astype(np.float64)
            elif a in ("PrePositionLocal", "PostPositionLocal", "PreDirectionLocal"):
                # coordinates of the attached volume (top transform of the touchable,
                # GateDigiAttributeList.cpp:241-254,282-288): p_local = R^T (p - t), d_local = R^T d
                R, t = self.attached_to_volume.world_transform()
                if a == "PreDirectionLocal":
                    loc = rows["Direction"].astype(np.float64) @ np.asarray(R, dtype=np.float64)
                else:
                    loc = (rows["Position"].astype(np.float64) - np.asarray(t, dtype=np.float64)) @ np.asarray(R, dtype=np.float64)
                for i, ax in enumerate("XYZ"):
                    cols[f"{a}_{ax}"] = np.ascontiguousarray(loc[:, i])
            elif a == "GlobalTime":
                cols[a] = rows["Time"].astype(np.float64) + (t_event if t_event is not None else 0.0)
            elif a in ("LocalTime", "TimeFromBeginOfEvent"):
                cols[a] = rows["Time"].astype(np.float64)
            elif a == "ParticleName":
                cols[a] = np.full(n, "gamma")
            elif a == "PDGCode":
                cols[a] = np.full(n, 22, dtype=np.int32)
            elif a == "TrackID":
                cols[a] = np.ones(n, dtype=np.int32)
            elif a == "ParentID":
                cols[a] = np.zeros(n, dtype=np.int32)
        for k, v in cols.items():
            self.data[k] = np.concatenate([self.data[k], v]) if k in self.data else v
        self.total_number_of_entries += n
        self.number_of_absorbed_events += int(n_absorbed)

    def end_of_simulation(self):
        if self.total_number_of_entries == 0:
            warning(f"Empty output, no particles stored in {self.get_output_path()}")
        if not self.write_to_disk:
            return
        path = str(self.get_output_path())
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if path.endswith(".root"):
            # a flat TTree named after the actor, one branch per column (GateDigiCollectionsRootManager: D -> double,
            # I -> int, 3 -> name_X / _Y / _Z), written by the package's own ROOT writer (rootio.py)
            from . import rootio
            rootio.write_tree(path, self.name, self.data, title=f"PhaseSpaceActor {self.name}")
            return
        np.savez(path, **self.data)


class SimulationStatisticsActor(ActorBase):
    """runs / events / tracks / steps, duration and primaries per second
    (GateSimulationStatisticsActor.cpp:42-226, dataitems.py:468-480)."""

    user_info_defaults = {"track_types_flag": (False, {"doc": "Should the type of tracks be counted?"})}

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.counts = {"runs": 0, "events": 0, "tracks": 0, "steps": 0, "duration": 0.0, "init": 0.0,
                       "track_types": {}, "nb_threads": 1}
        self._freeze()

    def store(self, runs, events, tracks, steps, duration_ns, init_ns=0.0, extra=None):
        c = self.counts
        c.update(runs=int(runs), events=int(events), tracks=int(tracks), steps=int(steps),
                 duration=float(duration_ns), init=float(init_ns))
        if self.track_types_flag:
            c["track_types"] = {"gamma": int(tracks)}
        if extra:
            c.update(extra)

    def _rate(self, key):
        sec = g4_units.s
        return self.counts[key] / (self.counts["duration"] / sec) if self.counts["duration"] > 0 else 0.0

    pps = property(lambda s: s._rate("events"))
    tps = property(lambda s: s._rate("tracks"))
    sps = property(lambda s: s._rate("steps"))

    def end_of_simulation(self):
        if self.write_to_disk and self.output_filename not in ("auto", None, ""):
            path = str(self.simulation.get_output_path(self.output_filename))
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            with open(path, "w") as f:
                json.dump({k: v for k, v in self.counts.items()}, f, indent=2)

    def __str__(self):
        c = self.counts
        return (f"runs     {c['runs']}\nevents   {c['events']}\ntracks   {c['tracks']}\nsteps    {c['steps']}\n"
                f"duration {c['duration'] / g4_units.s:.4f} s\npps      {self.pps:.1f}\n")


actor_types = {"DoseActor": DoseActor, "TLEDoseActor": TLEDoseActor, "PhaseSpaceActor": PhaseSpaceActor,
               "SimulationStatisticsActor": SimulationStatisticsActor}
