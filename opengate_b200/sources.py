"""GenericSource / VoxelSource: host-side mirror of the reference's source user interface
(opengate/sources/base.py:13-296, generic.py:24-77,333-550, voxelsources.py:15-129), gammas only.

The objects only hold and validate parameters; `to_abi(...)` converts one source to the
`gate_b200_source` descriptor of the C-ABI (frames composed to WORLD as GateVSource.cpp:95-107 /
GateGenericSource.cpp:160-195 do; spectrum CDFs built as GateGenericSource.cpp:545-644).
"""

from __future__ import annotations

import numpy as np

from . import abi
from . import image as gimage
from .base import Box, GateObject
from .exception import fatal, warning
from .units import g4_units
from .volumes import __world_name__


def _position_parameters():
    return Box({"type": "point", "radius": 0, "sigma_x": 0, "sigma_y": 0, "size": [0, 0, 0],
                "translation": [0, 0, 0], "rotation": np.eye(3), "confine": None, "dz": None})


def _direction_parameters():
    return Box({"type": "iso", "theta": [0, 180 * g4_units.deg], "phi": [0, 360 * g4_units.deg],
                "momentum": [0, 0, 1], "focus_point": [0, 0, 0], "sigma": [0, 0],
                "angular_acceptance": Box({"policy": None, "skip_policy": "SkipEvents", "volumes": []}),
                "accolinearity_flag": False, "accolinearity_fwhm": 0.5 * g4_units.deg,
                "histogram_theta_weights": [], "histogram_theta_angles": [],
                "histogram_phi_weights": [], "histogram_phi_angles": []})


def _energy_parameters():
    return Box({"type": "mono", "mono": 0, "sigma_gauss": 0, "is_cdf": False, "min_energy": None,
                "max_energy": None, "spectrum_type": None, "spectrum_weights": [], "spectrum_energies": [],
                "histogram_weight": [], "histogram_energy": [], "spectrum_energy_bin_edges": [],
                "spectrum_histogram_interpolation": None})


def _hook_particle(self, value):
    if value != "gamma":
        fatal(f"Source '{self.user_info.get('name')}': the B200 GPU engine transports gammas only "
              f"(particle = '{value}' is not supported)")
    return value


class SourceBase(GateObject):
    user_info_defaults = {
        "attached_to": (__world_name__, {"doc": "Name of the volume to which the source is attached."}),
        "mother": (None, {"deprecated": "The user input parameter 'mother' is deprecated. Use 'attached_to' instead. "}),
        "start_time": (None, {"doc": "Starting time of the source"}),
        "end_time": (None, {"doc": "End time of the source"}),
        "number_of_primaries": (0, {"doc": "Number of primaries to generate (exclusive with 'activity')"}),
        "n": (0, {"deprecated": "The user input parameter 'n' is deprecated. Use 'number_of_primaries' instead."}),
        "activity": (0, {"doc": "Activity of the source in Bq (exclusive with 'number_of_primaries')"}),
        "half_life": (-1, {"doc": "Half-life decay (-1 if no decay). Only when used with 'activity'"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.run_timing_intervals = None

    def __setattr__(self, key, value):
        if key == "attached_to" and hasattr(value, "name"):
            value = value.name
        super().__setattr__(key, value)

    def resolve_and_validate_config(self, run_timing_intervals):
        """Same rules and messages as SourceBase.resolve_and_validate_config (base.py:257-293)."""
        self.run_timing_intervals = run_timing_intervals
        ui = self.user_info
        if ui.start_time is None:
            ui.start_time = run_timing_intervals[0][0]
        if ui.end_time is None:
            ui.end_time = run_timing_intervals[-1][1]
        n = np.atleast_1d(np.array(ui.number_of_primaries)).astype(int)
        ui.number_of_primaries = n
        if ui.activity == 0 and len(n) != len(run_timing_intervals):
            fatal("source.number_of_primaries and run_timing_intervals do not have the same length.")
        if np.any(n > 0) and ui.activity > 0:
            fatal("Cannot use both the two parameters 'number_of_primaries' and 'activity' at the same time. ")
        if np.all(n == 0) and ui.activity == 0:
            fatal("You must set one of the two parameters 'number_of_primaries' or 'activity'.")
        if ui.activity > 0:
            ui.number_of_primaries = np.zeros(len(run_timing_intervals), dtype=int)

    def primaries_for_run(self, run_index, rng=None):
        """Histories of this source in run `run_index`.  `number_of_primaries[run]` as is; an
        `activity` source emits N ~ Poisson(A_eff * dt) events where the reference draws
        exponential inter-event times until the interval ends (GateVSource.cpp:55-89): the same
        count distribution, without per-event time stamps (documented difference)."""
        ui = self.user_info
        if ui.activity > 0:
            t0, t1 = self.run_timing_intervals[run_index]
            t0, t1 = max(t0, ui.start_time), min(t1, ui.end_time)
            if t1 <= t0:
                return 0
            a = ui.activity * self._activity_scale()
            if ui.half_life > 0:
                lam = np.log(2.0) / ui.half_life
                mean = a / lam * (np.exp(-lam * (t0 - ui.start_time)) - np.exp(-lam * (t1 - ui.start_time)))
            else:
                mean = a * (t1 - t0)
            rng = rng or np.random.default_rng(0)
            return int(rng.poisson(mean))
        return int(ui.number_of_primaries[run_index])

    def _activity_scale(self):
        return 1.0


class GenericSource(SourceBase):
    user_info_defaults = {
        "particle": ("gamma", {"doc": "Name of the particle generated by the source", "setter_hook": _hook_particle}),
        "weight": (-1, {"doc": "Particle initial weight (for variance reduction technique)"}),
        "weight_sigma": (-1, {"doc": "not supported on the GPU engine"}),
        "direction_relative_to_attached_volume": (False, {"doc": "Rotate the direction with the attached volume"}),
        "position": (_position_parameters(), {"doc": "Define the position of the primary particles"}),
        "direction": (_direction_parameters(), {"doc": "Define the direction of the primary particles"}),
        "energy": (_energy_parameters(), {"doc": "Define the energy of the primary particles"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._freeze()

    # -- validation (messages follow generic.py:90-330) ----------------------------------------
    def resolve_and_validate_config(self, run_timing_intervals):
        super().resolve_and_validate_config(run_timing_intervals)
        p, d, e = self.position, self.direction, self.energy
        ctx = f"source '{self.name}'"
        valid = {"sphere", "point", "box", "disc", "cylinder", "surface_sphere"}
        if p.type not in valid:
            fatal(f"In {ctx}: '{p.type}' is not a valid position type. Must be one of {valid}.")
        if p.radius < 0 or p.sigma_x < 0 or p.sigma_y < 0:
            fatal(f"In {ctx}: radius, sigma_x and sigma_y must be >= 0")
        if len(p.size) != 3 or len(p.translation) != 3:
            fatal(f"In {ctx}: size and translation must be 3-vectors")
        if np.asarray(p.rotation).shape != (3, 3):
            fatal(f"In {ctx}: 'rotation' must be convertible to a 3x3 matrix/array, but got: {p.rotation}")
        if p.confine and p.confine != "None":
            # generic.py:500-508: a confined point source is almost certainly a mistake
            if p.type == "point":
                warning(f"In source {self.name}, confine is used, while position.type is point ... really ?")
            vols = self.simulation.volume_manager.volumes
            if p.confine not in vols:
                fatal(f"In {ctx}: position.confine = '{p.confine}' is not a volume of the simulation")
            if not hasattr(vols[p.confine], "half_size_mm"):
                fatal(f"In {ctx}: position.confine needs a Box or Image volume on the B200 GPU engine, "
                      f"'{p.confine}' is a {vols[p.confine].volume_type}")
        valid_dir = ["iso", "histogram", "momentum", "focused", "beam2d", "cos"]   # generic.py:160
        if d.type not in valid_dir:
            fatal(f"In {ctx}: Cannot find the direction type '{d.type}'. Available types are {valid_dir}")
        if d.type == "histogram":
            # generic.py:194-205 / GateGenericSource.cpp:440-458: n weights for n + 1 angles, both histograms
            for nm in ("theta", "phi"):
                w, a = getattr(d, f"histogram_{nm}_weights"), getattr(d, f"histogram_{nm}_angles")
                if len(w) != len(a) - 1 or len(w) < 1:
                    fatal(f"In {ctx}: histogram_{nm}_weights must have -1 elements than histogram_{nm}_angles")
                lim = np.pi if nm == "theta" else 2 * np.pi
                if np.any(np.diff(np.asarray(a, float)) < 0) or np.min(a) < 0 or np.max(a) > lim * (1 + 1e-12):
                    fatal(f"In {ctx}: histogram_{nm}_angles must increase within [0, {'pi' if nm == 'theta' else '2 pi'}]")
                if np.any(np.asarray(w, float) < 0) or not np.sum(w) > 0:
                    fatal(f"In {ctx}: histogram_{nm}_weights must be >= 0 and not all zero")
        if d.angular_acceptance.get("policy") not in (None, "None"):
            fatal(f"In {ctx}: direction.angular_acceptance is not supported by the B200 GPU engine")
        et = "CDF" if e.is_cdf else e.type
        if et not in {"mono", "gauss", "range", "spectrum_discrete", "spectrum_histogram", "CDF"}:
            fatal(f"In {ctx}: energy.type '{e.type}' is not supported by the B200 GPU engine")
        if e.type == "spectrum_discrete":
            if len(e.spectrum_weights) == 0:
                fatal(f"The weights for {self.name} is zero length. Abort")
            if len(e.spectrum_weights) != len(e.spectrum_energies):
                fatal(f"For {self.name}, the spectrum vectors weights and energies must have the same size "
                      f"({len(e.spectrum_weights)} ≠ {len(e.spectrum_energies)})")
        if e.type == "spectrum_histogram":
            if len(e.spectrum_weights) == 0:
                fatal(f"The weights for {self.name} is zero length. Abort")
            if len(e.spectrum_weights) + 1 != len(e.spectrum_energy_bin_edges):
                fatal(f"For {self.name}, the spectrum vector energy_bin_edges must have exactly one more element "
                      f"than the vector weights ({len(e.spectrum_energy_bin_edges)} ≠ {len(e.spectrum_weights)} + 1)")
            if e.spectrum_histogram_interpolation not in (None, "None", "none", "linear"):
                fatal(f"For {self.name}, invalid spectrum interpolation type: {e.spectrum_histogram_interpolation}")
        if self.weight_sigma is not None and self.weight_sigma > 0:
            fatal(f"In {ctx}: weight_sigma is not supported by the B200 GPU engine")

    def max_energy(self):
        e = self.energy
        if e.is_cdf:
            return float(np.max(e.spectrum_energies))
        if e.type == "mono":
            return float(e.mono)
        if e.type == "gauss":
            return float(e.mono + 6.0 * e.sigma_gauss)
        if e.type == "range":
            return float(e.max_energy)
        if e.type == "spectrum_discrete":
            return float(np.max(e.spectrum_energies))
        return float(np.max(e.spectrum_energy_bin_edges))

    def _activity_scale(self):
        """The reference multiplies the activity by the sum of the spectrum weights before
        normalising them (GateGenericSource.cpp:560-566, 622-626)."""
        e = self.energy
        if e.type == "spectrum_discrete":
            return float(np.sum(e.spectrum_weights))
        if e.type == "spectrum_histogram":
            return float(self._histogram_cdf()[1])
        return 1.0

    def _histogram_cdf(self):
        e = self.energy
        w = np.asarray(e.spectrum_weights, dtype=np.float64).copy()
        edges = np.asarray(e.spectrum_energy_bin_edges, dtype=np.float64)
        interp = e.spectrum_histogram_interpolation
        acc = 0.0
        if interp in (None, "None", "none"):
            for i in range(len(w)):
                acc += w[i] * (edges[i + 1] - edges[i])
                w[i] = acc
        else:
            # GateGenericSource.cpp:613-621 (reads weights[i+1]; the last bin has no right neighbour
            # in the reference either, it reads one element past the user vector: treat as equal)
            src = np.append(np.asarray(e.spectrum_weights, dtype=np.float64), e.spectrum_weights[-1])
            for i in range(len(w)):
                de = edges[i + 1] - edges[i]
                dw = src[i + 1] - src[i]
                acc += de * src[i] - 0.5 * de * dw
                w[i] = acc
        total = w[-1]
        return w / total, total

    # -- conversion to the C-ABI descriptor ------------------------------------------------------
    def _frames(self, attached_volume):
        """(R_pos, t_pos, R_dir): source-local -> world."""
        Rv, tv = attached_volume.world_transform() if attached_volume is not None else (np.eye(3), np.zeros(3))
        p = self.position
        Rl = np.asarray(p.rotation, dtype=np.float64)
        tl = np.asarray(p.translation, dtype=np.float64)
        return Rv @ Rl, Rv @ tl + tv, Rv

    def to_abi(self, attached_volume):
        p, d, e = self.position, self.direction, self.energy
        R_pos, t_pos, R_vol = self._frames(attached_volume)
        kw = {}
        kw["pos_type"] = {"point": abi.POS_POINT, "box": abi.POS_BOX, "sphere": abi.POS_SPHERE,
                          "disc": abi.POS_DISC, "cylinder": abi.POS_CYLINDER,
                          "surface_sphere": abi.POS_SURFACE_SPHERE}[p.type]
        half = [0.5 * float(s) for s in p.size]
        if p.type == "cylinder":
            half[2] = float(p.dz) if p.dz is not None else 0.0   # G4 SetHalfZ(dz), GateGenericSource.cpp:350-352
        kw.update(half_size=half, radius=p.radius, sigma_x=p.sigma_x, sigma_y=p.sigma_y,
                  pos_rotation=R_pos, pos_translation=t_pos)
        # direction: world frame unless direction_relative_to_attached_volume, in which case the
        # GLOBAL rotation of the source (attached volume x position.rotation) is applied
        # (GateGenericSource.cpp:176-195, GateSPSAngDistribution.cpp:10-18)
        relative = bool(self.direction_relative_to_attached_volume)
        if d.type in ("iso", "cos"):
            kw.update(dir_type=abi.DIR_ISO if d.type == "iso" else abi.DIR_COS, theta=d.theta, phi=d.phi,
                      dir_relative=relative, dir_rotation=R_pos)
        elif d.type == "beam2d":
            # two Gaussian angular deviations around -z of the source frame (GateGenericSource.cpp:428-433)
            kw.update(dir_type=abi.DIR_BEAM2D, dir_sigma=[float(x) for x in d.sigma], dir_relative=relative,
                      dir_rotation=R_pos)
        elif d.type == "histogram":
            # user histograms of theta and phi (GateGenericSource.cpp:435-464): weights are bin contents
            kw.update(dir_type=abi.DIR_HISTOGRAM, dir_relative=relative, dir_rotation=R_pos,
                      theta_hist=(np.asarray(d.histogram_theta_angles, float), np.asarray(d.histogram_theta_weights, float)),
                      phi_hist=(np.asarray(d.histogram_phi_angles, float), np.asarray(d.histogram_phi_weights, float)))
        elif d.type == "momentum":
            m = np.asarray(d.momentum, dtype=np.float64)
            if relative:
                m = R_pos @ m
            kw.update(dir_type=abi.DIR_MOMENTUM, momentum=m)
        else:
            f = np.asarray(d.focus_point, dtype=np.float64)
            if relative:
                f = R_pos @ (f - np.asarray(p.translation, dtype=np.float64)) + t_pos
            kw.update(dir_type=abi.DIR_FOCUSED, focus_point=f)
        if e.is_cdf:
            kw.update(ene_type=abi.ENE_CDF, spectrum_energy=np.asarray(e.spectrum_energies, float),
                      spectrum_cdf=np.asarray(e.spectrum_weights, float))
        elif e.type == "mono":
            kw.update(ene_type=abi.ENE_MONO, mono=e.mono)
        elif e.type == "gauss":
            kw.update(ene_type=abi.ENE_GAUSS, mono=e.mono, sigma=e.sigma_gauss)
        elif e.type == "range":
            kw.update(ene_type=abi.ENE_RANGE, e_min=e.min_energy, e_max=e.max_energy)
        elif e.type == "spectrum_discrete":
            w = np.cumsum(np.asarray(e.spectrum_weights, dtype=np.float64))
            kw.update(ene_type=abi.ENE_SPECTRUM_DISCRETE, spectrum_energy=np.asarray(e.spectrum_energies, float),
                      spectrum_cdf=w / w[-1])
        else:
            cdf, _ = self._histogram_cdf()
            linear = e.spectrum_histogram_interpolation == "linear"
            kw.update(ene_type=abi.ENE_SPECTRUM_HISTOGRAM_LINEAR if linear else abi.ENE_SPECTRUM_HISTOGRAM,
                      spectrum_energy=np.asarray(e.spectrum_energy_bin_edges, float), spectrum_cdf=cdf)
        kw["weight"] = 1.0 if self.weight is None or self.weight < 0 else float(self.weight)
        if p.confine and p.confine != "None":
            # the confining volume as a box in the world (GateGenericSource.cpp:223-227: ConfineSourceToVolume)
            cv = self.simulation.volume_manager.volumes[p.confine]
            Rc, tc = cv.world_transform()
            w2c = np.zeros((3, 4))
            w2c[:, :3] = Rc.T
            w2c[:, 3] = -Rc.T @ tc
            hs = cv.half_size_mm() if callable(cv.half_size_mm) else cv.half_size_mm
            kw.update(confine_half_size=[float(x) for x in hs], world_to_confine=w2c)
        return abi.make_source(**kw)


class VoxelSource(GenericSource):
    """3-D activity distribution sampled with cumulative distribution functions
    (voxelsources.py:15-129, GateSPSVoxelsPosDistribution.cpp:29-72)."""

    user_info_defaults = {"image": (None, {"doc": "Filename (or Image) of the 3D activity distribution"})}

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        object.__setattr__(self, "_frozen", False)
        self._activity_image = None
        self._freeze()

    def resolve_and_validate_config(self, run_timing_intervals):
        self.position.type = "point" if self.position.type not in ("point",) else self.position.type
        super().resolve_and_validate_config(run_timing_intervals)
        if self.image is None:
            fatal(f"VoxelSource '{self.name}': no activity image set")

    def load_image(self):
        if self._activity_image is None:
            src = self.image
            if isinstance(src, gimage.Image):
                self._activity_image = src
            elif isinstance(src, np.ndarray):
                self._activity_image = gimage.Image(src)
            else:
                self._activity_image = gimage.read_image(str(src))
        return self._activity_image

    def to_abi(self, attached_volume):
        s, keep = super().to_abi(attached_volume)
        img = self.load_image()
        if float(np.sum(img.array)) <= 0:
            fatal(f"VoxelSource '{self.name}': the activity image is empty")
        cdf_x, cdf_y, cdf_z = gimage.compute_image_3D_CDF(img)
        size = img.size.astype(np.float64)
        # origin of voxel (0,0,0): half size + translation + half-pixel shift (voxelsources.py:67-81);
        # only the attached volume's transform is applied afterwards (GateSPSVoxelsPosDistribution.cpp:68-70)
        org = -size / 2.0 * img.spacing + np.asarray(self.position.translation, dtype=np.float64) + img.spacing / 2.0
        Rv, tv = attached_volume.world_transform() if attached_volume is not None else (np.eye(3), np.zeros(3))
        import ctypes as C
        s.pos_type = abi.POS_VOXELS
        s.pos_rotation = (C.c_double * 9)(*Rv.reshape(-1))
        s.pos_translation = (C.c_double * 3)(*tv)
        s.vox_size = (C.c_int32 * 3)(*[int(x) for x in img.size])
        s.vox_spacing = (C.c_double * 3)(*[float(x) for x in img.spacing])
        s.vox_origin = (C.c_double * 3)(*[float(x) for x in org])
        cz = np.ascontiguousarray(cdf_z, dtype=np.float32)
        cy = np.ascontiguousarray(cdf_y, dtype=np.float32)
        cx = np.ascontiguousarray(cdf_x, dtype=np.float32)
        keep += [cz, cy, cx]
        s.cdf_z, s.cdf_y, s.cdf_x = abi.fptr(cz), abi.fptr(cy), abi.fptr(cx)
        return s, keep


def read_phase_space_columns(path, keys):
    """Columns `keys` (those present) of a phase-space file -> dict of numpy arrays, and the names of all columns.
    `.root`: first tree of the file, branches by name, as the reference reads it through uproot
    (sources/phspsources.py:59-75,153-158) -- here through the package's own reader of flat TTrees (rootio.py:
    uncompressed / zlib / lzma, fixed-size leaves), falling back to uproot, where installed, for what that reader
    does not cover (LZ4 / ZSTD compression, split or jagged branches); `.npz`: a column archive."""
    path = str(path)
    if path.endswith(".npz") or path.endswith(".npy"):
        with np.load(path, allow_pickle=False) as f:
            cols = {k: np.asarray(f[k]) for k in keys if k in f.files}
            avail = list(f.files)
        return cols, avail
    from . import rootio
    try:
        return rootio.read_tree(path, keys=list(keys))
    except KeyError:
        fatal(f"PhaseSpaceSourceGenerator: No usable branches in the root file {path}. Aborting.")
    except ValueError as why:
        try:
            import uproot
        except ImportError:
            fatal(f"PhaseSpaceSource: cannot read '{path}': {why}; the uproot package would be needed for this file")
        f = uproot.open(path)
        branches = f.keys()
        if len(branches) == 0:
            fatal(f"PhaseSpaceSourceGenerator: No usable branches in the root file {path}. Aborting.")
        tree = f[branches[0]]
        avail = list(tree.keys())
        cols = tree.arrays([k for k in keys if k in avail], library="numpy")
        return cols, avail


class PhaseSpaceSource(SourceBase):
    """Source of particles replayed from a phase-space file: position + direction + energy + weight of
    every event are read from the file (sources/phspsources.py:262-590, GatePhaseSpaceSource.cpp:92-178).
    The whole file is uploaded to the device once; the event with global id i replays entry
    (entry_start + i) % num_entries, the order in which the reference's generator walks the file
    (sequential batches, restart from the beginning at the end: phspsources.py:119-146)."""

    user_info_defaults = {
        "phsp_file": (None, {"doc": "Phase-space file to replay (.root through uproot, or a .npz column archive). Required.",
                             "is_input_file": True}),
        "entry_start": (None, {"doc": "Index of the first entry to replay (taken modulo the number of entries)"}),
        "particle": ("", {"doc": "'gamma', or empty to check the PDGCode column (the GPU engine replays gammas only)"}),
        "global_flag": (False, {"doc": "True: positions / directions in the file are world coordinates; "
                                       "False: coordinates of the volume the source is attached to"}),
        "isotropic_direction": (False, {"doc": "Take position, energy and weight from the file, draw the direction isotropically"}),
        "translate_position": (False, {"doc": "Shift every position by position.translation"}),
        "rotate_direction": (False, {"doc": "Turn every direction by position.rotation"}),
        "batch_size": (10000, {"doc": "Kept for compatibility: the GPU engine uploads the whole file once"}),
        "position_key": ("PrePositionLocal", {"doc": "Column prefix of the position (_X, _Y, _Z are appended)"}),
        "position_key_x": (None, {"doc": "Column of the x position (default: position_key + '_X')"}),
        "position_key_y": (None, {"doc": "Column of the y position"}),
        "position_key_z": (None, {"doc": "Column of the z position"}),
        "position": (Box({"translation": [0, 0, 0], "rotation": np.eye(3)}),
                     {"doc": "Translation / rotation used by translate_position and rotate_direction"}),
        "direction_key": ("PreDirectionLocal", {"doc": "Column prefix of the direction (_X, _Y, _Z are appended)"}),
        "direction_key_x": (None, {"doc": "Column of the x direction (default: direction_key + '_X')"}),
        "direction_key_y": (None, {"doc": "Column of the y direction"}),
        "direction_key_z": (None, {"doc": "Column of the z direction"}),
        "energy_key": ("KineticEnergy", {"doc": "Column of the kinetic energy (MeV)"}),
        "weight_key": ("Weight", {"doc": "Column of the statistical weight (all ones when absent)"}),
        "PDGCode_key": ("PDGCode", {"doc": "Column of the PDG code, read when 'particle' is empty"}),
        "generate_until_next_primary": (False, {"doc": "Several particles per event: not supported on the GPU engine"}),
        "primary_lower_energy_threshold": (0, {"doc": "Only with generate_until_next_primary"}),
        "primary_PDGCode": (0, {"doc": "Only with generate_until_next_primary"}),
        "verbose": (False, {"doc": "Accepted for compatibility"}),
        "verbose_batch": (False, {"doc": "Accepted for compatibility"}),
    }

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.num_entries = None
        self._columns = None
        self._freeze()

    def resolve_and_validate_config(self, run_timing_intervals):
        super().resolve_and_validate_config(run_timing_intervals)
        ui = self.user_info
        ctx = f"PhaseSpaceSource '{self.name}'"
        if ui.phsp_file is None:
            fatal(f"{ctx}: phsp_file is required")
        if int(ui.batch_size) < 1:
            fatal("PhaseSpaceSourceGenerator: Batch size should be > 0")
        if ui.particle not in ("", None, "None", "gamma"):
            fatal(f"{ctx}: the B200 GPU engine transports gammas only (particle = '{ui.particle}' is not supported)")
        if ui.generate_until_next_primary:
            fatal(f"{ctx}: generate_until_next_primary is not supported by the B200 GPU engine")
        if ui.activity > 0:
            fatal(f"{ctx}: 'activity' is not supported for a phase-space source on the B200 GPU engine")
        if ui.entry_start is None:
            ui.entry_start = 0
        if not np.isscalar(ui.entry_start):
            fatal("entry_start must be a simple number is mono-thread mode")
        for k3 in ("position", "direction"):
            for ax in "xyz":
                if ui[f"{k3}_key_{ax}"] is None:
                    ui[f"{k3}_key_{ax}"] = f"{ui[k3 + '_key']}_{ax.upper()}"

    def load_columns(self):
        """float32 columns after the Python-side transforms of PhaseSpaceSourceGenerator.generate
        (phspsources.py:195-235)."""
        if self._columns is not None:
            return self._columns
        ui = self.user_info
        keys = [ui.position_key_x, ui.position_key_y, ui.position_key_z, ui.energy_key]
        if not ui.isotropic_direction:
            keys += [ui.direction_key_x, ui.direction_key_y, ui.direction_key_z]
        opt = [k for k in (ui.weight_key, ui.PDGCode_key) if k]
        raw, avail = read_phase_space_columns(ui.phsp_file, keys + opt)
        for k in keys:
            if k not in raw:
                fatal(f"PhaseSpaceSource: Key '{k}' not found. Available: {avail}")
        f32 = lambda k: np.array(raw[k], dtype=np.float32, copy=True, order="C")
        c = {"pos_x": f32(ui.position_key_x), "pos_y": f32(ui.position_key_y), "pos_z": f32(ui.position_key_z),
             "energy": f32(ui.energy_key)}
        n = len(c["pos_x"])
        if n == 0:
            fatal(f"PhaseSpaceSource '{self.name}': the phase-space file {ui.phsp_file} is empty")
        if len(c["energy"]) != n:
            fatal(f"Size mismatch: Pos {n} vs Energy {len(c['energy'])}")
        if not ui.isotropic_direction:
            c.update(dir_x=f32(ui.direction_key_x), dir_y=f32(ui.direction_key_y), dir_z=f32(ui.direction_key_z))
        c["weight"] = f32(ui.weight_key) if ui.weight_key and ui.weight_key in raw else np.ones(n, dtype=np.float32)
        if ui.particle in ("", None, "None"):
            if ui.PDGCode_key not in raw:
                fatal(f"PhaseSpaceSource: Key '{ui.PDGCode_key}' not found. Available: {avail}")
            pdg = np.asarray(raw[ui.PDGCode_key]).astype(np.int32)
            if np.any(pdg != 22):
                bad = int(pdg[pdg != 22][0])
                fatal(f"PhaseSpaceSource '{self.name}': the phase space holds particles other than gammas "
                      f"(PDGCode {bad}); the B200 GPU engine transports gammas only")
        if ui.translate_position:
            t = ui.position.translation
            c["pos_x"] += float(t[0])
            c["pos_y"] += float(t[1])
            c["pos_z"] += float(t[2])
        if ui.rotate_direction and not ui.isotropic_direction:
            pts = np.column_stack((c["dir_x"], c["dir_y"], c["dir_z"])).astype(np.float64)
            rot = pts @ np.asarray(ui.position.rotation, dtype=np.float64).T
            for i, ax in enumerate("xyz"):
                c[f"dir_{ax}"] = np.ascontiguousarray(rot[:, i], dtype=np.float32)
        if not ui.isotropic_direction:
            nrm = c["dir_x"].astype(np.float64) ** 2 + c["dir_y"].astype(np.float64) ** 2 + c["dir_z"].astype(np.float64) ** 2
            if np.any(nrm <= 0):
                fatal(f"PhaseSpaceSource '{self.name}': entry {int(np.argmax(nrm <= 0))} has a null direction")
        self.num_entries = n
        self._columns = c
        return c

    def max_energy(self):
        return float(np.max(self.load_columns()["energy"]))

    @property
    def cycle_count(self):
        """How many times the replay passed the end of the file (phspsources.py:133-146)."""
        if self.num_entries is None:
            return 0
        n_events = int(np.sum(self.user_info.number_of_primaries))
        return int((int(self.user_info.entry_start) % self.num_entries + n_events) // self.num_entries)

    def to_abi(self, attached_volume):
        ui = self.user_info
        c = self.load_columns()
        n = self.num_entries
        start = int(ui.entry_start)
        if start > n:
            warning(f"In source {self.name} entry_start = {start} while the phsp contains {n}. "
                    f"We consider {start % n} instead (modulo)")
        if ui.global_flag or attached_volume is None:
            R, t = np.eye(3), np.zeros(3)
        else:
            R, t = attached_volume.world_transform()
        kw = dict(pos_type=abi.POS_PHSP, ene_type=abi.ENE_PHSP, pos_rotation=R, pos_translation=t,
                  dir_relative=not ui.global_flag, dir_rotation=R, weight=1.0, phsp=c, phsp_entry_start=start % n)
        if ui.isotropic_direction:
            kw.update(dir_type=abi.DIR_ISO, theta=(0.0, np.pi), phi=(0.0, 2 * np.pi))
        else:
            kw.update(dir_type=abi.DIR_PHSP)
        return abi.make_source(**kw)


source_types = {"GenericSource": GenericSource, "VoxelSource": VoxelSource, "PhaseSpaceSource": PhaseSpaceSource}


def scale_counts_for_rank(counts, rank, world):
    """n//T + (t < n%T): the reference's per-thread split (sources/base.py:190-197), used per GPU."""
    counts = np.asarray(counts, dtype=np.int64)
    return counts // world + np.asarray(rank < counts % world, dtype=np.int64)


__all__ = ["GenericSource", "VoxelSource", "PhaseSpaceSource", "SourceBase", "source_types", "scale_counts_for_rank", "warning"]
