"""Python front of the CPU oracle (TEST INFRASTRUCTURE — never imported by the product).

* `TransportOracle`: ctypes wrapper of oracle/transport_oracle.c (built by oracle/Makefile),
  same call sequence as the product C-ABI.
* numpy restatements of the reference's host-side algorithms on the path (HU -> materials,
  label image, scoring algebra, 3-D CDF); each cites the reference file:line it follows.
  They are checked against golden vectors generated from the reference's own Python
  (tests/golden/, tools/gen_golden_from_reference.py).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from opengate_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_LIB64 = os.path.join(_HERE, "_build", "liboracle64.so")   # -DORACLE_FP64: the same algorithm in double precision


def build(force=False):
    src = os.path.join(_HERE, "transport_oracle.c")
    if force or any(not os.path.exists(l) or os.path.getmtime(l) < os.path.getmtime(src) for l in (_LIB, _LIB64)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


class PhspRow(C.Structure):
    _fields_ = [("kinetic_energy", C.c_float), ("position", C.c_float * 3),
                ("direction", C.c_float * 3), ("weight", C.c_float), ("time_ns", C.c_float),
                ("event_id", C.c_int64), ("flags", C.c_int32)]


TRACE_MAX = 64


class Trace(C.Structure):
    _fields_ = [("E0", C.c_float), ("p0", C.c_float * 3), ("d0", C.c_float * 3),
                ("E_end", C.c_float), ("edep", C.c_float), ("n_tentative", C.c_int),
                ("n_interactions", C.c_int), ("n_dep", C.c_int), ("n_iter", C.c_int), ("dep_e", C.c_float * TRACE_MAX)]


_libs = {}


def lib(fp64=False):
    """The C oracle: the single-precision build the CUDA engine matches bit for bit, or (fp64=True) the
    double-precision build of the same source (-DORACLE_FP64)."""
    if fp64 not in _libs:
        build()
        L = C.CDLL(_LIB64 if fp64 else _LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_uint64]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_tables.argtypes = [C.c_void_p, C.POINTER(abi.Tables)]
        L.oracle_set_world.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.oracle_set_image_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                              C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_add_source.argtypes = [C.c_void_p, C.POINTER(abi.Source)]
        L.oracle_add_dose_actor.argtypes = [C.c_void_p, C.POINTER(abi.DoseDesc)]
        L.oracle_add_tle_dose_actor.argtypes = [C.c_void_p, C.POINTER(abi.DoseDesc), C.POINTER(abi.TleDesc)]
        L.oracle_add_phsp_actor.argtypes = [C.c_void_p, C.POINTER(abi.PhspDesc)]
        L.oracle_add_box_region.argtypes = [C.c_void_p, C.POINTER(abi.RegionDesc)]
        L.oracle_begin_run.argtypes = [C.c_void_p]
        L.oracle_run_source.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        L.oracle_set_tle_fold.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_tle_fold.restype = C.c_int
        L.oracle_set_tle_rec_cap.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_tle_rec_cap.restype = C.c_int
        L.oracle_trace_history.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(Trace)]
        L.oracle_dose_grid.restype = C.POINTER(C.c_double)
        L.oracle_dose_grid.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.Stats)]
        L.oracle_phsp_rows.restype = C.c_uint64
        L.oracle_phsp_rows.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(PhspRow))]
        L.oracle_logf.restype = C.c_float
        L.oracle_logf.argtypes = [C.c_float]
        L.oracle_expf.restype = C.c_float
        L.oracle_expf.argtypes = [C.c_float]
        L.oracle_sincos2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]
        _libs[fp64] = L
    return _libs[fp64]


class TransportOracle:
    """Same call sequence as opengate_b200.core.Engine, executed by the scalar C oracle."""

    def __init__(self, seed, fp64=False):
        self.L = lib(fp64)
        self.h = C.c_void_p(self.L.oracle_create(C.c_uint64(seed)))
        self._keep = []
        self._dose_sizes = []

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tables(self, tables):
        s, keep = abi.make_tables(tables)
        self._keep.append(keep)
        self.L.oracle_set_tables(self.h, C.byref(s))

    def set_world(self, half_size, material):
        self.L.oracle_set_world(self.h, (C.c_double * 3)(*half_size), int(material))

    def set_image_volume(self, labels_zyx, spacing, world_to_volume12):
        lab = np.ascontiguousarray(labels_zyx)
        assert lab.dtype in (np.uint8, np.uint16)
        nz, ny, nx = lab.shape
        self.L.oracle_set_image_volume(self.h, lab.ctypes.data_as(C.c_void_p), lab.dtype.itemsize,
                                       (C.c_int * 3)(nx, ny, nz), (C.c_double * 3)(*spacing),
                                       (C.c_double * 12)(*np.asarray(world_to_volume12, float).reshape(-1)))

    def add_source(self, src_struct, keep=None):
        self._keep.append(keep)
        return self.L.oracle_add_source(self.h, C.byref(src_struct))

    def add_dose_actor(self, desc):
        self._dose_sizes.append(tuple(desc.size))
        return self.L.oracle_add_dose_actor(self.h, C.byref(desc))

    def add_tle_dose_actor(self, desc, tle_desc, keep=None):
        self._keep.append(keep)
        self._dose_sizes.append(tuple(desc.size))
        return self.L.oracle_add_tle_dose_actor(self.h, C.byref(desc), C.byref(tle_desc))

    def set_tle_fold(self, on=True):
        """score the (one, image-aligned, squares) TLE actor by the device's record-fold statement instead of the
        per-history table: a CPU cross-check of that algorithm (transport_oracle.c tle_fold_*)"""
        if self.L.oracle_set_tle_fold(self.h, int(on)) != 0:
            raise ValueError("the record-fold statement needs exactly one TLE actor, with squares, on the image grid")

    def set_tle_rec_cap(self, cap):
        """records (scored flights + point deposits) a history may hold before its sums are written out: 64 on the
        device; tests lower it on both sides to reach that overflow path"""
        if self.L.oracle_set_tle_rec_cap(self.h, int(cap)) != 0:
            raise ValueError("record cap out of range")

    def add_phsp_actor(self, desc):
        return self.L.oracle_add_phsp_actor(self.h, C.byref(desc))

    def add_box_region(self, desc):
        r = self.L.oracle_add_box_region(self.h, C.byref(desc))
        if r < 0:
            raise ValueError("box region refused (more than 4, or a material outside the tables)")
        return r

    def begin_run(self):
        self.L.oracle_begin_run(self.h)

    def run_source(self, source_id, first, n, n_threads=0):
        self.L.oracle_run_source(self.h, source_id, C.c_uint64(first), C.c_uint64(n), n_threads)

    def dose_grid(self, actor, which):
        nx, ny, nz = self._dose_sizes[actor]
        p = self.L.oracle_dose_grid(self.h, actor, which)
        if not p:
            raise ValueError(f"grid {which} is not scored by actor {actor}")
        return np.ctypeslib.as_array(p, shape=(nz, ny, nx)).copy()

    def stats(self):
        s = abi.Stats()
        self.L.oracle_get_stats(self.h, C.byref(s))
        return s.as_dict()

    def phsp(self, actor):
        rows = C.POINTER(PhspRow)()
        n = self.L.oracle_phsp_rows(self.h, actor, C.byref(rows))
        out = {"KineticEnergy": np.zeros(n, np.float32), "Position": np.zeros((n, 3), np.float32),
               "Direction": np.zeros((n, 3), np.float32), "Weight": np.zeros(n, np.float32),
               "Time": np.zeros(n, np.float32), "EventID": np.zeros(n, np.int64),
               "Flags": np.zeros(n, np.int32)}
        for i in range(n):
            r = rows[i]
            out["KineticEnergy"][i] = r.kinetic_energy
            out["Position"][i] = r.position[:]
            out["Direction"][i] = r.direction[:]
            out["Weight"][i] = r.weight
            out["Time"][i] = r.time_ns
            out["EventID"][i] = r.event_id
            out["Flags"][i] = r.flags
        return out

    def trace(self, source_id, hist):
        t = Trace()
        self.L.oracle_trace_history(self.h, source_id, C.c_uint64(hist), C.byref(t))
        return t


def uniforms(seed, hist, n):
    out = np.zeros(n, np.float32)
    lib().oracle_uniforms(C.c_uint64(seed), C.c_uint64(hist), n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def philox(seed, hist, block):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(C.c_uint64(seed), C.c_uint64(hist), C.c_uint32(block), out)
    return np.array(out[:], dtype=np.uint32)


# ---------------------------------------------------------------------------------------
# numpy restatements of the reference's host-side algorithms
# ---------------------------------------------------------------------------------------

G4_G_CM3 = 6.241509074460762e18   # opengate/utility.py:230 ("g/cm3")


def hu_read_materials_table(path):
    """Restates opengate/geometry/materials.py:61-100 (HU_read_materials_table)."""
    elements = ["HU"]
    materials = []
    section = None
    with open(path) as f:
        for line in f:
            cur = None
            for i, word in enumerate(line.split()):
                if word[0] == "#":
                    break
                if word == "[Elements]":
                    section = "element"
                    break
                if word == "[/Elements]":
                    elements.append("name")
                    section = "table"
                    break
                if section is None:
                    break
                if section == "element":
                    elements.append(word)
                if section == "table":
                    if cur is None:
                        cur = {}
                    if i == 0:
                        cur[elements[i]] = int(word)
                    elif i == len(elements) - 1:
                        cur[elements[i]] = word
                    else:
                        cur[elements[i]] = float(word)
            if cur:
                materials.append(cur)
    return materials, elements


def hu_read_density_table(path):
    """Restates materials.py:102-113."""
    out = []
    with open(path) as f:
        for line in f:
            w = line.split()
            if len(w) < 1 or w[0][0] == "#":
                continue
            out.append({"HU": float(w[0]), "density": float(w[1])})
    return out


def hu_interp_density(hu, densities):
    """Restates materials.py:116-129 (HU_linear_interpolate_densities)."""
    n = len(densities)
    i = 0
    while i < n and hu > densities[i]["HU"]:
        i += 1
    i -= 1
    if i < 0:
        return densities[0]["density"]
    if i >= n - 1:
        return densities[n - 1]["density"]
    return ((hu - densities[i]["HU"]) / (densities[i + 1]["HU"] - densities[i]["HU"])) * (
        densities[i + 1]["density"] - densities[i]["density"]) + densities[i]["density"]


def hu_max_density_difference(hu_min, hu_max, d_min, d_max, densities):
    """Restates materials.py:132-146."""
    n = len(densities)
    i = 0
    while i < n and hu_min > densities[i]["HU"]:
        i += 1
    j = 0
    while j < n and hu_max > densities[j]["HU"]:
        j += 1
    j -= 1
    for x in range(i, j):
        d_min = min(d_min, densities[x]["density"])
        d_max = max(d_max, densities[x]["density"])
    return d_max - d_min


def hounsfield_to_materials(density_tolerance_g_cm3, file_mat, file_density):
    """Restates materials.py:149-230 (HounsfieldUnit_to_material), densities in g/cm3.

    Returns (voxel_materials [[h1,h2,name]], materials {name: (density, {element: weight})}).
    """
    materials, elements = hu_read_materials_table(file_mat)
    densities = hu_read_density_table(file_density)
    elems = elements[1:-1]
    voxel_materials = []
    created = {}
    num = 0
    last = len(materials) - 1
    for i, mat in enumerate(materials):
        hu_min = mat["HU"]
        hu_max = hu_min + 1 if i == last else materials[i + 1]["HU"]
        if hu_max <= hu_min:
            raise ValueError(f"HU interval not valid: {mat}")
        dmin = hu_interp_density(hu_min, densities)
        dmax = hu_interp_density(hu_max, densities)
        ddiff = hu_max_density_difference(hu_min, hu_max, dmin, dmax, densities)
        # same arithmetic as the reference, which carries Geant4 units on both sides
        # (ddiff * gcm3 / density_tolerance with density_tolerance = tol * gcm3): the rounding of
        # the sub-interval bounds depends on it (pinned by tests/golden/schneider_reference.json)
        n = max(1, ddiff * G4_G_CM3 / (density_tolerance_g_cm3 * G4_G_CM3))
        if "Air" in mat["name"] or "AIR" in mat["name"]:
            n = 1
        htol = (hu_max - hu_min) / n
        if n != 1:
            n = int(np.ceil(n))
        for j in range(n):
            h1 = hu_min + j * htol
            h2 = min(hu_min + (j + 1) * htol, hu_max)
            d = hu_interp_density(h1 + (h2 - h1) / 2.0, densities)
            w = {e: mat[e] for e in elems if mat[e] > 0}
            s = sum(w.values())
            w = {e: v / s for e, v in w.items()}
            name = f'{mat["name"]}_{num}'
            created[name] = (d, w)
            voxel_materials.append([h1, h2, name])
            num += 1
    return voxel_materials, created


def material_to_label_lut(default_material, voxel_materials):
    """Restates opengate/geometry/volumes.py:1144-1168."""
    lut = {default_material: 0}
    vs = sorted(voxel_materials, key=lambda x: x[0])
    lo = np.array([r[0] for r in vs])
    hi = np.array([r[1] for r in vs])
    if not (lo[1:] >= hi[:-1]).all():
        raise ValueError("Overlapping intervals in voxel_materials")
    i = 1
    for r in vs:
        if r[2] not in lut:
            lut[r[2]] = i
            i += 1
    return lut


def label_image(hu_zyx, default_material, voxel_materials):
    """Restates volumes.py:1200-1243 (create_label_image): np.digitize over lower bounds plus
    upper bounds that are not lower bounds (those map to label 0)."""
    lut = material_to_label_lut(default_material, voxel_materials)
    vs = sorted(voxel_materials, key=lambda x: x[0])
    bins = [r[0] for r in vs]
    extra = set(r[1] for r in vs).difference(bins)
    bins.extend(extra)
    labels = [lut[r[2]] for r in vs]
    labels.extend(len(extra) * [0])
    bins_sorted, labels_sorted = [], [0]
    for b, l in sorted(zip(bins, labels), key=lambda p: p[0]):
        bins_sorted.append(b)
        labels_sorted.append(l)
    return np.array(labels_sorted, dtype=np.ushort)[np.digitize(hu_zyx, bins=bins_sorted)], lut


def calculate_variance(value, squared, n):
    """Restates opengate/utility.py:551-560."""
    return np.clip((squared / n - np.power(value / n, 2)) / (n - 1), 0, None)


def uncertainty(value, squared, n):
    """Restates opengate/actors/dataitems.py:1956-1993 ('uncertainty')."""
    value = np.asarray(value, dtype=np.float64)
    if not n > 1:
        return np.ones_like(value)
    if squared is None:
        return np.zeros_like(value)
    out = np.sqrt(calculate_variance(value, np.asarray(squared, dtype=np.float64), n))
    return np.divide(out, value / n, out=np.zeros_like(out), where=value != 0)


def image_3d_cdf(array_zyx):
    """Restates opengate/image.py:313-358 (compute_image_3D_CDF)."""
    array = array_zyx / np.sum(array_zyx)
    sumx = np.sum(array, axis=2)
    sumxy = np.sum(sumx, axis=1)
    cdf_x = []
    for i in range(array.shape[0]):
        cdf_x.append([])
        for j in range(array.shape[1]):
            t = np.cumsum(array[i][j])
            if t[-1] != 0:
                t = t / t[-1]
            cdf_x[i].append(t)
    cdf_y = []
    for i in range(len(sumx)):
        t = np.cumsum(sumx[i])
        if t[-1] != 0:
            t = t / t[-1]
        cdf_y.append(t)
    cdf_z = np.cumsum(sumxy) / np.sum(sumxy)
    return np.array(cdf_x), np.array(cdf_y), cdf_z


def mean_uncertainty(edep, edep_sq, n_events, top_voxels, threshold):
    """Restates GateDoseActor::ComputeMeanUncertainty (GateDoseActor.cpp:280-330)."""
    n = max(float(n_events), 2.0)
    flat = edep.ravel()
    k = min(int(top_voxels), flat.size)
    top = np.partition(flat, flat.size - k)[flat.size - k:]
    mean_max = top.mean() if k > 0 else 0.0
    sel = flat > mean_max * threshold
    if not sel.any():
        return 1.0
    val = flat[sel] / n
    sq = edep_sq.ravel()[sel] / n
    unc = np.clip((sq - val ** 2) / (n - 1.0), 0, None)
    unc = np.minimum(np.sqrt(unc) / val, 1.0)
    m = unc.sum()
    return m / sel.sum() if m > 0 else 1.0
