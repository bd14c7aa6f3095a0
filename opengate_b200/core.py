"""ctypes binding of libgate_b200.so — the thin C-ABI layer (`opengate_core` side of the seam).

The reference reaches its run phase through the pybind11 module `opengate_core`
(core/opengate_core/opengate_core.cpp); this module plays that role for the GPU engine and
binds exactly the entry points declared in include/gate_b200.h.  There is NO CPU fallback:
if the shared library is missing or no CUDA device is present, construction raises.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libgate_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
]

EXPORTS = [
    "gate_b200_abi_version", "gate_b200_create", "gate_b200_destroy", "gate_b200_release_cached_memory", "gate_b200_last_error",
    "gate_b200_set_stream", "gate_b200_set_tables", "gate_b200_set_world",
    "gate_b200_set_image_volume", "gate_b200_set_image_volume_hu", "gate_b200_add_source", "gate_b200_add_dose_actor",
    "gate_b200_add_tle_dose_actor", "gate_b200_add_phsp_actor", "gate_b200_add_box_region", "gate_b200_begin_run", "gate_b200_run_source", "gate_b200_sync",
    "gate_b200_get_dose_grid", "gate_b200_dose_buffers", "gate_b200_set_dose_grid",
    "gate_b200_mean_uncertainty", "gate_b200_phsp_count", "gate_b200_get_phsp",
    "gate_b200_get_stats", "gate_b200_timed_run", "gate_b200_kernel_name",
    "gate_b200_get_dose_grid_scaled", "gate_b200_get_dose_grid_range", "gate_b200_get_uncertainty", "gate_b200_host_register", "gate_b200_host_unregister",
]


class GateB200Error(RuntimeError):
    pass


def build_library(force=False, verbose=False):
    """Compile csrc/gate_b200.cu for sm_100a into csrc/libgate_b200.so (in-tree)."""
    srcs = [os.path.join(CSRC, f) for f in ("gate_b200.cu", "transport.cuh", "device_math.cuh", "uncertainty.cuh")]
    srcs.append(os.path.join(_HERE, "..", "include", "gate_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, srcs[0]]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load_library():
    """dlopen the C-ABI library and declare prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = os.environ.get("GATE_B200_LIB", LIB_PATH)   # development: A/B builds of the same ABI
    if not os.path.exists(lib_path):
        raise GateB200Error(
            f"{lib_path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The GPU engine has no CPU fallback."
        )
    L = C.CDLL(lib_path)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    L.gate_b200_abi_version.restype = C.c_int
    L.gate_b200_create.argtypes = [C.c_int, u64, C.POINTER(vp)]
    L.gate_b200_destroy.argtypes = [vp]
    L.gate_b200_release_cached_memory.argtypes = [C.c_int]
    L.gate_b200_release_cached_memory.restype = C.c_int
    L.gate_b200_destroy.restype = None
    L.gate_b200_last_error.argtypes = [vp]
    L.gate_b200_last_error.restype = C.c_char_p
    L.gate_b200_kernel_name.argtypes = [vp]
    L.gate_b200_kernel_name.restype = C.c_char_p
    L.gate_b200_set_stream.argtypes = [vp, vp]
    L.gate_b200_set_tables.argtypes = [vp, C.POINTER(abi.Tables)]
    L.gate_b200_set_world.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.gate_b200_set_image_volume.argtypes = [vp, vp, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.gate_b200_set_image_volume_hu.argtypes = [vp, vp, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double), i32,
                                                C.POINTER(C.c_double), C.POINTER(C.c_uint16), i32, vp]
    L.gate_b200_add_source.argtypes = [vp, C.POINTER(abi.Source), C.POINTER(i32)]
    L.gate_b200_add_dose_actor.argtypes = [vp, C.POINTER(abi.DoseDesc), C.POINTER(i32)]
    L.gate_b200_add_tle_dose_actor.argtypes = [vp, C.POINTER(abi.DoseDesc), C.POINTER(abi.TleDesc), C.POINTER(i32)]
    L.gate_b200_add_phsp_actor.argtypes = [vp, C.POINTER(abi.PhspDesc), C.POINTER(i32)]
    L.gate_b200_add_box_region.argtypes = [vp, C.POINTER(abi.RegionDesc), C.POINTER(i32)]
    L.gate_b200_begin_run.argtypes = [vp, i32]
    L.gate_b200_run_source.argtypes = [vp, i32, u64, u64]
    L.gate_b200_sync.argtypes = [vp]
    L.gate_b200_get_dose_grid.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
    L.gate_b200_set_dose_grid.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
    L.gate_b200_get_dose_grid_scaled.argtypes = [vp, i32, i32, C.c_double, C.POINTER(C.c_double)]
    L.gate_b200_get_dose_grid_range.argtypes = [vp, i32, i32, C.c_double, u64, u64, C.POINTER(C.c_double)]
    L.gate_b200_get_uncertainty.argtypes = [vp, i32, i32, u64, C.POINTER(C.c_double)]
    L.gate_b200_host_register.argtypes = [vp, u64]
    L.gate_b200_host_unregister.argtypes = [vp]
    L.gate_b200_dose_buffers.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(u64), C.POINTER(i32)]
    L.gate_b200_mean_uncertainty.argtypes = [vp, i32, u64, i32, C.c_double, C.POINTER(C.c_double)]
    L.gate_b200_phsp_count.argtypes = [vp, i32, C.POINTER(u64)]
    L.gate_b200_get_phsp.argtypes = [vp, i32, u64, C.POINTER(abi.PhspColumns), C.POINTER(u64)]
    L.gate_b200_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    L.gate_b200_timed_run.argtypes = [vp, i32, u64, u64, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name not in ("gate_b200_destroy", "gate_b200_last_error", "gate_b200_kernel_name"):
            getattr(L, name).restype = C.c_int
    v = L.gate_b200_abi_version()
    if v != abi.ABI_VERSION:
        raise GateB200Error(f"ABI mismatch: library {v}, python {abi.ABI_VERSION}")
    _lib = L
    return L


def release_cached_memory(device=0):
    """Give the device memory cached from destroyed engines back to the driver (gate_b200_release_cached_memory)."""
    rc = load_library().gate_b200_release_cached_memory(int(device))
    if rc != 0:
        raise GateB200Error(f"gate_b200_release_cached_memory failed ({rc})")


def host_register(array):
    """page-lock a numpy array in place (gate_b200_host_register); False if the driver refuses"""
    return load_library().gate_b200_host_register(array.ctypes.data_as(C.c_void_p), C.c_uint64(array.nbytes)) == 0


def host_unregister(array):
    load_library().gate_b200_host_unregister(array.ctypes.data_as(C.c_void_p))


class Engine:
    """Object wrapper of one gate_b200_engine handle (one CUDA device)."""

    def __init__(self, device=0, seed=0):
        self.L = load_library()
        self.h = C.c_void_p()
        self._keep = []
        self._dose_sizes = []
        rc = self.L.gate_b200_create(int(device), C.c_uint64(int(seed) & (2**64 - 1)), C.byref(self.h))
        if rc != 0:
            msg = self.L.gate_b200_last_error(self.h).decode() if self.h else "allocation failed"
            if self.h:
                self.L.gate_b200_destroy(self.h)
                self.h = None
            raise GateB200Error(f"gate_b200_create failed ({rc}): {msg}")
        self.device = int(device)

    def _check(self, rc):
        if rc != 0:
            raise GateB200Error(f"[{rc}] {self.L.gate_b200_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.gate_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # configuration -------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.gate_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_tables(self, tables):
        s, keep = abi.make_tables(tables)
        self._keep.append(keep)
        self._check(self.L.gate_b200_set_tables(self.h, C.byref(s)))

    def set_world(self, half_size, material):
        self._check(self.L.gate_b200_set_world(self.h, (C.c_double * 3)(*[float(x) for x in half_size]), int(material)))

    def set_image_volume(self, labels_zyx, spacing, world_to_volume12):
        lab = np.ascontiguousarray(labels_zyx)
        if lab.dtype not in (np.uint8, np.uint16):
            raise GateB200Error("labels must be uint8 or uint16")
        nz, ny, nx = lab.shape
        self._check(self.L.gate_b200_set_image_volume(
            self.h, lab.ctypes.data_as(C.c_void_p), lab.dtype.itemsize, (C.c_int32 * 3)(nx, ny, nz),
            (C.c_double * 3)(*[float(x) for x in spacing]),
            (C.c_double * 12)(*np.asarray(world_to_volume12, float).reshape(-1))))

    def set_image_volume_hu(self, hu_zyx, spacing, world_to_volume12, bins, bin_labels, label_bytes=1, want_labels=False):
        """HU image -> label volume on the device (np.digitize rule of ImageVolume.create_label_image).
        Returns the label array (z,y,x) when want_labels, else None."""
        hu = np.ascontiguousarray(hu_zyx)
        types = {np.dtype(np.int16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
        if hu.dtype not in types:
            hu = hu.astype(np.float64)
        nz, ny, nx = hu.shape
        b = np.ascontiguousarray(bins, dtype=np.float64)
        bl = np.ascontiguousarray(bin_labels, dtype=np.uint16)
        if len(bl) != len(b) + 1:
            raise GateB200Error("bin_labels must have len(bins) + 1 entries")
        out = np.empty((nz, ny, nx), dtype=np.uint8 if label_bytes == 1 else np.uint16) if want_labels else None
        self._check(self.L.gate_b200_set_image_volume_hu(
            self.h, hu.ctypes.data_as(C.c_void_p), types[hu.dtype], (C.c_int32 * 3)(nx, ny, nz),
            (C.c_double * 3)(*[float(x) for x in spacing]),
            (C.c_double * 12)(*np.asarray(world_to_volume12, float).reshape(-1)), len(b),
            b.ctypes.data_as(C.POINTER(C.c_double)), bl.ctypes.data_as(C.POINTER(C.c_uint16)), int(label_bytes),
            out.ctypes.data_as(C.c_void_p) if out is not None else None))
        return out

    def add_source(self, src_struct, keep=None):
        self._keep.append(keep)
        sid = C.c_int32(-1)
        self._check(self.L.gate_b200_add_source(self.h, C.byref(src_struct), C.byref(sid)))
        return sid.value

    def add_dose_actor(self, desc):
        aid = C.c_int32(-1)
        self._check(self.L.gate_b200_add_dose_actor(self.h, C.byref(desc), C.byref(aid)))
        self._dose_sizes.append(tuple(desc.size))
        return aid.value

    def add_tle_dose_actor(self, desc, tle_desc, keep=None):
        self._keep.append(keep)
        aid = C.c_int32(-1)
        self._check(self.L.gate_b200_add_tle_dose_actor(self.h, C.byref(desc), C.byref(tle_desc), C.byref(aid)))
        self._dose_sizes.append(tuple(desc.size))
        return aid.value

    def add_phsp_actor(self, desc):
        aid = C.c_int32(-1)
        self._check(self.L.gate_b200_add_phsp_actor(self.h, C.byref(desc), C.byref(aid)))
        return aid.value

    def add_box_region(self, desc):
        """a homogeneous Box of its own material in the world (gate_b200_add_box_region)"""
        rid = C.c_int32(-1)
        self._check(self.L.gate_b200_add_box_region(self.h, C.byref(desc), C.byref(rid)))
        return rid.value

    # run -----------------------------------------------------------------------------
    def begin_run(self, run_id=0):
        self._check(self.L.gate_b200_begin_run(self.h, int(run_id)))

    def run_source(self, source_id, first, n):
        self._check(self.L.gate_b200_run_source(self.h, int(source_id), C.c_uint64(int(first)), C.c_uint64(int(n))))

    def sync(self):
        self._check(self.L.gate_b200_sync(self.h))

    def timed_run(self, source_id, first, n):
        ms = C.c_double(0.0)
        self._check(self.L.gate_b200_timed_run(self.h, int(source_id), C.c_uint64(int(first)), C.c_uint64(int(n)), C.byref(ms)))
        return ms.value

    # results -------------------------------------------------------------------------
    def dose_grid(self, actor, which, out=None, scale=1.0):
        """one accumulator grid as a (z,y,x) float64 array, every value multiplied by `scale` on the device"""
        nx, ny, nz = self._dose_sizes[actor]
        if out is None:
            out = np.empty((nz, ny, nx), dtype=np.float64)
        self._check(self.L.gate_b200_get_dose_grid_scaled(self.h, actor, which, float(scale), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def dose_grid_range(self, actor, which, first, count, out, scale=1.0):
        """voxels [first, first + count) of one accumulator grid (flat, x fastest) into out[first:first + count] of a
        flat float64 array `out` that spans the whole grid (e.g. one shared by the ranks of a node)"""
        flat = out.reshape(-1)
        ptr = C.cast(flat.ctypes.data + 8 * int(first), C.POINTER(C.c_double))
        self._check(self.L.gate_b200_get_dose_grid_range(self.h, actor, which, float(scale), C.c_uint64(int(first)), C.c_uint64(int(count)), ptr))
        return out

    def uncertainty_grid(self, actor, which, n_events, out=None):
        """relative uncertainty image of edep (which = GRID_EDEP) or dose (GRID_DOSE) of the run, computed on the device"""
        nx, ny, nz = self._dose_sizes[actor]
        if out is None:
            out = np.empty((nz, ny, nx), dtype=np.float64)
        self._check(self.L.gate_b200_get_uncertainty(self.h, actor, which, C.c_uint64(int(n_events)), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def dose_grid_shape(self, actor):
        nx, ny, nz = self._dose_sizes[actor]
        return (nz, ny, nx)

    def pin(self, array):
        """page-lock a numpy array in place (gate_b200_host_register); False if the driver refuses"""
        return host_register(array)

    def unpin(self, array):
        host_unregister(array)

    def set_dose_grid(self, actor, which, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        self._check(self.L.gate_b200_set_dose_grid(self.h, actor, which, a.ctypes.data_as(C.POINTER(C.c_double))))

    def dose_buffers(self, actor):
        """[(device pointer, number of doubles)]: the voxel-record array (multigpu.record_layout) and, when
        scored, the counts image of one dose actor (what a multi-GPU sum-reduce has to touch)."""
        ptrs = (C.c_void_p * 2)()
        ns = (C.c_uint64 * 2)()
        nb = C.c_int32()
        self._check(self.L.gate_b200_dose_buffers(self.h, actor, 2, ptrs, ns, C.byref(nb)))
        return [(ptrs[i], ns[i]) for i in range(nb.value)]

    def mean_uncertainty(self, actor, n_events, top_voxels, threshold):
        out = C.c_double()
        self._check(self.L.gate_b200_mean_uncertainty(self.h, actor, C.c_uint64(int(n_events)), int(top_voxels), float(threshold), C.byref(out)))
        return out.value

    def phsp(self, actor):
        n = C.c_uint64()
        self._check(self.L.gate_b200_phsp_count(self.h, actor, C.byref(n)))
        n = n.value
        out = {"KineticEnergy": np.zeros(n, np.float32), "Position": np.zeros((n, 3), np.float32),
               "Direction": np.zeros((n, 3), np.float32), "Weight": np.zeros(n, np.float32),
               "Time": np.zeros(n, np.float32), "EventID": np.zeros(n, np.int64),
               "Flags": np.zeros(n, np.int32)}
        cols = abi.PhspColumns()
        fp = C.POINTER(C.c_float)
        cols.kinetic_energy = out["KineticEnergy"].ctypes.data_as(fp)
        cols.position = out["Position"].ctypes.data_as(fp)
        cols.direction = out["Direction"].ctypes.data_as(fp)
        cols.weight = out["Weight"].ctypes.data_as(fp)
        cols.time_ns = out["Time"].ctypes.data_as(fp)
        cols.event_id = out["EventID"].ctypes.data_as(C.POINTER(C.c_int64))
        cols.flags = out["Flags"].ctypes.data_as(C.POINTER(C.c_int32))
        got = C.c_uint64()
        self._check(self.L.gate_b200_get_phsp(self.h, actor, C.c_uint64(n), C.byref(cols), C.byref(got)))
        return out

    def kernel_name(self):
        """instantiation of the most recent transport launch, e.g. 'gb2::transport_kernel<unsigned char, 2>'"""
        return self.L.gate_b200_kernel_name(self.h).decode()

    def stats(self):
        s = abi.Stats()
        self._check(self.L.gate_b200_get_stats(self.h, C.byref(s)))
        return s.as_dict()
