#!/usr/bin/env python
"""Generate tests/golden/*.json|npz by RUNNING the reference's own Python (read-only, from
/root/reference) for the host-side algorithms on the hot path.  Runs only in the build
container (the GPU box has no /root/reference); the fixtures it writes are committed.

The reference package cannot be imported as a whole here (`opengate_core`, `itk`, `box`,
`SimpleITK`, `colored`, `loguru` ... are absent).  The functions used below are
pure Python/numpy; they are imported from their real source files with the missing binary
dependencies replaced by inert stub modules (nothing of the stubs is ever executed by the
functions we call, except `box.Box` = attribute dict and the itk array<->image identity
shims used by `ImageVolume.create_label_image`).

    python tools/gen_golden_from_reference.py [--reference /root/reference]
"""
import argparse
import importlib
import json
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            # python-box exposes keys such as "g/cm3" as the attribute g_cm3
            for key in self:
                if isinstance(key, str) and re.sub(r"\W", "_", key) == k:
                    return self[key]
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _DummyMeta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _dummy(k)


def _dummy(name):
    return _DummyMeta(name, (), {"__init__": lambda self, *a, **kw: None, "__call__": lambda self, *a, **kw: None})


class _Anything(types.ModuleType):
    """Module stub: any attribute is a dummy class/callable."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        d = _dummy(k)
        setattr(self, k, d)
        return d


class _FakeItkImage:
    def __init__(self, arr):
        self.arr = arr

    def CopyInformation(self, other):
        pass


# third-party modules the reference imports at module level that are absent here
STUBBED = {"matplotlib", "awkward", "meshio", "pyvista", "stl", "radioactivedecay", "gatetools", "garf", "gaga_phsp", "icrp107_database",
           "requests", "jsonpickle", "git", "pint", "numpy_stl", "trimesh", "psutil_", "tomlkit", "xmltodict"}


class _StubFinder:
    """Last-resort finder: any top-level module that does not exist becomes an inert stub."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if name.split(".")[0] not in STUBBED:
            return None
        import importlib.machinery

        class L:
            @staticmethod
            def create_module(spec):
                m = _Anything(spec.name)
                m.__path__ = []
                return m

            @staticmethod
            def exec_module(mod):
                pass

        return importlib.machinery.ModuleSpec(name, L(), is_package=True)


def install_stubs(ref):
    sys.meta_path.append(_StubFinder)
    for name in ("opengate_core", "colored", "SimpleITK", "anytree", "uproot", "loguru", "tqdm", "click"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["loguru"].logger = _Anything("logger")
    sys.modules["colored"].fore = lambda *a, **k: ""
    sys.modules["colored"].style = lambda *a, **k: ""
    box = types.ModuleType("box")
    box.Box = _AttrDict
    box.BoxList = list
    sys.modules["box"] = box
    itk = _Anything("itk")
    itk.array_view_from_image = lambda im: im.arr if isinstance(im, _FakeItkImage) else im
    itk.image_from_array = lambda a: _FakeItkImage(a)
    itk.array_from_image = lambda im: np.array(im.arr if isinstance(im, _FakeItkImage) else im)
    sys.modules["itk"] = itk
    # a bare `opengate` package object: sub-modules are loaded from the real files, the
    # package __init__ (which imports the compiled core) is not executed
    pkg = types.ModuleType("opengate")
    pkg.__path__ = [os.path.join(ref, "opengate")]
    pkg.get_site_packages_dir = lambda: ""
    sys.modules["opengate"] = pkg
    for sub in ("geometry", "data", "actors", "sources"):
        m = types.ModuleType(f"opengate.{sub}")
        m.__path__ = [os.path.join(ref, "opengate", sub)]
        sys.modules[f"opengate.{sub}"] = m


def jsonable(x):
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    return x


def gen_schneider(ref, out):
    mats = importlib.import_module("opengate.geometry.materials")
    util = importlib.import_module("opengate.utility")
    f_mat = os.path.join(ref, "opengate", "data", "Schneider2000MaterialsTable.txt")
    f_den = os.path.join(ref, "opengate", "data", "Schneider2000DensitiesTable.txt")
    res = {"g_cm3": util.g4_units.g_cm3, "Gy": util.g4_units.Gy, "units": {k: getattr(util.g4_units, k) for k in
           ("mm", "cm", "m", "keV", "MeV", "deg", "Bq", "s", "ns", "g", "kg", "joule", "g_cm3", "Gy")}}
    for tol in (0.05, 0.1, 0.01):
        added = []

        class DB:
            def add_material_weights(self, name, elems, weights, density):
                added.append([name, list(elems), [float(w) for w in weights], float(density)])

        sim = _AttrDict(volume_manager=_AttrDict(material_database=DB()))
        vm, created = mats.HounsfieldUnit_to_material(sim, tol * util.g4_units.g_cm3, f_mat, f_den)
        key = f"tol_{tol}"
        if tol == 0.01:
            res[key] = {"n": len(vm), "first": jsonable(vm[0]), "last": jsonable(vm[-1])}
        else:
            res[key] = {"voxel_materials": jsonable(vm), "created": created, "materials": jsonable(added)}
    # density interpolation probes
    dens = mats.HU_read_density_table(f_den)
    hus = [-1200, -1000, -999.5, -98, -97.9, 0, 14, 23, 100, 101, 1600, 1601, 2999, 3000, 5000]
    res["density_probe"] = {"hu": hus, "rho": [mats.HU_linear_interpolate_densities(h, dens) for h in hus]}
    json.dump(res, open(os.path.join(out, "schneider_reference.json"), "w"), indent=0)
    print("schneider:", {k: (v["n"] if "n" in v else len(v["voxel_materials"])) for k, v in res.items() if k.startswith("tol_")})


def gen_label_image(ref, out):
    vols = None
    try:
        vols = importlib.import_module("opengate.geometry.volumes")
    except Exception as e:  # pragma: no cover
        print("volumes.py not importable with stubs:", repr(e)[:200])
        return
    rng = np.random.default_rng(20260101)
    hu = rng.integers(-1100, 3200, size=(6, 7, 8)).astype(np.int16)
    cases = {
        "contiguous": [[-2000, -900, "G4_AIR"], [-900, 100, "Lung"], [100, 4000, "Bone"]],
        "gaps": [[-900, -300, "Lung"], [-100, 300, "Water"], [300, 1000, "Bone"], [1000, 1500, "Water"]],
        "unsorted_repeat": [[300, 3100, "Bone"], [-2000, -500, "G4_AIR"], [-500, 300, "Water"]],
    }
    res = {"hu": hu.tolist(), "cases": {}}
    for name, vm in cases.items():
        fake = _AttrDict(voxel_materials=vm, material="G4_AIR", name="vol", itk_image=None, material_to_label_lut=None)
        fake.create_material_to_label_lut = lambda f=fake: vols.ImageVolume.create_material_to_label_lut(f)
        lut = vols.ImageVolume.create_material_to_label_lut(fake)
        fake.material_to_label_lut = lut
        lab = vols.ImageVolume.create_label_image(fake, itk_image=_FakeItkImage(hu))
        res["cases"][name] = {"voxel_materials": vm, "lut": lut, "labels": np.asarray(lab.arr).tolist(),
                              "dtype": str(np.asarray(lab.arr).dtype)}
    json.dump(res, open(os.path.join(out, "label_image_reference.json"), "w"))
    print("label image cases:", list(res["cases"]))


def gen_scoring(ref, out):
    util = importlib.import_module("opengate.utility")
    rng = np.random.default_rng(7)
    n = 1000
    x = rng.gamma(2.0, 1.0, size=(3, 4, 5)) * n * 0.1
    x2 = (x / n) ** 2 * n * (1.0 + rng.random(x.shape))
    x[0, 0, 0] = 0.0
    var = util.calculate_variance(x, x2, n)
    res = {"value": x.tolist(), "squared": x2.tolist(), "n": n, "variance": np.asarray(var).tolist()}
    # 3-D CDF of the voxel source
    try:
        img = importlib.import_module("opengate.image")
        a = rng.random((4, 5, 6)).astype(np.float32)
        a[1, 2, :] = 0.0
        fake = _FakeItkImage(a)
        cdf_x, cdf_y, cdf_z = img.compute_image_3D_CDF(fake)
        res["cdf"] = {"activity": a.tolist(), "cdf_x": np.asarray(cdf_x).tolist(), "cdf_y": np.asarray(cdf_y).tolist(),
                      "cdf_z": np.asarray(cdf_z).tolist()}
    except Exception as e:  # pragma: no cover
        print("image.py not importable with stubs:", repr(e)[:300])
    json.dump(res, open(os.path.join(out, "scoring_reference.json"), "w"))
    print("scoring: variance + cdf", "cdf" in res)


def gen_nist(ref, out):
    pa = importlib.import_module("opengate.data.PhotonAttenuation")
    e = np.array([0.001, 0.0015, 0.004, 0.01, 0.02, 0.03, 0.06, 0.08, 0.1405, 0.2, 0.5, 0.662, 1.0, 1.25, 2.0, 6.0, 10.0, 20.0])
    res = {"energy_MeV": e.tolist(), "mac": {}, "meac": {}}
    res["api"] = "PhotonAttenuationEl(Z, E, 1) = mu/rho, (Z, E, 2) = mu_en/rho [cm2/g], log-log PCHIP + edges"
    res["grid_energy_MeV"] = np.asarray(pa.energy).reshape(-1).tolist()
    for z in (1, 6, 7, 8, 11, 12, 13, 15, 16, 17, 18, 19, 20, 22, 26, 53, 74, 82):
        res["mac"][str(z)] = [float(np.asarray(pa.PhotonAttenuationEl(z, float(x), 1)[0]).reshape(-1)[0]) for x in e]
        res["meac"][str(z)] = [float(np.asarray(pa.PhotonAttenuationEl(z, float(x), 2)[0]).reshape(-1)[0]) for x in e]
    json.dump(res, open(os.path.join(out, "nist_reference.json"), "w"))
    print("nist: water-ish check O mac(1MeV)=", res["mac"]["8"][12])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    install_stubs(a.reference)
    gen_schneider(a.reference, OUT)
    gen_label_image(a.reference, OUT)
    gen_scoring(a.reference, OUT)
    gen_nist(a.reference, OUT)


if __name__ == "__main__":
    main()
