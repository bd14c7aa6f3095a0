"""Materials for the GPU engine: element table, GATE material database files, Schneider
HU -> material conversion, label images.

Mirrors, for the photon-in-image-volume path only, the reference's
`opengate/geometry/materials.py` (HounsfieldUnit_to_material :149-230, density helpers
:116-146, MaterialDatabase :726-941) and `opengate/geometry/volumes.py`
(create_material_to_label_lut :1144-1168, create_label_image :1200-1243).  Materials are
resolved to `physics.MaterialSpec` (density + element mass fractions) instead of G4Material.
"""

from __future__ import annotations

import os

import numpy as np

from . import physics_models as pm
from .exception import fatal
from .physics import MaterialSpec
from .units import g4_units

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
SCHNEIDER_MATERIALS = os.path.join(DATA_DIR, "schneider2000_materials.txt")
SCHNEIDER_DENSITIES = os.path.join(DATA_DIR, "schneider2000_densities.txt")
DEFAULT_DB = os.path.join(DATA_DIR, "materials_b200.db")

# element name (as written in GATE database / Schneider files) -> symbol
_ALIASES = {"Phosphor": "P", "Phosphorus": "P", "Sulfur": "S", "Sulphur": "S", "Aluminium": "Al",
            "Aluminum": "Al", "Cesium": "Cs", "Caesium": "Cs", "Wolfram": "W"}


def element_z(name_or_symbol: str) -> int:
    d = pm.nist()
    s = str(name_or_symbol)
    if s in _ALIASES:
        s = _ALIASES[s]
    syms = list(d["symbol"])
    if s in syms:
        return syms.index(s) + 1
    names = list(d["name"])
    if s in names:
        return names.index(s) + 1
    low = [n.lower() for n in names]
    if s.lower() in low:
        return low.index(s.lower()) + 1
    fatal(f"Unknown element '{name_or_symbol}'")


# A handful of Geant4-NIST materials by name (G4NistMaterialBuilder compositions, mass fractions).
NIST_MATERIALS = {
    "G4_WATER": (1.0, {1: 0.111894, 8: 0.888106}),
    "G4_AIR": (0.00120479, {6: 0.000124, 7: 0.755268, 8: 0.231781, 18: 0.012827}),
    "G4_Galactic": (1.0e-25, {1: 1.0}),
    "G4_LUNG_ICRP": (1.04, {1: 0.105, 6: 0.083, 7: 0.023, 8: 0.779, 11: 0.002, 15: 0.001, 16: 0.002,
                            17: 0.003, 19: 0.002}),
    "G4_BONE_CORTICAL_ICRP": (1.92, {1: 0.034, 6: 0.155, 7: 0.042, 8: 0.435, 11: 0.001, 12: 0.002,
                                     15: 0.103, 16: 0.003, 20: 0.225}),
    "G4_BONE_COMPACT_ICRU": (1.85, {1: 0.064, 6: 0.278, 7: 0.027, 8: 0.41, 12: 0.002, 15: 0.07,
                                    16: 0.002, 20: 0.147}),
    "G4_TISSUE_SOFT_ICRP": (1.03, {1: 0.105, 6: 0.256, 7: 0.027, 8: 0.602, 11: 0.001, 15: 0.002,
                                   16: 0.003, 17: 0.002, 19: 0.002}),
    "G4_ADIPOSE_TISSUE_ICRP": (0.95, {1: 0.114, 6: 0.598, 7: 0.007, 8: 0.278, 11: 0.001, 16: 0.001,
                                      17: 0.001}),
    "G4_Al": (2.699, {13: 1.0}),
    "G4_Ti": (4.54, {22: 1.0}),
    "G4_Fe": (7.874, {26: 1.0}),
    "G4_Cu": (8.96, {29: 1.0}),
    "G4_Pb": (11.35, {82: 1.0}),
    "G4_W": (19.3, {74: 1.0}),
    "G4_PLEXIGLASS": (1.19, {1: 0.080538, 6: 0.599848, 8: 0.319614}),
    "G4_POLYETHYLENE": (0.94, {1: 0.143711, 6: 0.856289}),
}


class MaterialDatabase:
    """Name -> MaterialSpec. Sources, in lookup order: materials added by weights
    (`add_material_weights`, as HounsfieldUnit_to_material does), materials added by number of
    atoms, database files in GATE .db syntax, the built-in NIST subset."""

    def __init__(self):
        self.filenames = []
        self.new_materials_weights = {}
        self.new_materials_nb_atoms = {}
        self._elements = {}     # name -> (symbol, Z, A)
        self._db = {}           # name -> MaterialSpec
        self._db_by_file = {}

    # -- user API (names follow the reference) -----------------------------------------
    def add_material_weights(self, name, elems_symbol, weights, density):
        """density in Geant4 internal units (g4_units.g_cm3 multiplied in), like the reference."""
        self.new_materials_weights[name] = (name, list(elems_symbol), list(weights), density)

    def add_material_nb_atoms(self, name, elems_symbol, nb_atoms, density):
        self.new_materials_nb_atoms[name] = (name, list(elems_symbol), list(nb_atoms), density)

    def read_from_file(self, filename):
        filename = str(filename)
        if not os.path.exists(filename):
            fatal(f"Material database file not found: {filename}")
        self.filenames.append(filename)
        self._db_by_file[filename] = []
        section = None
        with open(filename) as f:
            lines = [ln.strip().replace("\t", " ") for ln in f]
        i = 0
        while i < len(lines):
            line = lines[i]
            i += 1
            if not line or line[0] == "#":
                continue
            w = line.split()[0]
            if w in ("[Isotopes]", "[Elements]", "[Materials]"):
                section = w
                continue
            if section is None:
                fatal(f"Error while reading the file {filename}: it must start with "
                      f"[Isotopes], [Elements] or [Materials]")
            if section == "[Elements]":
                if "Z=" in line.replace(" ", ""):
                    name, rest = line.split(":", 1)
                    tags = _parse_tags(rest)
                    self._elements[name.strip()] = (tags["S"], int(float(tags["Z"])), _num(tags["A"]))
                # isotopic element definitions (n=...) are skipped: natural abundance is used
                elif "n=" in line.replace(" ", ""):
                    n = int(float(_parse_tags(line.split(":", 1)[1])["n"]))
                    i += n
            elif section == "[Materials]":
                name, rest = line.split(":", 1)
                tags = _parse_tags(rest)
                dval, dunit = tags["d"].split()[:2]
                density = float(dval) * g4_units[dunit] / g4_units.g_cm3
                n = int(float(tags["n"]))
                comps = []
                while len(comps) < n and i < len(lines):
                    cl = lines[i]
                    i += 1
                    if not cl or cl[0] == "#":
                        continue
                    if not (cl.startswith("+el:") or cl.startswith("+mat:")):
                        fatal(f"Error in {filename}: expected '+el:' or '+mat:' in material {name}, got '{cl}'")
                    kind, crest = cl.split(":", 1)
                    ctags = _parse_tags(crest)
                    comps.append((kind, ctags))
                self._db[name.strip()] = ("raw", density, comps)
                self._db_by_file[filename].append(name.strip())

    def get_database_material_names(self, db=None):
        if db is None:
            return list(self._db.keys())
        if db not in self._db_by_file:
            fatal(f"The database '{db}' is not in the list of read database: {self.filenames}")
        return list(self._db_by_file[db])

    # -- resolution ---------------------------------------------------------------------
    def _element_z_a(self, name):
        if name in self._elements:
            _, z, a = self._elements[name]
            return z, a
        z = element_z(name)
        return z, pm.atomic_mass(z)

    def find_or_build_material(self, name) -> MaterialSpec:
        if name in self.new_materials_weights:
            _, syms, wts, dens = self.new_materials_weights[name]
            fr = {}
            for s, w in zip(syms, wts):
                z = element_z(s)
                fr[z] = fr.get(z, 0.0) + float(w)
            return MaterialSpec(name, dens / g4_units.g_cm3, fr)
        if name in self.new_materials_nb_atoms:
            _, syms, nb, dens = self.new_materials_nb_atoms[name]
            fr = {}
            for s, n in zip(syms, nb):
                z = element_z(s)
                fr[z] = fr.get(z, 0.0) + float(n) * pm.atomic_mass(z)
            return MaterialSpec(name, dens / g4_units.g_cm3, fr)
        if name in self._db:
            entry = self._db[name]
            if isinstance(entry, MaterialSpec):
                return entry
            _, density, comps = entry
            fr = {}
            by_atoms = any("n" in t and "f" not in t for _, t in comps)
            for kind, t in comps:
                if kind == "+el":
                    z, a = self._element_z_a(t["name"])
                    w = float(t["n"]) * a if by_atoms else float(t["f"])
                    fr[z] = fr.get(z, 0.0) + w
                else:  # +mat: sub-material by mass fraction
                    sub = self.find_or_build_material(t["name"])
                    for z, w in sub.normalised().items():
                        fr[z] = fr.get(z, 0.0) + float(t["f"]) * w
            spec = MaterialSpec(name, density, fr)
            self._db[name] = spec
            return spec
        if name in NIST_MATERIALS:
            d, fr = NIST_MATERIALS[name]
            return MaterialSpec(name, d, dict(fr))
        fatal(f'Cannot find nor build material named "{name}"')

    FindOrBuildMaterial = find_or_build_material


def _num(s):
    return float(s.split()[0])


def _parse_tags(s):
    out = {}
    for part in s.split(";"):
        if "=" in part:
            k, v = part.split("=", 1)
            out[k.strip()] = v.strip()
    return out


# -------------------------------------------------------------------------------------------
# Schneider 2000: HU -> materials (reference materials.py:61-230)
# -------------------------------------------------------------------------------------------

def HU_read_materials_table(file_mat):
    elements = ["HU"]
    rows = []
    section = None
    with open(os.path.abspath(file_mat)) as f:
        for line in f:
            words = line.split()
            row = None
            for i, word in enumerate(words):
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
                elif section == "table":
                    if row is None:
                        row = {}
                    key = elements[i]
                    if i == 0:
                        row[key] = int(word)
                    elif i == len(elements) - 1:
                        row[key] = word
                    else:
                        row[key] = float(word)
            if row:
                rows.append(row)
    return rows, elements


def HU_read_density_table(file_density):
    out = []
    with open(os.path.abspath(file_density)) as f:
        for line in f:
            w = line.split()
            if not w or w[0][0] == "#":
                continue
            out.append({"HU": float(w[0]), "density": float(w[1])})
    return out


def HU_linear_interpolate_densities(hu, densities):
    n = len(densities)
    i = 0
    while i < n and hu > densities[i]["HU"]:
        i += 1
    i -= 1
    if i < 0:
        return densities[0]["density"]
    if i >= n - 1:
        return densities[n - 1]["density"]
    lo, hi = densities[i], densities[i + 1]
    return ((hu - lo["HU"]) / (hi["HU"] - lo["HU"])) * (hi["density"] - lo["density"]) + lo["density"]


def HU_find_max_density_difference(hu_min, hu_max, d_min, d_max, densities):
    n = len(densities)
    i = 0
    while i < n and hu_min > densities[i]["HU"]:
        i += 1
    j = 0
    while j < n and hu_max > densities[j]["HU"]:
        j += 1
    j -= 1
    for x in range(i, j, 1):
        d_min = min(d_min, densities[x]["density"])
        d_max = max(d_max, densities[x]["density"])
    return d_max - d_min


def HounsfieldUnit_to_material(simulation, density_tolerance, file_mat=None, file_density=None):
    """Same signature and result as the reference function: returns (voxel_materials,
    created_material_names) and registers every material with
    simulation.volume_manager.material_database.add_material_weights.
    `density_tolerance` carries Geant4 units (e.g. 0.05 * g4_units.g_cm3)."""
    file_mat = file_mat or SCHNEIDER_MATERIALS
    file_density = file_density or SCHNEIDER_DENSITIES
    materials, elements = HU_read_materials_table(file_mat)
    densities = HU_read_density_table(file_density)
    gcm3 = g4_units.g_cm3
    elems = elements[1:-1]
    symbols = [str(pm.nist()["symbol"][element_z(x) - 1]) for x in elems]
    db = simulation.volume_manager.material_database
    voxel_materials, created = [], []
    num = 0
    last_i = len(materials) - 1
    for i, mat in enumerate(materials):
        hu_min = mat["HU"]
        hu_max = hu_min + 1 if i == last_i else materials[i + 1]["HU"]
        if hu_max <= hu_min:
            fatal(f"Error, HU interval not valid: {mat}")
        dmin = HU_linear_interpolate_densities(hu_min, densities)
        dmax = HU_linear_interpolate_densities(hu_max, densities)
        ddiff = HU_find_max_density_difference(hu_min, hu_max, dmin, dmax, densities)
        n = max(1, ddiff * gcm3 / density_tolerance)
        if "Air" in mat["name"] or "AIR" in mat["name"]:
            n = 1
        htol = (hu_max - hu_min) / n
        if n != 1:
            n = int(np.ceil(n))
        for j in range(n):
            h1 = hu_min + j * htol
            h2 = min(hu_min + (j + 1) * htol, hu_max)
            d = HU_linear_interpolate_densities(h1 + (h2 - h1) / 2.0, densities)
            pairs = [(s, mat[e]) for s, e in zip(symbols, elems) if mat[e] > 0]
            total = sum(w for _, w in pairs)
            name = f'{mat["name"]}_{num}'
            db.add_material_weights(name, [s for s, _ in pairs], [w / total for _, w in pairs], d * gcm3)
            voxel_materials.append([h1, h2, name])
            created.append(name)
            num += 1
    return voxel_materials, created


# -------------------------------------------------------------------------------------------
# label image (reference volumes.py:1144-1243)
# -------------------------------------------------------------------------------------------

def create_material_to_label_lut(material, voxel_materials, volume_name="?"):
    lut = {material: 0}
    rows = sorted(voxel_materials, key=lambda x: x[0])
    lower = np.array([r[0] for r in rows])
    upper = np.array([r[1] for r in rows])
    if not (lower[1:] >= upper[:-1]).all():
        fatal(f"Overlapping intervals in voxel_materials of volume {volume_name}.")
    nxt = 1
    for r in rows:
        if r[2] not in lut:
            lut[r[2]] = nxt
            nxt += 1
    return lut


def create_label_array(hu_array, material, voxel_materials, lut=None):
    """HU array (numpy z,y,x) -> uint16 labels with np.digitize on the interval bounds:
    [lo, hi) intervals, values outside every interval (or in a gap) get label 0."""
    if lut is None:
        lut = create_material_to_label_lut(material, voxel_materials)
    bins_sorted, labels_sorted = label_bins(voxel_materials, lut)
    arr = np.array(labels_sorted, dtype=np.ushort)[np.digitize(hu_array, bins=bins_sorted)]
    return arr, lut


def label_bins(voxel_materials, lut):
    """(sorted bin edges, label per np.digitize index) of the label rule above: one table of
    len(bins) + 1 labels.  The GPU engine sends it with the HU image and does the digitize on the
    device (gate_b200_set_image_volume_hu)."""
    rows = sorted(voxel_materials, key=lambda x: x[0])
    bins = [r[0] for r in rows]
    gaps = set(r[1] for r in rows).difference(bins)
    bins.extend(gaps)
    labels = [lut[r[2]] for r in rows] + len(gaps) * [0]
    order = sorted(zip(bins, labels), key=lambda p: p[0])
    bins_sorted = [b for b, _ in order]
    labels_sorted = [0] + [l for _, l in order]
    return bins_sorted, labels_sorted
