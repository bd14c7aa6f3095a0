"""Volumes the GPU engine understands: the world Box, Box daughters (water box of config 1,
phase-space planes) and ONE ImageVolume (voxelised CT).

Host-side mirror of the reference's `opengate/geometry/volumes.py` for these types only
(`BoxVolume`, `ImageVolume` :882-1347): same user parameters, same label-image rule, same
placement convention (`p_mother = R p_local + t`, composition up the tree as in
`geometry/utility.py:178-215` / `GateHelpersGeometry.cpp:16-49`).  No Geant4 objects are built.
"""

from __future__ import annotations

import json

import numpy as np

from . import image as gimage
from . import materials as gmat
from .base import GateObject
from .exception import fatal
from .units import g4_units

__world_name__ = "world"


def _as_rotation(r):
    a = np.asarray(r, dtype=np.float64)
    if a.shape != (3, 3):
        fatal(f"rotation must be a 3x3 matrix, got shape {a.shape}")
    return a


class VolumeBase(GateObject):
    user_info_defaults = {
        "mother": (__world_name__, {"doc": "Name of the mother volume"}),
        "material": ("G4_AIR", {"doc": "Name of the material"}),
        "translation": ([0.0, 0.0, 0.0], {"doc": "Translation with respect to the mother volume"}),
        "rotation": (np.eye(3), {"doc": "Rotation matrix with respect to the mother volume"}),
        "color": ([1, 1, 1, 1], {"doc": "unused by the GPU engine"}),
    }
    volume_type = "VolumeBase"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.volume_manager = None

    @property
    def mother_volume(self):
        if self.name == __world_name__ or self.volume_manager is None:
            return None
        return self.volume_manager.get_volume(self.mother)

    def world_transform(self):
        """(R_cum, t_cum): local -> world, composed up the tree
        (R_cum = R_mother R, t_cum = R_mother t + t_mother)."""
        R = _as_rotation(self.rotation)
        t = np.asarray(self.translation, dtype=np.float64)
        v = self.mother_volume
        while v is not None:
            Rm, tm = _as_rotation(v.rotation), np.asarray(v.translation, dtype=np.float64)
            t = Rm @ t + tm
            R = Rm @ R
            v = v.mother_volume
        return R, t


class BoxVolume(VolumeBase):
    user_info_defaults = {"size": ([10 * g4_units.cm] * 3, {"doc": "Full size of the box along x, y, z"})}
    volume_type = "BoxVolume"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._freeze()

    @property
    def half_size_mm(self):
        return 0.5 * np.asarray(self.size, dtype=np.float64)


class ImageVolume(VolumeBase):
    """Voxelised volume: an HU (or label) image + `voxel_materials` intervals -> label image.
    The box is centred on its placement; the input image's own origin is NOT used for
    placement (reference volumes.py:1060-1061, GateImageNestedParameterisation.cpp:19-31)."""

    user_info_defaults = {
        "voxel_materials": ([[-np.inf, np.inf, "G4_AIR"]], {"doc": "List of [HU_low, HU_high, material]"}),
        "image": ("", {"doc": "Path to the image file (.mhd/.mha/.npy) or an opengate_b200.image.Image"}),
        "dump_label_image": (None, {"doc": "Path at which the label image should be saved"}),
    }
    volume_type = "ImageVolume"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.material_to_label_lut = None
        self._image = None
        self.label_image = None
        self._freeze()

    # -- input image ------------------------------------------------------------------------
    def load_input_image(self, path=None):
        src = self.image if path is None else path
        if isinstance(src, gimage.Image):
            img = src
        elif isinstance(src, np.ndarray):
            img = gimage.Image(src)
        else:
            if not src:
                fatal(f"ImageVolume '{self.name}': no image set")
            img = gimage.read_image(str(src))
        if path is None:
            self._image = img
        return img

    @property
    def itk_image(self):
        if self._image is None:
            self.load_input_image()
        return self._image

    @property
    def size_pix(self):
        return self.itk_image.size.astype(int)

    @property
    def spacing(self):
        return self.itk_image.spacing.copy()

    @property
    def half_size_mm(self):
        return 0.5 * self.size_pix * self.spacing

    @property
    def native_translation(self):
        img = self.itk_image
        center = (img.size - 1.0) * img.spacing / 2.0
        return img.origin + img.direction @ center

    @property
    def native_rotation(self):
        return self.itk_image.direction.copy()

    # -- materials --------------------------------------------------------------------------
    def set_materials_from_voxelisation(self, labels_json):
        with open(labels_json) as f:
            labels = json.load(f)
        self.voxel_materials = [[l["label"], l["label"] + 1, l["material"]] for l in labels.values()]

    def read_label_to_material(self, labels_filename):
        with open(labels_filename) as f:
            self.voxel_materials = json.load(f)

    def create_material_to_label_lut(self, material=None, voxel_materials=None):
        return gmat.create_material_to_label_lut(self.material if material is None else material,
                                                 self.voxel_materials if voxel_materials is None else voxel_materials,
                                                 self.name)

    def create_label_image(self, image=None):
        img = self.itk_image if image is None else image
        if self.material_to_label_lut is None:
            self.material_to_label_lut = self.create_material_to_label_lut()
        arr, _ = gmat.create_label_array(img.array, self.material, self.voxel_materials, self.material_to_label_lut)
        lab = gimage.Image(arr, img.spacing, img.origin, img.direction)
        return lab

    def construct(self):
        """What `ImageVolume.construct` does up to the Geant4 objects (volumes.py:1049-1067)."""
        self.material_to_label_lut = self.create_material_to_label_lut()
        self.load_input_image()
        self.label_image = self.create_label_image()
        if self.dump_label_image:
            self.save_label_image()
        return self.label_image

    def save_label_image(self, path=None):
        p = self.dump_label_image if path is None else path
        if self.volume_manager is not None and self.volume_manager.simulation is not None:
            p = self.volume_manager.simulation.get_output_path(p)
        if self.label_image is None:
            self.construct()
        gimage.write_image(self.label_image, str(p))
        with open(str(p).rsplit(".", 1)[0] + ".json", "w") as f:
            json.dump(self.material_to_label_lut, f)

    def create_density_image(self):
        """Density (g/cm3, float32) per voxel, reference materials.py:294-326."""
        if self.label_image is None:
            self.construct()
        db = self.volume_manager.material_database
        rho = np.zeros(len(self.material_to_label_lut), dtype=np.float32)
        for name, lab in self.material_to_label_lut.items():
            rho[lab] = db.find_or_build_material(name).density_g_cm3
        img = gimage.Image(rho[self.label_image.array], self.label_image.spacing, self.label_image.origin,
                           self.label_image.direction)
        return img


volume_types = {"Box": BoxVolume, "BoxVolume": BoxVolume, "Image": ImageVolume, "ImageVolume": ImageVolume}
