"""Synthetic phantoms, spectra and ABI-level run descriptions for the BASELINE.json configs.

`LowLevelSetup` is a complete description of one run at the level of the C-ABI (tables, label
volume, sources, actors).  `apply(engine)` replays it on anything that offers the Engine call
sequence — the CUDA engine (`core.Engine`) or, in tests, the CPU oracle — so both sides are
fed bit-identical inputs.  The configs follow SURVEY.md section 8(d) / BASELINE.md section 4.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import abi, materials, physics
from . import physics_models as pm
from .units import g4_units


# ------------------------------------------------------------------------------- spectra

def spectrum_6mv(n_bins=24):
    """Fixed tabulated 6 MV linac photon spectrum (synthetic, bremsstrahlung-like shape:
    rises from 0.25 MeV, peaks near 1 MeV, falls to 0 at 6 MeV). Histogram edges in MeV and
    weights per unit energy (the reference's `spectrum_histogram` convention,
    GateGenericSource.cpp:583-644)."""
    edges = np.linspace(0.25, 6.0, n_bins)
    mid = 0.5 * (edges[1:] + edges[:-1])
    w = mid ** 0.9 * np.exp(-mid / 1.05) * (1.0 - mid / 6.3)
    return edges, w / w.sum()


def spectrum_120kvp(bin_kev=2.0):
    """Fixed tabulated 120 kVp tungsten-anode spectrum: Kramers continuum filtered by 2.5 mm Al
    (+ inherent 1 mm Al) with the W K-alpha/K-beta lines. Edges in MeV, weights per MeV."""
    e_kev = np.arange(10.0, 120.0 + 1e-9, bin_kev)
    mid = 0.5 * (e_kev[1:] + e_kev[:-1])
    kram = (120.0 - mid) / mid
    mu_al = pm.nist_total_cm2_g(13, mid * 1e-3) * 2.699  # 1/cm
    w = kram * np.exp(-mu_al * 0.35)
    for line, rel in ((58.0, 0.6), (59.3, 1.0), (67.2, 0.35), (69.1, 0.1)):
        k = np.searchsorted(e_kev, line) - 1
        w[k] += rel * 0.12 * w.max()
    return e_kev * 1e-3, w / w.sum()


def histogram_cdf(edges, weights):
    """Cumulative weights of `spectrum_histogram` without interpolation:
    accumulate w_i * dE_i, normalise (GateGenericSource.cpp:606-612, 633-637)."""
    acc = np.cumsum(np.asarray(weights, float) * np.diff(np.asarray(edges, float)))
    return acc / acc[-1]


# ------------------------------------------------------------------------------- phantoms

def slab_phantom_hu(n=256):
    """C2: z-slabs water [0,n/4) / lung [n/4,n/2) / bone [n/2,5n/8) / water, int16 HU."""
    hu = np.zeros((n, n, n), dtype=np.int16)
    hu[n // 4: n // 2] = -740
    hu[n // 2: (5 * n) // 8] = 1000
    return hu


SLAB_VOXEL_MATERIALS = [[-2000, -900, "Air"], [-900, -300, "Lung"], [-300, 300, "Water"], [300, 4000, "RibBone"]]


def synthetic_ct_hu(nx=512, ny=512, nz=300, seed=987654321):
    """C3: air -1000, elliptical body +40 (+-20 noise), two lung ellipsoids -750, spine and rib
    cylinders +300..+1200. Returns int16 (z,y,x)."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, nz), np.linspace(-1, 1, ny), np.linspace(-1, 1, nx),
                          indexing="ij", sparse=True)
    hu = np.full((nz, ny, nx), -1000, dtype=np.int16)
    body = (x / 0.82) ** 2 + (y / 0.62) ** 2 < 1.0
    body = np.broadcast_to(body, hu.shape)
    noise = rng.integers(-20, 21, size=hu.shape, dtype=np.int16)
    hu = np.where(body, (40 + noise).astype(np.int16), hu)
    for cx in (-0.38, 0.38):
        lung = ((x - cx) / 0.27) ** 2 + ((y + 0.05) / 0.36) ** 2 + (z / 0.8) ** 2 < 1.0
        hu = np.where(lung, np.int16(-750) + (noise // 2), hu)
    spine = (x / 0.09) ** 2 + ((y - 0.36) / 0.09) ** 2 < 1.0
    hu = np.where(np.broadcast_to(spine, hu.shape), np.int16(900), hu)
    cord = (x / 0.03) ** 2 + ((y - 0.36) / 0.03) ** 2 < 1.0
    hu = np.where(np.broadcast_to(cord, hu.shape), np.int16(30), hu)
    # ribs: thin elliptical shell segments repeated along z
    shell = np.abs(np.sqrt((x / 0.72) ** 2 + (y / 0.52) ** 2) - 1.0) < 0.035
    stripes = (np.mod(z * 7.5, 1.0) < 0.35)
    ribs = np.broadcast_to(shell, hu.shape) & np.broadcast_to(stripes, hu.shape)
    hu = np.where(ribs, np.int16(300) + ((np.abs(noise) * 45) % 900).astype(np.int16), hu)
    return np.ascontiguousarray(hu.astype(np.int16))


def activity_image(n=128, seed=42):
    """C4: warm body + 3 hot spheres, float32 (z,y,x)."""
    g = np.linspace(-1, 1, n)
    z, y, x = np.meshgrid(g, g, g, indexing="ij", sparse=True)
    a = np.where((x / 0.8) ** 2 + (y / 0.6) ** 2 + (z / 0.9) ** 2 < 1.0, 1.0, 0.0)
    for (cx, cy, cz, r, v) in ((0.3, 0.1, 0.0, 0.12, 8.0), (-0.35, -0.1, 0.3, 0.08, 12.0), (0.0, 0.25, -0.4, 0.15, 5.0)):
        a = np.where((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 < r * r, v, a)
    return a.astype(np.float32)


# ------------------------------------------------------------------------------- setup object

@dataclass
class LowLevelSetup:
    name: str
    seed: int
    tables: physics.PhotonTables
    world_half: tuple
    world_material: int
    labels: np.ndarray            # (z,y,x) uint8/uint16
    spacing: tuple
    world_to_volume: np.ndarray   # 3x4
    sources: list = field(default_factory=list)       # (Source struct, keepalive)
    dose: list = field(default_factory=list)          # DoseDesc
    phsp: list = field(default_factory=list)          # PhspDesc
    n_primaries: int = 0
    tle: list = field(default_factory=list)           # (DoseDesc, TleDesc, keepalive): TLE dose actors
    materials: list = field(default_factory=list)     # MaterialSpec per label (for the mu_en database)
    # the volume as an HU image + the np.digitize table of ImageVolume.create_label_image (labels is then None
    # until someone asks for label_array()): the CUDA engine makes the label volume on the device
    hu: np.ndarray = None
    hu_bins: list = None
    hu_bin_labels: list = None
    label_bytes: int = 1
    regions: list = field(default_factory=list)       # RegionDesc: homogeneous Boxes of their own material in the world

    def label_array(self):
        """(z,y,x) label volume; made on the host from the HU image when it was not given."""
        if self.labels is None:
            lab = np.array(self.hu_bin_labels, dtype=np.ushort)[np.digitize(self.hu, bins=self.hu_bins)]
            self.labels = np.ascontiguousarray(lab.astype(np.uint8 if self.label_bytes == 1 else np.uint16))
        return self.labels

    def apply(self, eng):
        eng.set_tables(self.tables)
        eng.set_world(self.world_half, self.world_material)
        if self.hu is not None and self.labels is None and hasattr(eng, "set_image_volume_hu"):
            eng.set_image_volume_hu(self.hu, self.spacing, self.world_to_volume, self.hu_bins, self.hu_bin_labels,
                                    self.label_bytes)
        else:
            eng.set_image_volume(self.label_array(), self.spacing, self.world_to_volume)
        for r in self.regions:
            eng.add_box_region(r)
        sids = [eng.add_source(s, keep) for s, keep in self.sources]
        dids = [eng.add_dose_actor(d) for d in self.dose]
        dids += [eng.add_tle_dose_actor(d, t, keep) for d, t, keep in self.tle]
        pids = [eng.add_phsp_actor(p) for p in self.phsp]
        return sids, dids, pids

    def add_tle_actor(self, size, world_to_index, flags, hit_type=abi.HIT_RANDOM, threshold=float("inf"),
                      energy_min=0.0, database="EPDL"):
        """Attach a TLEDoseActor (GateTLEDoseActor): mu_en/rho per material from the reference's element
        database evaluated on the table's energy nodes (mu_handler)."""
        from . import mu_handler
        muen = mu_handler.mu_en_table_on_grid(self.materials, self.tables.energy, database)
        d = abi.make_dose_desc(size, world_to_index, flags, hit_type)
        t, keep = abi.make_tle_desc(muen, threshold, energy_min)
        self.tle.append((d, t, keep))
        return len(self.dose) + len(self.tle) - 1


def identity34():
    a = np.zeros((3, 4))
    a[:, :3] = np.eye(3)
    return a


def dose_world_to_index(size, spacing, rotation=None, translation=(0, 0, 0), vol_rotation=None, vol_translation=(0, 0, 0)):
    """world -> continuous index of a dose grid attached to a volume placed at
    (vol_rotation, vol_translation) with actor translation `translation`
    (GateHelpersImage.txx:20-58: origin = R_vol (-size*spacing/2 + spacing/2 + translation) + t_vol,
    direction = R_vol; index = D^-1 (p - origin) / spacing)."""
    size = np.asarray(size, float)
    spacing = np.asarray(spacing, float)
    R = np.eye(3) if vol_rotation is None else np.asarray(vol_rotation, float)
    t = np.asarray(vol_translation, float)
    origin = R @ (-size * spacing / 2.0 + spacing / 2.0 + np.asarray(translation, float)) + t
    A = np.zeros((3, 4))
    A[:, :3] = (R.T) / spacing[:, None]
    A[:, 3] = -(R.T @ origin) / spacing
    return A


def _materials_from_db(names, db=None):
    if db is None:
        db = materials.MaterialDatabase()
        db.read_from_file(materials.DEFAULT_DB)
    return [db.find_or_build_material(n) for n in names], db


def make_c1(n_primaries=1_000_000, flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ, hit_type=abi.HIT_POST,
            grid=(150, 150, 150), spacing_mm=2.0, seed=12345678, energy=1.0):
    """C1: 1 MeV pencil beam into a 30 cm water box in air, 2 mm dose grid."""
    mats, _ = _materials_from_db(["G4_WATER", "G4_AIR"])
    T = physics.build_photon_tables(mats, e_max_mev=energy)
    labels = np.zeros((1, 1, 1), np.uint8)
    s, keep = abi.make_source(pos_type=abi.POS_POINT, pos_translation=(0, 0, -150.0), dir_type=abi.DIR_MOMENTUM,
                              momentum=(0, 0, 1), ene_type=abi.ENE_MONO, mono=energy)
    A = dose_world_to_index(grid, (spacing_mm,) * 3)
    d = abi.make_dose_desc(grid, A, flags, hit_type)
    return LowLevelSetup("C1 1MeV pencil beam / 30 cm water box / 2 mm DoseActor", seed, T, (500, 500, 500), 1,
                         labels, (300.0, 300.0, 300.0), identity34(), [(s, keep)], [d], [], n_primaries,
                         materials=mats)


def make_c2(n=256, n_primaries=100_000_000, flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE | abi.SCORE_DOSE_SQ,
            seed=123456, field_mm=100.0):
    """C2: 6 MV spectrum, 10x10 cm2 field, into an n^3 @ 1 mm water/lung/bone slab phantom."""
    hu = slab_phantom_hu(n)
    db = materials.MaterialDatabase()
    db.read_from_file(materials.DEFAULT_DB)
    lab16, lut = materials.create_label_array(hu, "Air", SLAB_VOXEL_MATERIALS)
    names = list(lut.keys())
    mats = [db.find_or_build_material(nm) for nm in names]
    edges, w = spectrum_6mv()
    T = physics.build_photon_tables(mats, e_max_mev=edges[-1])
    labels = lab16.astype(np.uint8)
    half = n * 1.0 / 2
    s, keep = abi.make_source(pos_type=abi.POS_BOX, half_size=(field_mm / 2, field_mm / 2, 0.0),
                              pos_translation=(0, 0, -half - 1.0), dir_type=abi.DIR_MOMENTUM, momentum=(0, 0, 1),
                              ene_type=abi.ENE_SPECTRUM_HISTOGRAM, spectrum_energy=edges,
                              spectrum_cdf=histogram_cdf(edges, w))
    A = dose_world_to_index((n, n, n), (1.0, 1.0, 1.0))
    d = abi.make_dose_desc((n, n, n), A, flags, abi.HIT_POST)
    wm = names.index("Air")
    return LowLevelSetup(f"C2 6MV 10x10 / {n}^3 slab water-lung-bone / DoseActor dose+uncertainty", seed, T,
                         (half + 200, half + 200, half + 200), wm, labels, (1.0, 1.0, 1.0), identity34(),
                         [(s, keep)], [d], [], n_primaries, materials=mats)


def schneider_materials(tol_g_cm3=0.05):
    class _S:
        pass
    sim = _S()
    sim.volume_manager = _S()
    db = materials.MaterialDatabase()
    db.read_from_file(materials.DEFAULT_DB)
    sim.volume_manager.material_database = db
    vm, created = materials.HounsfieldUnit_to_material(sim, tol_g_cm3 * g4_units.g_cm3)
    return vm, created, db


def make_c3(nx=512, ny=512, nz=300, n_primaries=1_000_000_000, seed=987654321,
            flags=abi.SCORE_EDEP | abi.SCORE_EDEP_SQ | abi.SCORE_DOSE, spacing=(0.9766, 0.9766, 1.0),
            source_distance=600.0, hu=None, spectrum="120kvp"):
    """C3: 120 kVp spectrum, point source 600 mm off-axis focused cone, into a synthetic CT
    converted with Schneider (74 materials at 0.05 g/cm3)."""
    if hu is None:
        hu = synthetic_ct_hu(nx, ny, nz, seed)
    nz, ny, nx = hu.shape
    vm, created, db = schneider_materials(0.05)
    lab16, lut = materials.create_label_array(hu, "G4_AIR", vm)
    names = list(lut.keys())
    mats = [db.find_or_build_material(nm) for nm in names]
    edges, w = spectrum_6mv() if spectrum == "6mv" else spectrum_120kvp()
    T = physics.build_photon_tables(mats, e_max_mev=edges[-1])
    labels = lab16.astype(np.uint8) if len(names) <= 256 else lab16
    half = np.array([nx * spacing[0], ny * spacing[1], nz * spacing[2]]) / 2
    # cone beam: isotropic emission restricted in theta around -x ... use a focused field: a
    # square 'virtual aperture' at the phantom centre plane; photons start at the focal spot
    # and aim at a uniformly sampled point of the aperture (position box + focused direction
    # would aim all at one point, so emit from the aperture plane backwards instead):
    # source = disc of radius 0 at (-d,0,0); direction iso within a cone that covers the body.
    theta_max = np.arctan(max(half[1], half[2]) * 0.95 / source_distance)
    # G4 iso convention: momentum = -(sin t cos p, sin t sin p, cos t); rotate so that -z -> +x
    rot = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    s, keep = abi.make_source(pos_type=abi.POS_POINT, pos_translation=(-source_distance, 0, 0), dir_type=abi.DIR_ISO,
                              theta=(0.0, theta_max), phi=(0.0, 2 * np.pi), dir_relative=True, dir_rotation=rot,
                              ene_type=abi.ENE_SPECTRUM_HISTOGRAM, spectrum_energy=edges,
                              spectrum_cdf=histogram_cdf(edges, w))
    A = dose_world_to_index((nx, ny, nz), spacing)
    d = abi.make_dose_desc((nx, ny, nz), A, flags, abi.HIT_POST)
    wh = tuple(float(max(h + 100.0, source_distance + 100.0)) for h in half)
    title = "C5 6MV cone" if spectrum == "6mv" else "C3 120kVp cone"
    return LowLevelSetup(f"{title} / {nx}x{ny}x{nz} synthetic CT Schneider({len(names)} mat) / DoseActor dose+edep2",
                         seed, T, wh, names.index("G4_AIR"), labels, tuple(spacing), identity34(), [(s, keep)], [d], [],
                         n_primaries, materials=mats)


def make_c4(nx=256, ny=256, nz=150, spacing=(2.0, 2.0, 2.0), n_act=128, act_spacing=4.0, n_primaries=1_000_000_000,
            seed=42, capacity=None, hu=None):
    """C4: Tc-99m 140.511 keV voxelised source in the CT phantom, phase-space plane
    540x400x1 mm at 250 mm from the centre (entering steps)."""
    from .image import compute_image_3D_CDF
    if hu is None:
        hu = synthetic_ct_hu(nx, ny, nz, 987654321)
    nz, ny, nx = hu.shape
    vm, created, db = schneider_materials(0.05)
    lab16, lut = materials.create_label_array(hu, "G4_AIR", vm)
    names = list(lut.keys())
    mats = [db.find_or_build_material(nm) for nm in names]
    e0 = 0.140511
    T = physics.build_photon_tables(mats, e_max_mev=e0)
    labels = lab16.astype(np.uint8)
    act = activity_image(n_act)
    cdf_x, cdf_y, cdf_z = compute_image_3D_CDF(act)
    org = -np.array([n_act] * 3) / 2.0 * act_spacing + act_spacing / 2.0
    s, keep = abi.make_source(pos_type=abi.POS_VOXELS, dir_type=abi.DIR_ISO, ene_type=abi.ENE_MONO, mono=e0,
                              vox_size=(n_act,) * 3, vox_spacing=(act_spacing,) * 3, vox_origin=org,
                              cdf_z=cdf_z, cdf_y=cdf_y, cdf_x=cdf_x)
    half = np.array([nx * spacing[0], ny * spacing[1], nz * spacing[2]]) / 2
    plane_y = 250.0 + 0.5
    B = identity34()
    B[:, 3] = -np.array([0.0, plane_y, 0.0])
    if capacity is None:
        capacity = max(int(n_primaries * 0.12) + 1024, 1024)
    p = abi.make_phsp_desc((270.0, 0.5, 200.0), B, abi.PHSP_ENTERING, capacity)
    A = dose_world_to_index((nx, ny, nz), spacing)
    d = abi.make_dose_desc((nx, ny, nz), A, abi.SCORE_EDEP, abi.HIT_POST)
    return LowLevelSetup(f"C4 Tc-99m VoxelSource {n_act}^3 / {nx}x{ny}x{nz} CT / PhaseSpace plane entering", seed, T,
                         (400.0, 400.0, 400.0), names.index("G4_AIR"), labels, tuple(spacing), identity34(),
                         [(s, keep)], [d], [p], n_primaries, materials=mats)
