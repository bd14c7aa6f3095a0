"""opengate_b200 — B200-native photon-in-voxel-volume dose engine behind the GATE 10 API."""
__version__ = "0.1.0"
