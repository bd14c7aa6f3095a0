"""opengate_b200 — B200-native photon-in-voxel-volume dose engine behind the GATE 10 API.

    import opengate_b200 as gate
    sim = gate.Simulation()              # same script surface as `import opengate as gate`
    sim.engine = "b200"                  # the engine switch (default here)
"""
__version__ = "0.2.0"

from .units import g4_units  # noqa: F401
from .exception import fatal, warning  # noqa: F401
from .managers import Simulation  # noqa: F401
from . import materials, image, physics, sources, actors, volumes  # noqa: F401
from . import materials as _materials

# `gate.geometry.materials.HounsfieldUnit_to_material(...)` as in reference scripts
import types as _types

geometry = _types.SimpleNamespace(materials=_materials, volumes=volumes)
