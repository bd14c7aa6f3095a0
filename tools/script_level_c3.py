#!/usr/bin/env python
"""Config 3 through the PUBLIC script-level API (Simulation / add_volume("Image") / GenericSource /
DoseActor, engine = "b200") at full size, with a wall-clock profile of sim.run(): where the host
time goes around the transport kernel (development aid)."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

import opengate_b200 as gate  # noqa: E402
from opengate_b200 import image as gimage, setups  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
mm, m, MeV, deg = gate.g4_units.mm, gate.g4_units.m, gate.g4_units.MeV, gate.g4_units.deg
hu = setups.synthetic_ct_hu()
sp = (0.9766, 0.9766, 1.0)
sim = gate.Simulation()
sim.random_seed = 987654321
sim.output_dir = "/tmp/c3_script"
sim.world.size = [2 * m, 2 * m, 2 * m]
ct = sim.add_volume("Image", "ct")
ct.image = gimage.Image(hu, sp)
ct.material = "G4_AIR"
ct.voxel_materials, _ = gate.geometry.materials.HounsfieldUnit_to_material(sim, 0.05 * gate.g4_units.g_cm3)
sim.physics_manager.set_production_cut("world", "all", 1 * m)
sim.physics_manager.set_production_cut("ct", "all", 1 * m)
src = sim.add_source("GenericSource", "xray")
edges, w = setups.spectrum_120kvp()
src.energy.type = "spectrum_histogram"
src.energy.spectrum_energy_bin_edges = list(edges * MeV)
src.energy.spectrum_weights = list(w)
src.position.type = "point"
src.position.translation = [-600 * mm, 0, 0]
src.direction.type = "iso"
src.direction.theta = [0, float(np.degrees(np.arctan(0.95 * 250.0 / 600.0))) * deg]
src.position.rotation = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
src.direction_relative_to_attached_volume = True
src.number_of_primaries = n
dose = sim.add_actor("DoseActor", "dose")
dose.attached_to = "ct"
dose.size = "like_image_volume"
dose.spacing = "like_image_volume"
dose.hit_type = "post"
dose.dose.active = True
dose.edep_uncertainty.active = True
dose.write_to_disk = False
stats = sim.add_actor("SimulationStatisticsActor", "stats")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sim.run()
pr.disable()
t1 = time.perf_counter()
print(f"sim.run(): {t1 - t0:.2f} s for {n:.3g} primaries -> {n / (t1 - t0):.3g} primaries/s script level; "
      f"statistics actor: {stats.counts.get('events')} events, duration {stats.counts.get('duration', 0) / gate.g4_units.s:.3f} s")
u = dose.edep_uncertainty.image.array
e = dose.edep.image.array
print("edep sum MeV", float(e.sum()), "median uncertainty where edep > 0:", float(np.median(u[e > 0])))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print("\n".join(s.getvalue().split("\n")[6:36]))
