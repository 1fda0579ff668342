"""Per-step latency of small worlds (the reference's interactive use: a handful of bodies at 60 Hz)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes

for name, b in (("default scene (3 bodies)", scenes.default_scene()),
                ("1K sphere cloud", scenes.random_cloud(1024, 1, spheres=True, density_l=0.6)),
                ("32K sphere cloud", scenes.random_cloud(32768, 1, spheres=True, density_l=0.6))):
    with tr.World(tr.default_params()) as w:
        w.add(b)
        w.step(50)
        t0 = time.perf_counter()
        w.step(1000)
        dt = time.perf_counter() - t0
        print(f"{name}: {dt / 1000 * 1e6:.1f} us per step (live reference step: drag + uniform gravity + integration + collisions), "
              f"contacts last step {w.stats().last_contacts}")
