"""Step latency of SMALL worlds - the reference's own use (a handful of bodies at 60 Hz).

    python tools/small_world_latency.py

Live reference step (drag + uniform gravity + integration + collisions).  Worlds of up to 256 bodies
run whole steps in one single-CTA kernel (tiny.cuh); larger ones take the grid pipeline (one CUDA graph
per step).  Reported: microseconds per step for tr_step(w, 1) issued back to back, and per step inside
one tr_step(w, K) call."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes


def run(name, b, k=2000, **pover):
    with tr.World(tr.default_params(**pover)) as w:
        w.add(b)
        w.step(50)
        t0 = time.perf_counter()
        for _ in range(k):
            w.step(1, sync=False)
        w.sync()
        t1 = time.perf_counter()
        w.step(k)
        t2 = time.perf_counter()
        print(f"{name:34s} n={len(b):6d}: {1e6 * (t1 - t0) / k:8.2f} us/step (k x tr_step(w,1)), {1e6 * (t2 - t1) / k:8.2f} us/step (tr_step(w,k)), "
              f"{w.stats().last_contacts} contacts in the last step")


run("default scene", scenes.default_scene())
run("default scene, grid pipeline", scenes.default_scene(), debug_flags=tr.DEBUG_NO_TINY)
run("200 spheres", scenes.random_cloud(200, 3, spheres=True, density_l=0.6), k=1000)
run("200 spheres, grid pipeline", scenes.random_cloud(200, 3, spheres=True, density_l=0.6), k=1000, debug_flags=tr.DEBUG_NO_TINY)
run("1K spheres", scenes.random_cloud(1024, 3, spheres=True, density_l=0.8), k=1000)
run("32K spheres", scenes.random_cloud(32768, 3, spheres=True, density_l=1.0), k=500)
