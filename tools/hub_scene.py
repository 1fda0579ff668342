"""Hub-shaped contact graph: the reference's 10000 x 5 x 10000 `earth` slab (src/main.cpp:1482-1491) with N boxes
inside its AABB at once - every box is in contact with the slab, so ONE body has N contacts.

    python tools/hub_scene.py [N ...]          (default: 25000 50000 100000 200000 400000)

Prints the first step's per-phase device times (the later steps scatter the boxes: the reference's inverted approach
test kicks them out of the slab).  The K6 prep sorts the slab's contact list with one CTA (bitonic, O(L log^2 L));
round 1 ranked every entry against every other (O(L^2): 10^10 loads at L = 10^5).  With `static` as the last argument
the slab is isStatic: its contacts then carry no dependency through it and the replay is wide instead of one chain."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes

args = [a for a in sys.argv[1:] if a != "static"]
static = "static" in sys.argv[1:]
sizes = [int(a) for a in args] or [25000, 50000, 100000, 200000, 400000]
for n in sizes:
    rs = np.random.RandomState(9)
    b = scenes.random_cloud(n + 1, 91, spheres=False)
    side = 900.0 * np.sqrt(n / 100000.0)                  # constant areal density of boxes on the slab
    b["pos"][1:, 0] = rs.uniform(-side, side, n)
    b["pos"][1:, 1] = rs.uniform(-24.5, -15.5, n)
    b["pos"][1:, 2] = rs.uniform(-side, side, n)
    b["size"][1:] = rs.uniform(0.3, 1.0, size=(n, 3)).astype(np.float32)
    b["mass"][1:] = rs.uniform(1, 10, n).astype(np.float32)
    b[0] = scenes.default_scene()[0]
    b["isStatic"][0] = 1 if static else 0
    b["center"] = b["pos"]
    with tr.World(tr.default_params()) as w:
        w.add(b)
        w.step(0)
        w.enable_timing(True)
        w.step(1)
        st = w.stats()
        print(f"slab{' (static)' if static else ''} + {n:7d} boxes: {st.last_contacts:8d} contacts, broad {st.broadphase_ms * 1e3:8.1f} us, "
              f"narrow + side list {st.narrowphase_ms * 1e3:8.1f} us, prep + ordered replay {st.response_ms * 1e3:10.1f} us "
              f"({st.response_ms * 1e6 / max(st.last_contacts, 1):6.1f} ns per contact), longest worker {st.last_response_rounds} hand-offs", flush=True)
