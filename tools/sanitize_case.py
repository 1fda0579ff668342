"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of the
library runs at least once - gravity with the fused integrator (incl. the guarded diagonal path
and ragged tails), standalone integrator, counting-sort broad phase, narrow phase, side list,
far list, every ordered-response executor, a crowded cell (K4's CTA sort, the K6 prep's grid-wide sort of a
contact list), pack/unpack, the per-frame call, the single-CTA small-world step, a contact-buffer overflow with
its recovery."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes

b = scenes.random_cloud(3001, 1, spheres=True, density_l=0.6)
b["isSphere"][::50] = 0
b["isStatic"][::7] = 1
b["stationary"][::13] = 1
with tr.World(tr.default_params(flags=tr.PAIR_GRAVITY | tr.COLLISIONS, G=np.float32(1e-3))) as w:
    w.add(b)
    w.step(2)
    w.read_bodies(); w.contacts(); w.debug_grid(); w.debug_candidates(); w.debug_gravity()
    host = w.read_bodies()
    w.step_host(host, 1)
    frame = np.zeros(len(b), dtype=tr.FRAME_DTYPE)
    w.step_frame(frame, np.zeros((len(b), 3), dtype=np.float32), 1)
with tr.World(tr.default_params()) as w:       # single-CTA path, history ring on
    w.enable_trajectory(8)
    w.add(scenes.default_scene()); w.step(20); w.read_state(); w.trajectory(1)
with tr.World(tr.default_params(debug_flags=tr.DEBUG_NO_TINY)) as w:   # the same world through the grid pipeline (side list)
    w.add(scenes.default_scene()); w.step(20); w.read_state()
with tr.World(tr.default_params(g=(0, 0, 0))) as w:                   # first step overflows the initial contact buffers
    w.add(scenes.dense_lattice(32))
    for _ in range(3):
        w.step(1, sync=False)
    w.read_state()
d = scenes.dense_lattice(8)
d["pos"][:6] += np.float32(4.0e4)          # a few bodies outside the grid domain: far list
d["center"] = d["pos"]
for mode in (0, 1, 2, 3):
    with tr.World(tr.default_params(g=(0, 0, 0), response_mode=mode)) as w:
        w.add(d); w.step(2); w.contacts()
c = scenes.random_cloud(3000, 2, spheres=True)
c["pos"][:2200] = 0.5; c["center"] = c["pos"]
with tr.World(tr.default_params(g=(0, 0, 0), max_contacts=3_000_000)) as w:
    w.add(c); w.debug_grid(); w.step(1)
    assert w.stats().last_contacts > 2400000
print("SANITIZE_CASE_OK")
