"""Long run of the packed lattice (config 5 shape): the reference's inverted approach test makes the
pack explode after a few dozen steps - bodies leave the grid domain (far list), positions turn NaN/inf ("lost").
Prints device time per step in windows of 10 steps with the contact / outside-the-grid counts.

    python tools/dense_longrun.py [N] [steps] [host]     (host: go through tr_step_host every step)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceit_b200 as tr
from traceit_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
host = len(sys.argv) > 3
b = scenes.cfg5_dense(n)
with tr.World(tr.default_params(flags=tr.COLLISIONS, g=(0, 0, 0))) as w:
    w.add(b)
    w.step(0)
    cur = b.copy()
    for s0 in range(0, steps, 10):
        t0 = time.perf_counter()
        if host:
            for _ in range(10):
                w.step_host(cur, 1)
        else:
            w.step(10)
        dt = (time.perf_counter() - t0) / 10
        st = w.stats()
        pos, _ = w.read_state(vel=False)
        print(f"steps {s0:4d}-{s0 + 9:4d}: {dt * 1e3:8.3f} ms/step, contacts {st.last_contacts:9d}, outside the grid {st.side_list_size:7d}, "
              f"non-finite bodies {int((~np.isfinite(pos).all(axis=1)).sum()):7d}, hand-offs on the longest worker {st.last_response_rounds}", flush=True)
