import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import traceit_b200 as tr
from traceit_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
b = scenes.cfg5_dense(n)
with tr.World(tr.default_params(g=(0, 0, 0))) as w:
    w.add(b)
    for s in range(steps):
        t0 = time.time(); w.step(1); dt = time.time() - t0
        if s % 10 == 9 or dt > 0.5:
            st = w.stats(); pos, vel = w.read_state()
            fin = np.isfinite(pos).all(axis=1)
            far = fin & (np.abs(pos).max(axis=1) > 3.0e4)
            print(f"step {s+1}: {dt*1e3:8.2f} ms  contacts {st.last_contacts:8d} side {st.side_list_size:8d} nonfinite {(~fin).sum():8d} far-finite {far.sum():8d} max|v| {np.nanmax(np.abs(vel[np.isfinite(vel).all(axis=1)])):.3g}", flush=True)
        if dt > 5: break
