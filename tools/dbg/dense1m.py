"""Config 5 at full size on the GPU vs the oracle (CPU grid contacts + ordered replay), bitwise."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import traceit_b200 as tr
from oracle import pyoracle as po
from traceit_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = scenes.cfg5_dense(n)
p = tr.default_params(g=(0, 0, 0))
op = po.default_params(g=(0, 0, 0))
want = b.copy()
with tr.World(p) as w:
    w.add(b); w.enable_timing(True)
    for s in range(steps):
        t0 = time.time(); w.step(1); t1 = time.time()
        st = w.stats()
        got = w.read_bodies(); gc = w.contacts(); g = w.debug_grid()
        po.integrate(want, op)
        _, oc, _, nk = po.grid_pairs(want, np.ones(n, np.uint8), (0, 0, 0), g["inv_cell"], g["bits"], want_candidates=False)
        po.resolve_list(want, oc)
        same = all(np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)) for f in ("pos", "vel", "center"))
        print(f"step {s}: {t1-t0:.3f}s wall, contacts {len(gc)} (oracle {nk}) equal={np.array_equal(gc, oc)} state bit-equal={same} rounds={st.last_response_rounds} broad {st.broadphase_ms:.3f} narrow {st.narrowphase_ms:.3f} response {st.response_ms:.3f} ms (cumulative)")
