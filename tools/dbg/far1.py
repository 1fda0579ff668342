import sys, numpy as np
sys.path.insert(0, '/root/repo')
import traceit_b200 as tr
from traceit_b200 import scenes
b = scenes.random_cloud(3000, 78, spheres=True, density_l=0.6)
cell = np.float32(np.float32(2.0 * 0.5) * np.float32(1.0078125))
edge = float(cell) * 32768.0
k = 100
for axis, sign in ((0, 1), (1, -1), (2, 1)):
    for d_in, d_out in ((0.45, 0.40), (0.2, 0.7), (1.6, 0.1)):
        p_in = np.zeros(3, np.float32); p_out = np.zeros(3, np.float32)
        p_in[axis] = sign * (edge - d_in); p_out[axis] = sign * (edge + d_out)
        b["pos"][k] = p_in; b["pos"][k + 1] = p_out
        k += 2
for q in range(12):
    b["pos"][k + q] = (2.0e5 + 0.75 * q, -3.0e5, 0.5 * (q % 2))
b["isStatic"][k + 3] = 1; b["isStatic"][k + 4] = 1
b["isSphere"][k + 7] = 0
b["center"] = b["pos"]
with tr.World(tr.default_params()) as w:
    w.add(b)
    g = w.debug_grid()
    print("cell", g["cell"], "bits", g["bits"], "m", len(g["sorted_ids"]))
    print("keys 100..130", [hex(x) for x in g["keys"][100:130]])
    w.step(1)
    st = w.stats()
    print("side", st.side_list_size, "contacts", st.last_contacts)
    print(w.contacts()[:40])
