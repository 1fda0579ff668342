import sys, numpy as np
sys.path.insert(0, '/root/repo')
import traceit_b200 as tr
from oracle import pyoracle as po
g = np.load('/root/repo/tests/golden/sphere_cloud_2048_s50.npz')
b = g['initial'].copy()
want = b.copy()
p = po.default_params()
with tr.World(tr.default_params()) as w:
    w.add(b)
    for s in range(50):
        w.step(1)
        got = w.read_bodies()
        gc = w.contacts()
        po.integrate(want, p)
        oc = po.contacts(want)
        po.resolve_list(want, oc)
        bad = np.nonzero((got['pos'].view(np.uint32) != want['pos'].view(np.uint32)).any(axis=1) | (got['vel'].view(np.uint32) != want['vel'].view(np.uint32)).any(axis=1))[0]
        same_c = np.array_equal(gc, oc)
        if len(bad) or not same_c:
            print("step", s, "bad bodies", bad, "contacts equal", same_c, len(gc), len(oc), "rounds", w.stats().last_response_rounds)
            for i in bad[:6]:
                print(i, got[i]['pos'], want[i]['pos'], got[i]['vel'], want[i]['vel'])
                print("   contacts of", i, oc[(oc[:,0]==i)|(oc[:,1]==i)].tolist())
            if not same_c:
                sg = set(map(tuple, gc.tolist())); so = set(map(tuple, oc.tolist()))
                print("gpu-only", sorted(sg-so)[:10], "oracle-only", sorted(so-sg)[:10])
            break
    else:
        print("all 50 steps identical")
