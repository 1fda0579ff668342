import sys, numpy as np, time
sys.path.insert(0, '/root/repo')
import traceit_b200 as tr
from oracle import pyoracle as po
from traceit_b200 import scenes
def errs(a, b):
    ext = np.abs(b['pos']).max(); vr = np.sqrt((b['vel'].astype(np.float64)**2).sum(1).mean())
    dp = a['pos'].astype(np.float64)-b['pos']; dv = a['vel'].astype(np.float64)-b['vel']
    return (np.abs(dp).max()/ext, np.abs(dv).max()/vr, np.sqrt((dp**2).sum(1).mean())/ext, np.sqrt((dv**2).sum(1).mean())/vr)
for n, K, G in ((8192,100,1.0),(8192,100,1e-3),(8192,20,1.0),(16384,10,1.0)):
    b = scenes.random_cloud(n, scenes.SEED_CFG2, spheres=True)
    with tr.World(tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(G), g=(0,0,0))) as w:
        w.add(b); w.step(K); got = w.read_bodies()
    t=time.time()
    w32 = b.copy(); po.step(w32, po.default_params(flags=po.PAIR_GRAVITY, G=np.float32(G), g=(0,0,0)), K)
    w64 = b.copy(); po.step(w64, po.default_params(flags=po.PAIR_GRAVITY|po.GRAVITY_F64, G=np.float32(G), g=(0,0,0)), K)
    print(f"n={n} K={K} G={G} cpu {time.time()-t:.1f}s")
    print("   gpu vs f32  max pos %.2e max vel %.2e | rms pos %.2e rms vel %.2e" % errs(got, w32))
    print("   gpu vs f64  max pos %.2e max vel %.2e | rms pos %.2e rms vel %.2e" % errs(got, w64))
    print("   f32 vs f64  max pos %.2e max vel %.2e | rms pos %.2e rms vel %.2e" % errs(w32, w64))
    print("   max speed", np.sqrt((w64['vel'].astype(np.float64)**2).sum(1)).max())
