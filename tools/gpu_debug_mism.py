import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
from oracle import ref_host
sky = scenes.synthetic_sky(64, 64)
P = ref_host.PortOracle()
sc = scenes.stress_field(20000, 480, 270)
P.load_scene(sc)
r = api.Renderer(sc, sky)
rays, valid = P.segment_rays(sc.camera, 0, 0, 8, seed=31)
pg, tg, ng, hg = r.trace(rays)
pr, tr, nr, hr = P.trace(rays)
mism = np.where(pg != pr)[0]
print("mismatches", len(mism), "t equal among them", (tg[mism] == tr[mism]).mean())
T = sc.prim_type
from collections import Counter
print(Counter((int(T[a]) if a >= 0 else -1, int(T[b]) if b >= 0 else -1, bool(x == y)) for a, b, x, y in zip(pg[mism], pr[mism], tg[mism], tr[mism])))
order = P.leaf_order()
for i in mism[:10]:
    a, b = pg[i], pr[i]
    print("ray", rays[i][:6], "gpu", a, tg[i], "ref", b, tr[i], "ranks", order[a] if a>=0 else None, order[b] if b>=0 else None,
          "gpu prim", sc.center[a], sc.radius[a], sc.height[a], "ref prim", sc.center[b], sc.radius[b], sc.height[b])
r.set_option(api.PT_OPT_TIE_ORDER, 0); r.build()
pg2 = r.trace(rays)[0]
print("with tie order off: mismatches", (pg2 != pr).sum())
