import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
from oracle import ref_host
sky = scenes.synthetic_sky(512, 512)
P = ref_host.PortOracle()
for name, sc in [("demo", scenes.demo_scene(640, 360)), ("sph", scenes.sphere_field(1000, 480, 270))]:
    P.load_scene(sc); P.set_sky(sky)
    r = api.Renderer(sc, sky)
    for seg in range(0, 6):
        rays, valid = P.segment_rays(sc.camera, seg, 0, 8, seed=5)
        rays = rays[valid]
        pg, tg, ng, hg = r.trace(rays)
        pr, tr, nr, hr = P.trace(rays)
        mism = np.where(pg != pr)[0]
        tdiff = np.where((pg == pr) & (tg != tr))[0]
        print(name, "seg", seg, "rays", len(rays), "prim mismatches", len(mism), "t mismatches", len(tdiff))
        for i in mism[:6]:
            print("    ray", rays[i].tolist(), "gpu", pg[i], tg[i], "ref", pr[i], tr[i])
    r.close()
