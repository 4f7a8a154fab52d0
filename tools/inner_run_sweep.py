"""Sweep of PT_OPT_TRAV_INNER_RUN (inner-node steps per lane between two regroupings of the warp) on C1 / C2 / C3, one box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
sky = scenes.synthetic_sky(1280, 1280)
for which in sys.argv[1:] or ["c1", "c2", "c3"]:
    sc = {"c1": lambda: scenes.demo_scene(640, 360), "c2": lambda: scenes.sphere_field(1000, 1920, 1080),
          "c3": lambda: scenes.stress_field(100000, 1920, 1080)}[which]()
    spp = {"c1": 64, "c2": 16, "c3": 8}[which]
    r = api.Renderer(sc, sky)
    out = []
    for rep in range(2):
        for run in (3, 4, 5, 6, 8, 12):
            r.set_option(api.PT_OPT_TRAV_INNER_RUN, run)
            for _ in range(3):
                r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
            st = r.stats()
            out.append(f"run {run:2d}: {st['paths'] / st['render_ms'] / 1e3:7.1f}")
    print(which, " | ".join(out), flush=True)
    r.close()
