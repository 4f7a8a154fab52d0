import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080) if which == "c3" else scenes.demo_scene(640, 360)
spp = 8 if which != "c1" else 64
r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
r.set_option(api.PT_OPT_STAGE_TIMING, 1)
print(which, r.stats())
for refill, steps, run in [(32, 8, 4), (16, 8, 4), (8, 8, 4), (32, 16, 4), (32, 4, 4), (32, 8, 2), (32, 8, 8), (16, 16, 8)]:
    r.set_option(api.PT_OPT_TRAV_REFILL_MIN, refill); r.set_option(api.PT_OPT_TRAV_STEPS, steps)
    r.set_option(api.PT_OPT_TRAV_INNER_RUN, run)
    for _ in range(2):
        r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
    st = r.stats()
    print(f"refill_min {refill:2d} steps {steps:4d} run {run:3d}: {st['paths']/st['render_ms']/1e3:7.1f} Mpaths/s  trav {st['traverse_ms']:.2f} shade {st['shade_ms']:.2f} total {st['render_ms']:.2f} segs/path {st['segments']/st['paths']:.2f}", flush=True)
