"""Profiling target: two traversal configurations back to back (refill 8 vs no refill)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
sc = scenes.sphere_field(1000, 1920, 1080)
r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
for refill, steps, run in [(8, 8, 4), (32, 8, 4)]:
    r.set_option(api.PT_OPT_TRAV_REFILL_MIN, refill); r.set_option(api.PT_OPT_TRAV_STEPS, steps); r.set_option(api.PT_OPT_TRAV_INNER_RUN, run)
    r.reset_accum(); r.render(sc.camera, 8, 8, seed=2026)
    st = r.stats(); print(refill, steps, run, f"{st['paths'] / st['render_ms'] / 1e3:.1f} Mpaths/s")
r.close()
