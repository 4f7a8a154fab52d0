"""A/B of the entry grid on one box: Mpaths/s, node visits and primitive tests per segment, image identity.
usage: python tools/ab_entry.py [c1 c2 c3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
sky = scenes.synthetic_sky(1280, 1280)
for which in sys.argv[1:] or ["c1", "c2", "c3"]:
    sc = {"c1": lambda: scenes.demo_scene(640, 360), "c2": lambda: scenes.sphere_field(1000, 1920, 1080),
          "c3": lambda: scenes.stress_field(100000, 1920, 1080)}[which]()
    spp = {"c1": 64, "c2": 16, "c3": 8}[which]
    base = None
    for levels, cpp, run in ((0, 2, 4), (12, 2, 4), (12, 8, 4), (12, 32, 4), (12, 8, 2), (12, 8, 3), (12, 8, 6), (0, 2, 3)):
        r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
        r.set_option(api.PT_OPT_ENTRY_GRID, levels); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, cpp); r.set_option(api.PT_OPT_TRAV_INNER_RUN, run); r.build()
        b = r.stats()
        for _ in range(3):
            r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
        st = r.stats(); img = r.read_accum().copy()
        r.set_option(api.PT_OPT_COUNT_STEPS, 1); r.reset_accum(); r.render(sc.camera, 2, 8, seed=2026); cs = r.stats()
        if base is None:
            base = img
        same = bool(np.array_equal(base, img))
        print(f"{which} levels {b['entry_levels']:2d} cells {b['entry_cells']:8d} (cap {cpp:2d}/prim, inner run {run}): {st['paths'] / st['render_ms'] / 1e3:8.1f} Mpaths/s  "
              f"nodes/seg {cs['node_steps'] / cs['segments']:6.2f} prims/seg {cs['prim_tests'] / cs['segments']:.3f}  build {b['build_ms']:.2f} ms  identical {same}", flush=True)
        r.close()
        if levels == 0 and which == "c3":
            pass
