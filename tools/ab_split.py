"""A/B of the split pipeline (PT_OPT_SPLIT_WALK: shade + classify by entry-grid level, then walk one level class per warp) in
libpt_b200_grid.so against the fused bounce kernel, one box: Mpaths/s at two wave sizes and image identity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
sky = scenes.synthetic_sky(1280, 1280)
for which in sys.argv[1:] or ["c1", "c2", "c3"]:
    sc = {"c1": lambda: scenes.demo_scene(640, 360), "c2": lambda: scenes.sphere_field(1000, 1920, 1080),
          "c3": lambda: scenes.stress_field(100000, 1920, 1080)}[which]()
    base = None
    rows = []
    for name, lib, grid, split, run in (("default library, fused", api.LIB_PATH, 0, 0, 6), ("grid library, grid off, fused", api.LIB_PATH_GRID, 0, 0, 6),
                                        ("grid on, fused", api.LIB_PATH_GRID, 12, 0, 6), ("grid on, split", api.LIB_PATH_GRID, 12, 1, 6),
                                        ("grid on, split, inner run 4", api.LIB_PATH_GRID, 12, 1, 4), ("grid on, split, inner run 8", api.LIB_PATH_GRID, 12, 1, 8)):
        r = api.Renderer(sc, sky, lib_path=lib)
        if lib == api.LIB_PATH_GRID:
            r.set_option(api.PT_OPT_ENTRY_GRID, grid); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.set_option(api.PT_OPT_SPLIT_WALK, split)
        r.set_option(api.PT_OPT_TRAV_INNER_RUN, run); r.build()
        out = []
        for spp in ({"c1": (16, 64), "c2": (16, 64), "c3": (8, 32)}[which]):
            for _ in range(3):
                r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
            st = r.stats(); out.append(f"{spp:2d} spp {st['paths'] / st['render_ms'] / 1e3:8.1f}")
        r.reset_accum(); r.render(sc.camera, 4, 8, seed=5); img = r.read_accum().copy()
        if base is None:
            base = img
        diff = int((base != img).any(axis=2).sum())
        print(f"{which} {name:32s}: " + "  ".join(out) + f" Mpaths/s   image vs default: {'identical' if diff == 0 else str(diff) + ' pixels differ'}", flush=True)
        r.close()
