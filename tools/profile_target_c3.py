"""Profiling target for config C3 (100 k spheres / cylinders, glass-heavy): 1920x1080, 8 spp = one wave per frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = scenes.stress_field(100000, 1920, 1080)
r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280), lib_path=api.LIB_PATH_GRID if os.environ.get("PT_ENTRY_GRID") else api.LIB_PATH)
if os.environ.get("PT_ENTRY_GRID"):          # A/B of the entry grid (profiles/r02_c3_variants.md)
    r.set_option(api.PT_OPT_ENTRY_GRID, int(os.environ["PT_ENTRY_GRID"])); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.build()
for f in range(frames):
    r.reset_accum()
    r.render(sc.camera, 8, 8, seed=2026)
st = r.stats()
print(f"{st['paths'] / st['render_ms'] / 1e3:.1f} Mpaths/s, render_ms {st['render_ms']:.2f}, segments {st['segments']}")
r.close()
