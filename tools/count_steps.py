import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
for name, sc, spp in [("C1 demo", scenes.demo_scene(640, 360), 16), ("C2 spheres1000", scenes.sphere_field(1000, 1920, 1080), 4),
                      ("C3 stress100k", scenes.stress_field(100000, 1920, 1080), 4)]:
    r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
    r.set_option(api.PT_OPT_COUNT_STEPS, 1)
    for depth in (1, 2, 8):
        r.reset_accum(); r.render(sc.camera, spp, depth, seed=2026)
        st = r.stats()
        print(f"{name} depth<= {depth}: segs/path {st['segments']/st['paths']:.2f} node steps/seg {st['node_steps']/st['segments']:.1f} prim tests/seg {st['prim_tests']/st['segments']:.2f} bvh_nodes {st['bvh_nodes']} build_ms {st['build_ms']:.2f}")
    r.close()
