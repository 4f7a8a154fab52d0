"""SAH refit passes (PT_OPT_SAH_PASSES): SAH cost, node visits per segment, throughput, and image equality with the
plain LBVH (closest hits do not depend on the tree, so the images must be identical on scenes without graze-class rays)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
for which in sys.argv[1:] or ["c2", "c3", "c1"]:
    sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080) if which == "c3" else scenes.demo_scene(640, 360)
    spp = 16 if which != "c1" else 64
    ref = None
    for passes in (0, 1, 2, 4, 8):
        r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
        r.set_option(api.PT_OPT_SAH_PASSES, passes); r.build()
        for _ in range(3):
            r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
        st = r.stats(); img = r.read_accum().copy()
        r.set_option(api.PT_OPT_COUNT_STEPS, 1); r.reset_accum(); r.render(sc.camera, 2, 8, seed=2026); cs = r.stats()
        if ref is None: ref = img
        print(f"{which} passes {passes}: SAH {st['bvh_sah_lbvh']:.3f} -> {st['bvh_sah']:.3f}  nodes/seg {cs['node_steps']/cs['segments']:.2f} prims/seg {cs['prim_tests']/cs['segments']:.3f}"
              f"  {st['paths']/st['render_ms']/1e3:7.1f} Mpaths/s  build {st['build_ms']:.2f} ms  identical {np.array_equal(img, ref)}"
              + ("" if np.array_equal(img, ref) else f" (differing pixels {(np.abs(img - ref).max(axis=2) > 0).sum()})"), flush=True)
        r.close()
