import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
for which in sys.argv[1:] or ["c2"]:
    sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080) if which == "c3" else scenes.demo_scene(640, 360)
    spp = 8 if which != "c1" else 64
    r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
    r.set_option(api.PT_OPT_STAGE_TIMING, 1)
    for _ in range(3):
        r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
    st = r.stats()
    print(f"{which}: {st['paths']/st['render_ms']/1e3:7.1f} Mpaths/s  trace-launches {st["traverse_ms"]:.2f} tail {st["shade_ms"]:.2f} total {st['render_ms']:.2f} build {st['build_ms']:.2f}", flush=True)
    import time
    r.update_prim(1, sc.center[1], float(sc.radius[1]))
    t0 = time.perf_counter(); r.build(); t1 = time.perf_counter()
    print(f"{which}: pt_build after one pt_update_prim: {(t1 - t0) * 1e3:.2f} ms wall ({sc.n_prims} primitives; host tie order + upload + device LBVH {r.stats()['build_ms']:.2f} ms)", flush=True)
    r.close()
