"""A/B of the 4-wide BVH build variant (variants/libpt_bvh4*.so, -DPT_BVH4=1) against the binary walk (variants/libpt_bvh2.so) on one box:
Mpaths/s, node visits per segment, identical hits and images.   usage: tools/build_variants.sh "bvh4:-DPT_BVH4=1" "bvh2:"; python tools/ab_bvh4.py"""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sky = scenes.synthetic_sky(1280, 1280)
libs = sorted(glob.glob(os.path.join(root, "variants", "libpt_bvh*.so")))
for which in sys.argv[1:] or ["c1", "c2", "c3"]:
    sc = {"c1": lambda: scenes.demo_scene(640, 360), "c2": lambda: scenes.sphere_field(1000, 1920, 1080),
          "c3": lambda: scenes.stress_field(100000, 1920, 1080)}[which]()
    spp = {"c1": 64, "c2": 16, "c3": 8}[which]
    rng = np.random.RandomState(3)
    xy = np.stack([rng.randint(0, sc.camera.width, 200000), rng.randint(0, sc.camera.height, 200000)], 1).astype(np.int32)
    base = None
    for so in libs:
        r = api.Renderer(sc, sky, lib_path=so)
        rays = r.raygen(sc.camera, xy, rng.random_sample((200000, 3)).astype(np.float32)) if base is None else base[2]
        prim, t, _, _ = r.trace(rays)
        for _ in range(3):
            r.reset_accum(); r.render(sc.camera, spp, 8, seed=2026)
        st = r.stats(); img = r.read_accum().copy()
        r.set_option(api.PT_OPT_COUNT_STEPS, 1); r.reset_accum(); r.render(sc.camera, 2, 8, seed=2026); cs = r.stats()
        if base is None:
            base = (prim, t, rays, img)
        same = bool(np.array_equal(base[0], prim) and np.array_equal(base[1], t)); same_img = bool(np.array_equal(base[3], img))
        print(f"{which} {os.path.basename(so):22s}: {st['paths'] / st['render_ms'] / 1e3:8.1f} Mpaths/s  node visits/seg {cs['node_steps'] / cs['segments']:6.2f} "
              f"prims/seg {cs['prim_tests'] / cs['segments']:.3f}  hits identical {same}  image identical {same_img}"
              + ("" if same_img else f" (differing pixels {int((base[3] != img).any(axis=2).sum())})"), flush=True)
        r.close()
