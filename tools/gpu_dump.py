import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
sky = scenes.synthetic_sky(512, 512)
for name, sc in [("demo", scenes.demo_scene(640, 360)), ("sph", scenes.sphere_field(1000, 480, 270))]:
    r = api.Renderer(sc, sky)
    for d in (1, 2, 8):
        r.reset_accum(); r.render(sc.camera, 1, d, seed=5)
        np.save(f"gpurun_out/{name}_d{d}.npy", r.read_accum())
    r.close()
