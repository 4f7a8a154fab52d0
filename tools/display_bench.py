"""Display path throughput: pt_dispatch_ray (1 spp + Denoising + Mix + BGRA8 + D2H per call) at 1080p on C2, next to the
reference's own per-sample service loop (SamplePixel + Denoising + Mix + 3 syncs + D2H) rebuilt for sm_100."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
sc = scenes.sphere_field(1000, 1920, 1080)
sky = scenes.synthetic_sky(1280, 1280)
r = api.Renderer(sc, sky)
for i in range(1, 6):
    r.dispatch_ray(sc.camera, i)
n = 60
t0 = time.perf_counter()
for i in range(6, 6 + n):
    r.dispatch_ray(sc.camera, i)
dt = time.perf_counter() - t0
out = {"dispatch_ms": dt / n * 1e3, "mpaths_per_s": 1920 * 1080 * n / dt / 1e6}
a = r.dispatch_ray(sc.camera, 200).copy()
b = r.dispatch_ray_frame(sc.camera, 200).copy()      # same sample index twice: the running mean moves, so compare loosely
t0 = time.perf_counter()
for i in range(300, 300 + n):
    r.dispatch_ray_frame(sc.camera, i)
dt = time.perf_counter() - t0
out["frame_dispatch_ms"] = dt / n * 1e3
out["frame_mpaths_per_s"] = 1920 * 1080 * n / dt / 1e6
r.dispatch_ray_begin(sc.camera, 400)
t0 = time.perf_counter()
for i in range(401, 401 + n):
    r.dispatch_ray_begin(sc.camera, i)
    r.dispatch_ray_end()
dt = time.perf_counter() - t0
r.dispatch_ray_end()
out["pipelined_dispatch_ms"] = dt / n * 1e3
out["pipelined_mpaths_per_s"] = 1920 * 1080 * n / dt / 1e6
out["frame_vs_copy_max_byte_diff"] = int(np.abs(a.astype(int) - b.astype(int)).max())
try:
    from oracle import ref_cuda
    R = ref_cuda.RefCuda(); R.load_scene(sc); R.set_sky(sky)
    o = R.render(sc.camera, 16, warmup=2)
    out["reference_service_loop_mpaths_per_s"] = o["mpaths_service"]
    out["reference_service_ms_per_sample"] = o["service_ms"] / 16
except Exception as e:
    out["reference"] = str(e)
print(json.dumps(out))
