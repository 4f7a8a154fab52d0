"""Per-step wall time of the end-to-end loop (update + build + render + read-back) with the three read-back styles."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpppathtracer_b200 import api, scenes
sc = scenes.sphere_field(1000, 1920, 1080); sky = scenes.synthetic_sky(1280, 1280)
cam = sc.camera
for use_torch_stream in (True, False):
    r = api.Renderer(sc, sky)
    if use_torch_stream:
        r.set_stream(torch.cuda.current_stream().cuda_stream)
    bufs = [torch.empty((1080, 1920, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    def step(k, mode, first):
        r.update_prim(1, sc.center[1], float(sc.radius[1])); r.build(); r.reset_accum()
        r.render(cam, 64, 8, seed=2026, sync=False)
        if mode == "sync": r.read_accum(bufs[k & 1])
        elif mode == "pipe":
            r.read_accum_begin(bufs[k & 1])
            if not first: r.read_accum_end()
        else: r.synchronize()
    for mode in ("none", "sync", "pipe", "sync", "pipe"):
        step(0, mode, True)
        if mode == "pipe": r.read_accum_end()
        torch.cuda.synchronize(); t0 = time.perf_counter(); marks = []
        for k in range(8):
            step(k, mode, k == 0); marks.append(time.perf_counter())
        if mode == "pipe": r.read_accum_end()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"stream {'torch current' if use_torch_stream else 'default'} read-back {mode:5s}: {(t1 - t0) / 8 * 1e3:7.3f} ms / step   render_ms {r.stats()['render_ms']:.2f}  steps: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip([t0] + marks[:-1], marks)), flush=True)
    r.close()
