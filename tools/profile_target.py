"""Profiling target: C2 scene, 1920x1080, 16 spp (= one 33.2 M-path wave) per frame, depth 8.  Same kernels as
bench.py's step, which is one 64-spp wave (launches four times as long; ncu replays each launch ~40 times, so the
profile is taken on the quarter-size wave)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = scenes.sphere_field(1000, 1920, 1080)
r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
for f in range(frames):
    r.reset_accum()
    r.render(sc.camera, 16, 8, seed=2026)
st = r.stats()
print(f"{st['paths'] / st['render_ms'] / 1e3:.1f} Mpaths/s, render_ms {st['render_ms']:.2f}, segments {st['segments']}")
r.close()
