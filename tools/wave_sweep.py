import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
sc = scenes.sphere_field(1000, 1920, 1080)
r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280))
for run in (2, 4, 8):
    r.set_option(api.PT_OPT_TRAV_INNER_RUN, run)
    r.reset_accum(); r.render(sc.camera, 16, 8, seed=1); r.reset_accum(); r.render(sc.camera, 16, 8, seed=1)
    st = r.stats(); print(f"inner_run {run}: {st['paths']/st['render_ms']/1e3:.1f} Mpaths/s")
r.set_option(api.PT_OPT_TRAV_INNER_RUN, 4)
for spp_wave in (1, 2, 4, 8, 16, 32, 64):
    r.set_option(api.PT_OPT_WAVE_PATHS, spp_wave * 1920 * 1080)
    for _ in range(2):
        r.reset_accum(); r.render(sc.camera, 64, 8, seed=1)
    st = r.stats(); print(f"wave {spp_wave} spp ({st['wave_paths']/1e6:.1f} M paths): {st['paths']/st['render_ms']/1e3:.1f} Mpaths/s, {st['render_ms']:.2f} ms / 64 spp frame")
