"""Times library variants built into variants/libpt_*.so (development aid)."""
import os, sys, glob, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for so in sorted(glob.glob(os.path.join(root, "variants", "libpt_*.so"))):
    code = f"""
import os, sys; sys.path.insert(0, {root!r})
from cpppathtracer_b200 import api, scenes
for which in ("c2", "c3"):
    sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080)
    r = api.Renderer(sc, scenes.synthetic_sky(1280, 1280), lib_path={so!r})
    for _ in range(3):
        r.reset_accum(); r.render(sc.camera, int(os.environ.get('PT_VARIANT_SPP', '16')), 8, seed=2026)
    st = r.stats()
    print({os.path.basename(so)!r}, which, f"{{st['paths']/st['render_ms']/1e3:7.1f}} Mpaths/s", flush=True)
    r.close()
"""
    subprocess.run([sys.executable, "-c", code])
