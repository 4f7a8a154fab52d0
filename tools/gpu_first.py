"""First end-to-end GPU check: parity vs the CPU oracle + timing on C1 / C2 (development aid)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
from oracle import ref_host

def compare_trace(r, orc, rays, tag):
    pg, tg, ng, hg = r.trace(rays)
    pr, tr, nr, hr = orc.trace(rays)
    mism = pg != pr
    both = (~mism) & (pr >= 0)
    rel_t = np.abs(tg[both] - tr[both]) / np.maximum(np.abs(tr[both]), 1e-20)
    dn = np.abs(ng[both] - nr[both]).max(axis=1) if both.any() else np.zeros(1)
    print(f"[{tag}] rays {len(rays)} hits {int((pr>=0).sum())} prim mismatches {int(mism.sum())} "
          f"t bit-exact {float((tg[both]==tr[both]).mean()):.6f} max rel t {rel_t.max() if rel_t.size else 0:.2e} "
          f"normal bit-exact {float((ng[both]==nr[both]).all(axis=1).mean()):.6f} max dn {dn.max():.2e}")
    return mism

def main():
    sky = scenes.synthetic_sky(512, 512)
    orc = ref_host.best_oracle()
    print("oracle kind:", orc.kind)
    for sc, spp in [(scenes.demo_scene(640, 360), 16), (scenes.sphere_field(1000, 480, 270), 8)]:
        orc.load_scene(sc); orc.set_sky(sky)
        r = api.Renderer(sc, sky)
        print(sc.name, "build stats", r.stats())
        rng = np.random.RandomState(1)
        n = 200000
        xy = np.stack([rng.randint(0, sc.camera.width, n), rng.randint(0, sc.camera.height, n)], 1)
        rays = orc.raygen(xy, rng.random_sample((n, 3)))
        rg = r.raygen(sc.camera, xy, rng.random_sample((n, 3)) * 0 + 0.5)
        compare_trace(r, orc, rays, sc.name + " primary")
        # secondary rays: start on the hit points, random directions
        pr, tr, nr, hr = orc.trace(rays)
        hit = pr >= 0
        d2 = rng.normal(size=(hit.sum(), 3)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        rays2 = np.concatenate([hr[hit], d2, np.full((hit.sum(), 1), 2e-5, np.float32), np.full((hit.sum(), 1), 1e30, np.float32)], 1).astype(np.float32)
        compare_trace(r, orc, rays2, sc.name + " bounce")
        t0 = time.time(); ref_sum, st = orc.render(sc.camera, spp, 8, seed=5); t_cpu = time.time() - t0
        r.render(sc.camera, spp, 8, seed=5)
        img = r.read_accum()
        gs = r.stats()
        ref = ref_sum / spp
        diff = img[..., :3] - ref
        exact = (img[..., :3] == ref).all(axis=2).mean()
        print(f"[{sc.name}] image rmse {np.sqrt((diff**2).mean()):.3e} pixels bit-exact {exact:.5f} "
              f"mean gpu {img[...,:3].mean(axis=(0,1))} ref {ref.mean(axis=(0,1))}")
        print(f"   cpu {st['paths']/t_cpu/1e6:.2f} Mpaths/s ({orc.max_threads()} thr)  gpu {gs['paths']/gs['render_ms']/1e3:.1f} Mpaths/s  stats {gs}")
        r.close()
    # throughput at the headline config
    sc = scenes.sphere_field(1000, 1920, 1080)
    r = api.Renderer(sc, sky)
    r.set_option(api.PT_OPT_STAGE_TIMING, 1)
    for it in range(3):
        r.reset_accum()
        r.render(sc.camera, 16, 8, seed=5)
        gs = r.stats()
        print(f"[C2 1080p 16spp] {gs['paths']/gs['render_ms']/1e3:.1f} Mpaths/s  render_ms {gs['render_ms']:.2f} trav {gs['traverse_ms']:.2f} shade {gs['shade_ms']:.2f} segs(last wave) {gs['segments']} launches {gs['launches']}")
    r.close()

if __name__ == "__main__":
    main()

def ref_cuda_baseline():
    from oracle import ref_cuda
    if not ref_cuda.available():
        print("ref cuda lib missing"); return
    sky = scenes.synthetic_sky(512, 512)
    R = ref_cuda.RefCuda()
    for sc, spp in [(scenes.sphere_field(1000, 1920, 1080), 8), (scenes.demo_scene(640, 360), 16)]:
        R.load_scene(sc); R.set_sky(sky)
        out = R.render(sc.camera, spp, warmup=2)
        print(f"[ref-cuda {sc.name} {sc.camera.width}x{sc.camera.height} {spp}spp]", out)

if __name__ == "__main__":
    ref_cuda_baseline()
