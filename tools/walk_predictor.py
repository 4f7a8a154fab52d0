"""How well does the entry-grid level of a ray (the extent of its clipped segment) predict the length of its walk?
Real path rays of segments 0-3 (oracle ray generator, test infrastructure) are traced with the instrumented kernel
(per-ray node visits + primitive tests); SIMD utilisation of a 32-ray chunk = mean / max of its rays' steps.
usage: python tools/walk_predictor.py [c2 c3]"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
from oracle import ref_host


def grid_of(sc):
    tree = sc.prim_type != scenes.PLATFORM
    c = sc.center[tree].astype(np.float64); ar = np.abs(sc.radius[tree]).astype(np.float64)
    hy = np.where(sc.prim_type[tree] == scenes.CYLINDER, np.abs(sc.height[tree]) / 2 + 1e-4, ar)
    e = np.stack([ar, hy, ar], 1)
    lo, hi = (c - e).min(0), (c + e).max(0)
    size = (hi - lo).max(); pad = size / 1024 + 1e-6
    lo -= pad; hi += pad; size += 2 * pad
    k = int(np.frexp(size)[1])
    return lo, hi, k


def levels_of(rays, sc, lo, hi, k, lmax):
    o = rays[:, :3].astype(np.float64); d = rays[:, 3:6].astype(np.float64)
    tmin, tmax = rays[:, 6].astype(np.float64), rays[:, 7].astype(np.float64).copy()
    for y in sc.y_pos[sc.prim_type == scenes.PLATFORM]:
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (float(y) - o[:, 1]) / d[:, 1]
        ok = (((o[:, 1] < y) & (d[:, 1] > 0)) | ((o[:, 1] > y) & (d[:, 1] < 0))) & (t > tmin) & (t < tmax)
        tmax = np.where(ok, t, tmax)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / np.where(np.abs(d) > 1e-18, d, 1e-18)
    a, b = (lo - o) * inv, (hi - o) * inv
    t0 = np.maximum(np.minimum(a, b).max(1), tmin); t1 = np.minimum(np.maximum(a, b).min(1), tmax)
    miss = t0 > t1
    p0, p1 = o + t0[:, None] * d, o + t1[:, None] * d
    ext = np.abs(p1 - p0).max(1)
    ext = np.where(miss, 0.0, np.maximum(ext, 1e-30))
    lvl = np.clip(k - (np.frexp(ext)[1]), 0, lmax)       # ext = m 2^e, 0.5 <= m < 1: cell 2^e
    return np.where(miss, lmax + 1, lvl), miss


def util(steps, order):
    s = steps[order]
    n = len(s) // 32 * 32
    g = s[:n].reshape(-1, 32)
    return g.sum() / max(g.max(1).sum() * 32, 1)


def main():
    sky = scenes.synthetic_sky(1280, 1280)
    for which in sys.argv[1:] or ["c2", "c3"]:
        sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080)
        port = ref_host.PortOracle(); port.load_scene(sc)
        cam = copy.deepcopy(sc.camera); cam.width, cam.height = 960, 540
        lo, hi, k = grid_of(sc)
        for grid in (0, 12):
            r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
            r.set_option(api.PT_OPT_ENTRY_GRID, grid); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.build()
            lmax = r.stats()["entry_levels"] if grid else 9
            r.set_option(api.PT_OPT_COUNT_STEPS, 1)
            for seg in (0, 1, 2):
                rays, valid = port.segment_rays(cam, seg, sample=0, max_depth=8, seed=1)
                rays = rays[valid]
                r.trace(rays)
                steps = r.last_trace_steps(len(rays)).astype(np.float64)
                lvl, miss = levels_of(rays, sc, lo, hi, k, lmax)
                if os.environ.get("PT_DUMP_STEPS"):
                    np.savez_compressed(os.path.join(os.environ["PT_DUMP_STEPS"], f"steps_{which}_grid{grid}_seg{seg}.npz"), steps=steps.astype(np.uint16), lvl=lvl.astype(np.int8))
                nat = np.arange(len(rays))
                by_lvl = np.argsort(lvl, kind="stable")
                ideal = np.argsort(steps, kind="stable")
                print(f"{which} grid {'on ' if grid else 'off'} seg {seg}: {len(rays)} rays, mean steps {steps.mean():.2f}, p90 {np.percentile(steps, 90):.0f}, max {steps.max():.0f};  "
                      f"chunk utilisation natural {util(steps, nat):.3f}  sorted by level {util(steps, by_lvl):.3f}  ideal {util(steps, ideal):.3f}", flush=True)
                for T in (2, 3, 4):
                    long_ = lvl <= T
                    if long_.sum() >= 64 and (~long_).sum() >= 64:
                        tot = steps.sum()
                        u_s, u_l = util(steps, nat[~long_]), util(steps, nat[long_])
                        # warp-iterations if short and long rays are chunked separately
                        it = steps[~long_].sum() / u_s + steps[long_].sum() / u_l
                        print(f"      split at level <= {T}: long {long_.mean():.3f} of rays carrying {steps[long_].sum() / tot:.3f} of the steps; utilisation short {u_s:.3f} long {u_l:.3f} combined {tot / it:.3f}")
                row = "      by level: " + "  ".join(f"L{l}: n {np.mean(lvl == l):.2f} steps {steps[lvl == l].mean():.1f}" for l in range(0, lmax + 2) if (lvl == l).sum() > 0)
                print(row, flush=True)
            r.close()


if __name__ == "__main__":
    main()
