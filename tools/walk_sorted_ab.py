"""Does ordering rays by entry-grid level pay in the walk kernel itself?  Real segment-1/2 rays (oracle generator, 1080p,
4 samples = ~8 M rays) through pt_trace_rays (k_bounce<SRC_RAYS>: load ray, walk, store hit) in natural order and sorted
by level, entry grid off / on; device time of the launch alone.   usage: python tools/walk_sorted_ab.py [c2 c3]"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpppathtracer_b200 import api, scenes
from oracle import ref_host
from walk_predictor import grid_of, levels_of   # noqa: E402  (module body runs only under __main__ there)

sky = scenes.synthetic_sky(1280, 1280)
for which in sys.argv[1:] or ["c2", "c3"]:
    sc = scenes.sphere_field(1000, 1920, 1080) if which == "c2" else scenes.stress_field(100000, 1920, 1080)
    port = ref_host.PortOracle(); port.load_scene(sc)
    lo, hi, k = grid_of(sc)
    for seg in (1, 2):
        batch = []
        for sample in range(4):
            rays, valid = port.segment_rays(sc.camera, seg, sample=sample, max_depth=8, seed=1)
            batch.append(rays[valid])
        rays = np.concatenate(batch)
        for grid in (0, 12):
            r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
            r.set_option(api.PT_OPT_ENTRY_GRID, grid); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.build()
            lmax = r.stats()["entry_levels"] if grid else 9
            lvl, miss = levels_of(rays, sc, lo, hi, k, lmax)
            orders = {"natural": np.arange(len(rays)), "by level (long first)": np.argsort(lvl, kind="stable"),
                      "by level, rays that miss the box removed": np.argsort(lvl, kind="stable")[: int((~miss).sum())]}
            ref = None
            for name, order in orders.items():
                rr = np.ascontiguousarray(rays[order])
                for run in (4, 6):
                    r.set_option(api.PT_OPT_TRAV_INNER_RUN, run)
                    for _ in range(3):
                        prim, t, _, _ = r.trace(rr)
                    ms = r.stats()["traverse_ms"]
                    if ref is None:
                        ref = (prim.copy(), t.copy())
                    elif name == "by level (long first)":
                        back = np.empty_like(order); back[order] = np.arange(len(order))
                        assert np.array_equal(prim[back], ref[0]) and np.array_equal(t[back], ref[1])
                    print(f"{which} seg {seg} grid {'on ' if grid else 'off'} {name:42s} inner run {run}: {len(rr) / 1e6:5.2f} M rays in {ms:6.3f} ms = {len(rr) / ms / 1e3:8.1f} Mrays/s", flush=True)
            r.close()
