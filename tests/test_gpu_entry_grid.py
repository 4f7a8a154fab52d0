"""-m gpu: the entry grid (PT_OPT_ENTRY_GRID) changes where a walk STARTS, never what it finds:
closest hits (primitive, t, normal) and images must be bit-identical with and without it, while the node visits drop."""
import numpy as np
import pytest

from cpppathtracer_b200 import api, scenes

pytestmark = pytest.mark.gpu


def _ray_batch(sc, rng, n):
    """Rays of every kind a path produces: from outside towards the scene, short ones born on primitives and on the
    floor, axis-parallel ones, zero direction components, bounded [tmin, tmax] intervals."""
    tree = np.where(sc.prim_type != scenes.PLATFORM)[0]
    c = sc.center[tree]
    lo, hi = c.min(0) - 5, c.max(0) + 5
    size = float((hi - lo).max())
    o_in = (lo + rng.random_sample((n, 3)) * (hi - lo)).astype(np.float32)
    o_floor = o_in.copy(); o_floor[:, 1] = 0.0
    pick = tree[rng.randint(0, len(tree), n)]
    o_surf = (sc.center[pick] + np.abs(sc.radius[pick])[:, None] * 1.001 * _unit(rng.normal(size=(n, 3)))).astype(np.float32)
    o_out = (o_in + _unit(rng.normal(size=(n, 3))) * size * 3).astype(np.float32)
    d_rand = _unit(rng.normal(size=(n, 3)))
    d_axis = np.eye(3, dtype=np.float32)[rng.randint(0, 3, n)] * rng.choice([-1.0, 1.0], (n, 1))
    d_planar = d_rand.copy(); d_planar[:, rng.randint(0, 3)] = 0.0; d_planar = _unit(d_planar + 1e-20)
    d_to = _unit(o_in - o_out)
    rays = []
    for o, d in ((o_in, d_rand), (o_floor, d_rand), (o_surf, d_rand), (o_out, d_to), (o_in, d_axis), (o_floor, d_planar), (o_surf, d_axis)):
        tmin = np.where(rng.random_sample(n) < 0.5, 0.0, 2e-5)
        tmax = np.where(rng.random_sample(n) < 0.7, 1e30, rng.random_sample(n) * size)
        rays.append(np.concatenate([o, d, tmin[:, None], tmax[:, None]], 1))
    return np.concatenate(rays).astype(np.float32)


def _unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


SCENES = {
    "c2": lambda: scenes.sphere_field(1000, 320, 180),
    "demo": lambda: scenes.demo_scene(320, 180),
    "stress20k": lambda: scenes.stress_field(20000, 320, 180),
    "two": lambda: _tiny(2), "three": lambda: _tiny(3),
    "far": lambda: _shifted(scenes.sphere_field(300, 160, 90), 3.0e5),
    "small": lambda: _scaled(scenes.sphere_field(300, 160, 90), 1e-3),
}


def _tiny(n):
    sc = scenes.sphere_field(n, 160, 90)
    return sc


def _shifted(sc, dx):
    sc.center[:, 0] += np.float32(dx); sc.center[:, 2] -= np.float32(dx)
    sc.camera.origin = (sc.camera.origin[0] + dx, sc.camera.origin[1], sc.camera.origin[2] - dx)
    sc.camera.look_at = (dx, 0.0, -dx)
    return sc


def _scaled(sc, k):
    sc.center *= np.float32(k); sc.radius *= np.float32(k)
    sc.camera.origin = tuple(v * k for v in sc.camera.origin)
    return sc


@pytest.mark.parametrize("name", list(SCENES))
def test_entry_grid_changes_no_hit(name, sky):
    sc = SCENES[name]()
    rng = np.random.RandomState(7)
    rays = _ray_batch(sc, rng, 6000)
    out = {}
    for levels in (0, 12):
        r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
        r.set_option(api.PT_OPT_ENTRY_GRID, levels); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.build()
        st = r.stats()
        prim, t, nrm, pos = r.trace(rays)
        r.render(sc.camera, 4, 8, seed=3)
        img = r.read_accum().copy()
        r.set_option(api.PT_OPT_COUNT_STEPS, 1); r.reset_accum(); r.render(sc.camera, 2, 8, seed=3); cs = r.stats()
        out[levels] = (prim, t, nrm, img, cs["node_steps"] / cs["segments"], st["entry_levels"], st["entry_cells"])
        r.close()
    p0, t0, n0, i0, v0, l0, c0 = out[0]
    p1, t1, n1, i1, v1, l1, c1 = out[12]
    assert l0 == 0 and c0 == 0
    assert np.array_equal(p0, p1) and np.array_equal(t0, t1) and np.array_equal(n0, n1)
    assert np.array_equal(i0, i1)
    assert (p0 >= 0).mean() > 0.2                        # the batch does hit things
    if name in ("c2", "stress20k"):
        assert l1 >= 3 and v1 < 0.85 * v0, f"node visits per segment {v0:.2f} -> {v1:.2f}"
    else:
        assert v1 <= v0 * 1.02


def test_entry_grid_is_off_for_reference_leaf_boxes_and_single_primitives(sky):
    sc = scenes.sphere_field(1, 64, 36)
    r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
    r.set_option(api.PT_OPT_ENTRY_GRID, 12); r.build()
    assert r.stats()["entry_levels"] == 0
    r.close()
    sc = scenes.sphere_field(500, 64, 36)
    r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
    assert r.stats()["entry_levels"] == 0                     # off by default
    r.set_option(api.PT_OPT_ENTRY_GRID, 12); r.build()
    assert r.stats()["entry_levels"] > 0
    r.set_option(api.PT_OPT_REF_LEAF_BOXES, 1); r.build()
    assert r.stats()["entry_levels"] == 0
    r.close()
    # the default library carries no entry-grid code in its kernels (it costs 4 % on C2 even when unused) and says so
    r = api.Renderer(sc, sky)
    with pytest.raises(api.PtError):
        r.set_option(api.PT_OPT_ENTRY_GRID, 12)
    r.set_option(api.PT_OPT_ENTRY_GRID, 0)
    r.close()


@pytest.mark.parametrize("maker", [lambda: scenes.sphere_field(1000, 320, 180), lambda: scenes.demo_scene(320, 180)])
def test_split_pipeline_renders_the_same_image(maker, sky):
    """PT_OPT_SPLIT_WALK (bounces as shade + classify by entry-grid level, then walk per level class) changes the order in
    which rays are walked, nothing else: images are bit-identical to the fused kernel's, with and without emitters, and the
    display path (G-buffer normals of the first segment) delivers the same bytes."""
    sc = maker()
    sc.materials = sc.materials.copy()
    out = {}
    for emit in (0.0, 0.8):
        sc.materials["emit_intensity"][3] = emit
        for split in (0, 1):
            r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_GRID)
            r.set_option(api.PT_OPT_ENTRY_GRID, 12); r.set_option(api.PT_OPT_ENTRY_CELLS_PER_PRIM, 8); r.set_option(api.PT_OPT_SPLIT_WALK, split); r.build()
            r.render(sc.camera, 6, 8, seed=4); img = r.read_accum().copy()
            r.reset_accum(); r.render(sc.camera, 2, 3, seed=4); img3 = r.read_accum().copy()       # a shallow depth too
            frames = [r.dispatch_ray(sc.camera, i, 8, 9).copy() for i in (1, 2)]
            st = r.stats()
            out[(emit, split)] = (img, img3, frames)
            r.close()
        a, b = out[(emit, 0)], out[(emit, 1)]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert all(np.array_equal(x, y) for x, y in zip(a[2], b[2]))
