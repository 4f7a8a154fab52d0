"""-m gpu: parity of the CUDA path (through the C ABI, libpt_b200.so) against the CPU oracle.

Level 1  fixed ray batches: closest-hit primitive ids bit-exact, t / normals within 1e-5 relative
Level 2  fixed (material, normal, in-dir, u1,u2,u3): scatter directions within 1e-5
Level 3  images: RMSE and channel means vs the oracle render of the same Philox stream
plus size-independent properties at the BASELINE size (determinism, sample-range additivity,
wave-size independence) and the edge cases (empty / single-primitive / platform-only scenes, n = 0).
"""
import numpy as np
import pytest

from conftest import golden
from cpppathtracer_b200 import api, scenes

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def oracle(port_oracle):
    from oracle import ref_host
    return ref_host.RefHost() if ref_host.available() else port_oracle


def _check_trace(r, orc, rays, max_mismatch_frac=0.0):
    pg, tg, ng, hg = r.trace(rays)
    pr, tr, nr, hr = orc.trace(rays)
    mism = pg != pr
    assert mism.mean() <= max_mismatch_frac, f"{mism.sum()} primitive-id mismatches of {len(rays)}"
    ok = (~mism) & (pr >= 0)
    assert np.all(np.abs(tg[ok] - tr[ok]) <= REL * np.abs(tr[ok]))
    assert np.all(np.abs(ng[ok] - nr[ok]) <= REL)
    assert np.all(np.abs(hg[ok] - hr[ok]) <= REL * np.maximum(1.0, np.abs(hr[ok])))
    miss = (~mism) & (pr < 0)
    assert np.all(tg[miss] == tr[miss])
    return int(mism.sum()), float((tg[ok] == tr[ok]).mean()) if ok.any() else 1.0


def test_philox_stream_matches_oracle_and_kat(tmp_path_factory):
    from test_oracle import _philox_lib
    import ctypes as C
    L = _philox_lib(tmp_path_factory)
    r = api.Renderer()
    rng = np.random.RandomState(4)
    pss = np.stack([rng.randint(0, 2**31, 4096), rng.randint(0, 1 << 20, 4096), rng.randint(0, 65, 4096)], 1).astype(np.uint32)
    seed = 0xDEADBEEFCAFEF00D
    got = r.philox(seed, pss)
    want = np.zeros_like(got)
    for i in range(len(pss)):
        L.uni(C.c_ulonglong(seed), C.c_uint(int(pss[i, 0])), C.c_uint(int(pss[i, 1])), C.c_uint(int(pss[i, 2])),
              want[i].ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, want)
    assert got.min() > 0.0 and got.max() <= 1.0
    # Random123 KAT through the (0,1] mapping: counter (0,0,0,0), key (0,0)
    kat = r.philox(0, np.zeros((1, 3), np.uint32))[0]
    w = np.array([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], np.uint32).astype(np.float32)
    assert np.array_equal(kat, w * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10))
    r.close()


def test_camera_basis_and_raygen_bit_exact(oracle):
    rng = np.random.RandomState(5)
    r = api.Renderer()
    for cam in (scenes.Camera(640, 360), scenes.Camera(1920, 1080, (50, 20, -70), (3, 1, 2), view_fov=55.0, lens_radius=0.3),
                scenes.Camera(77, 131, (0, 5, 0), (0, 4, 9), view_fov=12.5, lens_radius=0.0)):
        assert np.array_equal(api.camera_basis(cam), oracle.set_camera(cam))
        n = 20000
        xy = np.stack([rng.randint(0, cam.width, n), rng.randint(0, cam.height, n)], 1).astype(np.int32)
        u3 = rng.random_sample((n, 3)).astype(np.float32)
        assert np.array_equal(r.raygen(cam, xy, u3), oracle.raygen(xy, u3))
    r.close()


@pytest.mark.parametrize("name,maker", [("demo", lambda: scenes.demo_scene(64, 36)),
                                        ("spheres200", lambda: scenes.sphere_field(200, 64, 36))])
def test_level1_golden_ray_batches(name, maker, sky):
    """The committed vectors generated from the reference's own code."""
    g = golden(f"trace_{name}.npz")
    sc = maker()
    r = api.Renderer(sc, sky)
    for rk, pk, tk, nk, hk in (("rays", "prim", "t", "normal", "pos"), ("rays2", "prim2", "t2", "normal2", "pos2")):
        p, t, n, h = r.trace(g[rk])
        assert np.array_equal(p, g[pk])
        hit = p >= 0
        assert np.all(np.abs(t[hit] - g[tk][hit]) <= REL * np.abs(g[tk][hit]))
        assert np.all(np.abs(n[hit] - g[nk][hit]) <= REL)
        assert (t[hit] == g[tk][hit]).mean() > 0.999      # in practice bit-exact
    r.close()


@pytest.mark.parametrize("maker", [lambda: scenes.demo_scene(640, 360), lambda: scenes.sphere_field(1000, 480, 270),
                                   lambda: scenes.stress_field(20000, 480, 270), lambda: scenes.demo_scene(320, 180, seed=99)])
def test_level1_closest_hit_on_path_rays(maker, oracle, port_oracle, sky):
    """Camera rays and the oracle's own bounce rays (segments 1..4 of real paths: grazing exits, glass interiors,
    hollow shells, caps resting on the floor).  Exact ties resolve in the reference's leaf order, so they are NOT
    excluded.  What remains is the graze class north_star allows to be excluded and counted: far from the camera the
    reference's fp32 `b*b - a*c` is so coarse that it reports hits on primitives whose AABB the ray misses -- it
    tests a leaf whenever the PARENT box is hit (bvh.cu:175-180) -- while this BVH culls by the primitive's own box.
    Measured: 10 of 404 123 rays on the 20 k stress scene (camera ~900 units away), 0 on the other scenes."""
    sc = maker()
    r = api.Renderer(sc, sky)
    oracle.load_scene(sc)
    port_oracle.load_scene(sc)
    total_mism, total = 0, 0
    for seg in range(0, 5):
        rays, valid = port_oracle.segment_rays(sc.camera, seg, sample=seg, max_depth=8, seed=31)
        rays = rays[valid][:150000]
        m, exact = _check_trace(r, oracle, rays, max_mismatch_frac=1e-4)
        total_mism += m; total += len(rays)
        assert exact > 0.999
    assert total_mism <= 1e-4 * total, f"{total_mism} graze-class mismatches of {total}"
    print(f"graze-class mismatches: {total_mism} of {total}")
    r.close()


def test_level1_random_rays_inside_and_outside(oracle, sky):
    sc = scenes.stress_field(5000, 64, 36, seed=11)
    r = api.Renderer(sc, sky)
    oracle.load_scene(sc)
    rng = np.random.RandomState(12)
    n = 200000
    o = rng.uniform(-250, 250, (n, 3)); o[:, 1] = rng.uniform(-5, 30, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 20, rng.randint(0, 3)] = 0.0          # axis-parallel components (bvh.cu:182-199 skips those slabs)
    d[: n // 20] /= np.maximum(np.linalg.norm(d[: n // 20], axis=1, keepdims=True), 1e-9)
    rays = np.concatenate([o, d, rng.choice([0.0, 2e-5], (n, 1)), rng.choice([1e30, 50.0], (n, 1))], 1).astype(np.float32)
    m, _ = _check_trace(r, oracle, rays, max_mismatch_frac=2e-5)
    r.close()


def test_level1_edge_scenes(oracle, sky):
    rays = np.array([[0, 5, -10, 0, -0.3, 1, 0, 1e30], [0, 5, -10, 0, 1, 0, 0, 1e30], [0, 1, 0, 0, -1, 0, 0, 1e30],
                     [0, 0.5, -10, 0, 0, 1, 0, 1e30], [0, 0.5, 0, 1, 0, 0, 2e-5, 1e30]], np.float32)
    base = scenes.demo_scene(32, 18)

    def scene_of(types, centers, radii, ypos, heights):
        sc = scenes.demo_scene(32, 18)
        sc.prim_type = np.array(types, np.int32); sc.center = np.array(centers, np.float32).reshape(-1, 3)
        sc.radius = np.array(radii, np.float32); sc.y_pos = np.array(ypos, np.float32)
        sc.height = np.array(heights, np.float32); sc.mat_idx = np.zeros(len(types), np.int32)
        sc.materials = base.materials
        return sc

    cases = [scene_of([scenes.PLATFORM], [[0, 0, 0]], [0], [0.0], [0]),                       # platform only: no tree
             scene_of([scenes.SPHERE], [[0, 0.5, 0]], [1.0], [0], [0]),                        # one-primitive tree
             scene_of([scenes.CYLINDER], [[0, 0.5, 0]], [1.0], [0], [1.0]),
             scene_of([scenes.SPHERE, scenes.SPHERE], [[0, 0.5, 0], [0, 0.5, 0]], [1.0, -0.99], [0, 0], [0, 0]),   # hollow shell
             scene_of([scenes.PLATFORM, scenes.CYLINDER, scenes.CYLINDER], [[0, 0, 0], [0, 0.5, 0], [0.5, 0.5, 0]],
                      [0, 1.0, 1.0], [0, 0, 0], [0, 1.0, 1.0])]                                   # caps resting on the floor: ties
    for sc in cases:
        r = api.Renderer(sc, sky)
        oracle.load_scene(sc)
        _check_trace(r, oracle, rays)
        assert r.trace(np.zeros((0, 8), np.float32))[0].shape == (0,)       # n = 0
        r.close()
    # empty scene: everything misses, render gives pure sky
    r = api.Renderer(); r.set_sky(sky); r.build()
    assert (r.trace(rays)[0] == -1).all()
    cam = scenes.Camera(32, 18)
    r.render(cam, 2, 8, seed=1)
    img = r.read_accum()
    assert (img[..., 3] == 2).all() and img[..., :3].min() > 0.0
    # call-order errors are reported, not ignored
    r.add_sphere((0, 0, 0), 1.0)
    with pytest.raises(api.PtError):
        r.render(cam, 1)
    r.close()


@pytest.mark.parametrize("treelets", [0, 1])
def test_level1_coincident_primitives_and_sah_refit(treelets, oracle, sky):
    """Equal Morton codes and exactly coinciding boxes (100 copies of one sphere, 50 of one cylinder, among 250 random
    primitives): the Karras tie-break by index, the SAH refit (rotations / treelets with zero-gain candidates) and the
    reference's tie order must still give the reference's closest hit on every ray."""
    base = scenes.demo_scene(32, 18)
    rng = np.random.RandomState(4)
    n_rand = 250
    types = [scenes.PLATFORM] + [scenes.SPHERE] * 100 + [scenes.CYLINDER] * 50 + list(rng.choice([scenes.SPHERE, scenes.CYLINDER], n_rand))
    centers = [[0, 0, 0]] + [[3.0, 2.0, -1.0]] * 100 + [[-6.0, 1.5, 4.0]] * 50 + (rng.uniform(-20, 20, (n_rand, 3)) * [1, 0, 1] + [0, 2, 0]).tolist()
    radii = [0] + [2.0] * 100 + [1.5] * 50 + rng.uniform(0.3, 2.0, n_rand).tolist()
    heights = [0] + [0] * 100 + [3.0] * 50 + rng.uniform(0.5, 4.0, n_rand).tolist()
    sc = scenes.demo_scene(32, 18)
    sc.prim_type = np.array(types, np.int32); sc.center = np.array(centers, np.float32).reshape(-1, 3)
    sc.radius = np.array(radii, np.float32); sc.y_pos = np.zeros(len(types), np.float32)
    sc.height = np.array(heights, np.float32); sc.mat_idx = (np.arange(len(types)) % len(base.materials)).astype(np.int32)
    sc.materials = base.materials
    r = api.Renderer(sc, sky)
    r.set_option(api.PT_OPT_SAH_TREELETS, treelets); r.build()
    oracle.load_scene(sc)
    o = rng.uniform(-25, 25, (20000, 3)).astype(np.float32); o[:, 1] = rng.uniform(0.1, 30, 20000)
    target = sc.center[rng.randint(1, len(types), 20000)] + rng.normal(size=(20000, 3)).astype(np.float32)
    d = target - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((20000, 1), 2e-5, np.float32), np.full((20000, 1), 1e30, np.float32)], 1).astype(np.float32)
    mism, exact = _check_trace(r, oracle, rays)
    assert mism == 0 and exact > 0.999
    st = r.stats()
    assert st["bvh_sah"] <= st["bvh_sah_lbvh"]
    r.close()


def test_level1_ties_follow_reference_leaf_order(oracle, sky):
    """Cylinder bottom caps coincide with the floor plane (video_renderer.cpp:78,105-106): rays from inside a
    cylinder towards the floor hit both at exactly the same t; the winner must be the reference's."""
    sc = scenes.demo_scene(64, 36)
    r = api.Renderer(sc, sky)
    oracle.load_scene(sc)
    cyl = np.where(sc.prim_type == scenes.CYLINDER)[0]
    rng = np.random.RandomState(8)
    rays = []
    for c in cyl:
        for _ in range(200):
            ang, rad = rng.uniform(0, 2 * np.pi), rng.uniform(0, 0.9) * abs(sc.radius[c])
            o = [sc.center[c, 0] + rad * np.cos(ang), sc.center[c, 1], sc.center[c, 2] + rad * np.sin(ang)]
            d = np.array([rng.normal() * 0.05, -1.0, rng.normal() * 0.05]); d /= np.linalg.norm(d)
            rays.append(o + d.tolist() + [2e-5, 1e30])
    rays = np.array(rays, np.float32)
    pr = oracle.trace(rays)[0]
    assert (sc.prim_type[pr] == scenes.PLATFORM).any() and (sc.prim_type[pr] == scenes.CYLINDER).any()   # both win somewhere
    _check_trace(r, oracle, rays)
    r.close()


def test_level2_scatter_golden_vectors():
    """Material::EvalAttenuationAndCreateRay with replayed uniforms vs the reference's own code (golden)."""
    g = golden("shade.npz")
    r = api.Renderer()
    b, a, rad, nd = r.shade(g["materials"], g["mat_idx"], g["pos"], g["normal"], g["in_dir"], g["u3"])
    assert np.array_equal(nd, g["draws"])
    assert np.array_equal(a, g["atten"]) and np.array_equal(rad, g["radiance"])
    fin = np.isfinite(g["bounce"]).all(axis=1)
    assert fin.mean() > 0.999
    err = np.abs(b[fin] - g["bounce"][fin]).max(axis=1)
    assert err.max() <= REL, f"max scatter-direction error {err.max():.3e}"
    assert (err == 0).mean() > 0.5
    r.close()


def test_level2_scatter_fresh_inputs_all_materials(oracle):
    rng = np.random.RandomState(21)
    mats = np.zeros(12, scenes.MATERIAL_DTYPE)
    for i in range(12):
        mats[i] = (i % 5, 0, tuple(rng.uniform(0.1, 1.0, 3)), rng.uniform(1.05, 3.2), rng.uniform(0, 2) * (i % 3 == 0),
                   rng.uniform(0.0, 6.0) if i % 2 else rng.uniform(0.0, 1.5), rng.uniform(0, 1))
    n = 200000
    mi = rng.randint(0, 12, n).astype(np.int32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    ind = rng.normal(size=(n, 3)).astype(np.float32); ind /= np.linalg.norm(ind, axis=1, keepdims=True)
    u3 = rng.random_sample((n, 3)).astype(np.float32)
    u3[:100, 0] = np.float32(1.1641532182693481e-10)      # smallest uniform the stream can produce
    u3[100:200, 0] = 1.0                                  # largest
    r = api.Renderer()
    gb, ga, gr, gd = r.shade(mats, mi, np.zeros((n, 3), np.float32), nrm, ind, u3)
    ob, oa, orad, od = oracle.shade(mats, mi, np.zeros((n, 3), np.float32), nrm, ind, u3)
    assert np.array_equal(gd, od) and np.array_equal(gr, orad)
    fin = np.isfinite(ob).all(axis=1) & np.isfinite(gb).all(axis=1)
    assert fin.mean() > 0.9999
    # attenuation flips only where dot(normal, bounce) is within rounding of 0 (material.cu:31,56,92)
    flips = (ga != oa).any(axis=1)
    assert flips.mean() < 1e-4
    err = np.abs(gb[fin] - ob[fin]).max(axis=1)
    assert np.quantile(err, 0.9999) <= REL and (err > 1e-4).mean() < 2e-5, f"max {err.max():.3e}"
    r.close()


def test_sky_lookup_matches_reference_texture_path(oracle, sky):
    g = golden("sky_camera.npz")
    r = api.Renderer(); r.set_sky(sky); r.build()
    got = r.sky_lookup(g["dirs"])
    # hardware bilinear filter (8-bit weights) vs the oracle's emulation: 8-bit texels, so <= ~1/255 per tap step;
    # texels next to the u = 0.25 seam of the quarter-row upload are excluded
    d = g["dirs"]
    u = np.abs(np.arctan(d[:, 1] / d[:, 0]) / 2 / np.pi)
    keep = u < 0.2495
    err = np.abs(got[keep] - g["rgb"][keep]).max(axis=1)
    assert keep.mean() > 0.99 and err.max() < 2.5e-3 and err.mean() < 2e-4
    r.close()


def _render_pair(sc, sky, oracle, spp, seed, depth=8):
    r = api.Renderer(sc, sky)
    r.render(sc.camera, spp, depth, seed=seed)
    img = r.read_accum()
    st = r.stats()
    r.close()
    oracle.load_scene(sc); oracle.set_sky(sky)
    ref, ost = oracle.render(sc.camera, spp, depth, seed=seed)
    return img, ref / spp, st, ost


@pytest.mark.parametrize("maker,spp", [(lambda: scenes.demo_scene(320, 180), 32), (lambda: scenes.sphere_field(1000, 240, 136), 32),
                                       (lambda: scenes.stress_field(20000, 240, 136), 16)])
def test_level3_image_vs_oracle(maker, spp, oracle, sky):
    """Same Philox stream on both sides: single-segment results are bit-comparable; longer paths decorrelate
    chaotically from 1-ulp sin/cos differences (measured 1.0e-2 RMSE per sample at depth 8), so the bound
    scales as 1.6e-2/sqrt(spp) -- 2.5e-4 at the 4096 spp of north_star, four times inside its 1e-3."""
    sc = maker()
    img, ref, st, ost = _render_pair(sc, sky, oracle, spp, seed=17)
    assert (img[..., 3] == spp).all()
    rmse = float(np.sqrt(np.mean((img[..., :3] - ref) ** 2)))
    assert rmse <= 1.6e-2 / np.sqrt(spp), f"rmse {rmse:.3e}"
    gm, rm = img[..., :3].mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    assert np.all(np.abs(gm - rm) <= 0.005 * rm), f"channel means {gm} vs {rm}"
    # the oracle keeps bouncing zero-throughput paths (path_tracer.cu:141-170); the CUDA path ends them (result-identical)
    assert 0.90 * ost["segments"] <= st["segments"] <= 1.002 * ost["segments"]


@pytest.mark.parametrize("w,h", [(320, 180), (322, 181), (37, 23)])
def test_level3_depth1_image_is_tight(w, h, oracle, sky):
    """One segment per path: no chaotic amplification, only the texture unit's filter differs.  320x180 takes the
    8x4-tile order of the camera chunks, 322x181 the row order (width not a multiple of 8), 37x23 a ragged last chunk."""
    sc = scenes.demo_scene(w, h)
    img, ref, st, ost = _render_pair(sc, sky, oracle, 4, seed=3, depth=1)
    assert np.abs(img[..., :3] - ref).max() < 2.5e-3
    assert st["segments"] == ost["segments"] == w * h * 4


def test_properties_at_baseline_size(sky):
    """1920x1080 (BASELINE size): determinism, sample-range additivity and wave-size independence are bit-exact
    because the stream is keyed by (pixel, sample, bounce) and the accumulation is in sample order."""
    sc = scenes.sphere_field(1000, 1920, 1080)
    r = api.Renderer(sc, sky)
    r.render(sc.camera, 8, 8, seed=5); a = r.read_accum().copy()
    r.reset_accum(); r.render(sc.camera, 8, 8, seed=5); b = r.read_accum().copy()
    assert np.array_equal(a, b)
    r.reset_accum()
    r.render(sc.camera, 3, 8, seed=5, first_sample=0); r.render(sc.camera, 5, 8, seed=5, first_sample=3)
    c = r.read_accum().copy()
    assert np.array_equal(a, c)
    r.reset_accum(); r.set_option(api.PT_OPT_WAVE_PATHS, 3 * 1920 * 1080)
    r.render(sc.camera, 8, 8, seed=5); d = r.read_accum().copy()
    assert np.array_equal(a, d)
    r.reset_accum(); r.render(sc.camera, 8, 8, seed=6); e = r.read_accum()
    assert not np.array_equal(a, e)
    st = r.stats()
    assert st["paths"] == 1920 * 1080 * 8 and 2.0 < st["segments"] / st["paths"] < 4.0
    assert np.isfinite(a).all() and a[..., :3].min() >= 0.0
    r.close()


def test_emitters_and_textures(oracle, sky):
    sc = scenes.demo_scene(160, 90)
    sc.materials = sc.materials.copy()
    sc.materials["emit_intensity"][3] = 0.7
    sc.materials["emit_intensity"][5] = 1.5
    img, ref, st, ost = _render_pair(sc, sky, oracle, 16, seed=2)
    rmse = float(np.sqrt(np.mean((img[..., :3] - ref) ** 2)))
    assert rmse < 1.6e-2 / 4 * 1.5 and abs(img[..., :3].mean() - ref.mean()) < 0.005 * ref.mean()
    # kd from a texture sampled at (0,0) (Material::GetKd, material.cu:11-18)
    r = api.Renderer()
    tex = r.create_texture(np.full((4, 4, 4), 128, np.uint8))
    m = api.to_pt_material(sc.materials[0], tex=tex)
    p = r.add_sphere((0, 0, 0), 1.0); r.set_material(p, m); r.build()
    r.close()


def test_reference_cuda_kernels_agree(sky):
    """Device-side oracle: the reference's own TraceRay compiled for sm_100 (FMA-contracted arithmetic)."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpt_ref_cuda.so not built")
    sc = scenes.demo_scene(640, 360)
    R = ref_cuda.RefCuda(); R.load_scene(sc)
    r = api.Renderer(sc, sky)
    rng = np.random.RandomState(6)
    n = 100000
    xy = np.stack([rng.randint(0, 640, n), rng.randint(0, 360, n)], 1).astype(np.int32)
    rays = r.raygen(sc.camera, xy, rng.random_sample((n, 3)).astype(np.float32))
    pg, tg, ng, hg = r.trace(rays)
    pr, tr, nr, hr = R.trace(rays)
    mism = pg != pr
    assert mism.mean() < 1e-4                       # grazing rays where contraction changes the decision
    ok = ~mism & (pr >= 0)
    # nvcc contracts the reference's `b*b - a*c` into FMAs, its host build does not: the reference disagrees with
    # ITSELF by a few ulp of the cancelling terms, so this comparison is statistical (the exact pins are the host ones)
    rel = np.abs(tg[ok] - tr[ok]) / np.abs(tr[ok])
    assert np.quantile(rel, 0.99) <= REL and rel.max() < 1e-3
    assert np.quantile(np.abs(ng[ok] - nr[ok]).max(axis=1), 0.99) <= 1e-4
    r.close()


@pytest.mark.parametrize("w,h,spp,bound", [(160, 90, 1024, 1.0e-3), (96, 54, 4096, 5.0e-4), (640, 360, 4096, 5.0e-4)])
def test_level3_converged_image(w, h, spp, bound, oracle, sky):
    """Convergence of the image error with spp: RMSE falls as 1/sqrt(spp) from the per-sample 1.0e-2.  north_star asks
    for RMSE <= 1e-3 and channel means within 0.5 % at 4096 spp: met at 1024 spp already, and with a factor of two to
    spare at 4096 spp (the demo scene, every material, the sky).  The last case is config C1 at its own size, 640x360,
    against the reference build's render of the same scene (944 M paths: about a minute of the box's host cores)."""
    sc = scenes.demo_scene(w, h)
    img, ref, st, ost = _render_pair(sc, sky, oracle, spp, seed=23)
    rmse = float(np.sqrt(np.mean((img[..., :3] - ref) ** 2)))
    assert rmse <= bound, f"rmse at {spp} spp {rmse:.3e}"
    gm, rm = img[..., :3].mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    assert np.all(np.abs(gm - rm) <= 0.005 * rm)
    print(f"{spp} spp: rmse {rmse:.3e}, means {gm} vs {rm}")


def test_full_size_stress_scene_c3(oracle, port_oracle, sky):
    """Config C3 at full size: 100 k mixed spheres / cylinders, glass-heavy, 1920x1080.  The oracle cannot render it in
    seconds, so: closest-hit parity on a 150 k sample of its real camera / bounce rays, plus render properties."""
    sc = scenes.stress_field(100000, 1920, 1080)
    r = api.Renderer(sc, sky)
    st0 = r.stats()
    assert st0["bvh_nodes"] == sc.n_prims - 2            # every non-platform primitive is a leaf of the LBVH
    oracle.load_scene(sc)
    port_oracle.load_scene(sc)
    small = scenes.Camera(480, 270, sc.camera.origin, sc.camera.look_at)
    total, mism = 0, 0
    for seg in (0, 1, 3):
        rays, valid = port_oracle.segment_rays(small, seg, sample=seg, max_depth=8, seed=77)
        rays = rays[valid][:50000]
        m, exact = _check_trace(r, oracle, rays, max_mismatch_frac=1.5e-3)
        total += len(rays); mism += m
        assert exact > 0.999
    # graze class: the camera is ~1100 units away, where the reference's fp32 quadratics are
    # coarse enough to report hits outside a primitive's own box and the winner depends on ITS traversal order.
    # Measured 37 of 148 225 (2.5e-4); PT_OPT_REF_LEAF_BOXES (test the primitives the reference tests) brings it to 15
    # at 41 % less throughput, which is why it is off by default.
    assert mism <= 5e-4 * total, f"{mism} graze-class mismatches of {total}"
    print(f"C3 graze-class mismatches: {mism} of {total}")
    r2 = api.Renderer(); r2.set_option(api.PT_OPT_REF_LEAF_BOXES, 1); r2.load_scene(sc); r2.set_sky(sky); r2.build()
    rays, valid = port_oracle.segment_rays(small, 0, sample=0, max_depth=8, seed=77)
    rays = rays[valid][:50000]
    m_on = int((r2.trace(rays)[0] != oracle.trace(rays)[0]).sum())
    m_off = int((r.trace(rays)[0] != oracle.trace(rays)[0]).sum())
    assert m_on <= m_off
    r2.close()
    r.render(sc.camera, 4, 8, seed=3); a = r.read_accum().copy()
    st = r.stats()
    assert 3.0 < st["segments"] / st["paths"] < 4.5      # glass-heavy: long paths (reference walk: 3.25 for spheres only)
    r.reset_accum(); r.render(sc.camera, 4, 8, seed=3)
    assert np.array_equal(a, r.read_accum())
    assert np.isfinite(a).all()
    r.close()


def test_level3_image_c3_full_scene(oracle, port_oracle, sky):
    """Config C3's scene (all 128 836 primitives, glass-heavy) at image level against the reference build: reduced
    resolution so that the host render takes seconds, same camera, same Philox stream.
    Asserted: per-channel means within 0.5 % (measured 2e-5 relative); whole-image RMSE <= 7e-3; the graze class stays
    below 1e-3 of the primary rays; testing the primitives the reference tests (PT_OPT_REF_LEAF_BOXES) lowers the RMSE.
    NOT met on this scene, and said so in DESIGN.md section 4: north_star's RMSE <= 1e-3 at 4096 spp.  The RMSE does not
    fall as 1/sqrt(spp) here (5.1e-3 at 64 spp, 4.7e-3 at 256): the camera is 1100 units away, where the reference's
    fp32 quadratic reports hits on spheres that the ray passes by (b*b - a*c cancels to noise of ~0.1 units) while this
    path culls by the primitive's own box and keeps the geometrically nearest hit -- the graze class that north_star
    excludes from the id comparison and wants counted.  Without pixel jitter (motional_camera.cu:204-207) and with
    near-specular glass (alpha = 1000^smoothness >= 1e6) a pixel whose primary ray OR specular chain meets such a case
    shows another surface in every sample: a floor that no sample count averages away.  On C1 / C2 (camera 200 units
    away, no such cases: bench.py counts 0 of 300 k rays) the bound is met with a factor of two to spare."""
    sc = scenes.stress_field(100000, 240, 136)
    oracle.load_scene(sc); port_oracle.load_scene(sc)
    rays, valid = port_oracle.segment_rays(sc.camera, 0, sample=0, max_depth=8, seed=31)
    r = api.Renderer(sc, sky)
    same_primary = (r.trace(rays)[0] == oracle.trace(rays)[0]).reshape(136, 240)
    r.close()
    n_graze = int((~same_primary).sum())
    assert n_graze <= 1e-3 * same_primary.size, f"{n_graze} graze-class primary rays of {same_primary.size}"
    out = {}
    for spp in (64, 256):
        img, ref, st, ost = _render_pair(sc, sky, oracle, spp, seed=31)
        assert (img[..., 3] == spp).all()
        d2 = (img[..., :3] - ref) ** 2
        rmse_all, rmse_same = float(np.sqrt(d2.mean())), float(np.sqrt(d2[same_primary].mean()))
        gm, rm = img[..., :3].mean(axis=(0, 1)), ref.mean(axis=(0, 1))
        print(f"C3 scene at 240x136 x {spp} spp: rmse {rmse_all:.3e} over all pixels, {rmse_same:.3e} without the {n_graze} "
              f"pixels whose primary ray is in the graze class; means {gm} vs {rm}")
        assert np.all(np.abs(gm - rm) <= 0.005 * rm), f"channel means {gm} vs {rm}"
        assert rmse_all <= 7e-3, f"rmse {rmse_all:.3e}"
        assert 0.90 * ost["segments"] <= st["segments"] <= 1.002 * ost["segments"]
        out[spp] = (rmse_all, rmse_same, ref)
    # the same frame with the reference's leaf boxes: fewer graze-class decisions, lower floor
    r2 = api.Renderer(); r2.set_option(api.PT_OPT_REF_LEAF_BOXES, 1); r2.load_scene(sc); r2.set_sky(sky); r2.build()
    r2.render(sc.camera, 256, 8, seed=31)
    img2 = r2.read_accum()[..., :3]
    r2.close()
    rmse_ref_boxes = float(np.sqrt(np.mean((img2 - out[256][2]) ** 2)))
    print(f"C3 scene, 256 spp, PT_OPT_REF_LEAF_BOXES = 1: rmse {rmse_ref_boxes:.3e} (default {out[256][0]:.3e})")
    assert rmse_ref_boxes < out[256][0]


def test_4k_frame_properties_c4(sky):
    """Config C4 frame size (3840x2160): sample-range additivity is what spp-sharding across GPUs relies on."""
    sc = scenes.sphere_field(1000, 3840, 2160)
    r = api.Renderer(sc, sky)
    r.render(sc.camera, 4, 8, seed=9); full = r.read_accum().copy()
    r.reset_accum()
    r.render(sc.camera, 2, 8, seed=9, first_sample=2)      # "rank 1" first: the SUM is order-dependent only in rounding
    r.render(sc.camera, 2, 8, seed=9, first_sample=0)
    split = r.read_accum()
    assert (split[..., 3] == 4).all()
    assert np.allclose(split[..., :3], full[..., :3], rtol=2e-6, atol=2e-6)
    assert full.shape == (2160, 3840, 4) and np.isfinite(full).all()
    r.close()


def test_material_texture_is_sampled_at_origin(sky):
    """Material::GetKd with have_tex_ (material.cu:11-18): the reference samples the texture at (0,0) for every hit
    (no UVs exist), so a constant-colour texture must render exactly like kd = that colour, and the attenuation of a
    non-constant texture is its texel (0,0) under the linear / mirror sampler."""
    sc = scenes.demo_scene(160, 90)
    col = np.array([200, 120, 40], np.uint8)
    r1 = api.Renderer(); r2 = api.Renderer()
    tex = r1.create_texture(np.tile(np.append(col, 255).astype(np.uint8), (8, 8, 1)))
    for r, use_tex in ((r1, True), (r2, False)):
        r.add_platform(0.0)
        m = scenes.np.zeros(1, scenes.MATERIAL_DTYPE)[0]
        m["type"] = scenes.DIFFUSE; m["kd"] = tuple(col.astype(np.float32) / np.float32(255))
        r.set_material(0, api.to_pt_material(m, tex=tex if use_tex else 0))
        p = r.add_sphere((0, 20, 0), 20.0)
        r.set_material(p, api.to_pt_material(sc.materials[0]))
        r.set_sky(sky); r.build()
        r.render(sc.camera, 8, 8, seed=4)
    a, b = r1.read_accum(), r2.read_accum()
    assert np.abs(a[..., :3] - b[..., :3]).max() < 1e-6 and a[..., :3].mean() > 0.05
    r1.close(); r2.close()


def test_level1_randomized_small_scenes(oracle, sky):
    """Robustness of the conservative culling (centre/half-extent nodes, FMA slabs, slack): random scenes over six
    decades of scale, negative radii / heights, stacked platforms, rays parallel to axes, origins inside primitives
    and far away.  Every mismatch is printed; none is allowed beyond exact-tie artefacts of degenerate inputs."""
    rng = np.random.RandomState(2027)
    base = scenes.demo_scene(32, 18)
    total = bad = 0
    for trial in range(40):
        scale = 10.0 ** rng.uniform(-2, 4)
        n = int(rng.randint(1, 40))
        types = rng.choice([scenes.SPHERE, scenes.CYLINDER, scenes.PLATFORM], n, p=[0.55, 0.35, 0.10])
        sc = scenes.demo_scene(32, 18)
        sc.prim_type = types.astype(np.int32)
        sc.center = (rng.uniform(-1, 1, (n, 3)) * scale).astype(np.float32)
        sc.radius = (rng.uniform(0.01, 0.4, n) * scale * rng.choice([1, 1, 1, -1], n)).astype(np.float32)
        # heights stay positive: a negative-height CYLINDER gets an inverted y-interval from the reference's
        # GetAABBMin/Max (object.cu:145,164) and its own BVH then culls hits on it depending on the tree -- that is the
        # reference failing on input it never generates (its negative heights sit on sphere-typed inner objects)
        sc.height = (rng.uniform(0.01, 0.8, n) * scale).astype(np.float32)
        sc.y_pos = (rng.uniform(-1, 1, n) * scale).astype(np.float32)
        sc.mat_idx = rng.randint(0, len(base.materials), n).astype(np.int32)
        r = api.Renderer(sc, sky)
        oracle.load_scene(sc)
        m = 20000
        o = rng.uniform(-1.5, 1.5, (m, 3)) * scale
        o[: m // 10] = sc.center[rng.randint(0, n, m // 10)] + rng.normal(size=(m // 10, 3)) * 0.01 * scale   # inside / on primitives
        o[m // 10: m // 5] *= 50.0                                                                       # far away
        d = rng.normal(size=(m, 3))
        ax = rng.randint(0, 3, m // 8)
        d[np.arange(m // 8), ax] = 0.0                                                                    # axis-parallel components
        d[m // 8: m // 6] = np.eye(3)[rng.randint(0, 3, m // 6 - m // 8)] * rng.choice([-1.0, 1.0], (m // 6 - m // 8, 1))
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
        rays = np.concatenate([o, d, rng.choice([0.0, 2e-5], (m, 1)), np.full((m, 1), 1e30)], 1).astype(np.float32)
        pg, tg, ng, hg = r.trace(rays)
        pr, tr, nr, hr = oracle.trace(rays)
        mism = np.where(pg != pr)[0]
        # a mismatch is tolerated only when both sides found a hit at (nearly) the same distance: a tie between
        # coincident surfaces of a degenerate random scene resolved by a different but equally near primitive
        hard = [i for i in mism if not (pg[i] >= 0 and pr[i] >= 0 and abs(tg[i] - tr[i]) <= 1e-4 * max(abs(tr[i]), 1e-30))]
        ok = (pg == pr) & (pr >= 0)
        assert np.all(np.abs(tg[ok] - tr[ok]) <= REL * np.abs(tr[ok]))
        total += m; bad += len(hard)
        for i in hard[:3]:
            print(f"trial {trial} scale {scale:.3g}: ray {rays[i].tolist()} gpu {pg[i]} (type {types[pg[i]] if pg[i] >= 0 else None}) {tg[i]} "
                  f"ref {pr[i]} (type {types[pr[i]] if pr[i] >= 0 else None}) {tr[i]}")
        r.close()
    assert bad <= 2e-5 * total, f"{bad} hard mismatches of {total}"
    print(f"randomized scenes: {bad} hard mismatches of {total}")


def _stats_vs_reference_cuda(sky, lib_path, spp=256):
    from oracle import ref_cuda
    sc = scenes.demo_scene(640, 368)                 # multiples of 16: the reference's grid covers the whole image
    R = ref_cuda.RefCuda(); R.load_scene(sc); R.set_sky(sky)
    ref = R.render(sc.camera, spp, warmup=0, seed=99, want_image=True)["sum_rgb"] / spp
    r = api.Renderer(sc, sky, lib_path=lib_path)
    r.render(sc.camera, spp, 8, seed=1234); img = r.read_accum()[..., :3].copy()
    r.reset_accum(); r.render(sc.camera, spp, 8, seed=4321); img2 = r.read_accum()[..., :3].copy()
    r.close()
    gm, rm = img.mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    rmse_ref = float(np.sqrt(np.mean((img - ref) ** 2)))
    rmse_self = float(np.sqrt(np.mean((img - img2) ** 2)))      # noise floor: two independent renders of this path
    return gm, rm, rmse_ref, rmse_self


def test_level3_statistics_vs_reference_cuda_megakernel(sky):
    """End to end against the reference's OWN CUDA kernels (SamplePixel megakernel, cuRAND XORWOW streams, device libm,
    the same texture unit), rebuilt for sm_100.  The streams are independent, so this is the statistical check of
    north_star: per-channel means within 0.5 %.  The RMSE is NOT pure Monte-Carlo noise for the default library: nvcc
    contracts the reference's `b*b - a*c`, `origin + t*dir`, ... into FMAs, its host build (the pinned oracle, which the
    default library follows bit for bit) does not, and the self-intersection pattern at tmin = 2e-5 on large spheres
    ("acne" stripes) differs between the two -- the reference disagrees with itself by ~1.2e-2 RMS on those surfaces.
    The FMA-contracted variant of this library removes that term (next test)."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpt_ref_cuda.so not built")
    gm, rm, rmse_ref, rmse_self = _stats_vs_reference_cuda(sky, api.LIB_PATH)
    assert np.all(np.abs(gm - rm) <= 0.005 * rm), f"channel means {gm} vs reference CUDA {rm}"
    assert rmse_ref <= 1.5 * rmse_self, f"rmse vs reference CUDA {rmse_ref:.3e}, own noise floor {rmse_self:.3e}"
    print(f"default library vs reference CUDA megakernel at 256 spp: means {gm} vs {rm}; rmse {rmse_ref:.3e} (noise floor {rmse_self:.3e})")


def test_fma_variant_matches_reference_cuda_build(sky, port_oracle):
    """libpt_b200_fma.so = the same sources compiled with --fmad=true, i.e. contracted by nvcc the way it contracts the
    reference's cuSrc/*.cu.  Against the reference's CUDA kernels: no primitive-id mismatches on real path rays, most
    distances bit-identical, and a 256-spp image that differs from the reference megakernel's by Monte-Carlo noise only."""
    import os
    from oracle import ref_cuda
    if not ref_cuda.available() or not os.path.exists(api.LIB_PATH_FMA):
        pytest.skip("needs oracle/_ref/libpt_ref_cuda.so and libpt_b200_fma.so")
    sc = scenes.demo_scene(640, 368)
    R = ref_cuda.RefCuda(); R.load_scene(sc); R.set_sky(sky)
    port_oracle.load_scene(sc)
    r = api.Renderer(sc, sky, lib_path=api.LIB_PATH_FMA)
    assert b"fma" in r.lib.pt_version()
    total = mism = 0
    for seg in (0, 1, 2):
        rays, valid = port_oracle.segment_rays(sc.camera, seg, 0, 8, seed=5)
        rays = rays[valid][:200000]
        pg, tg, ng, hg = r.trace(rays)
        pr, tr, nr, hr = R.trace(rays)
        ok = (pg == pr) & (pr >= 0)
        total += len(rays); mism += int((pg != pr).sum())
        assert (tg[ok] == tr[ok]).mean() > 0.85
        assert np.quantile(np.abs(tg[ok] - tr[ok]) / np.abs(tr[ok]), 0.999) <= REL
    assert mism <= 2e-5 * total, f"{mism} id mismatches of {total} against the reference's CUDA TraceRay"
    r.close()
    gm, rm, rmse_ref, rmse_self = _stats_vs_reference_cuda(sky, api.LIB_PATH_FMA)
    assert np.all(np.abs(gm - rm) <= 0.005 * rm)
    assert rmse_ref <= 1.10 * rmse_self, f"rmse vs reference CUDA {rmse_ref:.3e}, own noise floor {rmse_self:.3e}"
    print(f"fma variant vs reference CUDA megakernel: id mismatches {mism}/{total}; rmse {rmse_ref:.3e} (noise floor {rmse_self:.3e})")
