"""-m gpu: dynamic scenes (pt_update_prim + pt_build replace SceneBVH::UpdateObject,
bvh.cu:122-157) and the headless video front-end (row 4, config C5)."""
import numpy as np
import pytest

from cpppathtracer_b200 import api, scenes, video

pytestmark = pytest.mark.gpu


def test_update_prim_rebuild_matches_oracle(port_oracle, sky):
    from oracle import ref_host
    expected = ref_host.RefHost() if ref_host.available() else port_oracle      # ids / t / normals from the reference build
    sc = scenes.sphere_field(500, 320, 180)
    r = api.Renderer(sc, sky)
    rng = np.random.RandomState(3)
    for frame in range(3):
        ids, centres = video.moving_spheres(sc, frame, fraction=0.3)
        for p, c in zip(ids, centres):
            r.update_prim(int(p), c, float(sc.radius[p]))
            sc.center[p] = c
        with pytest.raises(api.PtError):          # edits need a rebuild before the next render
            r.render(sc.camera, 1)
        r.build()
        port_oracle.load_scene(sc)
        if expected is not port_oracle:
            expected.load_scene(sc)
        rays, valid = port_oracle.segment_rays(sc.camera, frame % 3, sample=frame, seed=5)
        rays = rays[valid]
        pg, tg, ng, _ = r.trace(rays)
        pr, tr, nr, _ = expected.trace(rays)
        assert np.array_equal(pg, pr) and np.array_equal(tg, tr) and np.array_equal(ng, nr)
    assert r.stats()["build_ms"] < 5.0
    r.close()


def test_video_frames_round_robin(sky, tmp_path):
    sc = scenes.sphere_field(200, 160, 90)
    a = video.render_video(sc, sky, n_frames=6, spp=4, rank=0, world=1, out_dir=str(tmp_path))
    assert [s["frame"] for s in a] == list(range(6))
    assert sorted(p.name for p in tmp_path.iterdir()) == [f"frame_{f:04d}.ppm" for f in range(6)]
    b0 = video.render_video(sc, sky, n_frames=6, spp=4, rank=0, world=2)
    b1 = video.render_video(sc, sky, n_frames=6, spp=4, rank=1, world=2)
    assert [s["frame"] for s in b0] == [0, 2, 4] and [s["frame"] for s in b1] == [1, 3, 5]
    # frames differ (camera and objects move) and are reproducible
    f0 = open(tmp_path / "frame_0000.ppm", "rb").read(); f1 = open(tmp_path / "frame_0005.ppm", "rb").read()
    assert f0 != f1
    d2 = tmp_path / "again"
    video.render_video(sc, sky, n_frames=1, spp=4, out_dir=str(d2))
    assert open(d2 / "frame_0000.ppm", "rb").read() == f0


@pytest.mark.parametrize("maker", [lambda: scenes.sphere_field(1000, 320, 180), lambda: scenes.demo_scene(320, 180),
                                   lambda: scenes.stress_field(5000, 320, 180)])
def test_sah_refit_lowers_cost_and_keeps_every_hit(maker, sky):
    """PT_OPT_SAH_PASSES (tree rotations after the Karras build): the closest hit does not depend on the tree, so
    hits and images are identical to the plain LBVH's; the SAH cost and the node visits go down; the tree stays a
    tree over the same primitives (every primitive is still found by rays aimed at it)."""
    sc = maker()
    out = {}
    for passes in (0, 4, -2):                                       # -2: two passes of treelet restructuring
        r = api.Renderer(sc, sky)
        r.set_option(api.PT_OPT_SAH_PASSES, abs(passes)); r.set_option(api.PT_OPT_SAH_TREELETS, int(passes < 0)); r.build()
        r.render(sc.camera, 4, 8, seed=3)
        img = r.read_accum().copy(); st = r.stats()
        rng = np.random.RandomState(1)
        tree = np.where(sc.prim_type != scenes.PLATFORM)[0]
        targets = sc.center[tree]                                  # one ray aimed at the centre of every primitive
        org = targets + rng.normal(size=targets.shape).astype(np.float32) * 50 + np.float32([0, 80, 0])
        d = targets - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([org, d, np.zeros((len(org), 1), np.float32), np.full((len(org), 1), 1e30, np.float32)], 1).astype(np.float32)
        prim, t, n, _ = r.trace(rays)
        r.set_option(api.PT_OPT_COUNT_STEPS, 1); r.reset_accum(); r.render(sc.camera, 1, 8, seed=3); cs = r.stats()
        out[passes] = (img, prim, t, st, cs["node_steps"] / cs["segments"])
        r.close()
    img0, prim0, t0, st0, v0 = out[0]
    img4, prim4, t4, st4, v4 = out[4]
    assert np.array_equal(prim0, prim4) and np.array_equal(t0, t4)
    assert (prim4 >= 0).all()                                       # aimed at a centre: something is always hit
    assert np.array_equal(img0, img4)
    assert st4["bvh_sah_lbvh"] == pytest.approx(st0["bvh_sah_lbvh"], rel=1e-4) and st0["bvh_sah"] == st0["bvh_sah_lbvh"]
    assert st4["bvh_sah"] < st4["bvh_sah_lbvh"] and v4 < v0
    imgt, primt, tt, stt, vt = out[-2]
    assert np.array_equal(prim0, primt) and np.array_equal(t0, tt) and np.array_equal(img0, imgt)
    assert stt["bvh_sah"] < st4["bvh_sah"] and vt < v0


def test_update_prim_matches_the_references_own_update_object(ref_oracle, sky):
    """pt_update_prim + pt_build against SceneBVH::UpdateObject of the reference build itself (bvh.cu:144-157: refit,
    no re-sort), on a scene full of exact ties: cylinders standing on the floor plane and hollow glass shells.  The
    reference's leaf visiting order survives its updates, so the product keeps the tie order of its last full build
    -- and skips the host-side restatement of the reference's sort on every rebuild after the first."""
    sc = scenes.demo_scene(64, 36)
    r = api.Renderer(sc, sky)
    ref_oracle.load_scene(sc)
    order0 = ref_oracle.leaf_order().copy()
    rng = np.random.RandomState(12)
    cyl = np.where(sc.prim_type == scenes.CYLINDER)[0]
    inner = np.where(sc.radius < 0)[0]                        # hollow shells: same centre as the primitive before them

    def tie_rays():
        rays = []
        for c in cyl:                                          # down through a cylinder onto the floor: cap == floor plane
            for _ in range(60):
                ang, rad = rng.uniform(0, 2 * np.pi), rng.uniform(0, 0.9) * abs(sc.radius[c])
                o = [sc.center[c, 0] + rad * np.cos(ang), sc.center[c, 1], sc.center[c, 2] + rad * np.sin(ang)]
                d = np.array([rng.normal() * 0.05, -1.0, rng.normal() * 0.05]); d /= np.linalg.norm(d)
                rays.append(o + d.tolist() + [2e-5, 1e30])
        for c in inner:                                        # from far away through a hollow shell: inner and outer surface
            for _ in range(60):
                o = sc.center[c] + np.float32([rng.uniform(-1, 1) * 900, rng.uniform(200, 600), rng.uniform(-1, 1) * 900])
                tgt = sc.center[c] + rng.normal(size=3) * abs(sc.radius[c]) * 0.5
                d = tgt - o; d /= np.linalg.norm(d)
                rays.append(o.tolist() + d.tolist() + [0.0, 1e30])
        return np.array(rays, np.float32)

    for step in range(3):
        move = np.concatenate([rng.choice(cyl, 6, replace=False), rng.choice(inner, min(4, len(inner)), replace=False)])
        for p in move:
            delta = np.float32([rng.uniform(-30, 30), 0.0, rng.uniform(-30, 30)])
            group = [p, p - 1] if sc.radius[p] < 0 else [p]    # a shell moves together with its outer primitive
            for q in group:
                sc.center[q] = sc.center[q] + delta
                r.update_prim(int(q), sc.center[q], float(sc.radius[q]), 0.0, float(sc.height[q]))
                ref_oracle.update_object(int(q), sc.center[q], float(sc.radius[q]), 0.0, float(sc.height[q]))
        r.build()
        assert np.array_equal(ref_oracle.leaf_order(), order0)          # the reference did not re-sort
        rays = np.concatenate([tie_rays(), ref_oracle.raygen(np.stack([rng.randint(0, 64, 4000), rng.randint(0, 36, 4000)], 1),
                                                             rng.random_sample((4000, 3)))])
        pg, tg, ng, _ = r.trace(rays)
        pr, tr, nr, _ = ref_oracle.trace(rays)
        assert (sc.prim_type[pr[pr >= 0]] == scenes.PLATFORM).any() and (sc.prim_type[pr[pr >= 0]] == scenes.CYLINDER).any()
        assert np.array_equal(pg, pr) and np.array_equal(tg, tr) and np.array_equal(ng, nr), f"step {step}: {(pg != pr).sum()} id mismatches"
    r.close()


def test_rebuild_after_update_is_cheap_at_100k_primitives(sky):
    """pt_build after one pt_update_prim at C3's size: no host-side sort any more (it took 35 ms in round 1), only the moved
    primitives are restaged and the two geometry arrays re-uploaded from page-locked memory: 1.0 ms measured (0.85 ms of it the device LBVH)."""
    import time
    sc = scenes.stress_field(100000, 320, 180)
    r = api.Renderer(sc, sky)
    walls = []
    for k in range(5):
        r.update_prim(7, sc.center[7] + np.float32([0.0, 0.0, 0.25 * k]), float(sc.radius[7]), 0.0, float(sc.height[7]))
        t0 = time.perf_counter(); r.build(); walls.append((time.perf_counter() - t0) * 1e3)
    r.render(sc.camera, 1, 8, seed=1)
    assert min(walls) < 3.0, f"pt_build after an update: {walls} ms"
    r.close()


@pytest.mark.parametrize("n", [1, 2, 31, 256, 2047, 2048, 2049, 8192, 8193, 200001])
def test_lbvh_radix_sort_is_a_stable_sort(n):
    """The build's hand-written radix sort (one-block path up to 2048 keys, histogram / scan / scatter passes above) against
    numpy's stable argsort: 30-bit codes like the Morton keys, plus heavy duplication (coincident primitives share a code and
    the Karras tie-break relies on their input order) and full 32-bit keys."""
    r = api.Renderer()
    rng = np.random.RandomState(n)
    for codes in (rng.randint(0, 1 << 30, n).astype(np.uint32), rng.randint(0, 7, n).astype(np.uint32),
                  rng.randint(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32), np.zeros(n, np.uint32),
                  np.arange(n, 0, -1).astype(np.uint32)):
        ks, vs = r.sort_code_pairs(codes)
        order = np.argsort(codes, kind="stable")
        assert np.array_equal(vs, order.astype(np.uint32)) and np.array_equal(ks, codes[order])
    r.close()


def test_pipelined_read_back_delivers_the_same_frames(sky):
    """pt_read_accum_begin / _end (copy of frame i on the copy stream under the render of frame i+1) == pt_read_accum, frame by frame."""
    import torch
    sc = scenes.sphere_field(300, 320, 180)
    a, b = api.Renderer(sc, sky), api.Renderer(sc, sky)
    want = []
    for f in range(5):
        a.reset_accum(); a.render(sc.camera, 3, 8, seed=f); want.append(a.read_accum().copy())
    bufs = [torch.empty((180, 320, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    got = []
    for f in range(5):
        b.reset_accum(); b.render(sc.camera, 3, 8, seed=f, sync=False)
        b.read_accum_begin(bufs[f & 1])
        if f >= 1:
            b.read_accum_end(); got.append(bufs[(f - 1) & 1].copy())
    with pytest.raises(api.PtError):                       # one in flight + two more = three: refused
        b.read_accum_begin(bufs[1]); b.read_accum_begin(bufs[0])
    b.read_accum_end(); got.append(bufs[0].copy())
    b.read_accum_end()
    with pytest.raises(api.PtError):
        b.read_accum_end()                                 # nothing in flight
    assert len(got) == 5 and all(np.array_equal(g, w) for g, w in zip(got, want))
    a.close(); b.close()
