"""-m gpu: dynamic scenes (pt_update_prim + pt_build replace SceneBVH::UpdateObject,
bvh.cu:122-157) and the headless video front-end (row 4, config C5)."""
import numpy as np
import pytest

from cpppathtracer_b200 import api, scenes, video

pytestmark = pytest.mark.gpu


def test_update_prim_rebuild_matches_oracle(port_oracle, sky):
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
        rays, valid = port_oracle.segment_rays(sc.camera, frame % 3, sample=frame, seed=5)
        rays = rays[valid]
        pg, tg, ng, _ = r.trace(rays)
        pr, tr, nr, _ = port_oracle.trace(rays)
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
