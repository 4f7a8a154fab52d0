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
