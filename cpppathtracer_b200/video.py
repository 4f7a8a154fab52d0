"""Headless video front-end: replaces the reference's VideoRenderer / MainWnd frame loop
(cppSrc/video_renderer.cpp:147-280, cppSrc/main_wnd.cpp) with a scripted camera path and moving objects.

Config C5: a camera orbit made of the reference's own controls (RotateAroundRight + MoveEyeLeft per frame,
motional_camera.cu:76-82,155-164) over a sphere field in which 10 % of the spheres bounce
(center.y = r + A*|sin(w*f)|), so every frame re-uploads the primitives and rebuilds the LBVH on the device
(pt_update_prim + pt_build, the replacement of SceneBVH::UpdateObject, bvh.cu:144-157).  Frames are independent:
rank g of G renders frames f with f % G == g; there is no exchange.
"""
from __future__ import annotations

import copy
import os
import time

import numpy as np

from . import api, scenes
from .distributed import frames_of_rank


def camera_path(cam0, n_frames: int, rot_per_frame: float = 0.02, move_per_frame: float = 0.02):
    """Camera poses of the orbit, frame 0 first.  Pure host math through the C ABI (no device needed)."""
    cams, cam = [], copy.deepcopy(cam0)
    for _ in range(n_frames):
        cams.append(copy.deepcopy(cam))
        api.camera_control(cam, api.CAM_ROT_RIGHT, rot_per_frame)
        api.camera_control(cam, api.CAM_LEFT, move_per_frame)
    return cams


def moving_spheres(scene, frame: int, fraction: float = 0.1, amplitude: float = 4.0, omega: float = 0.15, seed: int = 7):
    """(prim ids, centres) of the bouncing spheres at `frame`; the same subset for every frame."""
    rng = np.random.RandomState(seed)
    cand = np.where(scene.prim_type == scenes.SPHERE)[0]
    ids = np.sort(rng.choice(cand, max(1, int(len(cand) * fraction)), replace=False))
    phase = rng.uniform(0, np.pi, len(ids))
    c = scene.center[ids].copy()
    c[:, 1] = scene.radius[ids] + np.float32(amplitude) * np.abs(np.sin(omega * frame + phase)).astype(np.float32)
    return ids, c.astype(np.float32)


def write_ppm(path: str, mean_rgb: np.ndarray):
    img = (np.clip(mean_rgb, 0.0, 1.0) * 255.99).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6 %d %d 255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def render_video(scene, sky, n_frames: int = 120, spp: int = 64, max_depth: int = 8, seed: int = 0, rank: int = 0,
                 world: int = 1, out_dir: str | None = None, renderer: api.Renderer | None = None):
    """Render this rank's frames.  Returns per-frame stats: frame, build_ms, render_ms, mpaths_s, wall_ms."""
    r = renderer or api.Renderer(scene, sky)
    cams = camera_path(scene.camera, n_frames)
    shape = (scene.camera.height, scene.camera.width, 4)
    try:                                    # page-locked frames make the read-back a true DMA; plain numpy works too
        import torch
        hosts = [torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
    except Exception:
        hosts = [np.empty(shape, np.float32) for _ in range(2)]
    stats = []
    pending = None                          # (frame, buffer index) whose read-back is in flight

    def deliver():
        # the frame handed out here was copied to the host while the NEXT frame was being built and rendered
        # (pt_read_accum_begin / _end, like the reference's render thread hands a finished frame to its callback)
        f_done, k_done = pending
        r.read_accum_end()
        os.makedirs(out_dir, exist_ok=True)
        write_ppm(os.path.join(out_dir, f"frame_{f_done:04d}.ppm"), hosts[k_done][..., :3])

    for n, f in enumerate(frames_of_rank(n_frames, world, rank)):
        t0 = time.perf_counter()
        ids, centres = moving_spheres(scene, f)
        for p, c in zip(ids, centres):
            r.update_prim(int(p), c, float(scene.radius[p]))
        r.build()
        r.reset_accum()
        r.render(cams[f], spp, max_depth, seed=seed)
        st = r.stats()
        if out_dir:
            r.read_accum_begin(hosts[n & 1])
            if pending is not None:
                deliver()
            pending = (f, n & 1)
        stats.append({"frame": f, "build_ms": st["build_ms"], "render_ms": st["render_ms"],
                      "mpaths_s": st["paths"] / st["render_ms"] / 1e3, "wall_ms": (time.perf_counter() - t0) * 1e3})
    if pending is not None:
        deliver()
    return stats
