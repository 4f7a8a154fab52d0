"""Generates tests/golden/*.npz from the reference's OWN code (oracle/_ref/libpt_ref_host.so, i.e.
/root/reference/cuSrc/{object,material,bvh,motional_camera}.cu compiled for the host).  Run in the
build container (where /root/reference exists):   python tests/golden/make_golden.py
The reference has no golden vectors of its own; these are the pins.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cpppathtracer_b200 import scenes  # noqa: E402
from oracle.ref_host import RefHost  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    from PIL import Image
    R = RefHost()
    rng = np.random.RandomState(2026)
    # the valid (left) quarter of the reference's sky.png under its quarter-row upload (textures.cu:32)
    sky_full = np.array(Image.open("/root/reference/textures/sky.png").convert("RGBA"))
    quarter = np.ascontiguousarray(sky_full[:, :320, :3])
    np.savez_compressed(os.path.join(OUT, "sky_quarter.npz"), rgb=quarter, full_width=np.int32(sky_full.shape[1]))
    sky = np.zeros_like(sky_full); sky[:, :320, :3] = quarter; sky[..., 3] = 255

    for name, sc in (("demo", scenes.demo_scene(64, 36)), ("spheres200", scenes.sphere_field(200, 64, 36))):
        R.load_scene(sc); R.set_sky(sky)
        n = 4096
        xy = np.stack([rng.randint(0, sc.camera.width, n), rng.randint(0, sc.camera.height, n)], 1).astype(np.int32)
        u3 = rng.random_sample((n, 3)).astype(np.float32)
        rays = R.raygen(xy, u3)
        prim, t, nrm, pos = R.trace(rays)
        hit = prim >= 0
        d2 = rng.normal(size=(int(hit.sum()), 3)).astype(np.float32)
        d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        rays2 = np.concatenate([pos[hit], d2, np.full((len(d2), 1), 2e-5), np.full((len(d2), 1), 1e30)], 1).astype(np.float32)
        prim2, t2, nrm2, pos2 = R.trace(rays2)
        img, st = R.render(sc.camera, 2, 8, seed=77)
        basis = R.set_camera(sc.camera)
        np.savez_compressed(os.path.join(OUT, f"trace_{name}.npz"), xy=xy, u3=u3, rays=rays, prim=prim, t=t, normal=nrm,
                            pos=pos, rays2=rays2, prim2=prim2, t2=t2, normal2=nrm2, pos2=pos2, image_sum=img,
                            image_spp=np.int32(2), image_seed=np.int64(77), segments=np.int64(st["segments"]),
                            leaf_order=R.leaf_order(), camera_basis=basis)
    # shading vectors over the demo material library + edge materials
    sc = scenes.demo_scene(64, 36)
    mats = np.concatenate([sc.materials, np.zeros(4, scenes.MATERIAL_DTYPE)])
    mats[-4] = (scenes.GLASS, 0, (1, 1, 1), 1.5, 0.0, 0.3, 0.0)      # rough glass: lobe far from degenerate
    mats[-3] = (scenes.METAL, 0, (.9, .8, .7), 0.0, 2.0, 0.5, 0.0)   # emissive rough metal
    mats[-2] = (scenes.MIRROR, 0, (.5, .6, .7), 0.0, 0.0, 1.2, 0.5)
    mats[-1] = (scenes.TEST, 0, (.2, .3, .4), 0.0, 0.5, 0.0, 0.0)    # Test -> Diffuse (material.cu:160-161)
    n = 8192
    mi = rng.randint(0, len(mats), n).astype(np.int32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    ind = rng.normal(size=(n, 3)).astype(np.float32); ind /= np.linalg.norm(ind, axis=1, keepdims=True)
    pos = rng.normal(size=(n, 3)).astype(np.float32) * 50
    u3 = rng.random_sample((n, 3)).astype(np.float32)
    b, a, r, nd = R.shade(mats, mi, pos, nrm, ind, u3)
    np.savez_compressed(os.path.join(OUT, "shade.npz"), materials=mats, mat_idx=mi, normal=nrm, in_dir=ind, pos=pos, u3=u3,
                        bounce=b, atten=a, radiance=r, draws=nd)
    # sky lookups and camera controls
    dirs = rng.normal(size=(4096, 3)).astype(np.float32); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    R.set_sky(sky)
    rgb = R.sky_lookup(dirs)
    pose = np.array([130, 103, 130, 0, 0, 0, 30], np.float32)
    ops, poses = [], [pose.copy()]
    for k in range(64):
        op = int(rng.randint(0, 11)); arg = float(rng.uniform(0.01, 0.5))
        pose = R.camera_op(pose, op, arg)
        ops.append((op, arg)); poses.append(pose.copy())
    np.savez_compressed(os.path.join(OUT, "sky_camera.npz"), dirs=dirs, rgb=rgb, ops=np.array(ops, np.float32),
                        poses=np.array(poses, np.float32))
    print("golden written:", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
