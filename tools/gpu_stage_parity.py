"""Stage-by-stage parity probe (development aid): philox, raygen, shade, sky vs the CPU oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpppathtracer_b200 import api, scenes
from oracle import ref_host

sky = scenes.synthetic_sky(512, 512)
orc = ref_host.best_oracle()
sc = scenes.demo_scene(640, 360)
orc.load_scene(sc); orc.set_sky(sky)
r = api.Renderer(sc, sky)
rng = np.random.RandomState(3)
n = 100000
# raygen
xy = np.stack([rng.randint(0, 640, n), rng.randint(0, 360, n)], 1)
u3 = rng.random_sample((n, 3)).astype(np.float32)
a = orc.raygen(xy, u3); b = r.raygen(sc.camera, xy, u3)
print("raygen bit-exact frac", (a == b).all(axis=1).mean(), "max abs", np.abs(a - b).max())
print("camera basis equal", np.array_equal(orc.set_camera(sc.camera), api.camera_basis(sc.camera)))
# shade
mats = sc.materials
mi = rng.randint(0, len(mats), n).astype(np.int32)
nrm = rng.normal(size=(n, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
ind = rng.normal(size=(n, 3)).astype(np.float32); ind /= np.linalg.norm(ind, axis=1, keepdims=True)
pos = rng.normal(size=(n, 3)).astype(np.float32)
ob, oa, orad, ond = orc.shade(mats, mi, pos, nrm, ind, u3)
gb, ga, grad, gnd = r.shade(mats, mi, pos, nrm, ind, u3)
for t in range(4):
    sel = mats["type"][mi] == t
    if not sel.any(): continue
    d = np.abs(ob[sel] - gb[sel]).max(axis=1)
    fin = np.isfinite(ob[sel]).all(axis=1)
    print(f"shade type {t}: n {sel.sum()} bounce bit-exact {(ob[sel]==gb[sel]).all(axis=1).mean():.5f} "
          f"max abs diff (finite) {np.nanmax(np.where(fin, d, 0)):.3e} frac>1e-5 {(d[fin]>1e-5).mean():.2e} "
          f"atten equal {(oa[sel]==ga[sel]).all(axis=1).mean():.5f} draws equal {(ond[sel]==gnd[sel]).mean():.3f} nan ref {(~fin).sum()}")
# sky
dirs = rng.normal(size=(n, 3)).astype(np.float32); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
so = orc.sky_lookup(dirs); sg = r.sky_lookup(dirs)
d = np.abs(so - sg).max(axis=1)
print("sky max abs diff", d.max(), "mean", d.mean(), "frac > 1e-4", (d > 1e-4).mean())
worst = np.argsort(-d)[:5]
for i in worst: print("   dir", dirs[i], "ref", so[i], "gpu", sg[i])
# philox
pss = np.stack([rng.randint(0, 2**31, 1000), rng.randint(0, 4096, 1000), rng.randint(0, 9, 1000)], 1).astype(np.uint32)
ug = r.philox(0x123456789abcdef, pss)
import ctypes as C
lib = C.CDLL(os.path.join(os.path.dirname(ref_host.__file__), "libpt_oracle.so"))
print("philox sample", ug[0], pss[0])
np.save("gpurun_out/philox_gpu.npy", ug); np.save("gpurun_out/philox_pss.npy", pss)
