"""CPU tests of the oracles (test infrastructure): Philox KATs, the restatement (oracle/pt_oracle.cpp)
against the committed golden vectors generated from the reference's own host build, and -- when that
build is present -- bit-exactness of the restatement against it on fresh inputs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from cpppathtracer_b200 import scenes


def _philox_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("philox")
    src = d / "ph.c"
    src.write_text('#include "pt_philox.h"\n'
                   "void kat(const unsigned* c, const unsigned* k, unsigned* o){ pt_philox4x32_10(c,k,o);}\n"
                   "void uni(unsigned long long s, unsigned p, unsigned q, unsigned r, float* u){ pt_philox_uniform4(s,p,q,r,u);}\n")
    so = d / "ph.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", f"-I{ROOT}/oracle", "-o", str(so), str(src)])
    return C.CDLL(str(so))


def test_philox_known_answers(tmp_path_factory):
    """Random123 v1.14 kat_vectors, philox4x32 10 rounds."""
    L = _philox_lib(tmp_path_factory)
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kats:
        c = np.array(c, np.uint32); k = np.array(k, np.uint32); o = np.zeros(4, np.uint32)
        L.kat(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert o.tolist() == want


def test_philox_uniform_range(tmp_path_factory):
    """curand_uniform convention: (0, 1], never 0 (quirk 10 of SURVEY.md 8a)."""
    L = _philox_lib(tmp_path_factory)
    u = np.zeros(4, np.float32)
    lo, hi = 1.0, 0.0
    for i in range(2000):
        L.uni(C.c_ulonglong(i * 7919), C.c_uint(i), C.c_uint(i % 64), C.c_uint(i % 9), u.ctypes.data_as(C.c_void_p))
        lo, hi = min(lo, u.min()), max(hi, u.max())
    assert lo > 0.0 and hi <= 1.0
    assert np.float32(0) * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10) > 0


@pytest.mark.parametrize("name,maker", [("demo", lambda: scenes.demo_scene(64, 36)),
                                        ("spheres200", lambda: scenes.sphere_field(200, 64, 36))])
def test_port_matches_golden_trace_and_image(port_oracle, sky, name, maker):
    g = golden(f"trace_{name}.npz")
    sc = maker()
    port_oracle.load_scene(sc)
    port_oracle.set_sky(sky)
    assert np.array_equal(port_oracle.set_camera(sc.camera), g["camera_basis"])
    assert np.array_equal(port_oracle.leaf_order(), g["leaf_order"])
    assert np.array_equal(port_oracle.raygen(g["xy"], g["u3"]), g["rays"])
    for rk, pk, tk, nk, hk in (("rays", "prim", "t", "normal", "pos"), ("rays2", "prim2", "t2", "normal2", "pos2")):
        p, t, n, h = port_oracle.trace(g[rk])
        assert np.array_equal(p, g[pk]) and np.array_equal(t, g[tk])
        assert np.array_equal(n, g[nk]) and np.array_equal(h, g[hk])
    img, st = port_oracle.render(sc.camera, int(g["image_spp"]), 8, seed=int(g["image_seed"]))
    assert st["segments"] == int(g["segments"])
    assert np.array_equal(img, g["image_sum"])


def test_port_matches_golden_shade(port_oracle):
    g = golden("shade.npz")
    b, a, r, nd = port_oracle.shade(g["materials"], g["mat_idx"], g["pos"], g["normal"], g["in_dir"], g["u3"])
    assert np.array_equal(nd, g["draws"])
    assert np.array_equal(a, g["atten"]) and np.array_equal(r, g["radiance"])
    assert np.array_equal(b, g["bounce"], equal_nan=True)
    # draw budget per shader (material.cu:23,44,68,77,103,133): Diffuse 2, type-Metal 2, type-Mirror 3, Glass 3
    want = {scenes.DIFFUSE: 2, scenes.METAL: 2, scenes.MIRROR: 3, scenes.GLASS: 3, scenes.TEST: 2}
    types = g["materials"]["type"][g["mat_idx"]]
    for t, k in want.items():
        assert (nd[types == t] == k).all()


def test_port_matches_golden_sky_and_camera(port_oracle, sky):
    g = golden("sky_camera.npz")
    port_oracle.set_sky(sky)
    assert np.array_equal(port_oracle.sky_lookup(g["dirs"]), g["rgb"])
    pose = g["poses"][0].copy()
    for (op, arg), want in zip(g["ops"], g["poses"][1:]):
        pose = port_oracle.camera_op(pose, int(op), float(arg))
        assert np.array_equal(pose, want)


def test_port_bit_exact_with_reference_build(port_oracle, ref_oracle, sky):
    """Fresh inputs (not the golden ones): the restatement equals the reference's own code, bit for bit."""
    rng = np.random.RandomState(99)
    for sc in (scenes.demo_scene(96, 54, seed=7), scenes.stress_field(3000, 96, 54, seed=3)):
        for o in (port_oracle, ref_oracle):
            o.load_scene(sc); o.set_sky(sky)
        assert np.array_equal(port_oracle.leaf_order(), ref_oracle.leaf_order())
        a, sa = ref_oracle.render(sc.camera, 3, 8, seed=5, first_sample=2)
        b, sb = port_oracle.render(sc.camera, 3, 8, seed=5, first_sample=2)
        assert sa == sb and np.array_equal(a, b)
        n = 20000
        rays = np.concatenate([rng.uniform(-200, 200, (n, 3)), rng.normal(size=(n, 3)), np.zeros((n, 1)),
                               np.full((n, 1), 1e30)], 1).astype(np.float32)
        for x, y in zip(ref_oracle.trace(rays), port_oracle.trace(rays)):
            assert np.array_equal(x, y)


def test_oracle_quirks(port_oracle, sky):
    """Observable reference quirks the oracle must keep."""
    mats = np.zeros(2, scenes.MATERIAL_DTYPE)
    mats[0] = (scenes.METAL, 0, (1, 1, 1), 0, 0, 5.0, 0.0)    # type Metal runs the 2-draw glossy shader
    mats[1] = (scenes.MIRROR, 0, (1, 1, 1), 0, 0, 5.0, 1.0)   # type Mirror runs the 3-draw mixing shader
    n = np.array([[0, 1, 0]] * 2, np.float32); d = np.array([[0.6, -0.8, 0]] * 2, np.float32)
    b, a, r, nd = port_oracle.shade(mats, [0, 1], n * 0, n, d, np.full((2, 3), 0.5, np.float32))
    assert nd.tolist() == [2, 3]
    # smoothness 5 -> alpha 1e15: perfect mirror direction
    assert np.allclose(b[0], [0.6, 0.8, 0], atol=1e-6)
    # negative-radius sphere: inward normal on the near root (object.cu:22-23)
    sc = scenes.demo_scene(32, 18)
    sc.prim_type = np.array([scenes.SPHERE], np.int32); sc.center = np.zeros((1, 3), np.float32)
    sc.radius = np.array([-2.0], np.float32); sc.y_pos = sc.height = np.zeros(1, np.float32)
    sc.mat_idx = np.zeros(1, np.int32)
    port_oracle.load_scene(sc)
    p, t, nrm, pos = port_oracle.trace(np.array([[0, 0, -10, 0, 0, 1, 0, 1e30]], np.float32))
    assert p[0] == 0 and abs(t[0] - 8) < 1e-5 and np.allclose(nrm[0], [0, 0, 1])
    # sky: only texels x < w/4 survive the upload (textures.cu:32)
    full = np.full((8, 8, 4), 255, np.uint8)
    port_oracle.set_sky(full, quarter_rows=True)
    dirs = np.array([[1.0, 0.02, 0.0]], np.float32)   # u ~ 0.003 -> inside the valid quarter
    assert port_oracle.sky_lookup(dirs)[0, 0] > 0.99
    dirs = np.array([[0.02, 1.0, 0.0]], np.float32)   # u ~ 0.247 -> bilinear tap at the seam sees zeros
    assert port_oracle.sky_lookup(dirs)[0, 0] < 0.99
