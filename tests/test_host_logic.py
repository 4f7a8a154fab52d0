"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/pt_api.h declares
(no compute without a GPU), scene generators are deterministic, sample / frame sharding is exact."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from cpppathtracer_b200 import api, scenes, distributed as ptd


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pt_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 25
    for path in (api.LIB_PATH, api.LIB_PATH_FMA, api.LIB_PATH_GRID):
        lib = api.load_library(path)
        for n in names:
            assert hasattr(lib, n), f"{path} does not export {n}"
        assert b"sm_100a" in lib.pt_version()
    assert sorted(api.EXPORTS) == names
    assert b"fma" in api.load_library(api.LIB_PATH_FMA).pt_version() and b"fma" not in api.load_library().pt_version()
    assert b"entry grid" in api.load_library(api.LIB_PATH_GRID).pt_version() and b"entry grid" not in api.load_library().pt_version()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = api.load_library()
    h = C.c_void_p()
    assert lib.pt_scene_create(C.byref(h)) == -4          # PT_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.pt_last_error()
    with pytest.raises(api.PtError):
        api.Renderer()


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under cpppathtracer_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "cpppathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                for needle in ("libpt_oracle", "ref_host", "_ref/", "pt_oracle", "libpt_ref"):
                    assert needle not in src, (f, needle)


def test_struct_layouts_match_header():
    assert C.sizeof(api.PtMaterial) == 48 and C.sizeof(api.PtCamera) == 52
    assert C.sizeof(api.PtStats) == 96
    # and against the C compiler's view of include/pt_api.h
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "pt_api.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(pt_material), sizeof(pt_camera), sizeof(pt_stats));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(api.PtMaterial), C.sizeof(api.PtCamera), C.sizeof(api.PtStats)]


def test_scene_generators_are_deterministic():
    a, b = scenes.demo_scene(64, 36), scenes.demo_scene(64, 36)
    for k in ("prim_type", "center", "radius", "y_pos", "height", "mat_idx", "materials"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert a.prim_type[0] == scenes.PLATFORM and a.n_prims >= 75
    # quirk 7: hollow inner objects are spheres with radius 0.01 - r, even inside cylinders
    inner = np.where(a.radius < 0)[0]
    assert (a.prim_type[inner] == scenes.SPHERE).all()
    assert (a.materials["type"][a.mat_idx[inner]] == scenes.GLASS).all()
    c2 = scenes.sphere_field(1000)
    assert c2.n_prims == 1001 and (c2.prim_type[1:] == scenes.SPHERE).all()
    assert np.allclose(c2.center[1:, 1], c2.radius[1:])        # spheres rest on the floor
    c3 = scenes.stress_field(2000)
    glass = c3.materials["type"][c3.mat_idx] == scenes.GLASS
    assert glass.mean() > 0.45 and (c3.prim_type == scenes.CYLINDER).mean() > 0.3


def test_shard_samples_partitions_exactly():
    for total in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            ranges = [ptd.shard_samples(total, world, r) for r in range(world)]
            assert sum(c for _, c in ranges) == total
            pos = 0
            for first, count in ranges:
                assert first == pos
                pos += count
            counts = [c for _, c in ranges]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        ptd.shard_samples(8, 2, 2)


def test_frame_round_robin():
    for world in (1, 2, 4, 8):
        owned = [ptd.frames_of_rank(120, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(120))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        assert all(ptd.frame_owner(f, world) == f % world for f in range(16))


def test_mean_from_sum():
    s = np.zeros((2, 2, 4), np.float32); s[..., :3] = 8.0; s[..., 3] = 4.0
    assert np.array_equal(ptd.mean_from_sum(s), np.full((2, 2, 3), 2.0, np.float32))


def test_camera_control_matches_reference_poses():
    """pt_camera_control (host math in the library) vs the poses the reference's own MotionalCamera produced."""
    from conftest import golden
    g = golden("sky_camera.npz")
    cam = scenes.Camera(16, 16, tuple(g["poses"][0][:3]), tuple(g["poses"][0][3:6]), view_fov=float(g["poses"][0][6]))
    for (op, arg), want in zip(g["ops"], g["poses"][1:]):
        api.camera_control(cam, int(op), float(np.float32(arg)))
        got = np.array(list(cam.origin) + list(cam.look_at) + [cam.view_fov], np.float32)
        assert np.array_equal(got, want)


def test_video_camera_path_and_moving_objects():
    from cpppathtracer_b200 import video
    sc = scenes.sphere_field(100, 64, 36)
    cams = video.camera_path(sc.camera, 5)
    assert len(cams) == 5 and cams[0].origin == sc.camera.origin and cams[1].origin != cams[0].origin
    ids, c = video.moving_spheres(sc, 3)
    ids2, c2 = video.moving_spheres(sc, 4)
    assert np.array_equal(ids, ids2) and len(ids) == 10 and not np.array_equal(c, c2)
    assert (c[:, 1] >= sc.radius[ids] - 1e-6).all()


def test_png_loader_matches_pil(tmp_path):
    """pt_image_load_png (replaces cv::imread + BGR2RGBA of textures.cu:15-17) against PIL on every supported layout."""
    from PIL import Image
    rng = np.random.RandomState(5)
    h, w = 37, 53
    rgb = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    rgb[:, : w // 2] = (np.linspace(0, 255, h)[:, None, None] * np.ones((1, w // 2, 3))).astype(np.uint8)   # smooth half: exercises the filters
    rgba = np.dstack([rgb, rng.randint(0, 256, (h, w, 1)).astype(np.uint8)])
    images = {"rgb": Image.fromarray(rgb, "RGB"), "rgba": Image.fromarray(rgba, "RGBA"), "grey": Image.fromarray(rgb[..., 0], "L"),
              "la": Image.fromarray(rgba[..., [0, 3]], "LA"), "pal": Image.fromarray(rgb, "RGB").quantize(64)}
    for name, im in images.items():
        path = str(tmp_path / f"{name}.png")
        im.save(path, optimize=(name == "rgb"))
        got = api.load_png(path)
        want = np.array(Image.open(path).convert("RGBA"))
        assert got.shape == (h, w, 4) and np.array_equal(got, want), name
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all" * 4)
    with pytest.raises(api.PtError):
        api.load_png(str(bad))
    with pytest.raises(api.PtError):
        api.load_png(str(tmp_path / "missing.png"))


def _write_png(path, samples, ctype, depth, interlace, palette=None):
    """Minimal PNG writer for the loader test: samples [h, w, channels] (ints < 2**depth), filter type 0, optional Adam7."""
    import struct, zlib
    h, w, ch = samples.shape

    def pack_rows(img):
        rows = []
        for row in img:
            flat = row.reshape(-1).astype(np.uint32)
            if depth == 16:
                b = np.stack([flat >> 8, flat & 255], 1).astype(np.uint8).tobytes()
            elif depth == 8:
                b = flat.astype(np.uint8).tobytes()
            else:
                bits = np.zeros(((len(flat) * depth + 7) // 8) * 8, np.uint8)
                for k in range(depth):
                    bits[np.arange(len(flat)) * depth + k] = (flat >> (depth - 1 - k)) & 1
                b = np.packbits(bits).tobytes()
            rows.append(b"\x00" + b)
        return b"".join(rows)

    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    raw = b"".join(pack_rows(samples[ys::yst, xs::xst]) for xs, ys, xst, yst in passes if samples[ys::yst, xs::xst].size)

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    out += chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    open(path, "wb").write(out)


@pytest.mark.parametrize("interlace", [0, 1])
def test_png_loader_every_depth_and_adam7(tmp_path, interlace):
    """Bit depths 1/2/4/16 and Adam7 interlacing: 16-bit samples keep their high byte and low-depth grey is scaled to
    0..255 -- the 8-bit image cv::imread hands the reference (textures.cu:15); cross-checked with cv2 when it is there."""
    rng = np.random.RandomState(9)
    h, w = 19, 23                                                  # odd sizes: ragged Adam7 passes and ragged packed rows
    cases = []
    for depth in (1, 2, 4, 8, 16):
        cases.append(("grey", 0, depth, rng.randint(0, 2 ** depth, (h, w, 1))))
    for depth in (8, 16):
        cases += [("rgb", 2, depth, rng.randint(0, 2 ** depth, (h, w, 3))), ("rgba", 6, depth, rng.randint(0, 2 ** depth, (h, w, 4))),
                  ("la", 4, depth, rng.randint(0, 2 ** depth, (h, w, 2)))]
    pal = rng.randint(0, 256, (16, 3))
    for depth in (1, 2, 4, 8):
        cases.append(("pal", 3, depth, rng.randint(0, min(2 ** depth, 16), (h, w, 1))))
    for name, ctype, depth, smp in cases:
        path = str(tmp_path / f"{name}{depth}_{interlace}.png")
        _write_png(path, smp, ctype, depth, interlace, palette=pal if ctype == 3 else None)
        got = api.load_png(path)
        s8 = (smp >> 8) if depth == 16 else smp
        if ctype == 0:
            g = (s8[..., 0] * 255 // (2 ** min(depth, 8) - 1)) if depth < 8 else s8[..., 0]
            want = np.dstack([g, g, g, np.full_like(g, 255)])
        elif ctype == 2:
            want = np.dstack([s8, np.full((h, w, 1), 255)])
        elif ctype == 6:
            want = s8
        elif ctype == 4:
            want = np.dstack([s8[..., 0], s8[..., 0], s8[..., 0], s8[..., 1]])
        else:
            want = np.dstack([pal[smp[..., 0]], np.full((h, w, 1), 255)])
        assert got.shape == (h, w, 4) and np.array_equal(got, want.astype(np.uint8)), (name, depth, interlace)
        try:
            import cv2
        except ImportError:
            continue
        bgr = cv2.imread(path)                                      # what the reference's loader starts from
        assert bgr is not None and np.array_equal(bgr[..., ::-1], got[..., :3]), (name, depth, interlace, "cv2")
