"""Seeded synthetic scenes of the shapes named in BASELINE.json `configs`.

Pure host-side data (numpy); no device code here.  The same `Scene` feeds the CUDA product
(`cpppathtracer_b200.api`) and, in tests / bench only, the CPU oracles under `oracle/`.

The demo recipe restates the scene construction of the reference's
`VideoRenderer::Create` (cppSrc/video_renderer.cpp:39-118) with a fixed seed instead of
`srand(time(0))` (include/ray_tracing_math.hpp:30-37).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# PrimitiveType::Enum (include/object.h:7-15)
SPHERE, PLATFORM, CYLINDER = 0, 1, 2
# MaterialType::Enum (include/material.h:5-15)
DIFFUSE, METAL, MIRROR, GLASS, TEST = 0, 1, 2, 3, 4

MATERIAL_DTYPE = np.dtype(
    [
        ("type", np.int32),
        ("have_tex", np.int32),
        ("kd", np.float32, (3,)),
        ("refractive_index", np.float32),
        ("emit_intensity", np.float32),
        ("smoothness", np.float32),
        ("reflectivity", np.float32),
    ]
)


@dataclass
class Camera:
    """Snapshot of the caller-visible MotionalCamera fields (include/motional_camera.h:8-19)."""

    width: int = 1920
    height: int = 1080
    origin: tuple = (130.0, 103.0, 130.0)
    look_at: tuple = (0.0, 0.0, 0.0)
    vup: tuple = (0.0, 1.0, 0.0)
    view_fov: float = 30.0
    lens_radius: float = 0.0005


@dataclass
class Scene:
    prim_type: np.ndarray  # int32 [N]
    center: np.ndarray  # float32 [N,3]
    radius: np.ndarray  # float32 [N]
    y_pos: np.ndarray  # float32 [N]
    height: np.ndarray  # float32 [N]
    mat_idx: np.ndarray  # int32 [N] into materials
    materials: np.ndarray  # MATERIAL_DTYPE [M]
    camera: Camera = field(default_factory=Camera)
    name: str = "scene"

    @property
    def n_prims(self) -> int:
        return int(self.prim_type.shape[0])


class _Builder:
    def __init__(self):
        self.t, self.c, self.r, self.y, self.h, self.m = [], [], [], [], [], []

    def add(self, ptype, center, radius, y_pos, height, mat):
        self.t.append(ptype)
        self.c.append([np.float32(v) for v in center])
        self.r.append(np.float32(radius))
        self.y.append(np.float32(y_pos))
        self.h.append(np.float32(height))
        self.m.append(int(mat))
        return len(self.t) - 1

    def scene(self, materials, camera, name):
        return Scene(
            np.asarray(self.t, np.int32),
            np.asarray(self.c, np.float32).reshape(-1, 3),
            np.asarray(self.r, np.float32),
            np.asarray(self.y, np.float32),
            np.asarray(self.h, np.float32),
            np.asarray(self.m, np.int32),
            materials,
            camera,
            name,
        )


def material_library(rng: np.random.RandomState, n: int = 20) -> np.ndarray:
    """The 20-material library of video_renderer.cpp:41-71."""
    f32 = np.float32
    U = lambda: f32(rng.random_sample())  # noqa: E731
    mats = np.zeros(n, MATERIAL_DTYPE)
    mats[0]["type"] = DIFFUSE
    mats[0]["kd"] = (0.95, 0.95, 0.95)
    for i in range(1, n):
        mats[i]["kd"] = (U(), U(), U())
        rnd = int(U() * f32(2048)) % 5
        if rnd == 1:
            mats[i]["type"] = METAL
            mats[i]["smoothness"] = U() * f32(4) + f32(1)
            mats[i]["reflectivity"] = U() * f32(0.8)
        elif rnd == 2:
            mats[i]["type"] = MIRROR
            mats[i]["kd"] = tuple(f32(0.5) + f32(0.5) * U() for _ in range(3))
            mats[i]["smoothness"] = U() * f32(4) + f32(0.5)
        elif rnd == 3:
            mats[i]["type"] = GLASS
            mats[i]["smoothness"] = U() * f32(4) + f32(2)
            mats[i]["refractive_index"] = U() * f32(2) + f32(1.2)
            mats[i]["kd"] = (1.0, 1.0, 1.0)
        else:
            mats[i]["type"] = DIFFUSE
    return mats


def demo_scene(width: int = 640, height: int = 360, seed: int = 12345) -> Scene:
    """Config C1: floor platform + 74 random spheres / cylinders (video_renderer.cpp:73-117).

    Inner "hollow" objects are created without a type assignment in the reference
    (video_renderer.cpp:94-98,109-114), so they are spheres even inside cylinders; reproduced.
    """
    f32 = np.float32
    rng = np.random.RandomState(seed)
    U = lambda: f32(rng.random_sample())  # noqa: E731
    mats = material_library(rng)
    b = _Builder()
    b.add(PLATFORM, (0.0, -10000.0, 0.0), 10000.0, 0.0, 0.0, 0)
    for i in range(-550, 550, 15):
        rnd = int(U() * f32(2048)) % 2
        m = int(rng.randint(0, 20))
        if rnd == 0:
            r = U() * f32(15) + f32(1)
            c = (U() * f32(300) - f32(150), r, f32(i))
            b.add(SPHERE, c, r, 0.0, 0.0, m)
            if mats[m]["type"] == GLASS and U() > f32(0.5):
                b.add(SPHERE, c, f32(0.01) - r, 0.0, 0.0, m)
        else:
            r = U() * f32(15) + f32(1)
            h = r / f32(2) + U() * f32(20)
            c = (U() * f32(300) - f32(150), h / f32(2), f32(i))
            b.add(CYLINDER, c, r, 0.0, h, m)
            if mats[m]["type"] == GLASS and U() > f32(0.5):
                b.add(SPHERE, c, f32(0.01) - r, 0.0, f32(0.01) - h, m)
    cam = Camera(width, height, (130.0, 103.0, 130.0), (0.0, 0.0, 0.0))
    return b.scene(mats, cam, "demo")


def sphere_field(n: int = 1000, width: int = 1920, height: int = 1080, seed: int = 12345) -> Scene:
    """Config C2: floor + n spheres, r~U[0.5,2.5], on a square of side 6*sqrt(n), library materials."""
    f32 = np.float32
    rng = np.random.RandomState(seed)
    mats = material_library(rng)
    ext = f32(6.0 * np.sqrt(n))
    b = _Builder()
    b.add(PLATFORM, (0.0, -10000.0, 0.0), 10000.0, 0.0, 0.0, 0)
    r = (rng.random_sample(n) * 2.0 + 0.5).astype(np.float32)
    x = (rng.random_sample(n).astype(np.float32) * ext - ext / f32(2)).astype(np.float32)
    z = (rng.random_sample(n).astype(np.float32) * ext - ext / f32(2)).astype(np.float32)
    m = rng.randint(0, 20, n)
    for i in range(n):
        b.add(SPHERE, (x[i], r[i], z[i]), r[i], 0.0, 0.0, m[i])
    s = float(ext) / 190.0
    cam = Camera(width, height, (130.0 * s, 103.0 * s, 130.0 * s), (0.0, 0.0, 0.0))
    return b.scene(mats, cam, f"spheres{n}")


def stress_field(n: int = 100_000, width: int = 1920, height: int = 1080, seed: int = 12345) -> Scene:
    """Config C3: floor + n prims, 50/50 spheres / cylinders, >= 50 % glass, half of the glass hollow."""
    f32 = np.float32
    rng = np.random.RandomState(seed)
    lib = material_library(rng)
    n_glass = 32
    glass = np.zeros(n_glass, MATERIAL_DTYPE)
    glass["type"] = GLASS
    glass["kd"] = 1.0
    glass["refractive_index"] = (rng.random_sample(n_glass) * 2.0 + 1.2).astype(np.float32)
    glass["smoothness"] = (rng.random_sample(n_glass) * 4.0 + 2.0).astype(np.float32)
    mats = np.concatenate([lib, glass])
    ext = f32(6.0 * np.sqrt(n))
    b = _Builder()
    b.add(PLATFORM, (0.0, -10000.0, 0.0), 10000.0, 0.0, 0.0, 0)
    kind = rng.randint(0, 2, n)
    is_glass = rng.random_sample(n) < 0.5
    m_lib = rng.randint(0, 20, n)
    m_gl = rng.randint(0, n_glass, n) + 20
    r = (rng.random_sample(n) * 2.0 + 0.5).astype(np.float32)
    hh = (r / f32(2) + (rng.random_sample(n) * 4.0).astype(np.float32)).astype(np.float32)
    x = (rng.random_sample(n).astype(np.float32) * ext - ext / f32(2)).astype(np.float32)
    z = (rng.random_sample(n).astype(np.float32) * ext - ext / f32(2)).astype(np.float32)
    hollow = rng.random_sample(n) < 0.5
    for i in range(n):
        m = int(m_gl[i]) if is_glass[i] else int(m_lib[i])
        g = mats[m]["type"] == GLASS
        if kind[i] == 0:
            c = (x[i], r[i], z[i])
            b.add(SPHERE, c, r[i], 0.0, 0.0, m)
            if g and hollow[i]:
                b.add(SPHERE, c, f32(0.01) - r[i], 0.0, 0.0, m)
        else:
            c = (x[i], hh[i] / f32(2), z[i])
            b.add(CYLINDER, c, r[i], 0.0, hh[i], m)
            if g and hollow[i]:
                b.add(SPHERE, c, f32(0.01) - r[i], 0.0, f32(0.01) - hh[i], m)
    cam = Camera(width, height, (0.35 * float(ext), 0.28 * float(ext), 0.35 * float(ext)), (0.0, 0.0, 0.0))
    return b.scene(mats, cam, f"stress{n}")


def synthetic_sky(w: int = 1280, h: int = 1280) -> np.ndarray:
    """A smooth RGBA8 gradient sky of sky.png's shape, for runs where no image file is available
    (the GPU box has no /root/reference; tests/golden/sky_quarter.npy holds the real left quarter)."""
    y = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None]
    x = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :]
    r = 0.35 + 0.55 * y + 0.05 * np.sin(12.0 * x)
    g = 0.55 + 0.40 * y + 0.03 * np.cos(9.0 * x)
    bl = 0.95 - 0.15 * y + 0.0 * x
    img = np.stack([r, g, bl, np.ones_like(r)], axis=-1)
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
