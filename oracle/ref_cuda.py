"""TEST / BASELINE INFRASTRUCTURE -- ctypes binding of oracle/_ref/libpt_ref_cuda.so: the reference's
own CUDA kernels (SamplePixel & co, cuSrc/path_tracer.cu) rebuilt for sm_100 by oracle/ref_build/Makefile.
Used by bench.py as the reported "reference CUDA on sm_100" baseline and by -m gpu tests as the
device-side level-1 oracle.  Never on the product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ref_host import RefMaterial, RefHost

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpt_ref_cuda.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


class RefCuda:
    def __init__(self, path: str = LIB_PATH):
        self.lib = C.CDLL(path)

    def load_scene(self, scene) -> None:
        mats = (RefMaterial * len(scene.materials))(*[RefHost._mat(m) for m in scene.materials])
        t = np.ascontiguousarray(scene.prim_type, np.int32); c = np.ascontiguousarray(scene.center, np.float32)
        r = np.ascontiguousarray(scene.radius, np.float32); y = np.ascontiguousarray(scene.y_pos, np.float32)
        h = np.ascontiguousarray(scene.height, np.float32); m = np.ascontiguousarray(scene.mat_idx, np.int32)
        rc = self.lib.refcuda_load_scene(C.c_int(scene.n_prims), C.c_void_p(t.ctypes.data), C.c_void_p(c.ctypes.data),
                                         C.c_void_p(r.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(h.ctypes.data),
                                         C.c_void_p(m.ctypes.data), mats)
        if rc:
            raise RuntimeError(f"refcuda_load_scene failed: {rc}")

    def set_sky(self, rgba: np.ndarray, quarter_rows: bool = True) -> None:
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        rc = self.lib.refcuda_set_sky(C.c_void_p(rgba.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(int(quarter_rows)))
        if rc:
            raise RuntimeError(f"refcuda_set_sky failed: {rc}")

    def trace(self, rays8: np.ndarray):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32)
        nrm = np.empty((n, 3), np.float32); pos = np.empty((n, 3), np.float32)
        rc = self.lib.refcuda_trace(C.c_int(n), C.c_void_p(rays8.ctypes.data), C.c_void_p(prim.ctypes.data),
                                    C.c_void_p(t.ctypes.data), C.c_void_p(nrm.ctypes.data), C.c_void_p(pos.ctypes.data))
        if rc:
            raise RuntimeError(f"refcuda_trace failed: {rc}")
        return prim, t, nrm, pos

    def denoise_mix(self, color: np.ndarray, normal: np.ndarray, mix: np.ndarray, bgra: np.ndarray, idx: int):
        """The reference's own Denoising + Mix kernels, in place on `mix` ([H,W,3] f32) and `bgra` ([H,W,4] u8)."""
        h, w = color.shape[:2]
        color = np.ascontiguousarray(color, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        rc = self.lib.refcuda_denoise_mix(C.c_int(w), C.c_int(h), C.c_void_p(color.ctypes.data), C.c_void_p(normal.ctypes.data),
                                          C.c_void_p(mix.ctypes.data), C.c_void_p(bgra.ctypes.data), C.c_int(idx))
        if rc:
            raise RuntimeError(f"refcuda_denoise_mix failed: {rc}")

    def render(self, cam, spp: int, warmup: int = 1, seed: int = 1234, want_image: bool = False):
        """Returns dict(kernel_ms, service_ms, traced_pixels, mpaths_kernel, mpaths_service[, sum_rgb])."""
        o = (C.c_float * 3)(*[float(v) for v in cam.origin]); a = (C.c_float * 3)(*[float(v) for v in cam.look_at])
        k = C.c_float(); s = C.c_float(); px = C.c_longlong()
        img = np.zeros((cam.height, cam.width, 3), np.float32) if want_image else None
        rc = self.lib.refcuda_render(C.c_int(cam.width), C.c_int(cam.height), o, a, C.c_float(cam.view_fov),
                                     C.c_float(cam.lens_radius), C.c_int(spp), C.c_int(warmup), C.c_ulonglong(seed),
                                     C.c_void_p(img.ctypes.data) if want_image else None, C.byref(k), C.byref(s), C.byref(px))
        if rc:
            raise RuntimeError(f"refcuda_render failed: {rc}")
        paths = px.value * spp
        out = {"kernel_ms": k.value, "service_ms": s.value, "traced_pixels": px.value,
               "mpaths_kernel": paths / k.value / 1e3, "mpaths_service": paths / s.value / 1e3}
        if want_image:
            out["sum_rgb"] = img
        return out
