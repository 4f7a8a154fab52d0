"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libpt_ref_host.so.

That library is the reference's OWN tracing code (cuSrc/{object,material,bvh,motional_camera}.cu
under /root/reference) compiled unmodified for the host by oracle/ref_build/Makefile, plus a
harness restating SamplePixel/Miss (cuSrc/path_tracer.cu:117-175).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpt_ref_host.so")      # the reference's own code, host build
PORT_LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")            # the restatement (oracle/pt_oracle.cpp)


class RefMaterial(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("have_tex", C.c_int32),
        ("kd", C.c_float * 3),
        ("refractive_index", C.c_float),
        ("emit_intensity", C.c_float),
        ("smoothness", C.c_float),
        ("reflectivity", C.c_float),
    ]


class RefStats(C.Structure):
    _fields_ = [("paths", C.c_ulonglong), ("segments", C.c_ulonglong), ("misses", C.c_ulonglong),
                ("depth_exhausted", C.c_ulonglong)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def port_available() -> bool:
    return os.path.exists(PORT_LIB_PATH)


class _Prefixed:
    """Resolves `lib.ref_xyz` to `<prefix>xyz` so one wrapper drives both oracle libraries."""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        if name.startswith("ref_"):
            name = self._prefix + name[4:]
        return getattr(self._lib, name)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefHost:
    """One process-wide scene (the reference keeps its BVH in file-static state, bvh.cu:16-20)."""

    def __init__(self, path: str = LIB_PATH, prefix: str = "ref_"):
        self.kind = "reference" if prefix == "ref_" else "port"
        self.lib = _Prefixed(C.CDLL(path), prefix)
        L = self.lib
        L.ref_add_object.restype = C.c_int
        L.ref_add_object.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float,
                                     C.POINTER(RefMaterial)]
        L.ref_build.restype = C.c_int
        L.ref_sizeof.restype = C.c_int
        L.ref_max_threads.restype = C.c_int
        self.n_prims = 0

    # -- scene ---------------------------------------------------------------------------------
    @staticmethod
    def _mat(m) -> RefMaterial:
        r = RefMaterial()
        r.type = int(m["type"]); r.have_tex = 0
        r.kd[0], r.kd[1], r.kd[2] = [float(v) for v in m["kd"]]
        r.refractive_index = float(m["refractive_index"]); r.emit_intensity = float(m["emit_intensity"])
        r.smoothness = float(m["smoothness"]); r.reflectivity = float(m["reflectivity"])
        return r

    def load_scene(self, scene) -> int:
        self.lib.ref_reset()
        mats = [self._mat(m) for m in scene.materials]
        for i in range(scene.n_prims):
            c = (C.c_float * 3)(*[float(v) for v in scene.center[i]])
            self.lib.ref_add_object(int(scene.prim_type[i]), c, float(scene.radius[i]), float(scene.y_pos[i]),
                                    float(scene.height[i]), C.byref(mats[int(scene.mat_idx[i])]))
        self.n_prims = scene.n_prims
        n_nodes = self.lib.ref_build()
        self.set_camera(scene.camera)
        return n_nodes

    def update_object(self, prim: int, center, radius: float, y_pos: float = 0.0, height: float = 0.0):
        """SceneBVH::UpdateObject on the reference's own tree (reference build only: refit, no re-sort)."""
        if self.kind != "reference":
            raise RuntimeError("update_object needs the reference build (oracle/_ref)")
        c = (C.c_float * 3)(*[float(v) for v in center])
        self.lib.ref_update_object.restype = C.c_int
        if self.lib.ref_update_object(C.c_int(int(prim)), c, C.c_float(radius), C.c_float(y_pos), C.c_float(height)) != 0:
            raise RuntimeError("ref_update_object failed")

    def leaf_order(self) -> np.ndarray:
        out = np.full(self.n_prims, -1, np.int32)
        self.lib.ref_leaf_order(_p(out, C.c_int32))
        return out

    def set_sky(self, rgba: np.ndarray, quarter_rows: bool = True):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        self.lib.ref_set_sky(_p(rgba, C.c_uint8), C.c_int(w), C.c_int(h), C.c_int(1 if quarter_rows else 0))

    def set_camera(self, cam) -> np.ndarray:
        out = np.zeros(19, np.float32)
        o = (C.c_float * 3)(*[float(v) for v in cam.origin])
        a = (C.c_float * 3)(*[float(v) for v in cam.look_at])
        self.lib.ref_camera_setup(C.c_int(cam.width), C.c_int(cam.height), o, a, C.c_float(cam.view_fov),
                                  C.c_float(cam.lens_radius), _p(out, C.c_float))
        return out

    def camera_op(self, pose7: np.ndarray, op: int, arg: float) -> np.ndarray:
        p = np.ascontiguousarray(pose7, np.float32).copy()
        self.lib.ref_camera_op(_p(p, C.c_float), C.c_int(op), C.c_float(arg))
        return p

    # -- level-1 / level-2 entry points ------------------------------------------------------------
    def trace(self, rays8: np.ndarray):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32)
        nrm = np.empty((n, 3), np.float32); pos = np.empty((n, 3), np.float32)
        self.lib.ref_trace(C.c_int(n), _p(rays8, C.c_float), _p(prim, C.c_int32), _p(t, C.c_float),
                           _p(nrm, C.c_float), _p(pos, C.c_float))
        return prim, t, nrm, pos

    def shade(self, materials: np.ndarray, mat_idx, pos, nrm, in_dir, u3):
        n = len(mat_idx)
        mats = (RefMaterial * len(materials))(*[self._mat(m) for m in materials])
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        pos, nrm, in_dir, u3 = [np.ascontiguousarray(a, np.float32) for a in (pos, nrm, in_dir, u3)]
        b = np.empty((n, 3), np.float32); a = np.empty((n, 3), np.float32); r = np.empty((n, 3), np.float32)
        nd = np.empty(n, np.int32)
        self.lib.ref_shade(C.c_int(n), mats, _p(mat_idx, C.c_int32), _p(pos, C.c_float), _p(nrm, C.c_float),
                           _p(in_dir, C.c_float), _p(u3, C.c_float), _p(b, C.c_float), _p(a, C.c_float),
                           _p(r, C.c_float), _p(nd, C.c_int32))
        return b, a, r, nd

    def raygen(self, xy, u3):
        xy = np.ascontiguousarray(xy, np.int32); u3 = np.ascontiguousarray(u3, np.float32)
        n = xy.shape[0]
        rays = np.empty((n, 8), np.float32)
        self.lib.ref_raygen(C.c_int(n), _p(xy, C.c_int32), _p(u3, C.c_float), _p(rays, C.c_float))
        return rays

    def sky_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.empty_like(dirs)
        self.lib.ref_sky_lookup(C.c_int(dirs.shape[0]), _p(dirs, C.c_float), _p(out, C.c_float))
        return out

    # -- full render -------------------------------------------------------------------------------
    def render(self, cam, spp: int, max_depth: int = 8, seed: int = 0, first_sample: int = 0, nthreads: int = 0,
               want_normals: bool = False):
        """Returns (sum_rgb [H,W,3] float32 -- per-pixel SUM over the samples, stats dict[, normals])."""
        self.set_camera(cam)
        out = np.zeros((cam.height, cam.width, 3), np.float32)
        nrm = np.zeros((cam.height, cam.width, 3), np.float32) if want_normals else None
        st = RefStats()
        self.lib.ref_render(C.c_int(spp), C.c_int(first_sample), C.c_int(max_depth), C.c_ulonglong(seed),
                            C.c_int(nthreads), _p(out, C.c_float),
                            _p(nrm, C.c_float) if want_normals else None, C.byref(st))
        stats = {"paths": st.paths, "segments": st.segments, "misses": st.misses,
                 "depth_exhausted": st.depth_exhausted}
        return (out, stats, nrm) if want_normals else (out, stats)

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())


class PortOracle(RefHost):
    """The CPU restatement oracle/pt_oracle.cpp (same entry points, prefix orc_)."""

    def __init__(self, path: str = PORT_LIB_PATH):
        super().__init__(path, "orc_")

    def denoise_mix(self, color: np.ndarray, normal: np.ndarray, mix: np.ndarray, bgra: np.ndarray, idx: int):
        """Denoising + Mix of one sample, in place on `mix` ([H,W,3] f32) and `bgra` ([H,W,4] u8)."""
        h, w = color.shape[:2]
        color = np.ascontiguousarray(color, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        self.lib.ref_denoise_mix(C.c_int(w), C.c_int(h), _p(color, C.c_float), _p(normal, C.c_float), _p(mix, C.c_float),
                                 _p(bgra, C.c_uint8), C.c_int(idx))

    def segment_rays(self, cam, seg: int, sample: int = 0, max_depth: int = 8, seed: int = 0):
        """Rays of path segment `seg` for every pixel: (rays8 [H*W,8], valid [H*W] bool)."""
        self.set_camera(cam)
        n = cam.width * cam.height
        rays = np.zeros((n, 8), np.float32); valid = np.zeros(n, np.int32)
        self.lib.ref_segment_rays(C.c_int(seg), C.c_int(sample), C.c_int(max_depth), C.c_ulonglong(seed),
                                  _p(rays, C.c_float), _p(valid, C.c_int32))
        return rays, valid.astype(bool)


def best_oracle():
    """The reference's own code when its host build is present, else the restatement."""
    if available():
        return RefHost()
    return PortOracle()
