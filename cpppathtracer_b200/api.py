"""ctypes binding of libpt_b200.so (C ABI in include/pt_api.h) -- the CUDA render path.

There is no CPU path: importing works anywhere, but creating a `Renderer` needs the built
library and a CUDA device, and raises otherwise.  The class mirrors what the reference's caller does
with `PathTracer` (include/path_tracer.h:15-26): add objects, set materials, build, dispatch samples,
read pixels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpt_b200.so")          # default: uncontracted arithmetic == the CPU oracle's
LIB_PATH_FMA = os.path.join(_HERE, "libpt_b200_fma.so")  # nvcc-contracted arithmetic == the reference's CUDA build
LIB_PATH_GRID = os.path.join(_HERE, "libpt_b200_grid.so")  # default arithmetic + the entry-grid code (PT_OPT_ENTRY_GRID)

PT_OPT_WAVE_PATHS, PT_OPT_STAGE_TIMING, PT_OPT_TIE_ORDER, PT_OPT_SKY_QUARTER_ROWS = 1, 2, 3, 4
PT_OPT_TRAV_INNER_RUN, PT_OPT_COUNT_STEPS, PT_OPT_REF_LEAF_BOXES, PT_OPT_SAH_PASSES, PT_OPT_SAH_TREELETS = 7, 8, 9, 10, 11
PT_OPT_ENTRY_GRID, PT_OPT_ENTRY_CELLS_PER_PRIM, PT_OPT_SPLIT_WALK = 12, 13, 14


class PtMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("have_tex", C.c_int32), ("kd", C.c_float * 3),
                ("refractive_index", C.c_float), ("emit_intensity", C.c_float), ("smoothness", C.c_float),
                ("reflectivity", C.c_float), ("tex", C.c_uint64)]


class PtCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("origin", C.c_float * 3), ("look_at", C.c_float * 3),
                ("vup", C.c_float * 3), ("view_fov_deg", C.c_float), ("lens_radius", C.c_float)]


class PtStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("launches", C.c_uint64), ("render_ms", C.c_float),
                ("traverse_ms", C.c_float), ("shade_ms", C.c_float), ("build_ms", C.c_float),
                ("bvh_nodes", C.c_uint32), ("wave_paths", C.c_uint32), ("traverse_launches", C.c_uint32),
                ("waves", C.c_uint32), ("node_steps", C.c_uint64), ("prim_tests", C.c_uint64), ("hit_records", C.c_uint64),
                ("bvh_sah_lbvh", C.c_float), ("bvh_sah", C.c_float),
                ("entry_levels", C.c_uint32), ("entry_cells", C.c_uint32)]


EXPORTS = [
    "pt_scene_create", "pt_scene_destroy", "pt_set_option", "pt_set_stream", "pt_add_sphere", "pt_add_platform",
    "pt_add_cylinder", "pt_set_material", "pt_update_prim", "pt_set_sky_rgba8", "pt_build", "pt_render_frame",
    "pt_render_frame_async", "pt_synchronize", "pt_reset_accum", "pt_ensure_frame", "pt_read_accum", "pt_read_accum_begin", "pt_read_accum_end", "pt_accum_device_ptr",
    "pt_get_stats", "pt_dispatch_ray", "pt_dispatch_ray_frame", "pt_dispatch_ray_begin", "pt_dispatch_ray_end", "pt_texture_create_rgba8", "pt_texture_destroy", "pt_trace_rays", "pt_last_trace_steps",
    "pt_shade_batch", "pt_raygen_batch", "pt_sky_lookup", "pt_philox_uniforms", "pt_sort_code_pairs", "pt_camera_basis", "pt_camera_control",
    "pt_last_error",
    "pt_version", "pt_image_load_png", "pt_image_free",
]


class PtError(RuntimeError):
    pass


_libs = {}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Load libpt_b200.so (or the variant at `path`); raises (never falls back) when it has not been built."""
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise PtError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the render path)")
    lib = C.CDLL(path)
    lib.pt_last_error.restype = C.c_char_p
    lib.pt_version.restype = C.c_char_p
    lib.pt_scene_destroy.restype = None
    lib.pt_add_sphere.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float]
    lib.pt_add_platform.argtypes = [C.c_void_p, C.c_float]
    lib.pt_add_cylinder.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_float]
    lib.pt_set_material.argtypes = [C.c_void_p, C.c_int, C.POINTER(PtMaterial)]
    lib.pt_update_prim.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float]
    lib.pt_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.pt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pt_render_frame.argtypes = [C.c_void_p, C.POINTER(PtCamera), C.c_int, C.c_int, C.c_uint64, C.c_uint32]
    lib.pt_render_frame_async.argtypes = lib.pt_render_frame.argtypes
    lib.pt_read_accum.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.pt_ensure_frame.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.pt_read_accum_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.pt_read_accum_end.argtypes = [C.c_void_p]
    lib.pt_set_sky_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.pt_dispatch_ray.argtypes = [C.c_void_p, C.POINTER(PtCamera), C.c_int, C.c_uint64, C.c_uint32, C.c_void_p,
                                    C.c_size_t]
    lib.pt_dispatch_ray_frame.argtypes = [C.c_void_p, C.POINTER(PtCamera), C.c_int, C.c_uint64, C.c_uint32,
                                          C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    lib.pt_dispatch_ray_begin.argtypes = [C.c_void_p, C.POINTER(PtCamera), C.c_int, C.c_uint64, C.c_uint32]
    lib.pt_dispatch_ray_end.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    lib.pt_trace_rays.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.pt_shade_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 9
    lib.pt_raygen_batch.argtypes = [C.c_void_p, C.POINTER(PtCamera), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pt_sky_lookup.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.pt_philox_uniforms.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.pt_sort_code_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pt_camera_basis.argtypes = [C.POINTER(PtCamera), C.c_void_p]
    lib.pt_camera_control.argtypes = [C.POINTER(PtCamera), C.c_int, C.c_float, C.c_float]
    lib.pt_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.pt_get_stats.argtypes = [C.c_void_p, C.POINTER(PtStats)]
    lib.pt_texture_create_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.pt_texture_destroy.argtypes = [C.c_void_p, C.c_uint64]
    for n in ("pt_build", "pt_synchronize", "pt_reset_accum", "pt_scene_destroy"):
        getattr(lib, n).argtypes = [C.c_void_p]
    _libs[path] = lib
    return lib


def to_pt_camera(cam) -> PtCamera:
    c = PtCamera()
    c.width, c.height = int(cam.width), int(cam.height)
    for k in range(3):
        c.origin[k] = float(cam.origin[k]); c.look_at[k] = float(cam.look_at[k]); c.vup[k] = float(cam.vup[k])
    c.view_fov_deg = float(cam.view_fov); c.lens_radius = float(cam.lens_radius)
    return c


def to_pt_material(m, tex: int = 0) -> PtMaterial:
    r = PtMaterial()
    r.type = int(m["type"]); r.have_tex = 1 if tex else 0
    r.kd[0], r.kd[1], r.kd[2] = [float(v) for v in m["kd"]]
    r.refractive_index = float(m["refractive_index"]); r.emit_intensity = float(m["emit_intensity"])
    r.smoothness = float(m["smoothness"]); r.reflectivity = float(m["reflectivity"]); r.tex = int(tex)
    return r


def load_png(path: str) -> np.ndarray:
    """RGBA8 [H,W,4] from an 8-bit PNG file (host only; the reference used cv::imread, textures.cu:15-17)."""
    lib = load_library()
    p = C.POINTER(C.c_uint8)(); w = C.c_int(); h = C.c_int()
    lib.pt_image_load_png.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pt_image_free.argtypes = [C.POINTER(C.c_uint8)]
    lib.pt_image_free.restype = None
    if lib.pt_image_load_png(path.encode(), C.byref(p), C.byref(w), C.byref(h)):
        raise PtError(lib.pt_last_error().decode())
    out = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    lib.pt_image_free(p)
    return out


def camera_basis(cam) -> np.ndarray:
    out = np.zeros(19, np.float32)
    c = to_pt_camera(cam)
    rc = load_library().pt_camera_basis(C.byref(c), out.ctypes.data)
    if rc:
        raise PtError(load_library().pt_last_error().decode())
    return out


CAM_LEFT, CAM_RIGHT, CAM_FORWARD, CAM_BACKWARD, CAM_UP, CAM_DOWN = 0, 1, 2, 3, 4, 5
CAM_ROT_UP, CAM_ROT_DOWN, CAM_ROT_LEFT, CAM_ROT_RIGHT, CAM_SCALE_FOV = 6, 7, 8, 9, 10


def camera_control(cam, op: int, arg: float, move_speed: float = 50.0):
    """MotionalCamera::MoveEye* / RotateAround* / ScaleFov (motional_camera.cu:76-168) on a scenes.Camera, in place."""
    c = to_pt_camera(cam)
    if load_library().pt_camera_control(C.byref(c), int(op), float(arg), float(move_speed)):
        raise PtError(load_library().pt_last_error().decode())
    cam.origin = tuple(np.float32(v) for v in c.origin)
    cam.look_at = tuple(np.float32(v) for v in c.look_at)
    cam.view_fov = float(np.float32(c.view_fov_deg))
    return cam


class Renderer:
    """One scene on the current CUDA device (the reference's `PathTracer` + `SceneBVH`)."""

    def __init__(self, scene=None, sky: np.ndarray | None = None, lib_path: str = LIB_PATH):
        self.lib = load_library(lib_path)
        h = C.c_void_p()
        self._check(self.lib.pt_scene_create(C.byref(h)))
        self.h = h
        self.stream_ptr = 0
        self.n_prims = 0
        if scene is not None:
            self.load_scene(scene)
        if sky is not None:
            self.set_sky(sky)
        if scene is not None:
            self.build()

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise PtError(f"pt error {rc}: {self.lib.pt_last_error().decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.pt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene --------------------------------------------------------------------------------------
    def add_sphere(self, center, radius) -> int:
        c = (C.c_float * 3)(*[float(v) for v in center])
        return self._check(self.lib.pt_add_sphere(self.h, c, float(radius)))

    def add_platform(self, y_pos) -> int:
        return self._check(self.lib.pt_add_platform(self.h, float(y_pos)))

    def add_cylinder(self, center, radius, height) -> int:
        c = (C.c_float * 3)(*[float(v) for v in center])
        return self._check(self.lib.pt_add_cylinder(self.h, c, float(radius), float(height)))

    def set_material(self, prim: int, mat: PtMaterial):
        self._check(self.lib.pt_set_material(self.h, int(prim), C.byref(mat)))

    def update_prim(self, prim, center, radius, y_pos=0.0, height=0.0):
        c = (C.c_float * 3)(*[float(v) for v in center])
        self._check(self.lib.pt_update_prim(self.h, int(prim), c, float(radius), float(y_pos), float(height)))

    def load_scene(self, scene):
        from . import scenes as S
        mats = [to_pt_material(m) for m in scene.materials]
        for i in range(scene.n_prims):
            t = int(scene.prim_type[i])
            if t == S.SPHERE:
                p = self.add_sphere(scene.center[i], scene.radius[i])
            elif t == S.PLATFORM:
                p = self.add_platform(scene.y_pos[i])
            else:
                p = self.add_cylinder(scene.center[i], scene.radius[i], scene.height[i])
            self.set_material(p, mats[int(scene.mat_idx[i])])
        self.n_prims = scene.n_prims

    def set_option(self, opt: int, value: int):
        self._check(self.lib.pt_set_option(self.h, int(opt), int(value)))

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.pt_set_stream(self.h, C.c_void_p(int(cuda_stream))))
        self.stream_ptr = int(cuda_stream)

    def set_sky(self, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        self._check(self.lib.pt_set_sky_rgba8(self.h, rgba.ctypes.data, w, h))

    def create_texture(self, rgba: np.ndarray) -> int:
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        out = C.c_uint64()
        self._check(self.lib.pt_texture_create_rgba8(self.h, rgba.ctypes.data, w, h, C.byref(out)))
        return int(out.value)

    def build(self):
        self._check(self.lib.pt_build(self.h))

    # -- render ---------------------------------------------------------------------------------------
    def render(self, cam, spp: int, max_depth: int = 8, seed: int = 0, first_sample: int = 0, sync: bool = True):
        c = to_pt_camera(cam)
        f = self.lib.pt_render_frame if sync else self.lib.pt_render_frame_async
        self._check(f(self.h, C.byref(c), int(spp), int(max_depth), int(seed), int(first_sample)))
        self._wh = (cam.width, cam.height)

    def synchronize(self):
        self._check(self.lib.pt_synchronize(self.h))

    def ensure_frame(self, width: int, height: int):
        """Allocate (zeroed) the SUM buffer for a width x height frame without rendering into it."""
        self._check(self.lib.pt_ensure_frame(self.h, int(width), int(height)))
        self._wh = (int(width), int(height))

    def reset_accum(self):
        self._check(self.lib.pt_reset_accum(self.h))

    def read_accum(self, out: np.ndarray | None = None) -> np.ndarray:
        w, h = self._wh
        if out is None:
            out = np.empty((h, w, 4), np.float32)
        self._check(self.lib.pt_read_accum(self.h, out.ctypes.data, out.size))
        return out

    def read_accum_begin(self, out: np.ndarray):
        """First half of a pipelined read-back into `out` (page-locked for a truly asynchronous copy); pair with read_accum_end."""
        self._check(self.lib.pt_read_accum_begin(self.h, out.ctypes.data, out.size))

    def read_accum_end(self):
        self._check(self.lib.pt_read_accum_end(self.h))

    def accum_device_ptr(self):
        p = C.c_void_p(); n = C.c_size_t()
        self._check(self.lib.pt_accum_device_ptr(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def stats(self) -> dict:
        st = PtStats()
        self._check(self.lib.pt_get_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in PtStats._fields_}

    def dispatch_ray(self, cam, sample_idx: int, max_depth: int = 8, seed: int = 0) -> np.ndarray:
        c = to_pt_camera(cam)
        out = np.empty((cam.height, cam.width, 4), np.uint8)
        self._check(self.lib.pt_dispatch_ray(self.h, C.byref(c), int(max_depth), int(seed), int(sample_idx),
                                             out.ctypes.data, out.size))
        return out

    def dispatch_ray_frame(self, cam, sample_idx: int, max_depth: int = 8, seed: int = 0) -> np.ndarray:
        """Like dispatch_ray, but returns a VIEW of the scene's page-locked frame (valid until the next dispatch)."""
        c = to_pt_camera(cam)
        ptr, nbytes = C.POINTER(C.c_uint8)(), C.c_size_t(0)
        self._check(self.lib.pt_dispatch_ray_frame(self.h, C.byref(c), int(max_depth), int(seed), int(sample_idx),
                                                   C.byref(ptr), C.byref(nbytes)))
        return np.ctypeslib.as_array(ptr, shape=(cam.height, cam.width, 4))

    def dispatch_ray_begin(self, cam, sample_idx: int, max_depth: int = 8, seed: int = 0):
        """First half of a pipelined dispatch (at most two in flight); pair with dispatch_ray_end."""
        c = to_pt_camera(cam)
        self._check(self.lib.pt_dispatch_ray_begin(self.h, C.byref(c), int(max_depth), int(seed), int(sample_idx)))
        self._pipe_shapes = getattr(self, "_pipe_shapes", []) + [(cam.height, cam.width, 4)]

    def dispatch_ray_end(self) -> np.ndarray:
        """Second half: a VIEW of the oldest frame begun (valid until the second begin after it)."""
        ptr, nbytes = C.POINTER(C.c_uint8)(), C.c_size_t(0)
        self._check(self.lib.pt_dispatch_ray_end(self.h, C.byref(ptr), C.byref(nbytes)))
        return np.ctypeslib.as_array(ptr, shape=self._pipe_shapes.pop(0))

    # -- parity entry points ----------------------------------------------------------------------------
    def trace(self, rays8: np.ndarray):
        rays8 = np.ascontiguousarray(rays8, np.float32)
        n = rays8.shape[0]
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32)
        nrm = np.empty((n, 3), np.float32); pos = np.empty((n, 3), np.float32)
        self._check(self.lib.pt_trace_rays(self.h, n, rays8.ctypes.data, prim.ctypes.data, t.ctypes.data,
                                           nrm.ctypes.data, pos.ctypes.data))
        return prim, t, nrm, pos

    def last_trace_steps(self, n: int) -> np.ndarray:
        out = np.empty(n, np.uint32)
        self._check(self.lib.pt_last_trace_steps(self.h, C.c_int(n), C.c_void_p(out.ctypes.data)))
        return out

    def shade(self, materials: np.ndarray, mat_idx, pos, nrm, in_dir, u3):
        n = len(mat_idx)
        mats = (PtMaterial * len(materials))(*[to_pt_material(m) for m in materials])
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        pos, nrm, in_dir, u3 = [np.ascontiguousarray(a, np.float32) for a in (pos, nrm, in_dir, u3)]
        b = np.empty((n, 3), np.float32); a = np.empty((n, 3), np.float32); r = np.empty((n, 3), np.float32)
        nd = np.empty(n, np.int32)
        self._check(self.lib.pt_shade_batch(self.h, n, C.addressof(mats), len(materials), mat_idx.ctypes.data,
                                            pos.ctypes.data, nrm.ctypes.data, in_dir.ctypes.data, u3.ctypes.data,
                                            b.ctypes.data, a.ctypes.data, r.ctypes.data, nd.ctypes.data))
        return b, a, r, nd

    def raygen(self, cam, xy, u3):
        xy = np.ascontiguousarray(xy, np.int32); u3 = np.ascontiguousarray(u3, np.float32)
        n = xy.shape[0]
        rays = np.empty((n, 8), np.float32)
        c = to_pt_camera(cam)
        self._check(self.lib.pt_raygen_batch(self.h, C.byref(c), n, xy.ctypes.data, u3.ctypes.data, rays.ctypes.data))
        return rays

    def sky_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.empty_like(dirs)
        self._check(self.lib.pt_sky_lookup(self.h, dirs.shape[0], dirs.ctypes.data, out.ctypes.data))
        return out

    def sort_code_pairs(self, codes: np.ndarray):
        """The LBVH build's radix sort on caller-supplied codes: (codes_sorted, input positions), stable."""
        codes = np.ascontiguousarray(codes, np.uint32)
        ks = np.empty_like(codes); vs = np.empty_like(codes)
        self._check(self.lib.pt_sort_code_pairs(self.h, codes.shape[0], codes.ctypes.data, ks.ctypes.data, vs.ctypes.data))
        return ks, vs

    def philox(self, seed: int, pixel_sample_slot: np.ndarray) -> np.ndarray:
        pss = np.ascontiguousarray(pixel_sample_slot, np.uint32)
        out = np.empty((pss.shape[0], 4), np.float32)
        self._check(self.lib.pt_philox_uniforms(self.h, pss.shape[0], int(seed), pss.ctypes.data, out.ctypes.data))
        return out
