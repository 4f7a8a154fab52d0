/*
 * pt_api.h -- C ABI of the B200-native render path (libpt_b200.so).
 *
 * Drop-in boundary for the render hot path of DearPoca/CppPathTracer: everything the reference's
 * `PathTracer` (include/path_tracer.h:15-26) does between "objects + camera in" and "pixels out"
 * is behind these entry points.  Plain pointers and sizes only; the caller owns every host buffer,
 * the library owns every device buffer.  All calls are synchronous unless stated, thread-compatible
 * (one caller per scene) and return PT_OK (0) or a negative pt_status; pt_last_error() gives the text.
 *
 * Reference paths below are relative to the reference repository root.
 */
#ifndef PT_API_H
#define PT_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pt_scene pt_scene;

typedef enum pt_status {
    PT_OK = 0,
    PT_ERR_INVALID = -1,   /* bad argument / bad handle / call order            */
    PT_ERR_CUDA = -2,      /* a CUDA runtime call failed (text in pt_last_error) */
    PT_ERR_NOT_BUILT = -3, /* render / trace before pt_build                     */
    PT_ERR_NO_DEVICE = -4  /* no CUDA device: this library has no CPU fallback   */
} pt_status;

/* PrimitiveType::Enum, include/object.h:7-15 */
enum { PT_PRIM_SPHERE = 0, PT_PRIM_PLATFORM = 1, PT_PRIM_CYLINDER = 2 };
/* MaterialType::Enum, include/material.h:5-15.  As in the reference's dispatcher
 * (cuSrc/material.cu:151-156) type METAL runs the glossy-lobe "Mirror" shader and type MIRROR runs
 * the reflectivity-mixing "Metal" shader; TEST and unknown values shade as DIFFUSE. */
enum { PT_MAT_DIFFUSE = 0, PT_MAT_METAL = 1, PT_MAT_MIRROR = 2, PT_MAT_GLASS = 3, PT_MAT_TEST = 4 };

/* class Material, include/material.h:17-35 (same fields, same meaning) */
typedef struct pt_material {
    int32_t type;
    int32_t have_tex;       /* 1: kd is read from texture `tex` at (0,0) (material.cu:11-18) */
    float kd[3];
    float refractive_index;
    float emit_intensity;
    float smoothness;
    float reflectivity;
    uint64_t tex;           /* handle from pt_texture_create_rgba8 */
} pt_material;

/* The caller-visible state of class MotionalCamera (include/motional_camera.h:8-19) that
 * MotionalCamera::GetCopy (cuSrc/motional_camera.cu:177-200) turns into a frame basis. */
typedef struct pt_camera {
    int32_t width, height;
    float origin[3];
    float look_at[3];
    float vup[3];
    float view_fov_deg;
    float lens_radius;
} pt_camera;

typedef struct pt_stats {
    uint64_t paths;           /* (pixel, sample) pairs traced by the last pt_render_frame     */
    uint64_t segments;        /* rays (path segments) traced                                  */
    uint64_t launches;        /* kernels launched by this scene since creation                */
    float    render_ms;       /* device time of the last pt_render_frame (CUDA events)        */
    float    traverse_ms;     /* ... of which the launches that trace a segment (only with PT_OPT_STAGE_TIMING) */
    float    shade_ms;        /* ... the final shade-only launch                              */
    float    build_ms;        /* device time of the last pt_build                             */
    uint32_t bvh_nodes;
    uint32_t wave_paths;      /* paths in flight per wave                                     */
    uint32_t traverse_launches; /* tracing launches of the last pt_render_frame (= waves x depth)    */
    uint32_t waves;           /* waves of the last pt_render_frame                            */
    uint64_t node_steps;      /* inner BVH nodes visited (only with PT_OPT_COUNT_STEPS)       */
    uint64_t prim_tests;      /* sphere / cylinder tests run (only with PT_OPT_COUNT_STEPS)   */
    uint64_t hit_records;     /* 48-byte hit records that went through the hit queues (segments that hit) */
    float    bvh_sah_lbvh;    /* SAH cost (sum of inner-node areas / root area) of the Karras tree ...        */
    float    bvh_sah;         /* ... and after the SAH refit passes (PT_OPT_SAH_PASSES)                        */
    uint32_t entry_levels;    /* finest level of the entry grid built by the last pt_build (0: none)          */
    uint32_t entry_cells;     /* blocks of the entry grid, all levels (16 bytes each)                          */
} pt_stats;

enum {
    PT_OPT_WAVE_PATHS = 1,    /* upper bound of paths in flight per wave (default 128 Mi = 64 spp at 1080p;
                                 halved at run time until the hit queues fit in free device memory) */
    PT_OPT_STAGE_TIMING = 2,  /* 1: record per-stage CUDA events (adds syncs at frame end only)    */
    PT_OPT_TIE_ORDER = 3,     /* 1 (default): exactly tied hits resolve in the reference's leaf
                                 visiting order (bvh.cu:31-90,167-205); 0: lowest primitive id wins */
    PT_OPT_SKY_QUARTER_ROWS = 4, /* 1 (default): reproduce the reference's upload that copies `width`
                                 BYTES per row (cuSrc/textures.cu:32); 0: upload whole rows         */
    PT_OPT_REF_LEAF_BOXES = 9,  /* 1: a primitive is tested whenever the box of its PARENT in the reference's median-split
                                   tree is hit, like SceneBVH::TraceRay does (bvh.cu:175-180); removes most of the
                                   long-range graze-class differences at the price of more primitive tests. default 0 */
    PT_OPT_COUNT_STEPS = 8,     /* 1: use the instrumented traversal kernel that counts node visits / primitive tests */
    PT_OPT_TRAV_INNER_RUN = 7,  /* tuning: inner-node steps a lane may take before the warp regroups for leaf tests (default 6) */
    PT_OPT_SAH_PASSES = 10,     /* SAH refit passes (tree rotations) after the LBVH build; default 4, 0 = plain LBVH */
    PT_OPT_ENTRY_GRID = 12,     /* finest level of the ENTRY GRID (default 0: off, every walk starts at the root; 12 = as fine as the
                                   scene allows).  Level l
                                   tiles the scene's box with overlapping 2x2x2 blocks of cubic cells of 2^-l of its size; a
                                   block stores the <= 4 tree links a ray segment inside it has to start from, so short rays
                                   skip the root-to-neighbourhood descent.  Results do not depend on it.  Measured (profiles/r02_c3_variants.md):
                                   node visits per segment 18.7 -> 7.7 on C3 and 9.7 -> 4.9 on C2, throughput +-1 % -- a warp walks until
                                   its LONGEST ray is done and long rays start at the root either way -- hence off by default. */
    PT_OPT_ENTRY_CELLS_PER_PRIM = 13, /* the finest level has at most this many blocks per primitive (default 2) */
    PT_OPT_SPLIT_WALK = 14,     /* libpt_b200_grid.so only, with the entry grid on: 1 = every bounce runs as two launches, shade + classify
                                   the rays by entry-grid level, then walk one level class per warp; images are identical */
    PT_OPT_SAH_TREELETS = 11    /* 1: the refit passes restructure 7-leaf treelets by dynamic programming instead of
                                   single rotations: lower SAH (C2 11.06 vs 11.24), but 1.5 ms instead of 0.06 ms per
                                   pass at 1 k primitives -- for scenes that are built once; default 0 */
};

/* ---- lifetime -------------------------------------------------------------------------------- */
/* replaces `new PathTracer` (cppSrc/video_renderer.cpp:33).  Uses the calling thread's current CUDA
 * device; fails with PT_ERR_NO_DEVICE when there is none. */
int pt_scene_create(pt_scene** out);
void pt_scene_destroy(pt_scene* s);
int pt_set_option(pt_scene* s, int option, int64_t value);
/* all work of this scene is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream) */
int pt_set_stream(pt_scene* s, void* cuda_stream);

/* ---- scene description: replaces PathTracer::AddObject / SceneBVH::AddObject
 *      (cuSrc/path_tracer.cu:29-34, cuSrc/bvh.cu:22-29).  Return the primitive id (insertion index). */
int pt_add_sphere(pt_scene* s, const float center[3], float radius);
int pt_add_platform(pt_scene* s, float y_pos);
int pt_add_cylinder(pt_scene* s, const float center[3], float radius, float height);
/* replaces the by-hand fill of Object::material_ (include/object.h:26) */
int pt_set_material(pt_scene* s, int prim, const pt_material* m);
/* replaces SceneBVH::UpdateObject (cuSrc/bvh.cu:144-157); takes effect at the next pt_build.  Like the reference, whose
 * UpdateObject refits the tree it has and never re-sorts, an update keeps the order in which exactly tied hits resolve
 * (PT_OPT_TIE_ORDER): that order is recomputed only by the first pt_build after a pt_add_* call (or after
 * pt_set_option(PT_OPT_TIE_ORDER, 1) is set again). */
int pt_update_prim(pt_scene* s, int prim, const float center[3], float radius, float y_pos, float height);
/* replaces PocaTextureUtils::AddTexByFile("textures\\sky.png") (cuSrc/path_tracer.cu:47,
 * cuSrc/textures.cu:14-62) minus the OpenCV decode: the caller passes decoded RGBA8 rows */
int pt_set_sky_rgba8(pt_scene* s, const uint8_t* rgba, int width, int height);
/* replaces PathTracer::InitPipeline -> SceneBVH::BuildBVH (cuSrc/path_tracer.cu:308-314,
 * cuSrc/bvh.cu:116-120): uploads the primitives and builds the LBVH on the device */
int pt_build(pt_scene* s);

/* ---- render: replaces the SamplePixel launches of PathTracer::PipelineLoop
 *      (cuSrc/path_tracer.cu:124-175,276-279).  Traces samples [first_sample, first_sample+spp) of
 *      every pixel and ADDS their radiance to the scene's accumulation buffer (fp32 sum, sample
 *      order).  Returns after the device finished. */
int pt_render_frame(pt_scene* s, const pt_camera* cam, int spp, int max_depth, uint64_t seed, uint32_t first_sample);
/* same, but only enqueues the work on the scene's stream */
int pt_render_frame_async(pt_scene* s, const pt_camera* cam, int spp, int max_depth, uint64_t seed, uint32_t first_sample);
int pt_synchronize(pt_scene* s);
/* make the accumulation buffer exist (zeroed) for a width x height frame without rendering into it: a rank of a
 * sample-sharded frame whose shard is empty still has to contribute a SUM buffer to the reduce.  A no-op when the
 * buffer already has that size (it is NOT cleared then: pt_reset_accum does that). */
int pt_ensure_frame(pt_scene* s, int width, int height);
/* zero the accumulation buffer (the reference's only reset is MotionalCamera::Refresh,
 * cuSrc/motional_camera.cu:43-45 + path_tracer.cu:250) */
int pt_reset_accum(pt_scene* s);
/* mean radiance per pixel: rgba[4*(y*width+x)+{0,1,2}] = sum/samples, [3] = samples accumulated.
 * nfloats must be 4*width*height. */
int pt_read_accum(pt_scene* s, float* rgba, size_t nfloats);
/* pt_read_accum in two halves, for callers that render frame after frame (the video path, config C5): _begin enqueues the
 * normalisation and the device-to-host copy (rgba should be page-locked for the copy to be asynchronous) and returns; _end waits
 * for the oldest read-back begun.  At most two may be in flight; with begin(i+1) before end(i) the copy of frame i runs under the
 * render of frame i+1.  rgba must stay valid until its _end. */
int pt_read_accum_begin(pt_scene* s, float* rgba, size_t nfloats);
int pt_read_accum_end(pt_scene* s);
/* the device-resident SUM buffer (float4 per pixel: r,g,b,samples) for the caller's own reduction
 * across GPUs (torch.distributed / NCCL); valid until the frame size changes or the scene dies */
int pt_accum_device_ptr(pt_scene* s, float** dptr, size_t* nfloats);
int pt_get_stats(pt_scene* s, pt_stats* out);

/* ---- display pipeline: replaces Denoising + Mix + the D2H copy of PipelineLoop
 *      (cuSrc/path_tracer.cu:177-254,285-304).  See pt_dispatch_ray in DESIGN.md. */
int pt_dispatch_ray(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx,
                    uint8_t* bgra_out, size_t nbytes);
/* Same dispatch, but the BGRA8 frame lands in a page-locked buffer OWNED BY THE SCENE (the reference's
 * PathTracer::output_buffer_, cuSrc/path_tracer.cu:303-304, which its callback receives and must not keep):
 * *bgra stays valid until the next dispatch on this scene or pt_scene_destroy.  One DMA, no staging copy. */
int pt_dispatch_ray_frame(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx,
                          const uint8_t** bgra, size_t* nbytes);
/* The same dispatch in two halves, for the reference's ASYNCHRONOUS use of DispatchRay (a task queue drained by a render
 * thread that calls back, cuSrc/path_tracer.cu:256-319): _begin enqueues sample + Denoising + Mix and the D2H copy of
 * the frame and returns; _end waits for the oldest frame begun and hands out its page-locked buffer (valid until the
 * second _begin after it).  At most two dispatches may be in flight; begin(i+1) before end(i) hides the copy and the
 * caller's work on frame i under the trace of sample i+1.  pt_get_stats is not updated by these calls. */
int pt_dispatch_ray_begin(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx);
int pt_dispatch_ray_end(pt_scene* s, const uint8_t** bgra, size_t* nbytes);

/* ---- textures: replaces PocaTextureUtils::{AddTexByFile,DestroyTexture} for material textures */
int pt_texture_create_rgba8(pt_scene* s, const uint8_t* rgba, int width, int height, uint64_t* handle);
int pt_texture_destroy(pt_scene* s, uint64_t handle);

/* ---- image files: replaces cv::imread + cvtColor(CV_BGR2RGBA) of AddTexByFile (cuSrc/textures.cu:15-17).
 *      Any PNG colour type / bit depth, sequential or Adam7 -> malloc'ed RGBA8 rows (16-bit samples keep their high
 *      byte, 1/2/4-bit grey is scaled to 0..255, like imread's 8-bit conversion); host only. */
int pt_image_load_png(const char* path, uint8_t** rgba_out, int* width, int* height);
void pt_image_free(uint8_t* rgba);

/* ---- parity entry points: the SAME device code as the render path, on caller-supplied batches.
 *      rays8: n x (ox oy oz dx dy dz tmin tmax).  Outputs may be NULL. */
/* closest hit (SceneBVH::TraceRay + Object::IntersectionTest, bvh.cu:167-205, object.cu:10-128) */
int pt_trace_rays(pt_scene* s, int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3);
/* diagnostics: BVH nodes + primitives visited by each ray of the last pt_trace_rays (needs PT_OPT_COUNT_STEPS) */
int pt_last_trace_steps(pt_scene* s, int n, uint32_t* steps);
/* Material::EvalAttenuationAndCreateRay with the uniforms given instead of drawn (material.cu:145-163) */
int pt_shade_batch(pt_scene* s, int n, const pt_material* mats, int n_mats, const int32_t* mat_idx,
                   const float* pos3, const float* normal3, const float* in_dir3, const float* u3,
                   float* bounce3, float* atten3, float* radiance3, int32_t* n_draws);
/* MotionalCamera::RayGen with the uniforms given (motional_camera.cu:202-213); xy: n x (x,y) */
int pt_raygen_batch(pt_scene* s, const pt_camera* cam, int n, const int32_t* xy, const float* u3, float* rays8);
/* Miss + sky texture (path_tracer.cu:117-122, textures.cu:68-71) */
int pt_sky_lookup(pt_scene* s, int n, const float* dir3, float* rgb3);
/* the uniforms the render path draws for (pixel, sample, slot): u4 = n x 4 */
int pt_philox_uniforms(pt_scene* s, int n, uint64_t seed, const uint32_t* pixel_sample_slot3, float* u4);
/* the LBVH build's own radix sort (pt_lbvh.cuh; round 1 used cub) on caller-supplied 32-bit codes: stable, ascending;
 * slots_sorted[i] = input position of codes_sorted[i].  The scene needs pt_build again afterwards. */
int pt_sort_code_pairs(pt_scene* s, int n, const uint32_t* codes, uint32_t* codes_sorted, uint32_t* slots_sorted);
/* MotionalCamera::GetCopy basis (host math): out19 = u v w top_left horizontal vertical dist_to_focus */
int pt_camera_basis(const pt_camera* cam, float out19[19]);

/* MotionalCamera::MoveEye{Left,Right,Forward,Backward,Up,Down} / RotateAround{Up,Down,Left,Right} / ScaleFov
 * (cuSrc/motional_camera.cu:76-168) applied to `cam` in place (host math, no device needed).
 * op: 0 left, 1 right, 2 forward, 3 backward, 4 up, 5 down (arg = coefficient, move_speed), 6 rotate up,
 * 7 rotate down, 8 rotate left, 9 rotate right (arg = delta), 10 scale fov (arg = d). */
int pt_camera_control(pt_camera* cam, int op, float arg, float move_speed);

const char* pt_last_error(void);
const char* pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PT_API_H */
