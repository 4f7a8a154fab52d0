/*
 * ref_host_harness.cpp -- TEST INFRASTRUCTURE.  C-ABI around the reference's own tracing
 * code compiled for the host (see Makefile in this directory).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
 * resulting oracle/_ref/libpt_ref_host.so.
 *
 * Linked against, UNMODIFIED, from /root/reference:
 *     cuSrc/object.cu  cuSrc/material.cu  cuSrc/bvh.cu  cuSrc/motional_camera.cu
 * Restated here (they use threadIdx / <<<>>> and cannot be compiled for the host):
 *     SamplePixel  cuSrc/path_tracer.cu:124-175
 *     Miss         cuSrc/path_tracer.cu:117-122
 *     tex2D        CUDA texture unit as configured at cuSrc/textures.cu:46-50
 *                  (mirror addressing, linear filter, normalized coords, normalized-float read)
 * RNG: the reference's clock-seeded XORWOW is replaced by the Philox stream of ../pt_philox.h.
 */
#include <cstdio>
#include <vector>
#include <mutex>
#include <omp.h>

#include "path_tracer.h"   /* reference header: Object, Material, SceneBVH, MotionalCamera, RayPayload */
#include "textures.h"
#include "logger.hpp"
#include "../pt_philox.h"

Logger logger("/dev/null");

/* ---- CUDA runtime stand-ins: BuildBVHInGpu (bvh.cu:97-114) "uploads" into host memory ---- */
extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host-stub"; }
}

/* ---- texture unit stand-in ---- */
struct HostTexture { int w = 0, h = 0; std::vector<uint8_t> rgba; };
static HostTexture g_sky;

static inline int mirror_index(int i, int n) {
    if (i < 0) i = -1 - i;
    if (i >= n) i = 2 * n - 1 - i;
    if (i < 0) i = 0;
    if (i >= n) i = n - 1;
    return i;
}
static inline float mirror_coord(float x) {
    /* cudaAddressModeMirror on a normalized coordinate */
    float f = floorf(x);
    float fr = x - f;
    long long k = (long long)f;
    return (k & 1LL) ? 1.0f - fr : fr;
}
template <> float4 tex2D<float4>(cudaTextureObject_t, float u, float v) {
    const HostTexture& t = g_sky;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.w == 0) return out;
    float x = mirror_coord(u) * (float)t.w - 0.5f;
    float y = mirror_coord(v) * (float)t.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    /* the texture unit keeps 8 fractional bits of the filter weights */
    float a = floorf((x - fx) * 256.0f + 0.5f) * (1.0f / 256.0f);
    float b = floorf((y - fy) * 256.0f + 0.5f) * (1.0f / 256.0f);
    int i0 = mirror_index((int)fx, t.w), i1 = mirror_index((int)fx + 1, t.w);
    int j0 = mirror_index((int)fy, t.h), j1 = mirror_index((int)fy + 1, t.h);
    float w00 = (1.f - a) * (1.f - b), w10 = a * (1.f - b), w01 = (1.f - a) * b, w11 = a * b;
    float c[4];
    for (int k = 0; k < 4; ++k) {
        float t00 = t.rgba[4 * ((size_t)j0 * t.w + i0) + k] * (1.0f / 255.0f);
        float t10 = t.rgba[4 * ((size_t)j0 * t.w + i1) + k] * (1.0f / 255.0f);
        float t01 = t.rgba[4 * ((size_t)j1 * t.w + i0) + k] * (1.0f / 255.0f);
        float t11 = t.rgba[4 * ((size_t)j1 * t.w + i1) + k] * (1.0f / 255.0f);
        c[k] = w00 * t00 + w10 * t10 + w01 * t01 + w11 * t11;
    }
    return make_float4(c[0], c[1], c[2], c[3]);
}
float4 PocaTextureUtils::GetTexture2D(cudaTextureObject_t tex_obj, float u, float v) {
    return tex2D<float4>(tex_obj, u, v);
}

/* ---- scene state ---- */
static std::vector<Object*> g_objs;           /* never freed: the reference keeps raw pointers (bvh.cu:27) */
static SceneBVHGPUHandle g_root = nullptr;
static MotionalCamera g_cam;
static bool g_built = false;

static inline float int_as_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline int32_t float_as_int(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
/* insertion index is stashed in a field the primitive's code never reads
 * (object.cu:10-112,134-170: Sphere/Cylinder never touch y_pos_, Platform never touches height_) */
static inline int32_t prim_id_of(const Object& o) {
    return o.type_ == PrimitiveType::Platform ? float_as_int(o.height_) : float_as_int(o.y_pos_);
}

struct ref_material {
    int32_t type; int32_t have_tex; float kd[3];
    float refractive_index, emit_intensity, smoothness, reflectivity;
};
struct ref_stats { unsigned long long paths, segments, misses, depth_exhausted; };

static void fill_material(Material& m, const ref_material* r) {
    memset(&m, 0, sizeof(Material));
    m.type_ = (MaterialType::Enum)r->type;
    m.have_tex_ = false;
    m.kd_ = make_float3(r->kd[0], r->kd[1], r->kd[2]);
    m.refractive_index_ = r->refractive_index;
    m.emit_intensity_ = r->emit_intensity;
    m.smoothness_ = r->smoothness;
    m.reflectivity_ = r->reflectivity;
}

/* make_float3(curand_uniform(), curand_uniform(), curand_uniform()) (ray_tracing_math.hpp:98-100)
 * leaves the order of the three draws to the compiler (g++ evaluates right to left, nvcc left to
 * right).  The stream is DEFINED as rd.x <- U[0], rd.y <- U[1], rd.z <- U[2]; the probe below finds
 * which draw feeds which component in this build so the harness can present U accordingly. */
static int g_comp_of_draw[3] = {-1, -1, -1};
static void probe_draw_order() {
    if (g_comp_of_draw[0] >= 0) return;
    curandState st; curand_init(0, 0, 0, &st);
    st.u[0] = 0.f; st.u[1] = 1.f; st.u[2] = 2.f;
    float3 f = device_create_random_float3(st);
    g_comp_of_draw[(int)f.x] = 0; g_comp_of_draw[(int)f.y] = 1; g_comp_of_draw[(int)f.z] = 2;
}
static inline void load_raygen_uniforms(curandState& st, const float* U) {
    probe_draw_order();
    for (int k = 0; k < 3; ++k) st.u[k] = U[g_comp_of_draw[k]];
    st.fixed = nullptr; st.cursor = 0;
}

/* Miss, path_tracer.cu:117-122 */
static inline void ref_miss(RayPayload& payload) {
    float3 d = normalize(payload.ray.dir);
    float v = asin(d.z) / M_PI + 0.5, u = atan(d.y / d.x) / 2 / M_PI;
    payload.radiance = make_float3(PocaTextureUtils::GetTexture2D(payload.sky_tex_obj, u, v));
    payload.recursion_depth = MAX_RECURSION_DEPTH_SET;
}

extern "C" {

void ref_reset(void) {
    if (g_built) { SceneBVH::ReleaseBVH(); g_built = false; g_root = nullptr; }
    else if (!g_objs.empty()) { SceneBVH::ReleaseBVH(); }
    g_objs.clear();
}

int ref_add_object(int prim_type, const float* center, float radius, float y_pos, float height,
                   const ref_material* m) {
    Object* o = new Object();
    memset(o, 0, sizeof(Object));
    o->type_ = (PrimitiveType::Enum)prim_type;
    fill_material(o->material_, m);
    o->center_ = make_float3(center[0], center[1], center[2]);
    o->radius_ = radius;
    o->y_pos_ = y_pos;
    o->height_ = height;
    int id = (int)g_objs.size();
    if (o->type_ == PrimitiveType::Platform) o->height_ = int_as_float(id);
    else o->y_pos_ = int_as_float(id);
    g_objs.push_back(o);
    SceneBVH::AddObject(o);
    return id;
}

int ref_build(void) {
    g_root = SceneBVH::BuildBVH();
    g_built = true;
    return (int)(2 * g_objs.size() - 1);
}

/* SceneBVH::UpdateObject (bvh.cu:144-157): the caller edits its Object and the reference copies it into the leaf and
 * refits the boxes up to the root -- the tree's topology, and with it the leaf visiting order, stays as built. */
int ref_update_object(int id, const float* center, float radius, float y_pos, float height) {
    if (!g_built || id < 0 || id >= (int)g_objs.size()) return -1;
    Object* o = g_objs[id];
    o->center_ = make_float3(center[0], center[1], center[2]);
    o->radius_ = radius;
    if (o->type_ == PrimitiveType::Platform) o->y_pos_ = y_pos;     /* height_ carries the id */
    else o->height_ = height;                                        /* y_pos_ carries the id */
    SceneBVH::UpdateObject(o);
    return 0;
}

/* Reference leaf visiting order: TraceRay (bvh.cu:167-205) pops the right son first, so among
 * exactly-tied hits the leaf met first in a right-first depth-first walk wins.  Returns, per
 * primitive id, its rank in that walk (0 = tested first). */
void ref_leaf_order(int32_t* rank_of_prim) {
    struct NodeMirror { float3 mn, mx; bool is_obj; int l, r; Object o; };
    static_assert(sizeof(NodeMirror) == sizeof(SceneBVH), "node mirror layout");
    const NodeMirror* nodes = (const NodeMirror*)g_root;
    std::vector<int> stack; stack.push_back(0);
    int rank = 0;
    while (!stack.empty()) {
        int n = stack.back(); stack.pop_back();
        if (n == -1) continue;
        if (nodes[n].is_obj) rank_of_prim[prim_id_of(nodes[n].o)] = rank++;
        stack.push_back(nodes[n].l);
        stack.push_back(nodes[n].r);
    }
}

void ref_set_sky(const uint8_t* rgba, int w, int h, int quarter_rows) {
    g_sky.w = w; g_sky.h = h;
    g_sky.rgba.assign((size_t)w * h * 4, 0);
    /* textures.cu:32 copies `width` BYTES of every `width*4`-byte row: only texels x < w/4 arrive */
    size_t row_bytes = quarter_rows ? (size_t)w : (size_t)w * 4;
    for (int y = 0; y < h; ++y) memcpy(&g_sky.rgba[(size_t)y * w * 4], rgba + (size_t)y * w * 4, row_bytes);
}

/* rays8: ox oy oz dx dy dz tmin tmax */
void ref_trace(int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        Ray ray;
        ray.origin = make_float3(rays8[8 * i + 0], rays8[8 * i + 1], rays8[8 * i + 2]);
        ray.dir = make_float3(rays8[8 * i + 3], rays8[8 * i + 4], rays8[8 * i + 5]);
        ray.tmin = rays8[8 * i + 6];
        ray.tmax = rays8[8 * i + 7];
        IntersectionAttributes attr;
        attr.normal = make_float3(0, 0, 0); attr.hit_pos = make_float3(0, 0, 0);
        Object obj;
        bool hit = g_root->TraceRay(ray, attr, obj);
        if (hit) {
            prim[i] = prim_id_of(obj);
            /* TraceRay takes the ray by value (bvh.h:22), so t is recovered from hit_pos:
             * recompute it the way the winning intersector did, via its own code on the returned object */
            Ray r2 = ray; IntersectionAttributes a2; a2.normal = attr.normal; a2.hit_pos = attr.hit_pos;
            obj.IntersectionTest(r2, a2);
            t[i] = r2.tmax;
            normal3[3 * i + 0] = attr.normal.x; normal3[3 * i + 1] = attr.normal.y; normal3[3 * i + 2] = attr.normal.z;
            hitpos3[3 * i + 0] = attr.hit_pos.x; hitpos3[3 * i + 1] = attr.hit_pos.y; hitpos3[3 * i + 2] = attr.hit_pos.z;
        } else {
            prim[i] = -1; t[i] = ray.tmax;
            for (int k = 0; k < 3; ++k) { normal3[3 * i + k] = 0.f; hitpos3[3 * i + k] = 0.f; }
        }
    }
}

/* One material evaluation per item with replayed uniforms (material.cu:145-163 and the 4 shaders). */
void ref_shade(int n, const ref_material* mats, const int32_t* mat_idx, const float* pos3, const float* nrm3,
               const float* in3, const float* u3, float* bounce3, float* atten3, float* rad3, int32_t* ndraws) {
    for (int i = 0; i < n; ++i) {
        Material m; fill_material(m, &mats[mat_idx[i]]);
        curandState st; curand_init(0, 0, 0, &st);
        st.fixed = u3 + 3 * i; st.n_fixed = 3;
        RayPayload payload; memset(&payload, 0, sizeof(payload));
        payload.d_rng_states = &st;
        m.EvalAttenuationAndCreateRay(payload, make_float3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]),
                                      make_float3(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2]),
                                      make_float3(in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]));
        bounce3[3 * i] = payload.bounce_dir.x; bounce3[3 * i + 1] = payload.bounce_dir.y; bounce3[3 * i + 2] = payload.bounce_dir.z;
        atten3[3 * i] = payload.attenuation.x; atten3[3 * i + 1] = payload.attenuation.y; atten3[3 * i + 2] = payload.attenuation.z;
        rad3[3 * i] = payload.radiance.x; rad3[3 * i + 1] = payload.radiance.y; rad3[3 * i + 2] = payload.radiance.z;
        ndraws[i] = st.cursor;
    }
}

/* MotionalCamera::GetCopy (motional_camera.cu:177-200).  out19: u v w top_left horizontal vertical dist */
void ref_camera_setup(int w, int h, const float* origin, const float* look_at, float fov_deg, float lens_radius,
                      float* out19) {
    MotionalCamera cam(w, h, make_float3(origin[0], origin[1], origin[2]),
                       make_float3(look_at[0], look_at[1], look_at[2]));
    cam.SetViewFov(fov_deg);
    cam.lens_radius_ = lens_radius;
    g_cam = cam.GetCopy();
    if (out19) {
        const float3* v[6] = {&g_cam.u_, &g_cam.v_, &g_cam.w_, &g_cam.top_left_corner_, &g_cam.horizontal_, &g_cam.vertical_};
        for (int k = 0; k < 6; ++k) { out19[3 * k] = v[k]->x; out19[3 * k + 1] = v[k]->y; out19[3 * k + 2] = v[k]->z; }
        out19[18] = g_cam.dist_to_focus_;
    }
}

/* Host camera controls (motional_camera.cu:76-168) applied to a caller-held pose.
 * op: 0 left 1 right 2 forward 3 backward 4 up 5 down 6 rot-up 7 rot-down 8 rot-left 9 rot-right 10 scale-fov
 * pose7: origin[3] look_at[3] fov */
void ref_camera_op(float* pose7, int op, float arg) {
    MotionalCamera cam(16, 16, make_float3(pose7[0], pose7[1], pose7[2]), make_float3(pose7[3], pose7[4], pose7[5]));
    cam.SetViewFov(pose7[6]);
    switch (op) {
        case 0: cam.MoveEyeLeft(arg); break;      case 1: cam.MoveEyeRight(arg); break;
        case 2: cam.MoveEyeForward(arg); break;   case 3: cam.MoveEyeBackward(arg); break;
        case 4: cam.MoveEyeUp(arg); break;        case 5: cam.MoveEyeDown(arg); break;
        case 6: cam.RotateAroundUp(arg); break;   case 7: cam.RotateAroundDown(arg); break;
        case 8: cam.RotateAroundLeft(arg); break; case 9: cam.RotateAroundRight(arg); break;
        case 10: cam.ScaleFov(arg); break;
    }
    pose7[0] = cam.origin_.x; pose7[1] = cam.origin_.y; pose7[2] = cam.origin_.z;
    pose7[3] = cam.look_at_.x; pose7[4] = cam.look_at_.y; pose7[5] = cam.look_at_.z;
    pose7[6] = cam.view_fov_;
}

/* MotionalCamera::RayGen (motional_camera.cu:202-213) with replayed uniforms */
void ref_raygen(int n, const int32_t* xy, const float* u3, float* rays8) {
    for (int i = 0; i < n; ++i) {
        curandState st; curand_init(0, 0, 0, &st);
        load_raygen_uniforms(st, u3 + 3 * i);
        Ray r = g_cam.RayGen(xy[2 * i], xy[2 * i + 1], st);
        rays8[8 * i + 0] = r.origin.x; rays8[8 * i + 1] = r.origin.y; rays8[8 * i + 2] = r.origin.z;
        rays8[8 * i + 3] = r.dir.x; rays8[8 * i + 4] = r.dir.y; rays8[8 * i + 5] = r.dir.z;
        rays8[8 * i + 6] = r.tmin; rays8[8 * i + 7] = r.tmax;
    }
}

void ref_sky_lookup(int n, const float* dir3, float* rgb3) {
    for (int i = 0; i < n; ++i) {
        RayPayload p; memset(&p, 0, sizeof(p));
        p.ray.dir = make_float3(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]);
        ref_miss(p);
        rgb3[3 * i] = p.radiance.x; rgb3[3 * i + 1] = p.radiance.y; rgb3[3 * i + 2] = p.radiance.z;
    }
}

/* SamplePixel (path_tracer.cu:124-175) for samples [first_sample, first_sample+spp) of every pixel,
 * summed per pixel in sample order (fp32).  All rows are rendered (quirk: the reference's grid
 * W/16 x H/16 drops trailing rows, path_tracer.cu:277).  sum_rgb: w*h*3; normal_rgb (optional): the
 * first-segment normal of the LAST sample (G-buffer of path_tracer.cu:159-163,174). */
void ref_render(int spp, int first_sample, int max_depth, unsigned long long seed, int nthreads,
                float* sum_rgb, float* normal_rgb, ref_stats* stats) {
    const int W = g_cam.width_, H = g_cam.height_;
    unsigned long long segs = 0, misses = 0, exhausted = 0;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : segs, misses, exhausted)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const uint32_t pixel = (uint32_t)(y * W + x);
            float3 sum = make_float3(0.f);
            float3 first_normal = make_float3(0.f);
            for (int s = first_sample; s < first_sample + spp; ++s) {
                curandState st; curand_init(0, 0, 0, &st);
                float U0[4];
                pt_philox_uniform4(seed, pixel, (uint32_t)s, 0u, U0);
                load_raygen_uniforms(st, U0);
                float3 radiance = make_float3(0.f);
                float3 normals = make_float3(0.f);
                RayPayload payload;
                payload.ray = g_cam.RayGen(x, y, st);
                float3 attenuation = make_float3(1.f);
                payload.recursion_depth = 0;
                payload.d_rng_states = &st;
                payload.sky_tex_obj = 0;
                bool first_recursion = false;
                bool missed = false;
                while (payload.recursion_depth < (uint)max_depth) {
                    IntersectionAttributes attr;
                    Object closet_hit_obj;
                    ++segs;
                    bool ret = g_root->TraceRay(payload.ray, attr, closet_hit_obj);
                    if (ret) {
                        pt_philox_uniform4(seed, pixel, (uint32_t)s, 1u + payload.recursion_depth, st.u); st.cursor = 0;
                        closet_hit_obj.ClosetHit(payload, attr);
                    } else {
                        attr.hit_pos = payload.ray.origin + DEFAULT_RAY_TMAX * payload.ray.dir;
                        attr.normal = -payload.ray.dir;
                        ref_miss(payload);
                        missed = true; ++misses;
                    }
                    radiance += attenuation * payload.radiance;
                    attenuation *= payload.attenuation;
                    if (!first_recursion) { normals += attr.normal; first_recursion = true; }
                    payload.ray.origin = payload.hit_pos;
                    payload.ray.dir = normalize(payload.bounce_dir);
                    payload.ray.tmin = BOUNCE_RAY_TMIN;
                    payload.ray.tmax = DEFAULT_RAY_TMAX;
                    payload.recursion_depth++;
                }
                if (!missed) ++exhausted;
                sum += radiance;
                first_normal = normals;
            }
            sum_rgb[3 * (size_t)pixel + 0] = sum.x; sum_rgb[3 * (size_t)pixel + 1] = sum.y; sum_rgb[3 * (size_t)pixel + 2] = sum.z;
            if (normal_rgb) {
                normal_rgb[3 * (size_t)pixel + 0] = first_normal.x; normal_rgb[3 * (size_t)pixel + 1] = first_normal.y;
                normal_rgb[3 * (size_t)pixel + 2] = first_normal.z;
            }
        }
    }
    if (stats) {
        stats->paths = (unsigned long long)W * H * spp; stats->segments = segs;
        stats->misses = misses; stats->depth_exhausted = exhausted;
    }
}

int ref_sizeof(int what) {
    switch (what) { case 0: return sizeof(Object); case 1: return sizeof(Material); case 2: return sizeof(SceneBVH);
                    case 3: return sizeof(MotionalCamera); case 4: return sizeof(Ray); default: return -1; }
}
int ref_max_threads(void) { return omp_get_max_threads(); }
}
