// pt_wavefront.cuh -- the wavefront integrator stages (sm_100a).
//
// The reference traces one full path per thread in a megakernel (SamplePixel,
// cuSrc/path_tracer.cu:124-175), so a warp runs until its longest path ends and every material
// branch is taken by every warp.  Here a wave of paths advances one segment per iteration through
//     ray-gen -> [ closest-hit traversal -> bin by shader -> shade / miss -> compact survivors ] x depth -> accumulate
// and the stages talk through SoA float4 queues in HBM:
//     ray_o[slot] = (origin.xyz, bits pixel)      ray_d[slot] = (dir.xyz, bits sample-in-wave)
//     thr[slot]   = (throughput.rgb, -)           hit[slot]   = (bits t, prim<<3 | part)   or code 0xFFFFFFFF = miss
// Every stage is a persistent grid (a multiple of the SM count) whose warps pull work with one atomic
// per warp and read the queue sizes from a device-side control block, so a whole frame is enqueued
// without a single host synchronisation.
#pragma once
#include "pt_math.cuh"

namespace pt {

#define PT_MISS_CODE 0xFFFFFFFFu
#define PT_STACK 64

// one control block per bounce, zeroed once per wave
struct BounceCtrl {
    uint32_t in_count;            // rays in the input queue of this bounce
    uint32_t trav_head;           // traversal work counter
    uint32_t shade_head;          // shade work counter (in 32-item chunks)
    uint32_t bin_count[NUM_BINS]; // items per shader bin
    uint32_t pad[8];
};
static_assert(sizeof(BounceCtrl) == 64, "control block is one 64-byte line");

// Primitives live on the device in the reference's leaf VISITING order (or insertion order when
// PT_OPT_TIE_ORDER is off): "device index" below always means that position.
struct SceneDev {
    const float4* nodes;      // LBVH, 4 x float4 per inner node
    const float4* geo0;       // (cx, cy, cz, radius)
    const float4* geo1;       // (height, y_pos, bits type, bits (bin<<28 | material slot))
    const MaterialRec* mats;
    const float* plat_y;      // platforms: side list tested for every ray
    const int* plat_id;
    const int* orig_id;       // device primitive index -> caller's primitive id (parity outputs only)
    int n_plat;
    int n_tree_prims;         // primitives in the LBVH (0: no tree)
    cudaTextureObject_t sky;
    int has_emitters;
};

struct WaveQueues {
    float4* ray_o[2];
    float4* ray_d[2];
    float4* thr[2];
    uint2* hit;
    uint32_t* bin_items;      // NUM_BINS x capacity slot indices
    float4* path_out;         // per path of the wave: radiance (rgb, -)
    BounceCtrl* ctrl;         // [max_depth + 1]
    uint32_t capacity;
};

// ---- stage 0: camera rays (MotionalCamera::RayGen, motional_camera.cu:202-213) ----------------------
__global__ void __launch_bounds__(256) k_raygen(CameraFrame cam, uint2 key, uint32_t npix, uint32_t n_paths,
                                                uint32_t first_sample, float4* __restrict__ ray_o,
                                                float4* __restrict__ ray_d, float4* __restrict__ thr,
                                                float4* __restrict__ path_out, int zero_out,
                                                BounceCtrl* __restrict__ ctrl0) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x; pid < n_paths; pid += stride) {
        const uint32_t s = pid / npix;
        const uint32_t pixel = pid - s * npix;
        const int y = pixel / (uint32_t)cam.width;
        const int x = pixel - y * (uint32_t)cam.width;
        float4 U = philox_uniform4(key, pixel, first_sample + s, 0u);
        float3 o, d;
        camera_ray(cam, x, y, U, o, d);
        ray_o[pid] = make_float4(o.x, o.y, o.z, __uint_as_float(pixel));
        ray_d[pid] = make_float4(d.x, d.y, d.z, __uint_as_float(s));
        thr[pid] = make_float4(1.f, 1.f, 1.f, 0.f);
        if (zero_out) path_out[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl0->in_count = n_paths;
}

// ---- stage 1: closest hit ------------------------------------------------------------------------------
// Replaces SceneBVH::TraceRay + Object::IntersectionTest (bvh.cu:167-205, object.cu:10-128): near-child-first
// stack traversal of the 64-byte LBVH nodes, tagged sphere / cylinder leaf test, platforms from the side list.
struct TraceResult { float t; uint32_t code; };

__device__ __forceinline__ TraceResult trace_closest(const SceneDev& sc, float3 o, float3 d, float tmin, float tmax) {
    float best_t = tmax;
    uint32_t best_code = PT_MISS_CODE;
    // Exact ties: the reference keeps whichever tied primitive its traversal meets first (strict `<` at
    // object.cu:19,27,40,57,71,94,103).  Ties are not rare: far from the camera `b*b - a*c` (object.cu:16)
    // cancels so coarsely that the two shells of a hollow glass sphere often get the same t, and cylinder
    // caps rest exactly on the floor plane (video_renderer.cpp:78,105-106).  Device primitive indices ARE
    // the reference's leaf visiting order (see upload_and_build), so a tied candidate wins iff its index
    // is lower than the current holder's.
    uint32_t best_idx = 0u;   // 0 = cannot be displaced (nothing yet, or the first-visited primitive)
    auto tie_tmax = [&](uint32_t idx) -> float {
        // the tmax handed to the intersector: one ulp above best_t when this primitive outranks the holder,
        // so that `t < tmax` also admits t == best_t
        return (idx < best_idx) ? __uint_as_float(__float_as_uint(best_t) + 1u) : best_t;
    };
    // platforms first: they also give the tree walk an early upper bound
    for (int k = 0; k < sc.n_plat; ++k) {
        const uint32_t prim = (uint32_t)sc.plat_id[k];
        float tm = tie_tmax(prim);
        if (hit_platform(sc.plat_y[k], o, d, tmin, tm)) { best_t = tm; best_code = (prim << 3) | PART_PLATFORM; best_idx = prim; }
    }
    if (sc.n_tree_prims > 0) {
        const float ix = 1.0f / (fabsf(d.x) > 1e-30f ? d.x : copysignf(1e-30f, d.x));
        const float iy = 1.0f / (fabsf(d.y) > 1e-30f ? d.y : copysignf(1e-30f, d.y));
        const float iz = 1.0f / (fabsf(d.z) > 1e-30f ? d.z : copysignf(1e-30f, d.z));
        int stack[PT_STACK];
        int sp = 0;
        int node = 0;
        const int SENTINEL = 0x7fffffff;
        stack[sp++] = SENTINEL;
        while (node != SENTINEL) {
            if (node >= 0) {
                const float4 n0 = __ldg(&sc.nodes[4 * node + 0]);
                const float4 n1 = __ldg(&sc.nodes[4 * node + 1]);
                const float4 n2 = __ldg(&sc.nodes[4 * node + 2]);
                const float4 n3 = __ldg(&sc.nodes[4 * node + 3]);
                const float a0x = (n0.x - o.x) * ix, a1x = (n0.y - o.x) * ix;
                const float a0y = (n0.z - o.y) * iy, a1y = (n0.w - o.y) * iy;
                const float a0z = (n2.x - o.z) * iz, a1z = (n2.y - o.z) * iz;
                const float b0x = (n1.x - o.x) * ix, b1x = (n1.y - o.x) * ix;
                const float b0y = (n1.z - o.y) * iy, b1y = (n1.w - o.y) * iy;
                const float b0z = (n2.z - o.z) * iz, b1z = (n2.w - o.z) * iz;
                const float ea = fmaxf(fmaxf(fminf(a0x, a1x), fminf(a0y, a1y)), fmaxf(fminf(a0z, a1z), tmin));
                const float xa = fminf(fminf(fmaxf(a0x, a1x), fmaxf(a0y, a1y)), fminf(fmaxf(a0z, a1z), best_t)) * 1.0000005f;
                const float eb = fmaxf(fmaxf(fminf(b0x, b1x), fminf(b0y, b1y)), fmaxf(fminf(b0z, b1z), tmin));
                const float xb = fminf(fminf(fmaxf(b0x, b1x), fmaxf(b0y, b1y)), fminf(fmaxf(b0z, b1z), best_t)) * 1.0000005f;
                const bool ha = ea <= xa, hb = eb <= xb;
                const int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
                if (ha && hb) {
                    const bool a_near = ea <= eb;
                    node = a_near ? ca : cb;
                    stack[sp++] = a_near ? cb : ca;
                } else if (ha) node = ca;
                else if (hb) node = cb;
                else node = stack[--sp];
            } else {
                const uint32_t leaf = (uint32_t)~node;
                const uint32_t prim = leaf >> 2;
                const float4 g0 = __ldg(&sc.geo0[prim]);
                int part = 0;
                float tm = tie_tmax(prim);
                bool h;
                if ((leaf & 3u) == 2u) h = hit_cylinder(g0, __ldg(&sc.geo1[prim]).x, o, d, tmin, tm, part);
                else h = hit_sphere(g0, o, d, tmin, tm, part);
                if (h) { best_t = tm; best_code = (prim << 3) | (uint32_t)part; best_idx = prim; }
                node = stack[--sp];
            }
        }
    }
    TraceResult r; r.t = best_t; r.code = best_code;
    return r;
}

template <bool PER_RAY_RANGE>
__global__ void __launch_bounds__(128) k_traverse(SceneDev sc, const float4* __restrict__ ray_o,
                                                  const float4* __restrict__ ray_d, const float2* __restrict__ ray_range,
                                                  float tmin_all, BounceCtrl* __restrict__ ctrl,
                                                  uint2* __restrict__ hit, uint32_t* __restrict__ bin_items,
                                                  uint32_t capacity) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = ctrl->in_count;
    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctrl->trav_head, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t slot = base + lane;
        const bool valid = slot < n;
        int bin = -1;
        if (valid) {
            const float4 o4 = ray_o[slot];
            const float4 d4 = ray_d[slot];
            float tmin = tmin_all, tmax = PT_TMAX;
            if (PER_RAY_RANGE) { float2 r = ray_range[slot]; tmin = r.x; tmax = r.y; }
            TraceResult r = trace_closest(sc, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), tmin, tmax);
            hit[slot] = make_uint2(__float_as_uint(r.t), r.code);
            bin = (r.code == PT_MISS_CODE) ? BIN_MISS : (int)(__float_as_uint(__ldg(&sc.geo1[r.code >> 3]).w) >> 28);
        }
        // bin the finished rays by shader: one atomic per warp and bin
#pragma unroll
        for (int b = 0; b < NUM_BINS; ++b) {
            const uint32_t m = __ballot_sync(0xffffffffu, bin == b);
            if (m) {
                uint32_t off = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) off = atomicAdd(&ctrl->bin_count[b], (uint32_t)__popc(m));
                off = __shfl_sync(0xffffffffu, off, __ffs(m) - 1);
                if (bin == b) bin_items[(size_t)b * capacity + off + __popc(m & ((1u << lane) - 1u))] = slot;
            }
        }
    }
}

// ---- stage 2: shade / miss, binned by shader; survivors are compacted into the next ray queue ------------
// Replaces Object::ClosetHit -> Material::EvalAttenuationAndCreateRay (object.cu:130-132, material.cu:145-163),
// Miss (path_tracer.cu:117-122) and the per-segment bookkeeping of SamplePixel (path_tracer.cu:156-170).
__global__ void __launch_bounds__(128) k_shade(SceneDev sc, uint2 key, uint32_t npix, uint32_t first_sample,
                                               int bounce, int max_depth, const float4* __restrict__ ray_o,
                                               const float4* __restrict__ ray_d, const float4* __restrict__ thr_in,
                                               const uint2* __restrict__ hit, const uint32_t* __restrict__ bin_items,
                                               uint32_t capacity, BounceCtrl* __restrict__ ctrl,
                                               BounceCtrl* __restrict__ ctrl_next, float4* __restrict__ out_o,
                                               float4* __restrict__ out_d, float4* __restrict__ out_thr,
                                               float4* __restrict__ path_out) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t chunk_end[NUM_BINS];
    uint32_t total = 0;
#pragma unroll
    for (int b = 0; b < NUM_BINS; ++b) { total += (ctrl->bin_count[b] + 31u) >> 5; chunk_end[b] = total; }
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&ctrl->shade_head, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= total) break;
        int bin = 0;
#pragma unroll
        for (int b = 0; b < NUM_BINS - 1; ++b) bin += (chunk >= chunk_end[b]) ? 1 : 0;
        const uint32_t first_chunk = bin ? chunk_end[bin - 1] : 0u;
        const uint32_t item = ((chunk - first_chunk) << 5) + lane;
        const bool valid = item < ctrl->bin_count[bin];
        bool alive = false;
        float3 pos = f3(0, 0, 0), ndir = f3(0, 0, 0), thr = f3(0, 0, 0);
        uint32_t pixel = 0, s = 0;
        if (valid) {
            const uint32_t slot = bin_items[(size_t)bin * capacity + item];
            const float4 o4 = ray_o[slot];
            const float4 d4 = ray_d[slot];
            const float4 t4 = thr_in[slot];
            const uint2 h = hit[slot];
            pixel = __float_as_uint(o4.w);
            s = __float_as_uint(d4.w);
            const size_t pid = (size_t)s * npix + pixel;
            const float3 o = f3(o4.x, o4.y, o4.z), d = f3(d4.x, d4.y, d4.z);
            thr = f3(t4.x, t4.y, t4.z);
            if (bin == BIN_MISS) {
                const float3 c = thr * sky_radiance(sc.sky, d);
                if (sc.has_emitters) { float4 p = path_out[pid]; path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f); }
                else path_out[pid] = make_float4(c.x, c.y, c.z, 0.f);
            } else {
                const uint32_t prim = h.y >> 3;
                const float4 g0 = __ldg(&sc.geo0[prim]);
                const uint32_t mslot = __float_as_uint(__ldg(&sc.geo1[prim]).w) & 0x0fffffffu;
                const MaterialRec m = sc.mats[mslot];
                float3 normal;
                hit_attributes(g0, (int)(h.y & 7u), o, d, __uint_as_float(h.x), pos, normal);
                const float4 U = philox_uniform4(key, pixel, first_sample + s, 1u + (uint32_t)bounce);
                Scatter sct;
                shade_bin(bin, m, normal, d, U, sct);
                if (sc.has_emitters) {
                    const float3 c = thr * sct.radiance;
                    float4 p = path_out[pid];
                    path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f);
                }
                thr = thr * sct.atten;
                ndir = normalize(sct.bounce);
                // a path whose throughput is exactly zero can add nothing more (finite sky and emission):
                // ending it here is result-identical to the reference, which keeps bouncing it (path_tracer.cu:141-170)
                alive = (bounce + 1 < max_depth) && (thr.x != 0.f || thr.y != 0.f || thr.z != 0.f);
                if (!alive && !sc.has_emitters) path_out[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, alive);
        if (m) {
            uint32_t off = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) off = atomicAdd(&ctrl_next->in_count, (uint32_t)__popc(m));
            off = __shfl_sync(0xffffffffu, off, __ffs(m) - 1);
            if (alive) {
                const uint32_t j = off + __popc(m & ((1u << lane) - 1u));
                out_o[j] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(pixel));
                out_d[j] = make_float4(ndir.x, ndir.y, ndir.z, __uint_as_float(s));
                out_thr[j] = make_float4(thr.x, thr.y, thr.z, 0.f);
            }
        }
    }
}

// ---- stage 3: accumulate the wave into the frame's SUM buffer, sample order, fp32 -------------------------
// (the reference's Mix keeps a running mean of CLAMPED, denoised samples, path_tracer.cu:241-254; the
// unclamped sum is what spp-sharding across GPUs needs -- the display pipeline is pt_dispatch_ray)
__global__ void __launch_bounds__(256) k_accumulate(uint32_t npix, uint32_t n_samples, const float4* __restrict__ path_out,
                                                    float4* __restrict__ accum) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        float4 a = accum[p];
        for (uint32_t s = 0; s < n_samples; ++s) {
            const float4 r = path_out[(size_t)s * npix + p];
            a.x += r.x; a.y += r.y; a.z += r.z;
        }
        a.w += (float)n_samples;
        accum[p] = a;
    }
}

// per-wave bookkeeping: rays traced = sum of the bounce input queues (diagnostic counter, 1 thread)
__global__ void k_wave_stats(int max_depth, const BounceCtrl* __restrict__ ctrl, unsigned long long* __restrict__ segments) {
    unsigned long long n = 0;
    for (int b = 0; b < max_depth; ++b) n += ctrl[b].in_count;
    *segments += n;
}

}  // namespace pt
