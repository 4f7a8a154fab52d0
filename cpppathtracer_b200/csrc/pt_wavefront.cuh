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
#define PT_FULL 0xffffffffu

// One control block per bounce, zeroed once per wave.  Every counter that takes atomics sits on its own
// 128-byte line (its own L2 slice queue): v1 kept them in one 64-byte line and k_shade spent its time
// waiting on ~300 k same-line atomics per launch (profiles/r01_v1_summary.md).
struct Counter128 { uint32_t v; uint32_t pad[31]; };
struct BounceCtrl {
    Counter128 in_count;            // slots in the input queue of this bounce (holes included), written once
    Counter128 rays;                // valid rays in that queue (statistics)
    Counter128 trav_head;           // traversal work counter
    Counter128 shade_head;          // shade work counter, in 32-item chunks
    Counter128 bin_count[NUM_BINS]; // items per shader bin
};
static_assert(sizeof(BounceCtrl) == 128 * (4 + NUM_BINS), "one line per counter");
#define PT_INVALID_SLOT 0xFFFFFFFFu   // ray_o.w of a queue slot whose path ended (hole)

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
    // hit queues, one per shader bin (NUM_BINS x capacity each): the complete shading input of a finished ray
    float4* hq_o; float4* hq_d; float4* hq_thr; uint2* hq_hit;
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
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl0->in_count.v = n_paths; ctrl0->rays.v = n_paths; }
}

// ---- stage 1: closest hit ------------------------------------------------------------------------------
// Replaces SceneBVH::TraceRay + Object::IntersectionTest (bvh.cu:167-205, object.cu:10-128): near-child-first
// stack traversal of the 64-byte LBVH nodes, tagged sphere / cylinder leaf test, platforms from the side list.
//
// Exact ties: the reference keeps whichever tied primitive its traversal meets first (strict `<` at
// object.cu:19,27,40,57,71,94,103).  Ties are not rare: far from the camera `b*b - a*c` (object.cu:16) cancels
// so coarsely that the two shells of a hollow glass sphere often get the same t, and cylinder caps rest exactly
// on the floor plane (video_renderer.cpp:78,105-106).  Device primitive indices ARE the reference's leaf
// visiting order (upload_and_build in pt_api.cu), so a tied candidate wins iff its index is lower than the
// current holder's: the intersector is then handed a tmax one ulp above best_t, so its `t < tmax` admits t == best_t.
// Persistent warps with lane refill: a lane whose ray finished takes the next ray of the queue while its
// neighbours keep walking, so a warp no longer idles until its longest ray ends (v1: 12.7 of 32 lanes active
// per instruction).  Finished rays are staged per warp in shared memory and binned by shader 32 at a time,
// one atomic per warp and bin.
#ifndef PT_TRAV_WARPS
#define PT_TRAV_WARPS 4
#endif
#ifndef PT_TRAV_MIN_BLOCKS
#define PT_TRAV_MIN_BLOCKS 1
#endif
#define PT_REFILL_MIN 32       // refill when at least this many lanes are idle (32: a warp refills as a whole;
                               // measured on C2: partial refills gain nothing, profiles/r01_v2_sweep.md)
#define PT_STEPS_PER_CHECK 8   // traversal rounds between two refill / retire checks

struct LaneRay {
    float3 o, d;
    float ix, iy, iz;      // 1/d
    float ox, oy, oz;      // o/d, so that a slab plane distance is one FMA: plane*(1/d) - o/d
    float slack;           // bound on the rounding error of those FMAs (added to the exit distance)
    float tmin, best_t;
    uint32_t best_code;
    uint32_t slot;
    int node, sp;
};
#define PT_SENTINEL 0x7fffffff

__device__ __forceinline__ float tie_tmax(uint32_t idx, float best_t, uint32_t best_code) {
    // see trace_closest: one ulp above best_t when primitive idx outranks the current holder
    const uint32_t holder = (best_code == PT_MISS_CODE) ? 0u : (best_code >> 3);
    return (idx < holder) ? __uint_as_float(__float_as_uint(best_t) + 1u) : best_t;
}

__device__ __forceinline__ void lane_begin(const SceneDev& sc, LaneRay& r, int* stack) {
    r.best_code = PT_MISS_CODE;
    for (int k = 0; k < sc.n_plat; ++k) {
        const uint32_t prim = (uint32_t)sc.plat_id[k];
        float tm = tie_tmax(prim, r.best_t, r.best_code);
        if (hit_platform(sc.plat_y[k], r.o, r.d, r.tmin, tm)) { r.best_t = tm; r.best_code = (prim << 3) | PART_PLATFORM; }
    }
    r.ix = 1.0f / (fabsf(r.d.x) > 1e-30f ? r.d.x : copysignf(1e-30f, r.d.x));
    r.iy = 1.0f / (fabsf(r.d.y) > 1e-30f ? r.d.y : copysignf(1e-30f, r.d.y));
    r.iz = 1.0f / (fabsf(r.d.z) > 1e-30f ? r.d.z : copysignf(1e-30f, r.d.z));
    r.ox = r.o.x * r.ix; r.oy = r.o.y * r.iy; r.oz = r.o.z * r.iz;
    // |fma(p, 1/d, -o/d) - (p-o)/d| <= ~3 ulp of |o/d| (rounding of 1/d, of o/d and of the fma); culling only has
    // to be conservative, so both interval ends get that much slack (plus the 4-ulp inflation of the leaf boxes)
    r.slack = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(r.ox), fabsf(r.oy)), fabsf(r.oz));
    r.sp = 0;
    stack[r.sp++] = PT_SENTINEL;
    r.node = sc.n_tree_prims > 0 ? 0 : PT_SENTINEL;
}

// One 64-byte node = two 256-bit loads (LDG.E.256, new on sm_100): half the LSU instructions and L1 wavefronts
// of four 128-bit loads -- the walk is bound by the L1 wavefront pipe on incoherent bounces.
__device__ __forceinline__ void load_node(const float4* __restrict__ nodes, int node, float4& n0, float4& n1, float4& n2,
                                          float4& n3) {
    const float4* p = nodes + 4 * (size_t)node;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w)
                 : "l"(p));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w)
                 : "l"(p + 2));
}

__device__ __forceinline__ void lane_inner(const SceneDev& sc, LaneRay& r, int* stack) {
    {
        float4 n0, n1, n2, n3;
        load_node(sc.nodes, r.node, n0, n1, n2, n3);
        const float a0x = fmaf(n0.x, r.ix, -r.ox), a1x = fmaf(n0.y, r.ix, -r.ox);
        const float a0y = fmaf(n0.z, r.iy, -r.oy), a1y = fmaf(n0.w, r.iy, -r.oy);
        const float a0z = fmaf(n2.x, r.iz, -r.oz), a1z = fmaf(n2.y, r.iz, -r.oz);
        const float b0x = fmaf(n1.x, r.ix, -r.ox), b1x = fmaf(n1.y, r.ix, -r.ox);
        const float b0y = fmaf(n1.z, r.iy, -r.oy), b1y = fmaf(n1.w, r.iy, -r.oy);
        const float b0z = fmaf(n2.z, r.iz, -r.oz), b1z = fmaf(n2.w, r.iz, -r.oz);
        const float lo = r.tmin - r.slack, hi = r.best_t * 1.0000005f + r.slack;
        const float ea = fmaxf(fmaxf(fminf(a0x, a1x), fminf(a0y, a1y)), fmaxf(fminf(a0z, a1z), lo));
        const float xa = fminf(fminf(fmaxf(a0x, a1x), fmaxf(a0y, a1y)), fminf(fmaxf(a0z, a1z), hi));
        const float eb = fmaxf(fmaxf(fminf(b0x, b1x), fminf(b0y, b1y)), fmaxf(fminf(b0z, b1z), lo));
        const float xb = fminf(fminf(fmaxf(b0x, b1x), fmaxf(b0y, b1y)), fminf(fmaxf(b0z, b1z), hi));
        const bool ha = ea <= xa, hb = eb <= xb;
        const int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
        if (ha && hb) {
            const bool a_near = ea <= eb;
            r.node = a_near ? ca : cb;
            stack[r.sp++] = a_near ? cb : ca;
        } else if (ha) r.node = ca;
        else if (hb) r.node = cb;
        else r.node = stack[--r.sp];
    }
}
__device__ __forceinline__ void lane_leaf(const SceneDev& sc, LaneRay& r, int* stack) {
    {
        const uint32_t leaf = (uint32_t)~r.node;
        const uint32_t prim = leaf >> 2;
        const float4 g0 = __ldg(&sc.geo0[prim]);
        int part = 0;
        float tm = tie_tmax(prim, r.best_t, r.best_code);
        bool h;
        if ((leaf & 3u) == 2u) h = hit_cylinder(g0, __ldg(&sc.geo1[prim]).x, r.o, r.d, r.tmin, tm, part);
        else h = hit_sphere(g0, r.o, r.d, r.tmin, tm, part);
        if (h) { r.best_t = tm; r.best_code = (prim << 3) | (uint32_t)part; }
        r.node = stack[--r.sp];
    }
}

// The hit queues: what a finished ray hands to the shade stage, one queue per shader bin, SoA float4:
//   hq_o = (origin, bits pixel)  hq_d = (dir, bits sample)  hq_thr = (throughput)  hq_hit = (bits t, prim<<3|part)
// The traversal warp that finished the rays also MOVES their records there (32 at a time, positions from one
// atomic per warp and bin), so the shade stage streams its bin in order instead of gathering 16-byte pieces
// through an index list (v3: 2x read amplification and a dependent DRAM round trip in k_shade).  The copy rides on
// the memory system the ALU/L1-bound walk leaves idle (k_traverse used 2-7 % of DRAM bandwidth).
struct HitQueues {
    float4* o; float4* d; float4* thr; uint2* hit;
    uint32_t capacity;     // entries per bin
};
struct StageEntry { uint32_t slot_bin; uint32_t t; uint32_t code; };   // slot_bin = bin<<29 | slot

__device__ __forceinline__ void flush_stage(StageEntry* stage, uint32_t& stage_n, uint32_t lane, BounceCtrl* ctrl,
                                            const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                            const float4* __restrict__ thr, const HitQueues& hq, bool everything) {
    while (stage_n >= 32u || (everything && stage_n > 0u)) {
        const uint32_t cnt = stage_n < 32u ? stage_n : 32u;
        StageEntry e; e.slot_bin = 0xFFFFFFFFu; e.t = 0u; e.code = 0u;
        if (lane < cnt) e = stage[lane];
        const int bin = lane < cnt ? (int)(e.slot_bin >> 29) : -1;
        const uint32_t slot = e.slot_bin & 0x1FFFFFFFu;
        float4 o4 = make_float4(0, 0, 0, 0), d4 = o4, t4 = o4;
        if (lane < cnt) { o4 = ray_o[slot]; d4 = ray_d[slot]; t4 = thr[slot]; }
#pragma unroll
        for (int b = 0; b < NUM_BINS; ++b) {
            const uint32_t m = __ballot_sync(PT_FULL, bin == b);
            if (m) {
                uint32_t off = 0;
                const int leader = __ffs(m) - 1;
                if ((int)lane == leader) off = atomicAdd(&ctrl->bin_count[b].v, (uint32_t)__popc(m));
                off = __shfl_sync(PT_FULL, off, leader);
                if (bin == b) {
                    const size_t j = (size_t)b * hq.capacity + off + __popc(m & ((1u << lane) - 1u));
                    hq.o[j] = o4; hq.d[j] = d4; hq.thr[j] = t4; hq.hit[j] = make_uint2(e.t, e.code);
                }
            }
        }
        __syncwarp();
        const uint32_t rest = stage_n - cnt;
        StageEntry moved; moved.slot_bin = 0u; moved.t = 0u; moved.code = 0u;
        if (lane < rest) moved = stage[32u + lane];
        __syncwarp();
        if (lane < rest) stage[lane] = moved;
        stage_n = rest;
        __syncwarp();
    }
}

template <bool PER_RAY_RANGE, bool COUNT_STEPS>
__global__ void __launch_bounds__(32 * PT_TRAV_WARPS, PT_TRAV_MIN_BLOCKS) k_traverse(SceneDev sc, const float4* __restrict__ ray_o,
                                                                const float4* __restrict__ ray_d,
                                                                const float4* __restrict__ ray_thr,
                                                                const float2* __restrict__ ray_range, float tmin_all,
                                                                BounceCtrl* __restrict__ ctrl, uint2* __restrict__ slot_hit,
                                                                HitQueues hq, uint32_t refill_min, int steps_per_check,
                                                                int inner_run,
                                                                unsigned long long* __restrict__ step_counters) {
    __shared__ StageEntry s_stage[PT_TRAV_WARPS][64];
    const uint32_t lane = threadIdx.x & 31u;
    StageEntry* stage = s_stage[threadIdx.x >> 5];
    uint32_t stage_n = 0;
    const uint32_t n = ctrl->in_count.v;
    int stack[PT_STACK];
    LaneRay r;
    r.node = PT_SENTINEL; r.sp = 0; r.slot = 0; r.best_t = 0.f; r.best_code = PT_MISS_CODE; r.tmin = 0.f;
    r.o = r.d = f3(0.f, 0.f, 0.f); r.ix = r.iy = r.iz = r.ox = r.oy = r.oz = r.slack = 0.f;
    bool active = false;
    bool exhausted = false;     // warp-uniform: the queue has no more rays to hand out
    uint32_t n_inner = 0, n_leaf = 0;   // COUNT_STEPS only
    while (true) {
        // ---- refill idle lanes ----
        const uint32_t idle = __ballot_sync(PT_FULL, !active);
        if (idle != 0u && !exhausted && ((uint32_t)__popc(idle) >= refill_min || idle == PT_FULL)) {
            const uint32_t want = (uint32_t)__popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctrl->trav_head.v, want);
            base = __shfl_sync(PT_FULL, base, 0);
            if (base + want >= n) exhausted = true;
            if (!active) {
                const uint32_t slot = base + __popc(idle & ((1u << lane) - 1u));
                if (slot < n) {
                    const float4 o4 = ray_o[slot];
                    if (__float_as_uint(o4.w) != PT_INVALID_SLOT) {
                        const float4 d4 = ray_d[slot];
                        r.o = f3(o4.x, o4.y, o4.z); r.d = f3(d4.x, d4.y, d4.z);
                        r.tmin = tmin_all; r.best_t = PT_TMAX;
                        if (PER_RAY_RANGE) { const float2 rr = ray_range[slot]; r.tmin = rr.x; r.best_t = rr.y; }
                        r.slot = slot;
                        lane_begin(sc, r, stack);
                        active = true;
                    }
                }
            }
        }
        if (!__any_sync(PT_FULL, active)) {
            if (exhausted) break;
            continue;                       // a refill that only found holes: fetch again
        }
        // ---- walk ----
        // while-while: a lane first runs down inner nodes (up to inner_run of them), then tests the leaf it
        // reached; the warp reconverges between the two so node steps and primitive tests each run with as many
        // lanes as possible instead of interleaving (v1 "if-if": 12.7 of 32 lanes active per instruction)
#pragma unroll 1
        for (int it = 0; it < steps_per_check; ++it) {
            if (active) {
                int k = 0;
#pragma unroll 1
                while (r.node >= 0 && r.node != PT_SENTINEL && k < inner_run) { lane_inner(sc, r, stack); ++k; if (COUNT_STEPS) ++n_inner; }
            }
            if (active && r.node < 0) { lane_leaf(sc, r, stack); if (COUNT_STEPS) ++n_leaf; }
        }
        // ---- retire finished lanes ----
        const bool fin = active && r.node == PT_SENTINEL;
        const uint32_t fm = __ballot_sync(PT_FULL, fin);
        if (fm) {
            if (fin) {
                // the parity entry point (pt_trace_rays) reads results by ray slot; the render path does not
                if (PER_RAY_RANGE) slot_hit[r.slot] = make_uint2(__float_as_uint(r.best_t), r.best_code);
                const uint32_t bin = (r.best_code == PT_MISS_CODE)
                                         ? (uint32_t)BIN_MISS
                                         : (__float_as_uint(__ldg(&sc.geo1[r.best_code >> 3]).w) >> 28);
                StageEntry e; e.slot_bin = (bin << 29) | r.slot; e.t = __float_as_uint(r.best_t); e.code = r.best_code;
                stage[stage_n + __popc(fm & ((1u << lane) - 1u))] = e;
                active = false;
            }
            stage_n += (uint32_t)__popc(fm);
            __syncwarp();
            flush_stage(stage, stage_n, lane, ctrl, ray_o, ray_d, ray_thr, hq, false);
        }
    }
    flush_stage(stage, stage_n, lane, ctrl, ray_o, ray_d, ray_thr, hq, true);
    if (COUNT_STEPS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { n_inner += __shfl_xor_sync(PT_FULL, n_inner, o); n_leaf += __shfl_xor_sync(PT_FULL, n_leaf, o); }
        if (lane == 0) { atomicAdd(&step_counters[0], (unsigned long long)n_inner); atomicAdd(&step_counters[1], (unsigned long long)n_leaf); }
    }
}

// ---- stage 2: shade / miss, binned by shader; survivors are compacted into the next ray queue ------------
// Replaces Object::ClosetHit -> Material::EvalAttenuationAndCreateRay (object.cu:130-132, material.cu:145-163),
// Miss (path_tracer.cu:117-122) and the per-segment bookkeeping of SamplePixel (path_tracer.cu:156-170).
__global__ void __launch_bounds__(128) k_shade(SceneDev sc, uint2 key, uint32_t npix, uint32_t first_sample,
                                               int bounce, int max_depth, HitQueues hq, BounceCtrl* __restrict__ ctrl,
                                               BounceCtrl* __restrict__ ctrl_next, float4* __restrict__ out_o,
                                               float4* __restrict__ out_d, float4* __restrict__ out_thr,
                                               float4* __restrict__ path_out, float4* __restrict__ gbuf_normal) {
    const uint32_t lane = threadIdx.x & 31u;
    // Work items are 32-item chunks of the five shader bins, so a warp never mixes shaders.  The output slot of
    // a hit item is a pure function of (bin, item): next queue = [diffuse | glossy | mix | glass] survivors in bin
    // order.  No atomics for the compaction (v1 took one same-address atomic per warp here); the few paths that
    // die with zero throughput leave a hole marked PT_INVALID_SLOT that the next traversal skips.
    uint32_t chunk_end[NUM_BINS], bin_n[NUM_BINS], out_base[NUM_BINS];
    uint32_t total = 0, next_n = 0;
#pragma unroll
    for (int b = 0; b < NUM_BINS; ++b) {
        bin_n[b] = ctrl->bin_count[b].v;
        total += (bin_n[b] + 31u) >> 5; chunk_end[b] = total;
        out_base[b] = next_n;
        if (b != BIN_MISS) next_n += bin_n[b];
    }
    const bool last_bounce = bounce + 1 >= max_depth;
    if (!last_bounce && blockIdx.x == 0 && threadIdx.x == 0) ctrl_next->in_count.v = next_n;
    // chunks taken per atomic: enough work per grab to keep the head counter cold, small enough to balance the tail
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t grab = total / (n_warps * 4u);
    grab = grab < 1u ? 1u : (grab > 8u ? 8u : grab);
    uint32_t alive_count = 0;
    while (true) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&ctrl->shade_head.v, grab);
        first = __shfl_sync(PT_FULL, first, 0);
        if (first >= total) break;
        const uint32_t end = first + grab < total ? first + grab : total;
        for (uint32_t chunk = first; chunk < end; ++chunk) {
            int bin = 0;
            uint32_t first_chunk = 0u, n_bin = bin_n[0], base_bin = out_base[0];
#pragma unroll
            for (int b = 1; b < NUM_BINS; ++b)
                if (chunk >= chunk_end[b - 1]) { bin = b; first_chunk = chunk_end[b - 1]; n_bin = bin_n[b]; base_bin = out_base[b]; }
            const uint32_t item = ((chunk - first_chunk) << 5) + lane;
            const bool valid = item < n_bin;
            if (!valid) continue;
            const size_t q = (size_t)bin * hq.capacity + item;
            const float4 o4 = hq.o[q];
            const float4 d4 = hq.d[q];
            const float4 t4 = hq.thr[q];
            const uint2 h = hq.hit[q];
            const uint32_t pixel = __float_as_uint(o4.w);
            const uint32_t s = __float_as_uint(d4.w);
            const size_t pid = (size_t)s * npix + pixel;
            const float3 o = f3(o4.x, o4.y, o4.z), d = f3(d4.x, d4.y, d4.z);
            float3 thr = f3(t4.x, t4.y, t4.z);
            if (bin == BIN_MISS) {
                const float3 c = thr * sky_radiance(sc.sky, d);
                if (sc.has_emitters) { float4 p = path_out[pid]; path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f); }
                else path_out[pid] = make_float4(c.x, c.y, c.z, 0.f);
                // G-buffer of the display pipeline: a first-segment miss records -dir (path_tracer.cu:152)
                if (gbuf_normal != nullptr && bounce == 0) gbuf_normal[pid] = make_float4(-d.x, -d.y, -d.z, 0.f);
                continue;
            }
            const uint32_t prim = h.y >> 3;
            const float4 g0 = __ldg(&sc.geo0[prim]);
            const uint32_t mslot = __float_as_uint(__ldg(&sc.geo1[prim]).w) & 0x0fffffffu;
            const MaterialRec m = sc.mats[mslot];
            float3 pos, normal;
            hit_attributes(g0, (int)(h.y & 7u), o, d, __uint_as_float(h.x), pos, normal);
            if (gbuf_normal != nullptr && bounce == 0) gbuf_normal[pid] = make_float4(normal.x, normal.y, normal.z, 0.f);
            const float4 U = philox_uniform4(key, pixel, first_sample + s, 1u + (uint32_t)bounce);
            Scatter sct;
            shade_bin(bin, m, normal, d, U, sct);
            if (sc.has_emitters) {
                const float3 c = thr * sct.radiance;
                float4 p = path_out[pid];
                path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f);
            }
            thr = thr * sct.atten;
            // a path whose throughput is exactly zero can add nothing more (finite sky and emission):
            // ending it here is result-identical to the reference, which keeps bouncing it (path_tracer.cu:141-170)
            const bool alive = !last_bounce && (thr.x != 0.f || thr.y != 0.f || thr.z != 0.f);
            if (!alive && !sc.has_emitters) path_out[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!last_bounce) {
                const uint32_t j = base_bin + item;
                if (alive) {
                    const float3 ndir = normalize(sct.bounce);
                    out_o[j] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(pixel));
                    out_d[j] = make_float4(ndir.x, ndir.y, ndir.z, __uint_as_float(s));
                    out_thr[j] = make_float4(thr.x, thr.y, thr.z, 0.f);
                    ++alive_count;
                } else {
                    out_o[j] = make_float4(0.f, 0.f, 0.f, __uint_as_float(PT_INVALID_SLOT));
                }
            }
        }
    }
    // statistics: valid rays handed to the next bounce (one atomic per warp per launch)
    if (!last_bounce) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) alive_count += __shfl_xor_sync(PT_FULL, alive_count, o);
        if (lane == 0 && alive_count) atomicAdd(&ctrl_next->rays.v, alive_count);
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
    // segments[0] rays traced, [1] inner-node steps, [2] primitive tests (the latter two only with PT_OPT_COUNT_STEPS)
    unsigned long long n = 0;
    for (int b = 0; b < max_depth; ++b) n += ctrl[b].rays.v;
    *segments += n;
}

}  // namespace pt
