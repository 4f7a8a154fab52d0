// pt_wavefront.cuh -- the wavefront integrator (sm_100a).
//
// The reference traces one full path per thread in a megakernel (SamplePixel, cuSrc/path_tracer.cu:124-175): a
// warp runs until its longest path ends, every warp executes every material branch, and the RNG state round-trips
// through global memory on every draw.  Here a wave of paths advances one segment per launch of k_bounce:
//     launch 0      : camera rays                          -> closest hit -> sky (path ends) | hit queue 0, binned by shader
//     launch k      : shade the hits of segment k-1        -> survivors traced as segment k -> sky | hit queue k
//     launch depth  : shade only (not enqueued in a scene without emitters: those hits end black)
//     k_accumulate  : wave radiance into the frame's SUM buffer, in sample order
// Paths therefore leave the wave the moment they end (no idle lanes waiting for long paths), a warp shades with
// ONE shader, and a ray lives in registers from the shader that creates it to the hit record it produces.
// Every launch is a persistent grid (a multiple of the SM count) whose warps pull 32-item chunks with one atomic
// per several chunks and read the queue sizes from a device-side control block, so a whole frame is enqueued
// without a single host synchronisation.
#pragma once
#include "pt_math.cuh"

namespace pt {

#define PT_MISS_CODE 0xFFFFFFFFu
#define PT_STACK 68   // Karras trees over 30-bit codes + index tie-break are <= 62 deep, + the sentinel, + 3 entry-grid links
#define PT_FULL 0xffffffffu
// build-time switches of k_bounce (tools/build_variants.sh times the alternatives on one box)
#ifndef PT_RCP_APPROX
#define PT_RCP_APPROX 1
#endif
#ifndef PT_CAMERA_TILES
#define PT_CAMERA_TILES 1
#endif
#ifndef PT_PREFETCH_FAR
#define PT_PREFETCH_FAR 1
#endif
#ifndef PT_ENTRY_GRID_CODE
#define PT_ENTRY_GRID_CODE 0   // 0: build variant without the entry-grid look-up in the kernels (A/B of its cost when unused)
#endif
#ifndef PT_BVH4
#define PT_BVH4 0             // 1: build variant that walks 4-wide nodes collapsed from the binary tree
#endif
#ifndef PT_STACK_TOP
#define PT_STACK_TOP 1
#endif
#ifndef PT_BOUNCE_WARPS
#define PT_BOUNCE_WARPS 4
#endif

// One control block per bounce, zeroed once per wave.  Every counter that takes atomics sits on its own
// 128-byte line (its own L2 slice queue): v1 kept them in one 64-byte line and k_shade spent its time
// waiting on ~300 k same-line atomics per launch (profiles/r01_v1_summary.md).
struct Counter128 { uint32_t v; uint32_t pad[31]; };
struct BounceCtrl {
    Counter128 in_count;            // (parity path) rays in the caller's batch
    Counter128 rays;                // rays traced as this segment (statistics)
    Counter128 trav_head;           // unused since the fused bounce kernel (kept: layout)
    Counter128 shade_head;          // work counter of the launch that CONSUMES this block's hit queue, in 32-item chunks
    Counter128 bin_count[NUM_BINS]; // hits of this segment per shader bin
};
static_assert(sizeof(BounceCtrl) == 128 * (4 + NUM_BINS), "one line per counter");
#define PT_INVALID_SLOT 0xFFFFFFFFu   // ray_o.w of a queue slot whose path ended (hole)

// Primitives live on the device in the reference's leaf VISITING order (or insertion order when
// PT_OPT_TIE_ORDER is off): "device index" below always means that position.
// Entry grid (pt_lbvh.cuh, k_entry_grid): where the walk of a SHORT ray starts.  A ray that is clipped to the scene's box
// and to its platform hit is a segment; every level l of the grid tiles the box with overlapping blocks of 2 x 2 x 2
// cubic cells of size 2^(k-l), and each block stores the frontier (<= 4 links) of the tree pruned by the block's box
// with every single-child descent followed.  A segment whose bounding box has extent < cell size fits in one block,
// so its walk starts at that frontier instead of the root: the root-to-neighbourhood descent (log2 N dependent node
// visits that mostly find one child) is paid once per block at build time instead of once per ray.
struct EntryGrid {
    const int4* cells;        // all levels, level l at levels[l].x
    const uint4* levels;      // (first cell, nx, ny, nz) per level; level 0 is one block = the root
    float3 org, hi;           // the gridded box: contains every leaf box, padded
    int k;                    // cell size of level l is 2^(k - l)
    int lmax;                 // finest level; < 0: no grid (walks start at the root)
};

struct SceneDev {
    EntryGrid grid;
    const float4* nodes;      // LBVH, 4 x float4 per inner node
    const float4* nodes4;     // build variant PT_BVH4: 8 x float4 per inner node (k_lbvh_collapse4)
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
#ifndef PT_BOUNCE_MIN_BLOCKS
#define PT_BOUNCE_MIN_BLOCKS 0   // 0: let ptxas choose (56 registers, 9 CTAs/SM). Same-box A/B: forcing 10 CTAs (48 regs) is 3 % slower on C2
#endif

#define PT_BOUNCE_SMEM ((size_t)PT_STACK_SMEM * 32 * PT_BOUNCE_WARPS * sizeof(int))   // dynamic shared memory per CTA

struct LaneRay {
    float3 o, d;
    float ix, iy, iz;      // 1/d
    float ox, oy, oz;      // o/d, so that a slab plane distance is one FMA: plane*(1/d) - o/d
    float slack;           // bound on the rounding error of those FMAs
    float lo, hi;          // [tmin - slack, best_t (1 + 2^-21) + slack]: the interval a child box must overlap
    float tmin, best_t;
    uint32_t best_code;
    uint32_t slot;
    int node, sp;
#if PT_STACK_TOP
    int top;               // top of the stack, in a register: a pop hands it out and starts loading the next entry,
                           // whose latency then hides behind the node visit instead of stalling the node fetch
#endif
};
#define PT_SENTINEL 0x7fffffff

__device__ __forceinline__ float tie_tmax(uint32_t idx, float best_t, uint32_t best_code) {
    // see trace_closest: one ulp above best_t when primitive idx outranks the current holder
    const uint32_t holder = (best_code == PT_MISS_CODE) ? 0u : (best_code >> 3);
    return (idx < holder) ? __uint_as_float(__float_as_uint(best_t) + 1u) : best_t;
}

// 1/d for the slab tests.  Culling only has to be conservative, so the one-instruction MUFU.RCP (relative error
// <= 2^-23) replaces the ~8-instruction IEEE division: a slab distance is then off by <= 1.5 ulp of |plane/d| +
// 2 ulp of |o/d|, inside the 4-ulp inflation of the boxes and the 4-ulp slack of the ray interval.
__device__ __forceinline__ float cull_rcp(float x) {
#if PT_RCP_APPROX
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

// The walk's stack.  Default: thread-local memory (L1-resident; the top entry lives in a register, PT_STACK_TOP).
// PT_STACK_SMEM = N > 0 (build variant, A/B in profiles/): the first N entries of every thread live in shared memory,
// laid out [entry][thread] so that a warp's accesses are conflict-free whatever the lanes' depths are (thread-local
// memory interleaves by lane too, but lanes at different depths touch different 128-byte lines: one L1 wavefront per
// distinct depth); deeper entries overflow to local memory.
#ifndef PT_STACK_SMEM
#define PT_STACK_SMEM 0
#endif
#if PT_STACK_SMEM > 0
__device__ __forceinline__ void stack_store(int* stack, int sp, int v) {
    extern __shared__ int pt_smem_stack[];
    if (sp < PT_STACK_SMEM) pt_smem_stack[sp * (32 * PT_BOUNCE_WARPS) + threadIdx.x] = v;
    else stack[sp - PT_STACK_SMEM] = v;
}
__device__ __forceinline__ int stack_load(const int* stack, int sp) {
    extern __shared__ int pt_smem_stack[];
    return sp < PT_STACK_SMEM ? pt_smem_stack[sp * (32 * PT_BOUNCE_WARPS) + threadIdx.x] : stack[sp - PT_STACK_SMEM];
}
#else
__device__ __forceinline__ void stack_store(int* stack, int sp, int v) { stack[sp] = v; }
__device__ __forceinline__ int stack_load(const int* stack, int sp) { return stack[sp]; }
#endif
__device__ __forceinline__ void stack_push(LaneRay& r, int* stack, int x) {
#if PT_STACK_TOP
    stack_store(stack, r.sp++, r.top); r.top = x;
#else
    stack_store(stack, r.sp++, x);
#endif
}
__device__ __forceinline__ int stack_pop(LaneRay& r, int* stack) {
#if PT_STACK_TOP
    const int x = r.top; r.top = stack_load(stack, --r.sp); return x;    // sp >= 1 here: entry 0 is the sentinel written by lane_setup
#else
    return stack_load(stack, --r.sp);
#endif
}

// platforms are kept out of the tree: tested first, they also give the walk its first tmax
__device__ __forceinline__ void lane_platforms(const SceneDev& sc, float3 o, float3 d, float tmin, float& best_t, uint32_t& best_code) {
    best_code = PT_MISS_CODE;
    for (int k = 0; k < sc.n_plat; ++k) {
        const uint32_t prim = (uint32_t)sc.plat_id[k];
        float tm = tie_tmax(prim, best_t, best_code);
        if (hit_platform(sc.plat_y[k], o, d, tmin, tm)) { best_t = tm; best_code = (prim << 3) | PART_PLATFORM; }
    }
}
// walk state of a ray whose origin, direction, tmin and (best_t, best_code) are set
__device__ __forceinline__ void lane_setup(const SceneDev& sc, LaneRay& r, int* stack) {
    // |1/d| is capped at 1e18 so that c*(1/d) and h*|1/d| stay finite (no inf - inf) for coordinates up to 1e20
    r.ix = cull_rcp(fabsf(r.d.x) > 1e-18f ? r.d.x : copysignf(1e-18f, r.d.x));
    r.iy = cull_rcp(fabsf(r.d.y) > 1e-18f ? r.d.y : copysignf(1e-18f, r.d.y));
    r.iz = cull_rcp(fabsf(r.d.z) > 1e-18f ? r.d.z : copysignf(1e-18f, r.d.z));
    r.ox = r.o.x * r.ix; r.oy = r.o.y * r.iy; r.oz = r.o.z * r.iz;
    // |fma(p, 1/d, -o/d) - (p-o)/d| <= ~3 ulp of |o/d| (rounding of 1/d, of o/d and of the fma); culling only has
    // to be conservative, so both interval ends get that much slack (plus the 4-ulp inflation of the leaf boxes)
    r.slack = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(r.ox), fabsf(r.oy)), fabsf(r.oz));
    r.lo = r.tmin - r.slack;
    r.hi = r.best_t * 1.0000005f + r.slack;
    r.sp = 0;
    stack_store(stack, r.sp++, PT_SENTINEL);
#if PT_STACK_TOP
    r.top = PT_SENTINEL;
#endif
    r.node = sc.n_tree_prims > 0 ? 0 : PT_SENTINEL;
#if PT_ENTRY_GRID_CODE
    const EntryGrid& g = sc.grid;
    if (g.lmax < 0 || sc.n_tree_prims <= 0) return;
    // ---- the ray as a segment: clipped to the gridded box and to [lo, hi] (tmin .. platform hit), same FMA slabs and
    //      the same slack as the node tests ----
    const float x0 = fmaf(g.org.x, r.ix, -r.ox), x1 = fmaf(g.hi.x, r.ix, -r.ox);
    const float y0 = fmaf(g.org.y, r.iy, -r.oy), y1 = fmaf(g.hi.y, r.iy, -r.oy);
    const float z0 = fmaf(g.org.z, r.iz, -r.oz), z1 = fmaf(g.hi.z, r.iz, -r.oz);
    const float t0 = fmaxf(fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1)) - r.slack, r.lo);
    const float t1 = fminf(fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) + r.slack, r.hi);
    if (!(t0 <= t1)) { r.node = PT_SENTINEL; return; }        // misses the box of everything in the tree
    const float ax = fmaf(t0, r.d.x, r.o.x), ay = fmaf(t0, r.d.y, r.o.y), az = fmaf(t0, r.d.z, r.o.z);
    const float bx = fmaf(t1, r.d.x, r.o.x), by = fmaf(t1, r.d.y, r.o.y), bz = fmaf(t1, r.d.z, r.o.z);
    const float ex = 2e-7f * (fabsf(ax) + fabsf(bx)) + 1e-30f, ey = 2e-7f * (fabsf(ay) + fabsf(by)) + 1e-30f,
                ez = 2e-7f * (fabsf(az) + fabsf(bz)) + 1e-30f;
    // bounding box of the segment, inside the gridded box (nothing to hit outside of it)
    const float mnx = fmaxf(fminf(ax, bx) - ex, g.org.x) - g.org.x, mny = fmaxf(fminf(ay, by) - ey, g.org.y) - g.org.y,
                mnz = fmaxf(fminf(az, bz) - ez, g.org.z) - g.org.z;
    const float mxx = fminf(fmaxf(ax, bx) + ex, g.hi.x) - g.org.x, mxy = fminf(fmaxf(ay, by) + ey, g.hi.y) - g.org.y,
                mxz = fminf(fmaxf(az, bz) + ez, g.hi.z) - g.org.z;
    const float ext = fmaxf(fmaxf(mxx - mnx, mxy - mny), fmaxf(mxz - mnz, 0.f));
    // finest level whose cell size 2^(k-l) exceeds the extent: ext = m 2^x, 1 <= m < 2  =>  cell 2^(x+1), l = k - x - 1
    int l = g.k + 126 - (__float_as_int(ext) >> 23);
    l = l > g.lmax ? g.lmax : l;
    if (l <= 0) return;                                        // long segment (or NaN): the root
    const float inv_s = __int_as_float((l - g.k + 127) << 23);
    const uint4 lv = __ldg(&g.levels[l]);
    const float fx = mnx * inv_s, fy = mny * inv_s, fz = mnz * inv_s;
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    ix = min(ix, (int)lv.y - 1); iy = min(iy, (int)lv.z - 1); iz = min(iz, (int)lv.w - 1);
    // block (ix, iy, iz) spans cells i .. i+2 on every axis: it must contain the whole box (rounding, clamping)
    const bool fits = fx >= (float)ix && fy >= (float)iy && fz >= (float)iz && mxx * inv_s <= (float)(ix + 2) &&
                      mxy * inv_s <= (float)(iy + 2) && mxz * inv_s <= (float)(iz + 2);
    if (!fits) return;
    const int4 f = __ldg(&g.cells[lv.x + ((uint32_t)ix * lv.z + (uint32_t)iy) * lv.w + (uint32_t)iz]);
    r.node = f.x;                                              // PT_SENTINEL: nothing overlaps the block
    if (f.y != PT_SENTINEL) {
        stack_push(r, stack, f.y);
        if (f.z != PT_SENTINEL) {
            stack_push(r, stack, f.z);
            if (f.w != PT_SENTINEL) stack_push(r, stack, f.w);
        }
    }
#endif
}
__device__ __forceinline__ void lane_begin(const SceneDev& sc, LaneRay& r, int* stack) {
    lane_platforms(sc, r.o, r.d, r.tmin, r.best_t, r.best_code);
    lane_setup(sc, r, stack);
}
#if PT_ENTRY_GRID_CODE
// The entry-grid level of a ray without starting its walk (the split pipeline classifies rays by it): the level lane_setup
// would pick -- 0 for a long segment that starts at the root --, or -1 when the ray misses the box of everything in the tree.
__device__ __forceinline__ int ray_level(const SceneDev& sc, float3 o, float3 d, float tmin, float best_t) {
    const EntryGrid& g = sc.grid;
    if (sc.n_tree_prims <= 0) return -1;
    if (g.lmax < 0) return 0;
    const float ix = cull_rcp(fabsf(d.x) > 1e-18f ? d.x : copysignf(1e-18f, d.x));
    const float iy = cull_rcp(fabsf(d.y) > 1e-18f ? d.y : copysignf(1e-18f, d.y));
    const float iz = cull_rcp(fabsf(d.z) > 1e-18f ? d.z : copysignf(1e-18f, d.z));
    const float ox = o.x * ix, oy = o.y * iy, oz = o.z * iz;
    const float slack = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    const float x0 = fmaf(g.org.x, ix, -ox), x1 = fmaf(g.hi.x, ix, -ox);
    const float y0 = fmaf(g.org.y, iy, -oy), y1 = fmaf(g.hi.y, iy, -oy);
    const float z0 = fmaf(g.org.z, iz, -oz), z1 = fmaf(g.hi.z, iz, -oz);
    const float t0 = fmaxf(fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1)) - slack, tmin - slack);
    const float t1 = fminf(fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) + slack, best_t * 1.0000005f + slack);
    if (!(t0 <= t1)) return -1;
    const float ax = fmaf(t0, d.x, o.x), ay = fmaf(t0, d.y, o.y), az = fmaf(t0, d.z, o.z);
    const float bx = fmaf(t1, d.x, o.x), by = fmaf(t1, d.y, o.y), bz = fmaf(t1, d.z, o.z);
    const float ex = fmaxf(fminf(fmaxf(ax, bx), g.hi.x) - fmaxf(fminf(ax, bx), g.org.x), 0.f);
    const float ey = fmaxf(fminf(fmaxf(ay, by), g.hi.y) - fmaxf(fminf(ay, by), g.org.y), 0.f);
    const float ez = fmaxf(fminf(fmaxf(az, bz), g.hi.z) - fmaxf(fminf(az, bz), g.org.z), 0.f);
    const float ext = fmaxf(fmaxf(ex, ey), ez) * 1.000001f + 1e-30f;
    int l = g.k + 126 - (__float_as_int(ext) >> 23);
    l = l > g.lmax ? g.lmax : l;
    return l < 0 ? 0 : l;
}
#endif

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

#if PT_PACKED_SLABS
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 v; asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(a), "f"(b)); return v; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif
__device__ __forceinline__ void lane_inner(const SceneDev& sc, LaneRay& r, int* stack) {
#if PT_PACKED_SLABS
    {
        // node = (cax cay hax hay | cbx cby hbx hby) (caz haz cbz hbz | link a, link b, -, -): the x and y slabs of a child are one
        // FFMA2 each for the centre distance, the near and the far plane (3 instead of 6 instructions), z stays scalar
        const float4* p = sc.nodes + 4 * (size_t)r.node;
        f32x2 ca, ha, cb, hb;
        float caz, haz, cbz, hbz, la, lb, pad0, pad1;
        asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(ca), "=l"(ha), "=l"(cb), "=l"(hb) : "l"(p));
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(caz), "=f"(haz), "=f"(cbz), "=f"(hbz), "=f"(la), "=f"(lb), "=f"(pad0), "=f"(pad1) : "l"(p + 2));
        const f32x2 ixy = pack2(r.ix, r.iy), noxy = pack2(-r.ox, -r.oy), axy = pack2(fabsf(r.ix), fabsf(r.iy)), naxy = pack2(-fabsf(r.ix), -fabsf(r.iy));
        const float az = fabsf(r.iz);
        const f32x2 ta = fma2(ca, ixy, noxy), tb = fma2(cb, ixy, noxy);
        float nax, nay, fax, fay, nbx, nby, fbx, fby;
        unpack2(fma2(ha, naxy, ta), nax, nay); unpack2(fma2(ha, axy, ta), fax, fay);
        unpack2(fma2(hb, naxy, tb), nbx, nby); unpack2(fma2(hb, axy, tb), fbx, fby);
        const float lo = r.lo, hi = r.hi;
#if PT_PACKED_SLABS == 2
        // (caz, haz, cbz, hbz) hold (caz, cbz, haz, hbz): the z slabs of both children as one pair
        const f32x2 czz = pack2(caz, haz), hzz = pack2(cbz, hbz), izz = pack2(r.iz, r.iz), nozz = pack2(-r.oz, -r.oz);
        const f32x2 tzz = fma2(czz, izz, nozz);
        float naz, nbz, faz, fbz;
        unpack2(fma2(hzz, pack2(-az, -az), tzz), naz, nbz); unpack2(fma2(hzz, pack2(az, az), tzz), faz, fbz);
        const float ea = fmaxf(fmaxf(nax, nay), fmaxf(naz, lo));
        const float xa = fminf(fminf(fax, fay), fminf(faz, hi));
        const float eb = fmaxf(fmaxf(nbx, nby), fmaxf(nbz, lo));
        const float xb = fminf(fminf(fbx, fby), fminf(fbz, hi));
#else
        const float taz = fmaf(caz, r.iz, -r.oz), tbz = fmaf(cbz, r.iz, -r.oz);
        const float ea = fmaxf(fmaxf(nax, nay), fmaxf(fmaf(-haz, az, taz), lo));
        const float xa = fminf(fminf(fax, fay), fminf(fmaf(haz, az, taz), hi));
        const float eb = fmaxf(fmaxf(nbx, nby), fmaxf(fmaf(-hbz, az, tbz), lo));
        const float xb = fminf(fminf(fbx, fby), fminf(fmaf(hbz, az, tbz), hi));
#endif
        const bool ha_ = ea <= xa, hb_ = eb <= xb;
        const int ca_ = __float_as_int(la), cb_ = __float_as_int(lb);
        if (ha_ && hb_) {
            const bool a_near = ea <= eb;
            r.node = a_near ? ca_ : cb_;
            const int far = a_near ? cb_ : ca_;
            stack_push(r, stack, far);
#if PT_PREFETCH_FAR
            if (far >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.nodes + 4 * (size_t)far));
#endif
        } else if (ha_) r.node = ca_;
        else if (hb_) r.node = cb_;
        else r.node = stack_pop(r, stack);
    }
#else
    {
        float4 n0, n1, n2, n3;
        load_node(sc.nodes, r.node, n0, n1, n2, n3);
        // child a: centre (n0.x n0.y n0.z), half (n0.w n1.x n1.y); child b: centre (n1.z n1.w n2.x), half (n2.y n2.z n2.w)
        const float ax = fabsf(r.ix), ay = fabsf(r.iy), az = fabsf(r.iz);
        const float tax = fmaf(n0.x, r.ix, -r.ox), tay = fmaf(n0.y, r.iy, -r.oy), taz = fmaf(n0.z, r.iz, -r.oz);
        const float tbx = fmaf(n1.z, r.ix, -r.ox), tby = fmaf(n1.w, r.iy, -r.oy), tbz = fmaf(n2.x, r.iz, -r.oz);
        const float lo = r.lo, hi = r.hi;
        const float ea = fmaxf(fmaxf(fmaf(-n0.w, ax, tax), fmaf(-n1.x, ay, tay)), fmaxf(fmaf(-n1.y, az, taz), lo));
        const float xa = fminf(fminf(fmaf(n0.w, ax, tax), fmaf(n1.x, ay, tay)), fminf(fmaf(n1.y, az, taz), hi));
        const float eb = fmaxf(fmaxf(fmaf(-n2.y, ax, tbx), fmaf(-n2.z, ay, tby)), fmaxf(fmaf(-n2.w, az, tbz), lo));
        const float xb = fminf(fminf(fmaf(n2.y, ax, tbx), fmaf(n2.z, ay, tby)), fminf(fmaf(n2.w, az, tbz), hi));
        const bool ha = ea <= xa, hb = eb <= xb;
        const int ca = __float_as_int(n3.x), cb = __float_as_int(n3.y);
        if (ha && hb) {
            const bool a_near = ea <= eb;
            r.node = a_near ? ca : cb;
            const int far = a_near ? cb : ca;
            stack_push(r, stack, far);
#if PT_PREFETCH_FAR
            // the far child is visited later (if at all): start pulling its node towards L1 now
            if (far >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.nodes + 4 * (size_t)far));
#endif
        } else if (ha) r.node = ca;
        else if (hb) r.node = cb;
        else r.node = stack_pop(r, stack);
    }
#endif
}
#if PT_BVH4
// 4-wide visit (build variant, profiles/r02_c3_variants.md): four slab tests on one 128-byte node, the hits sorted by entry
// distance together with their links (5 compare-exchanges), nearest next, the others pushed far-to-near.
__device__ __forceinline__ void cmpx(float& ka, int& la, float& kb, int& lb) {
    const bool sw = kb < ka;
    const float k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const int l0 = sw ? lb : la, l1 = sw ? la : lb;
    ka = k0; kb = k1; la = l0; lb = l1;
}
__device__ __forceinline__ void lane_inner4(const SceneDev& sc, LaneRay& r, int* stack) {
    const float4* p = sc.nodes4 + 8 * (size_t)r.node;
    float4 cx, cy, cz, hx, hy, hz, lk, pad;
    load_node(p, 0, cx, cy, cz, hx);
    load_node(p + 4, 0, hy, hz, lk, pad);
    const float ax = fabsf(r.ix), ay = fabsf(r.iy), az = fabsf(r.iz);
    const float lo = r.lo, hi = r.hi;
    const float BIG = 3.0e38f;
    float e[4]; int l[4];
    const float c_x[4] = {cx.x, cx.y, cx.z, cx.w}, c_y[4] = {cy.x, cy.y, cy.z, cy.w}, c_z[4] = {cz.x, cz.y, cz.z, cz.w};
    const float h_x[4] = {hx.x, hx.y, hx.z, hx.w}, h_y[4] = {hy.x, hy.y, hy.z, hy.w}, h_z[4] = {hz.x, hz.y, hz.z, hz.w};
    const float lks[4] = {lk.x, lk.y, lk.z, lk.w};
    int n = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float tx = fmaf(c_x[k], r.ix, -r.ox), ty = fmaf(c_y[k], r.iy, -r.oy), tz = fmaf(c_z[k], r.iz, -r.oz);
        const float en = fmaxf(fmaxf(fmaf(-h_x[k], ax, tx), fmaf(-h_y[k], ay, ty)), fmaxf(fmaf(-h_z[k], az, tz), lo));
        const float ex = fminf(fminf(fmaf(h_x[k], ax, tx), fmaf(h_y[k], ay, ty)), fminf(fmaf(h_z[k], az, tz), hi));
        const bool hit = en <= ex;
        e[k] = hit ? en : BIG;
        l[k] = __float_as_int(lks[k]);
        n += hit ? 1 : 0;
    }
    cmpx(e[0], l[0], e[1], l[1]); cmpx(e[2], l[2], e[3], l[3]);
    cmpx(e[0], l[0], e[2], l[2]); cmpx(e[1], l[1], e[3], l[3]);
    cmpx(e[1], l[1], e[2], l[2]);
    if (n == 0) { r.node = stack_pop(r, stack); return; }
    if (n > 3) stack_push(r, stack, l[3]);
    if (n > 2) stack_push(r, stack, l[2]);
    if (n > 1) stack_push(r, stack, l[1]);
    r.node = l[0];
}
#endif

__device__ __forceinline__ void lane_leaf_test(const SceneDev& sc, LaneRay& r, int code) {
    const uint32_t leaf = (uint32_t)~code;
    const uint32_t prim = leaf >> 2;
    const float4 g0 = __ldg(&sc.geo0[prim]);
    int part = 0;
    float tm = tie_tmax(prim, r.best_t, r.best_code);
    bool h;
    if ((leaf & 3u) == 2u) h = hit_cylinder(g0, __ldg(&sc.geo1[prim]).x, r.o, r.d, r.tmin, tm, part);
    else h = hit_sphere(g0, r.o, r.d, r.tmin, tm, part);
    if (h) { r.best_t = tm; r.best_code = (prim << 3) | (uint32_t)part; r.hi = tm * 1.0000005f + r.slack; }
}
__device__ __forceinline__ void lane_leaf(const SceneDev& sc, LaneRay& r, int* stack) {
    lane_leaf_test(sc, r, r.node);
    r.node = stack_pop(r, stack);
}

// ---- the bounce kernel ---------------------------------------------------------------------------------------
// One launch per path segment k.  A warp takes a 32-item chunk of its source, turns it into rays IN REGISTERS,
// walks them to completion and files the results, binned by shader, into the hit queue of segment k:
//   SRC_CAMERA : path ids        -> MotionalCamera::RayGen                      -> trace segment 0
//   SRC_HITS   : hit queue k-1   -> shade (one shader per chunk)                -> trace segment k (survivors)
//   SRC_RAYS   : caller's rays   (parity entry point pt_trace_rays)             -> trace, results by ray slot
// Rays never touch HBM: the only stage-to-stage traffic is the hit queue, one 48-byte record per segment that hits,
// SoA float4, written and read in order:
//   hq.p = (hit position, bits pixel)  hq.d = (dir, bits sample)  hq.thr = (throughput, bits prim<<3|part)
// (round 1 stored origin and t, 56 bytes; the position origin + t*dir is the same float whichever launch adds it up)
// The launch after the last segment runs with trace = 0: it only shades what is left.
// (v1-v4 kept rays in double-buffered HBM queues between a traverse and a shade kernel: 208 B per segment and
// twice the launches; profiles/r01_v5_summary.md.)
// Layout: the four shader bins share TWO regions of `capacity` records each, one bin growing up from either end of a
// region (bins 1, 2 in the first, bins 3, 4 in the second).  The two bins of a region can never collide -- together
// they hold at most one record per path of the wave -- and an address is base + direction * item: no tables, no
// allocation.  Round 1 gave every bin a region of its own (all rays may land in one bin): 4 x 2 buffers x 56 B x 128 Mi
// = 59 GB for the default wave; now 2 x 2 x 48 B x 128 Mi = 26 GB.  (A paged pool shared by all four bins, 16 GB, was
// built and measured: correct, but 6 % slower -- the loop that waits for a page to be published costs the kernel the
// 9th CTA per SM; profiles/ab/r2_paged_pool.log)
struct HitQueues {
    float4* p; float4* d; float4* thr;
    uint32_t capacity;     // records per region
};
__device__ __forceinline__ size_t hit_slot(uint32_t capacity, int bin, uint32_t item) {
    // bin 1: 0 + i   bin 2: cap-1 - i   bin 3: cap + i   bin 4: 2cap-1 - i
    const size_t region = bin > 2 ? (size_t)capacity : (size_t)0;
    return region + ((bin & 1) ? (size_t)item : (size_t)(capacity - 1u - item));
}
enum { SRC_CAMERA = 0, SRC_HITS = 1, SRC_RAYS = 2 };

struct BounceArgs {
    SceneDev sc;
    uint2 key;
    uint32_t npix, first_sample;
    int segment;              // index k of the segment traced by this launch (shaded hits belong to k-1)
    int max_depth;
    int trace;                // 0: shade only (the launch after the last segment)
    int drop_last_hits;       // this launch traces the last allowed segment of a scene without emitters (and no G-buffer is due)
    int cam_tiled;            // SRC_CAMERA: width % 8 == 0 and height % 4 == 0, chunks are 8x4 pixel tiles
    int inner_run;            // inner nodes a lane may walk before the warp regroups for primitive tests
    CameraFrame cam;          // SRC_CAMERA
    uint32_t n_paths;
    HitQueues hq_in, hq_out;  // SRC_HITS reads hq_in (ctrl_in), every traced segment writes hq_out (ctrl_out)
    BounceCtrl* ctrl_in;
    BounceCtrl* ctrl_out;
    float4* path_out;
    float4* gbuf_normal;      // display pipeline only: first-segment normal per pixel
    const float4* ray_o;      // SRC_RAYS
    const float4* ray_d;
    const float2* ray_range;
    uint2* slot_hit;
    uint32_t n_rays;
    // split pipeline (PT_OPT_SPLIT_WALK, entry-grid library only): k_shade_classify files the rays of this segment, by
    // entry-grid level class, into rq (control block rctrl); k_walk reads them from there
    HitQueues rq;             // same layout as the hit queues: p = (origin, pixel), d = (dir, sample), thr = (throughput, -)
    BounceCtrl* rctrl;
    unsigned long long* step_counters;   // COUNT_STEPS: [0] inner nodes visited, [1] primitive tests
    uint32_t* ray_steps;                 // COUNT_STEPS + SRC_RAYS: visits per ray (diagnostics)
};

// Miss, path_tracer.cu:117-122: the path ends with the sky seen through its throughput
__device__ __forceinline__ void shade_miss(const BounceArgs& a, float3 d, float3 thr, size_t pid, bool first_segment) {
    const float3 c = thr * sky_radiance(a.sc.sky, d);
    if (a.sc.has_emitters) { float4 p = a.path_out[pid]; a.path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f); }
    else a.path_out[pid] = make_float4(c.x, c.y, c.z, 0.f);
    // G-buffer of the display pipeline: a first-segment miss records -dir (path_tracer.cu:152)
    if (a.gbuf_normal != nullptr && first_segment) a.gbuf_normal[pid] = make_float4(-d.x, -d.y, -d.z, 0.f);
}

// Shade item `item` of shader bin `bin` of the input hit queue.  Returns true and the next ray when the
// path goes on; paths that end write their radiance.
__device__ __forceinline__ bool shade_hit_item(const BounceArgs& a, int bin, uint32_t item, bool last_shade, float3& ray_o,
                                               float3& ray_d, float3& thr, uint32_t& pixel, uint32_t& s) {
    const SceneDev& sc = a.sc;
    const size_t q = hit_slot(a.hq_in.capacity, bin, item);            // bin >= 1: bin 0 (miss) has no queue
    // streamed once: evict-first loads / stores keep the L1 and L2 lines for the BVH nodes
    const float4 p4 = __ldcs(&a.hq_in.p[q]);
    const float4 d4 = __ldcs(&a.hq_in.d[q]);
    const float4 t4 = __ldcs(&a.hq_in.thr[q]);
    pixel = __float_as_uint(p4.w);
    s = __float_as_uint(d4.w);
    const uint32_t code = __float_as_uint(t4.w);
    const size_t pid = (size_t)s * a.npix + pixel;
    const float3 d = f3(d4.x, d4.y, d4.z);
    thr = f3(t4.x, t4.y, t4.z);
    // Bin 0 has no queue, so this branch is never taken; it stays because ptxas allocates 56 registers with it and 58
    // without (8 instead of 9 CTAs per SM: -2.2 % on C2, same-box A/B profiles/ab/variants_r2f.log).  The Makefile
    // checks the register count of this kernel after every build.
    if (bin == BIN_MISS) { shade_miss(a, d, thr, pid, a.segment == 1); return false; }
    const uint32_t prim = code >> 3;
    const float4 g0 = __ldg(&sc.geo0[prim]);
    const uint32_t mslot = __float_as_uint(__ldg(&sc.geo1[prim]).w) & 0x0fffffffu;
    const MaterialRec m = sc.mats[mslot];
    const float3 pos = f3(p4.x, p4.y, p4.z);
    float3 normal;
    hit_normal(g0, (int)(code & 7u), pos, d, normal);
    if (a.gbuf_normal != nullptr && a.segment == 1) a.gbuf_normal[pid] = make_float4(normal.x, normal.y, normal.z, 0.f);
    const float4 U = philox_uniform4(a.key, pixel, a.first_sample + s, (uint32_t)a.segment);   // slot 1 + (k-1)
    Scatter sct;
    shade_bin(bin, m, normal, d, U, sct);
    if (sc.has_emitters) {
        const float3 c = thr * sct.radiance;
        float4 p = a.path_out[pid];
        a.path_out[pid] = make_float4(p.x + c.x, p.y + c.y, p.z + c.z, 0.f);
    }
    thr = thr * sct.atten;
    // a path whose throughput is exactly zero can add nothing more (finite sky and emission): ending it here is
    // result-identical to the reference, which keeps bouncing it (path_tracer.cu:141-170)
    const bool alive = !last_shade && (thr.x != 0.f || thr.y != 0.f || thr.z != 0.f);
    if (!alive && !sc.has_emitters) a.path_out[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (alive) { ray_o = pos; ray_d = normalize(sct.bounce); }
    return alive;
}

// Files the result of a traced segment: a ray that left the scene is shaded by the sky in place (its path ends), a
// hit goes to the hit queue of its shader bin.  Called by all 32 lanes.
__device__ __forceinline__ void file_hit(const BounceArgs& a, uint32_t lane, bool alive, float3 o, float3 d, float3 thr,
                                         uint32_t pixel, uint32_t s, float best_t, uint32_t best_code) {
    const SceneDev& sc = a.sc;
    int obin = -1;
    if (alive)
        obin = (best_code == PT_MISS_CODE) ? BIN_MISS : (int)(__float_as_uint(__ldg(&sc.geo1[best_code >> 3]).w) >> 28);
    // rays that leave the scene end here: the sky is looked up by the lanes that missed instead of sending a
    // 48-byte record through a miss bin to the next launch (v7 did: +9 % on C2 without it)
    if (obin == BIN_MISS) { shade_miss(a, d, thr, (size_t)s * a.npix + pixel, a.segment == 0); obin = -1; }
    // a hit of the last allowed segment can only add emission: in a scene without emitters its path ends black, and
    // the shade-only launch that would find that out is not enqueued (render_async)
    if (a.drop_last_hits && obin >= 0) { a.path_out[(size_t)s * a.npix + pixel] = make_float4(0.f, 0.f, 0.f, 0.f); obin = -1; }
    // One atomic INSTRUCTION per chunk: lane b reserves the slots of bin b (up to NUM_BINS lanes issue together, one
    // round trip to L2 instead of one per non-empty bin), then every lane writes its record once.  (v6 looped over
    // the bins: the warp waited for up to five dependent atomics per chunk, 6 % of the stall samples of a bounce
    // launch, and ran each bin's stores with that bin's lanes only.)
    uint32_t my_mask = 0u, lane_mask = 0u;
#pragma unroll
    for (int b = 0; b < NUM_BINS; ++b) {
        const uint32_t m = __ballot_sync(PT_FULL, obin == b);
        if (obin == b) my_mask = m;
        if ((int)lane == b) lane_mask = m;
    }
    uint32_t base = 0u;
    if (lane_mask) base = atomicAdd(&a.ctrl_out->bin_count[lane & 7u].v, (uint32_t)__popc(lane_mask));
    base = __shfl_sync(PT_FULL, base, obin & 7);
    if (obin >= 0) {
        const size_t j = hit_slot(a.hq_out.capacity, obin, base + (uint32_t)__popc(my_mask & ((1u << lane) - 1u)));
        const float3 pos = o + best_t * d;                           // attr.hit_pos of the winning test (object.cu:21,29,42,...)
        __stcs(&a.hq_out.p[j], make_float4(pos.x, pos.y, pos.z, __uint_as_float(pixel)));
        __stcs(&a.hq_out.d[j], make_float4(d.x, d.y, d.z, __uint_as_float(s)));
        __stcs(&a.hq_out.thr[j], make_float4(thr.x, thr.y, thr.z, __uint_as_float(best_code)));
    }
}

#if PT_BOUNCE_MIN_BLOCKS > 0
#define PT_BOUNCE_BOUNDS __launch_bounds__(32 * PT_BOUNCE_WARPS, PT_BOUNCE_MIN_BLOCKS)
#else
#define PT_BOUNCE_BOUNDS __launch_bounds__(32 * PT_BOUNCE_WARPS)
#endif
template <int SRC, bool COUNT_STEPS>
__global__ void PT_BOUNCE_BOUNDS k_bounce(const BounceArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const SceneDev& sc = a.sc;
    // ---- work list: 32-item chunks; SRC_HITS chunks never straddle two shader bins ----
    uint32_t chunk_end[NUM_BINS], bin_n[NUM_BINS];
    uint32_t total = 0;
    if (SRC == SRC_HITS) {
#pragma unroll
        for (int b = 0; b < NUM_BINS; ++b) { bin_n[b] = a.ctrl_in->bin_count[b].v; total += (bin_n[b] + 31u) >> 5; chunk_end[b] = total; }
    } else {
        const uint32_t n = SRC == SRC_CAMERA ? a.n_paths : a.n_rays;
        total = (n + 31u) >> 5;
#pragma unroll
        for (int b = 0; b < NUM_BINS; ++b) { bin_n[b] = n; chunk_end[b] = total; }
    }
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t grab = total / (n_warps * 8u);          // chunks per atomic: keeps the head counter cold, balances the tail
    grab = grab < 1u ? 1u : (grab > 4u ? 4u : grab);
    const bool last_shade = a.segment >= a.max_depth;   // the hits being shaded belong to the last allowed segment
    const float tmin = a.segment == 0 ? 0.f : PT_BOUNCE_TMIN;
    uint32_t traced = 0, n_inner = 0, n_leaf = 0;
    int stack[PT_STACK];
    while (true) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&a.ctrl_in->shade_head.v, grab);
        first = __shfl_sync(PT_FULL, first, 0);
        if (first >= total) break;
        const uint32_t end = first + grab < total ? first + grab : total;
        for (uint32_t chunk = first; chunk < end; ++chunk) {
            // ================= 1. a ray per lane, in registers =================
            bool alive = false;
            float3 thr = f3(1.f, 1.f, 1.f);
            uint32_t pixel = 0, s = 0, slot = 0;
            LaneRay r;
            r.tmin = tmin; r.best_t = PT_TMAX; r.best_code = PT_MISS_CODE; r.node = PT_SENTINEL; r.sp = 0;
            r.o = r.d = f3(0.f, 0.f, 0.f); r.ix = r.iy = r.iz = r.ox = r.oy = r.oz = r.slack = r.lo = r.hi = 0.f;
            if (SRC == SRC_CAMERA) {
                const uint32_t pid = (chunk << 5) + lane;
                if (pid < a.n_paths) {
                    s = pid / a.npix;
                    pixel = pid - s * a.npix;
#if PT_CAMERA_TILES
                    // a chunk covers an 8x4 pixel tile instead of 32 pixels of one row: the 32 primary rays of a warp
                    // then share more of their walk (fewer distinct nodes per fetch, closer walk lengths)
                    if (a.cam_tiled) {
                        const uint32_t tile = pixel >> 5, in_tile = pixel & 31u, tiles_x = (uint32_t)a.cam.width >> 3;
                        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
                        pixel = (ty * 4u + (in_tile >> 3)) * (uint32_t)a.cam.width + tx * 8u + (in_tile & 7u);
                    }
#endif
                    const int y = pixel / (uint32_t)a.cam.width;
                    const int x = pixel - y * (uint32_t)a.cam.width;
                    const float4 U = philox_uniform4(a.key, pixel, a.first_sample + s, 0u);
                    camera_ray(a.cam, x, y, U, r.o, r.d);
                    if (sc.has_emitters) a.path_out[(size_t)s * a.npix + pixel] = make_float4(0.f, 0.f, 0.f, 0.f);
                    alive = true;
                }
            } else if (SRC == SRC_RAYS) {
                slot = (chunk << 5) + lane;
                if (slot < a.n_rays) {
                    const float4 o4 = a.ray_o[slot], d4 = a.ray_d[slot];
                    const float2 rr = a.ray_range[slot];
                    r.o = f3(o4.x, o4.y, o4.z); r.d = f3(d4.x, d4.y, d4.z); r.tmin = rr.x; r.best_t = rr.y;
                    alive = true;
                }
            } else {
                // which bin does this chunk belong to
                int bin = 0;
                uint32_t first_chunk = 0u, n_bin = bin_n[0];
#pragma unroll
                for (int b = 1; b < NUM_BINS; ++b)
                    if (chunk >= chunk_end[b - 1]) { bin = b; first_chunk = chunk_end[b - 1]; n_bin = bin_n[b]; }
                const uint32_t item = ((chunk - first_chunk) << 5) + lane;
                if (item < n_bin) alive = shade_hit_item(a, bin, item, last_shade, r.o, r.d, thr, pixel, s);
            }
            if (!a.trace) continue;
            // ================= 2. walk: while-while, the warp regroups between node runs and primitive tests =================
            uint32_t lane_steps = 0;    // COUNT_STEPS: inner nodes + primitive tests of this lane's ray
            if (alive) { lane_begin(sc, r, stack); ++traced; }
            while (__any_sync(PT_FULL, alive && r.node != PT_SENTINEL)) {
                if (alive) {
                    int k = 0;
#pragma unroll 1
#if PT_BVH4
                    while (r.node >= 0 && r.node != PT_SENTINEL && k < a.inner_run) { lane_inner4(sc, r, stack); ++k; if (COUNT_STEPS) { ++n_inner; ++lane_steps; } }
#else
                    while (r.node >= 0 && r.node != PT_SENTINEL && k < a.inner_run) { lane_inner(sc, r, stack); ++k; if (COUNT_STEPS) { ++n_inner; ++lane_steps; } }
#endif
                }
                if (alive && r.node < 0) { lane_leaf(sc, r, stack); if (COUNT_STEPS) { ++n_leaf; ++lane_steps; } }
            }
            // ================= 3. file the results =================
            if (SRC == SRC_RAYS) {
                if (alive) a.slot_hit[slot] = make_uint2(__float_as_uint(r.best_t), r.best_code);
                if (COUNT_STEPS && alive && a.ray_steps != nullptr) a.ray_steps[slot] = lane_steps;
                continue;
            }
            file_hit(a, lane, alive, r.o, r.d, thr, pixel, s, r.best_t, r.best_code);
        }
    }
    // statistics: rays traced by this launch (one atomic per warp), optional visit counters
    if (a.trace) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) traced += __shfl_xor_sync(PT_FULL, traced, o);
        if (lane == 0 && traced) atomicAdd(&a.ctrl_out->rays.v, traced);
    }
    if (COUNT_STEPS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { n_inner += __shfl_xor_sync(PT_FULL, n_inner, o); n_leaf += __shfl_xor_sync(PT_FULL, n_leaf, o); }
        if (lane == 0) { atomicAdd(&a.step_counters[0], (unsigned long long)n_inner); atomicAdd(&a.step_counters[1], (unsigned long long)n_leaf); }
    }
}

#if PT_ENTRY_GRID_CODE
// ---- the split pipeline (PT_OPT_SPLIT_WALK) ---------------------------------------------------
// A warp walks until its LONGEST ray is done, and with the entry grid the level of a ray predicts the length of its walk.
// So a bounce becomes two launches: k_shade_classify shades the hits of segment k-1 (one shader per chunk, as k_bounce
// does) and files each surviving ray into one of four LEVEL CLASSES (two two-ended regions, like the hit queues; a ray
// that misses the box of the tree and every platform ends at the sky right there); k_walk takes 32-ray chunks of ONE
// class, longest class first, walks them and files the hits by shader.  The price is one 48-byte ray record through HBM
// per ray and a second launch per segment; the camera launch stays fused (its rays are coherent already).
__device__ __forceinline__ int level_class(const SceneDev& sc, int level) {
    // class 1: levels <= lmax - 3 (long rays, and the root), 2: lmax - 2, 3: lmax - 1, 4: lmax (the shortest)
    const int c = level - (sc.grid.lmax - 3);
    return 1 + (c < 0 ? 0 : (c > 3 ? 3 : c));
}
template <bool COUNT_STEPS>
__global__ void __launch_bounds__(32 * PT_BOUNCE_WARPS) k_shade_classify(const BounceArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const SceneDev& sc = a.sc;
    uint32_t chunk_end[NUM_BINS], bin_n[NUM_BINS];
    uint32_t total = 0;
#pragma unroll
    for (int b = 0; b < NUM_BINS; ++b) { bin_n[b] = a.ctrl_in->bin_count[b].v; total += (bin_n[b] + 31u) >> 5; chunk_end[b] = total; }
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t grab = total / (n_warps * 8u);
    grab = grab < 1u ? 1u : (grab > 4u ? 4u : grab);
    const bool last_shade = a.segment >= a.max_depth;
    while (true) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&a.ctrl_in->shade_head.v, grab);
        first = __shfl_sync(PT_FULL, first, 0);
        if (first >= total) break;
        const uint32_t end = first + grab < total ? first + grab : total;
        for (uint32_t chunk = first; chunk < end; ++chunk) {
            int bin = 0;
            uint32_t first_chunk = 0u, n_bin = bin_n[0];
#pragma unroll
            for (int b = 1; b < NUM_BINS; ++b)
                if (chunk >= chunk_end[b - 1]) { bin = b; first_chunk = chunk_end[b - 1]; n_bin = bin_n[b]; }
            const uint32_t item = ((chunk - first_chunk) << 5) + lane;
            bool alive = false;
            float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 0.f), thr = f3(1.f, 1.f, 1.f);
            uint32_t pixel = 0, s = 0;
            if (item < n_bin) alive = shade_hit_item(a, bin, item, last_shade, o, d, thr, pixel, s);
            int cls = -1;
            if (alive) {
                float best_t = PT_TMAX; uint32_t best_code;
                lane_platforms(sc, o, d, PT_BOUNCE_TMIN, best_t, best_code);
                const int level = ray_level(sc, o, d, PT_BOUNCE_TMIN, best_t);
                if (level < 0 && best_code == PT_MISS_CODE) shade_miss(a, d, thr, (size_t)s * a.npix + pixel, false);   // nothing in its way
                else cls = level < 0 ? NUM_BINS - 1 : level_class(sc, level);      // (a platform hit without a walk: shortest class)
            }
            uint32_t my_mask = 0u, lane_mask = 0u;
#pragma unroll
            for (int b = 1; b < NUM_BINS; ++b) {
                const uint32_t m = __ballot_sync(PT_FULL, cls == b);
                if (cls == b) my_mask = m;
                if ((int)lane == b) lane_mask = m;
            }
            uint32_t base = 0u;
            if (lane_mask) base = atomicAdd(&a.rctrl->bin_count[lane & 7u].v, (uint32_t)__popc(lane_mask));
            base = __shfl_sync(PT_FULL, base, cls & 7);
            if (cls >= 1) {
                const size_t j = hit_slot(a.rq.capacity, cls, base + (uint32_t)__popc(my_mask & ((1u << lane) - 1u)));
                __stcs(&a.rq.p[j], make_float4(o.x, o.y, o.z, __uint_as_float(pixel)));
                __stcs(&a.rq.d[j], make_float4(d.x, d.y, d.z, __uint_as_float(s)));
                __stcs(&a.rq.thr[j], make_float4(thr.x, thr.y, thr.z, 0.f));
            }
        }
    }
}
template <bool COUNT_STEPS>
__global__ void __launch_bounds__(32 * PT_BOUNCE_WARPS) k_walk(const BounceArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const SceneDev& sc = a.sc;
    uint32_t chunk_end[NUM_BINS], bin_n[NUM_BINS];
    uint32_t total = 0;
#pragma unroll
    for (int b = 0; b < NUM_BINS; ++b) { bin_n[b] = a.rctrl->bin_count[b].v; total += (bin_n[b] + 31u) >> 5; chunk_end[b] = total; }
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t grab = total / (n_warps * 8u);
    grab = grab < 1u ? 1u : (grab > 4u ? 4u : grab);
    uint32_t traced = 0, n_inner = 0, n_leaf = 0;
    int stack[PT_STACK];
    while (true) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&a.rctrl->shade_head.v, grab);
        first = __shfl_sync(PT_FULL, first, 0);
        if (first >= total) break;
        const uint32_t end = first + grab < total ? first + grab : total;
        for (uint32_t chunk = first; chunk < end; ++chunk) {
            int bin = 0;
            uint32_t first_chunk = 0u, n_bin = bin_n[0];
#pragma unroll
            for (int b = 1; b < NUM_BINS; ++b)
                if (chunk >= chunk_end[b - 1]) { bin = b; first_chunk = chunk_end[b - 1]; n_bin = bin_n[b]; }
            const uint32_t item = ((chunk - first_chunk) << 5) + lane;
            const bool alive = item < n_bin;
            LaneRay r;
            r.tmin = PT_BOUNCE_TMIN; r.best_t = PT_TMAX; r.best_code = PT_MISS_CODE; r.node = PT_SENTINEL; r.sp = 0;
            r.o = r.d = f3(0.f, 0.f, 0.f); r.ix = r.iy = r.iz = r.ox = r.oy = r.oz = r.slack = r.lo = r.hi = 0.f;
            float3 thr = f3(1.f, 1.f, 1.f);
            uint32_t pixel = 0, s = 0;
            if (alive) {
                const size_t q = hit_slot(a.rq.capacity, bin, item);
                const float4 p4 = __ldcs(&a.rq.p[q]), d4 = __ldcs(&a.rq.d[q]), t4 = __ldcs(&a.rq.thr[q]);
                r.o = f3(p4.x, p4.y, p4.z); r.d = f3(d4.x, d4.y, d4.z); thr = f3(t4.x, t4.y, t4.z);
                pixel = __float_as_uint(p4.w); s = __float_as_uint(d4.w);
                lane_begin(sc, r, stack); ++traced;
            }
            while (__any_sync(PT_FULL, alive && r.node != PT_SENTINEL)) {
                if (alive) {
                    int k = 0;
#pragma unroll 1
                    while (r.node >= 0 && r.node != PT_SENTINEL && k < a.inner_run) { lane_inner(sc, r, stack); ++k; if (COUNT_STEPS) ++n_inner; }
                }
                if (alive && r.node < 0) { lane_leaf(sc, r, stack); if (COUNT_STEPS) ++n_leaf; }
            }
            file_hit(a, lane, alive, r.o, r.d, thr, pixel, s, r.best_t, r.best_code);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_xor_sync(PT_FULL, traced, o);
    if (lane == 0 && traced) atomicAdd(&a.ctrl_out->rays.v, traced);
    if (COUNT_STEPS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { n_inner += __shfl_xor_sync(PT_FULL, n_inner, o); n_leaf += __shfl_xor_sync(PT_FULL, n_leaf, o); }
        if (lane == 0) { atomicAdd(&a.step_counters[0], (unsigned long long)n_inner); atomicAdd(&a.step_counters[1], (unsigned long long)n_leaf); }
    }
}
#endif

// ---- stage 3: accumulate the wave into the frame's SUM buffer, sample order, fp32 -------------------------
// (the reference's Mix keeps a running mean of CLAMPED, denoised samples, path_tracer.cu:241-254; the
// unclamped sum is what spp-sharding across GPUs needs -- the display pipeline is pt_dispatch_ray)
__global__ void __launch_bounds__(256, 4) k_accumulate(uint32_t npix, uint32_t n_samples, const float4* __restrict__ path_out,
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
    // segments[0] rays traced, [1] inner-node steps, [2] primitive tests (the latter two only with PT_OPT_COUNT_STEPS),
    // [3] hit records that went through the hit queues (one per segment that hit something; misses end in place)
    unsigned long long n = 0, h = 0;
    for (int b = 0; b < max_depth; ++b) {
        n += ctrl[b].rays.v;   // ctrl[k].rays = rays traced as segment k
        for (int q = 0; q < NUM_BINS; ++q) h += ctrl[b].bin_count[q].v;
    }
    segments[0] += n;
    segments[3] += h;
}

}  // namespace pt
