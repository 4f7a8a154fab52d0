// pt_lbvh.cuh -- GPU LBVH build (Morton codes -> radix sort (own) -> Karras 2012 hierarchy -> bottom-up
// fit) over the compact SoA primitive records.  Replaces the reference's host-side recursive
// median-split build with a std::sort per level (SceneBVH::Divide, cuSrc/bvh.cu:31-90; 0.66 s for
// 100 k objects) and its 112-byte AoS nodes that embed a copy of the Object (include/bvh.h:32-38).
//
// Node layout (64 B, four float4; both child boxes live in the parent so one fetch decides both).  Boxes are kept
// as CENTRE + HALF-EXTENT: a slab pair is then  t_c = c*(1/d) - o/d,  near = t_c - h*|1/d|,  far = t_c + h*|1/d|
// -- three FMAs and no min/max to sort the two plane distances (the walk is bound by the ALU pipe that executes
// FMNMX; this moves 12 of its 28 operations per node to the FMA pipe, profiles/r01_v6_summary.md):
//   n0 = (c0.x, c0.y, c0.z, h0.x)   n1 = (h0.y, h0.z, c1.x, c1.y)   n2 = (c1.z, h1.x, h1.y, h1.z)
//   n3 = (bits child0, bits child1, -, -)       child >= 0: inner node index; child < 0: ~(prim<<2 | type)
// h is rounded up by 4 ulp of (|c| + h) so that [c-h, c+h] contains the box after every rounding of the FMAs.
// Platforms (infinite AABB, object.cu:141-142,160-161) are kept out of the tree in a side list.
#pragma once
#include "pt_math.cuh"

namespace pt {

struct Aabb { float3 mn, mx; };

__device__ __forceinline__ float ulp_out(float v, float dir) {
    // push v outwards by ~4 ulp (+ an absolute floor) so that slab rounding can never cull a primitive
    // that the exact intersector accepts
    float a = fabsf(v);
    return v + dir * (a * 4.8e-7f + 1e-30f);
}

// Object::GetAABBMin/Max, cuSrc/object.cu:134-170, made orientation-proof (|radius|, |height|) and inflated
__device__ __forceinline__ Aabb prim_aabb(float4 g0, float4 g1) {
    const int type = __float_as_int(g1.z);
    const float ar = fabsf(g0.w);
    Aabb b;
    if (type == 2) {
        const float hh = fabsf(g1.x) / 2 + 1e-4f;   // tolerance = BOUNCE_RAY_TMIN * 5
        b.mn = f3(g0.x - ar, g0.y - hh, g0.z - ar);
        b.mx = f3(g0.x + ar, g0.y + hh, g0.z + ar);
    } else {
        b.mn = f3(g0.x - ar, g0.y - ar, g0.z - ar);
        b.mx = f3(g0.x + ar, g0.y + ar, g0.z + ar);
    }
    b.mn = f3(ulp_out(b.mn.x, -1.f), ulp_out(b.mn.y, -1.f), ulp_out(b.mn.z, -1.f));
    b.mx = f3(ulp_out(b.mx.x, 1.f), ulp_out(b.mx.y, 1.f), ulp_out(b.mx.z, 1.f));
    return b;
}

// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// pass 1: leaf boxes + scene bounds of the centroids.  bvh_prims[i] = primitive id of tree leaf slot i.
// test_mn/test_mx (optional, PT_OPT_REF_LEAF_BOXES): the box that decides whether the primitive gets TESTED.  The
// reference tests a leaf's object whenever the leaf's parent box is hit (bvh.cu:175-180), and far from the camera
// its fp32 quadratics report hits on rays that miss the primitive's own box; with these boxes the walk tests the
// same primitives the reference does (a finite platform-free parent box; primitives whose reference parent also
// holds a platform keep their own box -- the platform's box is infinite).
__global__ void k_lbvh_bounds(int n, const int* __restrict__ bvh_prims, const float4* __restrict__ geo0,
                              const float4* __restrict__ geo1, const float4* __restrict__ test_mn,
                              const float4* __restrict__ test_mx, float4* __restrict__ leaf_mn,
                              float4* __restrict__ leaf_mx, float4* __restrict__ leaf_c, uint32_t* __restrict__ bounds6) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 c = f3(0, 0, 0);
    bool valid = i < n;
    if (valid) {
        int p = bvh_prims[i];
        Aabb b = prim_aabb(geo0[p], geo1[p]);
        c = f3(0.5f * (b.mn.x + b.mx.x), 0.5f * (b.mn.y + b.mx.y), 0.5f * (b.mn.z + b.mx.z));
        if (test_mn != nullptr) {
            const float4 tmn = test_mn[p], tmx = test_mx[p];
            const bool finite = fabsf(tmn.x) < 1e29f && fabsf(tmx.x) < 1e29f && fabsf(tmn.z) < 1e29f && fabsf(tmx.z) < 1e29f;
            if (finite) {
                b.mn = f3(fminf(b.mn.x, ulp_out(tmn.x, -1.f)), fminf(b.mn.y, ulp_out(tmn.y, -1.f)), fminf(b.mn.z, ulp_out(tmn.z, -1.f)));
                b.mx = f3(fmaxf(b.mx.x, ulp_out(tmx.x, 1.f)), fmaxf(b.mx.y, ulp_out(tmx.y, 1.f)), fmaxf(b.mx.z, ulp_out(tmx.z, 1.f)));
            }
        }
        leaf_mn[i] = make_float4(b.mn.x, b.mn.y, b.mn.z, 0.f);
        leaf_mx[i] = make_float4(b.mx.x, b.mx.y, b.mx.z, 0.f);
        leaf_c[i] = make_float4(c.x, c.y, c.z, 0.f);     // Morton codes come from the primitive's own centre
    }
    // warp-reduce then one atomic per warp and component
    float mnx = valid ? c.x : 3e38f, mny = valid ? c.y : 3e38f, mnz = valid ? c.z : 3e38f;
    float mxx = valid ? c.x : -3e38f, mxy = valid ? c.y : -3e38f, mxz = valid ? c.z : -3e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bounds6[0], float_to_ordered(mnx)); atomicMin(&bounds6[1], float_to_ordered(mny));
        atomicMin(&bounds6[2], float_to_ordered(mnz)); atomicMax(&bounds6[3], float_to_ordered(mxx));
        atomicMax(&bounds6[4], float_to_ordered(mxy)); atomicMax(&bounds6[5], float_to_ordered(mxz));
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// pass 2: 30-bit Morton code of the box centre on a cubic grid over the centroid bounds
__global__ void k_lbvh_morton(int n, const float4* __restrict__ leaf_c, const uint32_t* __restrict__ bounds6,
                              uint32_t* __restrict__ codes, uint32_t* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 lo = f3(ordered_to_float(bounds6[0]), ordered_to_float(bounds6[1]), ordered_to_float(bounds6[2]));
    float3 hi = f3(ordered_to_float(bounds6[3]), ordered_to_float(bounds6[4]), ordered_to_float(bounds6[5]));
    const float4 c = leaf_c[i];
    // ONE scale for the three axes: the grid cells stay cubes.  Normalising each axis by its own extent makes the
    // code of a flat scene (spheres standing on a floor: y extent = radius spread) split on y as often as on x and
    // z, grouping primitives by radius instead of by position: SAH cost 25.9 vs 11.6 on C2, 71.8 vs 17.2 on a
    // 20 k-primitive field (2.0x / 3.4x fewer node visits per ray, profiles/r01_v4_summary.md).
    const float e = fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), fmaxf(hi.z - lo.z, 1e-20f));
    float x = (c.x - lo.x) / e, y = (c.y - lo.y) / e, z = (c.z - lo.z) / e;
    uint32_t xi = (uint32_t)fminf(fmaxf(x * 1024.f, 0.f), 1023.f);
    uint32_t yi = (uint32_t)fminf(fmaxf(y * 1024.f, 0.f), 1023.f);
    uint32_t zi = (uint32_t)fminf(fmaxf(z * 1024.f, 0.f), 1023.f);
    codes[i] = (expand_bits10(xi) << 2) | (expand_bits10(yi) << 1) | expand_bits10(zi);
    ids[i] = (uint32_t)i;
}

// ---- radix sort of the (Morton code, leaf slot) pairs: 30-bit keys, least-significant digit first, 8 bits per pass, stable ----
// Round 1 called cub::DeviceRadixSort here, the only library kernel on the path.  Two hand-written paths:
//   n <= PT_SORT_SMALL : ONE block sorts keys and values entirely in shared memory, four passes, one launch (every scene of
//                        C1 / C2 / C5 size: the sort is then a single 10-us launch of the per-frame rebuild);
//   larger             : per pass  k_sort_hist (per-tile digit histogram)  ->  k_sort_scan (exclusive scan, digit-major)  ->
//                        k_sort_scatter (stable ranks inside the tile by warp match + warp / round prefixes in shared memory).
// A tile is 256 threads x PT_SORT_ITEMS keys; element e of a tile is handled in round e / 256 by thread e % 256, so "earlier round,
// lower warp, lower lane" is input order and the scatter is stable.
#define PT_SORT_SMALL 2048
#define PT_SORT_ITEMS 8
#define PT_SORT_TILE (256 * PT_SORT_ITEMS)

// rank of this thread's digit among the threads of the block that hold the same digit and come before it, and the block's
// count per digit (in cnt[256], valid after the call); `valid` threads only.  Uses warp_hist[8][256].
__device__ __forceinline__ uint32_t block_digit_rank(uint32_t digit, bool valid, uint32_t (*warp_hist)[256], uint32_t* cnt) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 8u * 256u; i += 256u) (&warp_hist[0][0])[i] = 0u;
    __syncthreads();
    // peers = lanes of this warp with the same digit (invalid lanes get a digit of their own: 256 + lane never matches)
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : 256u + lane);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0u) warp_hist[warp][digit] = __popc(peers);
    __syncthreads();
    // thread d turns column d into exclusive prefixes over the warps and leaves the total in cnt[d]
    uint32_t run = 0u;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = warp_hist[w][threadIdx.x]; warp_hist[w][threadIdx.x] = run; run += c; }
    cnt[threadIdx.x] = run;
    __syncthreads();
    return valid ? warp_hist[warp][digit] + rank_in_warp : 0u;
}

__global__ void __launch_bounds__(256) k_sort_small(int n, const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                                                    uint32_t* __restrict__ vals_out) {
    // values are the identity 0..n-1 on input (k_lbvh_morton), so only the keys are read
    __shared__ uint32_t k[2][PT_SORT_SMALL], v[2][PT_SORT_SMALL];
    __shared__ uint32_t warp_hist[8][256], cnt[256], base[256], seen[256];
    for (int i = threadIdx.x; i < n; i += 256) { k[0][i] = keys_in[i]; v[0][i] = (uint32_t)i; }
    __syncthreads();
    int cur = 0;
    for (int shift = 0; shift < 32; shift += 8) {
        // digit totals of the whole array, then their exclusive scan
        cnt[threadIdx.x] = 0u; seen[threadIdx.x] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 256) atomicAdd(&cnt[(k[cur][i] >> shift) & 255u], 1u);
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t run = 0u; for (int d = 0; d < 256; ++d) { base[d] = run; run += cnt[d]; } }
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += 256) {
            const int i = i0 + (int)threadIdx.x;
            const bool valid = i < n;
            const uint32_t key = valid ? k[cur][i] : 0u, val = valid ? v[cur][i] : 0u, digit = (key >> shift) & 255u;
            const uint32_t r = block_digit_rank(digit, valid, warp_hist, cnt);
            if (valid) { const uint32_t pos = base[digit] + seen[digit] + r; k[cur ^ 1][pos] = key; v[cur ^ 1][pos] = val; }
            __syncthreads();
            seen[threadIdx.x] += cnt[threadIdx.x];
            __syncthreads();
        }
        cur ^= 1;
    }
    for (int i = threadIdx.x; i < n; i += 256) { keys_out[i] = k[cur][i]; vals_out[i] = v[cur][i]; }
}

// hist[digit * tiles + tile] = keys of the tile with that digit
__global__ void __launch_bounds__(256) k_sort_hist(int n, int shift, const uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int tile0 = blockIdx.x * PT_SORT_TILE;
    for (int e = threadIdx.x; e < PT_SORT_TILE; e += 256) {
        const int i = tile0 + e;
        if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}
// exclusive scan of hist (digit-major) in place, one block: each thread scans a contiguous slice, then the slice totals
__global__ void __launch_bounds__(1024) k_sort_scan(int m, uint32_t* __restrict__ hist) {
    __shared__ uint32_t part[1024];
    const int per = (m + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, m);
    uint32_t run = 0u;
    for (int i = lo; i < hi; ++i) run += hist[i];
    part[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t acc = 0u; for (int t = 0; t < 1024; ++t) { const uint32_t c = part[t]; part[t] = acc; acc += c; } }
    __syncthreads();
    run = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
}
__global__ void __launch_bounds__(256) k_sort_scatter(int n, int shift, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                      const uint32_t* __restrict__ hist, uint32_t* __restrict__ keys_out,
                                                      uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t warp_hist[8][256], cnt[256], base[256];
    base[threadIdx.x] = hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x];   // first output slot of this tile's keys with digit d
    const int tile0 = blockIdx.x * PT_SORT_TILE;
    for (int round = 0; round < PT_SORT_ITEMS; ++round) {
        const int i = tile0 + round * 256 + (int)threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keys_in[i] : 0u, val = valid ? vals_in[i] : 0u, digit = (key >> shift) & 255u;
        const uint32_t r = block_digit_rank(digit, valid, warp_hist, cnt);     // (its first barrier also orders the write of base[])
        if (valid) { const uint32_t pos = base[digit] + r; keys_out[pos] = key; vals_out[pos] = val; }
        __syncthreads();
        base[threadIdx.x] += cnt[threadIdx.x];
        __syncthreads();
    }
}

__device__ __forceinline__ int lbvh_delta(const uint32_t* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

// pass 3: Karras 2012 -- one thread per internal node finds its range and split.
// child encoding here: >= 0 internal node, < 0: ~(sorted leaf position)
__global__ void k_lbvh_hierarchy(int n, const uint32_t* __restrict__ codes, int2* __restrict__ children,
                                 int* __restrict__ parent_of_inner, int* __restrict__ parent_of_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (lbvh_delta(codes, n, i, i + 1) - lbvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(codes, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(codes, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left, right;
    if (lo == gamma) { left = ~gamma; parent_of_leaf[gamma] = i; } else { left = gamma; parent_of_inner[gamma] = i; }
    if (hi == gamma + 1) { right = ~(gamma + 1); parent_of_leaf[gamma + 1] = i; } else { right = gamma + 1; parent_of_inner[gamma + 1] = i; }
    children[i] = make_int2(left, right);
    if (i == 0) parent_of_inner[0] = -1;
}

// pass 4: bottom-up boxes of the internal nodes; the second thread to arrive at a node fits it.
//
// SAH refit (`rotate`): before a node's box is stored, the thread that owns the node -- the only thread alive in its
// subtree, every other one ended at its first arrival somewhere below -- tries the four tree rotations of Kensler
// 2008 on it: one child X trades places with a grandchild under the other child Y, which turns Y = (Y1, Y2) into
// (X, Y2) or (Y1, X).  The node's own box does not change (same leaves); the only surface area that does is Y's, so
// the rotation with the smallest new area of Y is taken when it beats the current one.  With one primitive per leaf
// the SAH cost of the tree is the sum of the inner nodes' areas, so every accepted rotation lowers it.  A Karras tree
// over Morton codes splits where the codes say, not where the boxes are small: the rotations recover most of the
// distance to a full-sweep SAH build (C2: SAH cost 11.6 -> see DESIGN.md section 3.3) in one or two extra passes of
// this kernel.  `height` (levels below a node) keeps the walk's stack bound: a rotation that would raise a subtree
// above PT_LBVH_MAX_ROT_HEIGHT levels is refused unless the subtree was already that high.
#ifndef PT_PACKED_SLABS
#define PT_PACKED_SLABS 0     // 1: node layout and slab tests for packed FP32x2 arithmetic (FFMA2, new on sm_100)
#endif
#define PT_LBVH_MAX_ROT_HEIGHT 48
__device__ __forceinline__ float box_area(float4 mn, float4 mx) {
    const float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ void box_union(float4 amn, float4 amx, float4 bmn, float4 bmx, float4& mn, float4& mx) {
    mn = make_float4(fminf(amn.x, bmn.x), fminf(amn.y, bmn.y), fminf(amn.z, bmn.z), 0.f);
    mx = make_float4(fmaxf(amx.x, bmx.x), fmaxf(amx.y, bmx.y), fmaxf(amx.z, bmx.z), 0.f);
}
struct LbvhTree {
    const uint32_t* sorted_ids; const float4* leaf_mn; const float4* leaf_mx;
    int2* children; int* parent_of_inner; int* parent_of_leaf;
    float4* inner_mn; float4* inner_mx; unsigned char* height;
};
__device__ __forceinline__ void lbvh_child_box(const LbvhTree& t, int c, float4& mn, float4& mx, int& h) {
    if (c < 0) { const uint32_t s = t.sorted_ids[~c]; mn = t.leaf_mn[s]; mx = t.leaf_mx[s]; h = 0; }
    else { mn = __ldcg(&t.inner_mn[c]); mx = __ldcg(&t.inner_mx[c]); h = (int)__ldcg(&t.height[c]); }
}
__device__ __forceinline__ void lbvh_set_parent(const LbvhTree& t, int c, int parent) {
    if (c < 0) t.parent_of_leaf[~c] = parent; else t.parent_of_inner[c] = parent;
}

// Treelet restructuring (Karras & Aila 2013, one thread per treelet): the subtree top under `node` is grown into a
// treelet of up to 7 "leaves" (always expanding the leaf with the largest area), the binary tree over those leaves
// with the smallest sum of inner areas is found by dynamic programming over the 2^7 subsets, and the treelet's inner
// nodes are rewired to it when it beats the current arrangement.  Subtrees below the treelet leaves are untouched, so
// their cost is a constant; the treelet root keeps its box.  Returns the new height of `node`.
#ifndef PT_TREELET
#define PT_TREELET 7
#endif
__device__ int lbvh_treelet_optimize(const LbvhTree& t, int node, int2 ch, int old_height) {
    int leaf[PT_TREELET];            // treelet leaves: tree links (inner >= 0, leaf < 0)
    float4 lmn[PT_TREELET], lmx[PT_TREELET];
    int lh[PT_TREELET];
    int inner[PT_TREELET - 1];       // inner nodes of the treelet, root first
    int n_leaf = 2, n_inner = 1;
    inner[0] = node;
    leaf[0] = ch.x; leaf[1] = ch.y;
    lbvh_child_box(t, ch.x, lmn[0], lmx[0], lh[0]);
    lbvh_child_box(t, ch.y, lmn[1], lmx[1], lh[1]);
    float old_cost = 0.f;            // sum of the areas of the treelet's inner nodes except the root
    while (n_leaf < PT_TREELET) {
        int pick = -1; float pa = -1.f;
        for (int i = 0; i < n_leaf; ++i)
            if (leaf[i] >= 0) { const float a = box_area(lmn[i], lmx[i]); if (a > pa) { pa = a; pick = i; } }
        if (pick < 0) break;
        const int x = leaf[pick];
        old_cost += pa;
        inner[n_inner++] = x;
        const int2 c = __ldcg(&t.children[x]);
        leaf[pick] = c.x; lbvh_child_box(t, c.x, lmn[pick], lmx[pick], lh[pick]);
        leaf[n_leaf] = c.y; lbvh_child_box(t, c.y, lmn[n_leaf], lmx[n_leaf], lh[n_leaf]);
        ++n_leaf;
    }
    if (n_leaf < 3) return old_height;
    const int full = (1 << n_leaf) - 1;
    float area[1 << PT_TREELET], cost[1 << PT_TREELET];
    unsigned char part[1 << PT_TREELET];
    for (int s = 1; s <= full; ++s) {
        float4 mn = make_float4(3e38f, 3e38f, 3e38f, 0.f), mx = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
        for (int i = 0; i < n_leaf; ++i)
            if (s >> i & 1) box_union(mn, mx, lmn[i], lmx[i], mn, mx);
        area[s] = box_area(mn, mx);
    }
    // subsets in increasing numeric order: every proper subset of s is smaller than s
    for (int s = 1; s <= full; ++s) {
        if ((s & (s - 1)) == 0) { cost[s] = 0.f; part[s] = 0; continue; }
        const int low = s & -s;
        float best = 3e38f; int bp = 0;
        for (int p = (s - 1) & s; p > 0; p = (p - 1) & s) {
            if (!(p & low)) continue;                       // each split once: the half that holds the lowest leaf
            const float c = cost[p] + cost[s ^ p];
            if (c < best) { best = c; bp = p; }
        }
        cost[s] = best + (s == full ? 0.f : area[s]);
        part[s] = (unsigned char)bp;
    }
    if (!(cost[full] < old_cost * 0.99999f)) return old_height;
    // rebuild: a stack of (subset, node slot); slot 0 is the treelet root
    int stack_s[PT_TREELET], stack_n[PT_TREELET], sp = 0, next_inner = 1;
    // heights need the children first: record the build order and finish bottom-up
    int order_s[PT_TREELET - 1], order_n[PT_TREELET - 1], n_order = 0;
    int2 new_ch[PT_TREELET - 1];
    stack_s[sp] = full; stack_n[sp] = inner[0]; ++sp;
    while (sp > 0) {
        --sp;
        const int s = stack_s[sp], nd = stack_n[sp];
        const int halves[2] = {(int)part[s], s ^ (int)part[s]};
        int links[2];
        for (int k = 0; k < 2; ++k) {
            const int h = halves[k];
            if ((h & (h - 1)) == 0) links[k] = leaf[__ffs(h) - 1];
            else { links[k] = inner[next_inner++]; stack_s[sp] = h; stack_n[sp] = links[k]; ++sp; }
        }
        order_s[n_order] = s; order_n[n_order] = nd; new_ch[n_order] = make_int2(links[0], links[1]); ++n_order;
    }
    // heights of the new inner nodes, children before parents (reverse build order), and the height test
    int new_h[PT_TREELET - 1];
    for (int k = n_order - 1; k >= 0; --k) {
        int hh[2];
        const int links[2] = {new_ch[k].x, new_ch[k].y};
        for (int c = 0; c < 2; ++c) {
            hh[c] = -1;
            for (int i = 0; i < n_leaf; ++i) if (leaf[i] == links[c]) hh[c] = lh[i];
            if (hh[c] < 0) for (int j = k + 1; j < n_order; ++j) if (order_n[j] == links[c]) hh[c] = new_h[j];
        }
        new_h[k] = 1 + max(hh[0], hh[1]);
    }
    if (new_h[0] > max(old_height, PT_LBVH_MAX_ROT_HEIGHT)) return old_height;
    for (int k = n_order - 1; k >= 0; --k) {
        const int nd = order_n[k], s = order_s[k];
        t.children[nd] = new_ch[k];
        lbvh_set_parent(t, new_ch[k].x, nd);
        lbvh_set_parent(t, new_ch[k].y, nd);
        t.height[nd] = (unsigned char)min(new_h[k], 255);
        if (k > 0) {                                        // the root's box is written by the caller
            float4 mn = make_float4(3e38f, 3e38f, 3e38f, 0.f), mx = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
            for (int i = 0; i < n_leaf; ++i)
                if (s >> i & 1) box_union(mn, mx, lmn[i], lmx[i], mn, mx);
            t.inner_mn[nd] = mn; t.inner_mx[nd] = mx;
        }
    }
    return new_h[0];
}

__global__ void k_lbvh_fit(int n, LbvhTree t, int* __restrict__ flags, int rotate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = __ldcg(&t.parent_of_leaf[i]);
    while (node >= 0) {
        if (atomicAdd(&flags[node], 1) == 0) return;   // first arrival: the sibling subtree is not done yet
        __threadfence();
        int2 ch = __ldcg(&t.children[node]);
        float4 amn, amx, bmn, bmx;
        int ha, hb;
        lbvh_child_box(t, ch.x, amn, amx, ha);
        lbvh_child_box(t, ch.y, bmn, bmx, hb);
        int hnode = 1 + max(ha, hb);
        if (rotate == 2) {
            hnode = lbvh_treelet_optimize(t, node, ch, hnode);
        } else if (rotate) {
            // candidates: (which child is Y, which of its children stays) -> new box of Y
            float best = 0.f;          // gain in area, > 0 to accept
            int best_y = -1, best_keep = 0;
            float4 ymn[2], ymx[2];     // boxes of the grandchildren of the best candidate's Y (scratch per Y)
            int yh[2] = {0, 0};
            int2 ych = make_int2(0, 0);
            float4 kmn = amn, kmx = amx; int2 kch = ch; int kh[2] = {0, 0};
            for (int y = 0; y < 2; ++y) {
                const int Y = y == 0 ? ch.x : ch.y;
                if (Y < 0) continue;                                  // a leaf has no children to trade
                const float4 xmn = y == 0 ? bmn : amn, xmx = y == 0 ? bmx : amx;   // X = the other child
                const int hx = y == 0 ? hb : ha;
                const float area_y = y == 0 ? box_area(amn, amx) : box_area(bmn, bmx);
                const int2 gc = __ldcg(&t.children[Y]);
                lbvh_child_box(t, gc.x, ymn[0], ymx[0], yh[0]);
                lbvh_child_box(t, gc.y, ymn[1], ymx[1], yh[1]);
                for (int keep = 0; keep < 2; ++keep) {               // Y keeps grandchild `keep`, the other moves up
                    float4 nmn, nmx;
                    box_union(ymn[keep], ymx[keep], xmn, xmx, nmn, nmx);
                    const float gain = area_y - box_area(nmn, nmx);
                    const int new_hy = 1 + max(yh[keep], hx);
                    const int new_h = 1 + max(new_hy, yh[1 - keep]);
                    if (gain > best && new_h <= max(hnode, PT_LBVH_MAX_ROT_HEIGHT)) {
                        best = gain; best_y = y; best_keep = keep;
                        kmn = nmn; kmx = nmx; kch = gc; kh[0] = yh[0]; kh[1] = yh[1];
                    }
                }
                ych = gc;
            }
            (void)ych;
            // two more candidates when both children are inner nodes: trade a grandchild of A for one of B
            // (A1 <-> B1 and A1 <-> B2; the other two pairings are the same trees with the children renamed)
            int best_swap = -1;
            int2 ga = make_int2(0, 0), gb2 = make_int2(0, 0);
            float4 a1mn, a1mx, a2mn, a2mx, b1mn, b1mx, b2mn, b2mx;
            int ha1 = 0, ha2 = 0, hb1 = 0, hb2 = 0;
            if (ch.x >= 0 && ch.y >= 0) {
                ga = __ldcg(&t.children[ch.x]); gb2 = __ldcg(&t.children[ch.y]);
                lbvh_child_box(t, ga.x, a1mn, a1mx, ha1); lbvh_child_box(t, ga.y, a2mn, a2mx, ha2);
                lbvh_child_box(t, gb2.x, b1mn, b1mx, hb1); lbvh_child_box(t, gb2.y, b2mn, b2mx, hb2);
                const float area_ab = box_area(amn, amx) + box_area(bmn, bmx);
                for (int w = 0; w < 2; ++w) {                        // w = 0: A1 <-> B1, w = 1: A1 <-> B2
                    float4 pmn, pmx, qmn, qmx;
                    box_union(w == 0 ? b1mn : b2mn, w == 0 ? b1mx : b2mx, a2mn, a2mx, pmn, pmx);      // A' = (B?, A2)
                    box_union(a1mn, a1mx, w == 0 ? b2mn : b1mn, w == 0 ? b2mx : b1mx, qmn, qmx);      // B' = (A1, B other)
                    const float gain = area_ab - box_area(pmn, pmx) - box_area(qmn, qmx);
                    const int nha = 1 + max(w == 0 ? hb1 : hb2, ha2), nhb = 1 + max(ha1, w == 0 ? hb2 : hb1);
                    if (gain > best && 1 + max(nha, nhb) <= max(hnode, PT_LBVH_MAX_ROT_HEIGHT)) { best = gain; best_swap = w; best_y = -1; }
                }
            }
            if (best_swap >= 0) {
                const int A = ch.x, B = ch.y;
                const int moved_to_a = best_swap == 0 ? gb2.x : gb2.y, stays_in_b = best_swap == 0 ? gb2.y : gb2.x;
                float4 pmn, pmx, qmn, qmx;
                box_union(best_swap == 0 ? b1mn : b2mn, best_swap == 0 ? b1mx : b2mx, a2mn, a2mx, pmn, pmx);
                box_union(a1mn, a1mx, best_swap == 0 ? b2mn : b1mn, best_swap == 0 ? b2mx : b1mx, qmn, qmx);
                t.children[A] = make_int2(moved_to_a, ga.y);
                t.children[B] = make_int2(ga.x, stays_in_b);
                t.inner_mn[A] = pmn; t.inner_mx[A] = pmx; t.inner_mn[B] = qmn; t.inner_mx[B] = qmx;
                const int nha = 1 + max(best_swap == 0 ? hb1 : hb2, ha2), nhb = 1 + max(ha1, best_swap == 0 ? hb2 : hb1);
                t.height[A] = (unsigned char)nha; t.height[B] = (unsigned char)nhb;
                lbvh_set_parent(t, moved_to_a, A);
                lbvh_set_parent(t, ga.x, B);
                hnode = 1 + max(nha, nhb);
            }
            if (best_y >= 0) {
                const int Y = best_y == 0 ? ch.x : ch.y;
                const int X = best_y == 0 ? ch.y : ch.x;
                const int hx = best_y == 0 ? hb : ha;
                const int kept = best_keep == 0 ? kch.x : kch.y;
                const int up = best_keep == 0 ? kch.y : kch.x;
                // Y = (kept, X), node = (up, Y)
                t.children[Y] = make_int2(kept, X);
                t.inner_mn[Y] = kmn; t.inner_mx[Y] = kmx;
                const int new_hy = 1 + max(kh[best_keep], hx);
                t.height[Y] = (unsigned char)new_hy;
                lbvh_set_parent(t, X, Y);
                lbvh_set_parent(t, up, node);
                ch = make_int2(up, Y);
                t.children[node] = ch;
                hnode = 1 + max(new_hy, kh[1 - best_keep]);
            }
        }
        float4 nmn, nmx;
        box_union(amn, amx, bmn, bmx, nmn, nmx);      // the same leaves before and after a rotation
        t.inner_mn[node] = nmn; t.inner_mx[node] = nmx;
        t.height[node] = (unsigned char)min(hnode, 255);
        __threadfence();
        node = __ldcg(&t.parent_of_inner[node]);
    }
}

// SAH cost of the tree relative to the root's area: sum of the inner nodes' areas / area(root) (one primitive per leaf)
__global__ void k_lbvh_sah(int n, const float4* __restrict__ inner_mn, const float4* __restrict__ inner_mx, float* __restrict__ out) {
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) acc += box_area(inner_mn[i], inner_mx[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc / fmaxf(box_area(inner_mn[0], inner_mx[0]), 1e-30f));
}

__device__ __forceinline__ void centre_half(float4 mn, float4 mx, float3& c, float3& h) {
    c = f3(0.5f * (mn.x + mx.x), 0.5f * (mn.y + mx.y), 0.5f * (mn.z + mx.z));
    h = f3(fmaxf(mx.x - c.x, c.x - mn.x), fmaxf(mx.y - c.y, c.y - mn.y), fmaxf(mx.z - c.z, c.z - mn.z));
    h = f3(h.x + 4.8e-7f * (fabsf(c.x) + h.x) + 1e-30f, h.y + 4.8e-7f * (fabsf(c.y) + h.y) + 1e-30f,
           h.z + 4.8e-7f * (fabsf(c.z) + h.z) + 1e-30f);
}

// pass 5: pack traversal nodes (child boxes in the parent, leaves as ~(prim<<2|type))
__global__ void k_lbvh_pack(int n, const uint32_t* __restrict__ sorted_ids, const int* __restrict__ bvh_prims,
                            const float4* __restrict__ geo1, const float4* __restrict__ leaf_mn,
                            const float4* __restrict__ leaf_mx, const int2* __restrict__ children,
                            const float4* __restrict__ inner_mn, const float4* __restrict__ inner_mx,
                            float4* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int2 ch = children[i];
    float4 amn, amx, bmn, bmx;
    int c0, c1;
    if (ch.x < 0) {
        uint32_t s = sorted_ids[~ch.x]; amn = leaf_mn[s]; amx = leaf_mx[s];
        int p = bvh_prims[s]; c0 = ~((p << 2) | (__float_as_int(geo1[p].z) & 3));
    } else { amn = inner_mn[ch.x]; amx = inner_mx[ch.x]; c0 = ch.x; }
    if (ch.y < 0) {
        uint32_t s = sorted_ids[~ch.y]; bmn = leaf_mn[s]; bmx = leaf_mx[s];
        int p = bvh_prims[s]; c1 = ~((p << 2) | (__float_as_int(geo1[p].z) & 3));
    } else { bmn = inner_mn[ch.y]; bmx = inner_mx[ch.y]; c1 = ch.y; }
    float3 ca, ha, cb, hb;
    centre_half(amn, amx, ca, ha);
    centre_half(bmn, bmx, cb, hb);
#if PT_PACKED_SLABS
    // (cx, cy) and (hx, hy) of a child sit in aligned pairs: the x and y slabs are evaluated by packed FFMA2 (pt_wavefront.cuh)
    nodes[4 * i + 0] = make_float4(ca.x, ca.y, ha.x, ha.y);
    nodes[4 * i + 1] = make_float4(cb.x, cb.y, hb.x, hb.y);
#if PT_PACKED_SLABS == 2
    nodes[4 * i + 2] = make_float4(ca.z, cb.z, ha.z, hb.z);     // z of both children as pairs too
#else
    nodes[4 * i + 2] = make_float4(ca.z, ha.z, cb.z, hb.z);
#endif
#else
    nodes[4 * i + 0] = make_float4(ca.x, ca.y, ca.z, ha.x);
    nodes[4 * i + 1] = make_float4(ha.y, ha.z, cb.x, cb.y);
    nodes[4 * i + 2] = make_float4(cb.z, hb.x, hb.y, hb.z);
#endif
    nodes[4 * i + 3] = make_float4(__int_as_float(c0), __int_as_float(c1), 0.f, 0.f);
}

// pass 6: the entry grid (struct EntryGrid, pt_wavefront.cuh).  One thread per block of one level: the tree is pruned
// by the block's box -- a node is replaced by those of its children whose boxes overlap the block, as long as the
// frontier stays within four links.  A node with ONE overlapping child is always replaced (that is the descent a ray
// would otherwise walk); a node with two is split when there is room (a block that straddles a high split plane of
// the tree then gets the deep nodes on both sides instead of their far-away common ancestor).  Conservative by
// construction: a primitive whose leaf box overlaps the block is below one of the frontier's links, because all of its
// ancestors' boxes overlap the block too.  Links: >= 0 inner node, < 0 leaf (as in the nodes), 0x7fffffff none.
struct EntryGridBuild {
    int4* cells;
    const uint4* levels;     // (first cell, nx, ny, nz)
    float3 org;
    int k, lmax;
    uint32_t n_cells;
};
__device__ __forceinline__ bool box_overlaps_block(float cx, float cy, float cz, float hx, float hy, float hz, float3 bc, float bh) {
    return fabsf(cx - bc.x) <= hx + bh && fabsf(cy - bc.y) <= hy + bh && fabsf(cz - bc.z) <= hz + bh;
}
__global__ void k_entry_grid(EntryGridBuild g, const float4* __restrict__ nodes) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= g.n_cells) return;
    int l = 0;
    while (l < g.lmax && cell >= g.levels[l + 1].x) ++l;
    const uint4 lv = g.levels[l];
    uint32_t rem = cell - lv.x;
    const uint32_t iz = rem % lv.w; rem /= lv.w;
    const uint32_t iy = rem % lv.z; const uint32_t ix = rem / lv.z;
    const float s = __int_as_float((g.k - l + 127) << 23);              // cell size 2^(k-l)
    // the block spans two cells per axis; its half-extent is padded by 2^-18 of itself against the rounding of the
    // ray side's (mn - org) * 2^(l-k) <= i + 2 test
    const float3 bc = f3(g.org.x + ((float)ix + 1.f) * s, g.org.y + ((float)iy + 1.f) * s, g.org.z + ((float)iz + 1.f) * s);
    const float bh = s * 1.000004f + 4e-7f * fmaxf(fmaxf(fabsf(bc.x), fabsf(bc.y)), fabsf(bc.z));
    int f[4] = {0, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    int n = 1;
    uint32_t closed = 0u;                                               // links that must not be expanded any further
    for (int iter = 0; iter < 4096; ++iter) {
        int j = -1;
#pragma unroll
        for (int q = 3; q >= 0; --q) if (q < n && f[q] >= 0 && !((closed >> q) & 1u)) j = q;
        if (j < 0) break;
        const float4* p = nodes + 4 * (size_t)f[j];
        const float4 n0 = p[0], n1 = p[1], n2 = p[2], n3 = p[3];
        const bool ha = box_overlaps_block(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, bc, bh);
        const bool hb = box_overlaps_block(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, bc, bh);
        const int add = (int)ha + (int)hb;
        const int la = __float_as_int(n3.x), lb = __float_as_int(n3.y);
        // a link to a LEAF never enters the frontier: the walk tests a primitive only after the ray has hit the
        // primitive's own box in its parent, and the frontier must not change which primitives get tested (round-2
        // A/B: with leaf links in the frontier the primitive tests per segment doubled and the images left the
        // grid-less ones in the graze class)
        if (n - 1 + add > 4 || (ha && la < 0) || (hb && lb < 0)) { closed |= 1u << j; continue; }
        // remove link j (the order of the others is kept, and with it their `closed` bits), append the children
        const uint32_t low = closed & ((1u << j) - 1u), high = (closed >> (j + 1)) << j;
        closed = low | high;
#pragma unroll
        for (int q = 0; q < 3; ++q) if (q >= j) f[q] = f[q + 1];
        --n;
        if (ha) { f[n] = la; ++n; }
        if (hb) { f[n] = lb; ++n; }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q >= n) f[q] = 0x7fffffff;
    }
    g.cells[cell] = make_int4(f[0], f[1], f[2], f[3]);
}

// pass 7 (build variant PT_BVH4): 4-wide nodes collapsed from the packed binary tree.  Node4[i] exists for EVERY binary inner node i (no
// depth pass, no compaction: a walk that starts at the root only ever reaches the even-depth ones): each child of i that is an inner node
// is replaced by its own two children, a leaf child stays.  128 bytes, component-major so that the four slab tests vectorise:
//   q0 = cx[4]  q1 = cy[4]  q2 = cz[4]  q3 = hx[4]  q4 = hy[4]  q5 = hz[4]  q6 = links[4]  (q7 unused)
// an unused slot has half-extent -1 (never hit) and link 0x7fffffff.
__global__ void k_lbvh_collapse4(int n, const float4* __restrict__ nodes, float4* __restrict__ nodes4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    float c[3][4], h[3][4]; int link[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { c[0][k] = c[1][k] = c[2][k] = 0.f; h[0][k] = h[1][k] = h[2][k] = -1.f; link[k] = 0x7fffffff; }
    int m = 0;
    auto put = [&](float cx, float cy, float cz, float hx, float hy, float hz, int l) {
        c[0][m] = cx; c[1][m] = cy; c[2][m] = cz; h[0][m] = hx; h[1][m] = hy; h[2][m] = hz; link[m] = l; ++m;
    };
    auto take = [&](const float4* p, int which, bool expand) {
        const float4 n0 = p[0], n1 = p[1], n2 = p[2], n3 = p[3];
        const int l = __float_as_int(which == 0 ? n3.x : n3.y);
        if (expand && l >= 0) return l;                       // caller descends into this inner child
        if (which == 0) put(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, l);
        else put(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, l);
        return -1;
    };
    const float4* p = nodes + 4 * (size_t)i;
    for (int which = 0; which < 2; ++which) {
        const int inner = take(p, which, true);
        if (inner >= 0) {
            const float4* q = nodes + 4 * (size_t)inner;
            // a never-entered second child (single-primitive tree) has a negative half-extent and stays out
            take(q, 0, false);
            take(q, 1, false);
        }
    }
    float4* o = nodes4 + 8 * (size_t)i;
    o[0] = make_float4(c[0][0], c[0][1], c[0][2], c[0][3]); o[1] = make_float4(c[1][0], c[1][1], c[1][2], c[1][3]);
    o[2] = make_float4(c[2][0], c[2][1], c[2][2], c[2][3]); o[3] = make_float4(h[0][0], h[0][1], h[0][2], h[0][3]);
    o[4] = make_float4(h[1][0], h[1][1], h[1][2], h[1][3]); o[5] = make_float4(h[2][0], h[2][1], h[2][2], h[2][3]);
    o[6] = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), __int_as_float(link[2]), __int_as_float(link[3]));
    o[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// a tree over a single primitive: one node whose second child can never be entered
__global__ void k_lbvh_single(const int* __restrict__ bvh_prims, const float4* __restrict__ geo0,
                              const float4* __restrict__ geo1, float4* __restrict__ nodes) {
    int p = bvh_prims[0];
    Aabb b = prim_aabb(geo0[p], geo1[p]);
    int c0 = ~((p << 2) | (__float_as_int(geo1[p].z) & 3));
    float3 c, h;
    centre_half(make_float4(b.mn.x, b.mn.y, b.mn.z, 0.f), make_float4(b.mx.x, b.mx.y, b.mx.z, 0.f), c, h);
#if PT_PACKED_SLABS
    nodes[0] = make_float4(c.x, c.y, h.x, h.y);
    nodes[1] = make_float4(0.f, 0.f, -1.f, -1.f);           // negative half-extent: near > far, never entered
#if PT_PACKED_SLABS == 2
    nodes[2] = make_float4(c.z, 0.f, h.z, -1.f);
#else
    nodes[2] = make_float4(c.z, h.z, 0.f, -1.f);
#endif
#else
    nodes[0] = make_float4(c.x, c.y, c.z, h.x);
    nodes[1] = make_float4(h.y, h.z, 0.f, 0.f);
    nodes[2] = make_float4(0.f, -1.f, -1.f, -1.f);          // negative half-extent: near > far, never entered
#endif
    nodes[3] = make_float4(__int_as_float(c0), __int_as_float(c0), 0.f, 0.f);
}

}  // namespace pt
