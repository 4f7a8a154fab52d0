// pt_lbvh.cuh -- GPU LBVH build (Morton codes -> radix sort -> Karras 2012 hierarchy -> bottom-up
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
#include <cub/cub.cuh>
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

// pass 4: bottom-up boxes of the internal nodes; the second thread to arrive at a node fits it
__global__ void k_lbvh_fit(int n, const uint32_t* __restrict__ sorted_ids, const float4* __restrict__ leaf_mn,
                           const float4* __restrict__ leaf_mx, const int2* __restrict__ children,
                           const int* __restrict__ parent_of_inner, const int* __restrict__ parent_of_leaf,
                           float4* __restrict__ inner_mn, float4* __restrict__ inner_mx, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_of_leaf[i];
    while (node >= 0) {
        if (atomicAdd(&flags[node], 1) == 0) return;   // first arrival: the sibling subtree is not done yet
        __threadfence();
        int2 ch = children[node];
        float4 amn, amx, bmn, bmx;
        if (ch.x < 0) { uint32_t s = sorted_ids[~ch.x]; amn = leaf_mn[s]; amx = leaf_mx[s]; }
        else { amn = __ldcg(&inner_mn[ch.x]); amx = __ldcg(&inner_mx[ch.x]); }
        if (ch.y < 0) { uint32_t s = sorted_ids[~ch.y]; bmn = leaf_mn[s]; bmx = leaf_mx[s]; }
        else { bmn = __ldcg(&inner_mn[ch.y]); bmx = __ldcg(&inner_mx[ch.y]); }
        inner_mn[node] = make_float4(fminf(amn.x, bmn.x), fminf(amn.y, bmn.y), fminf(amn.z, bmn.z), 0.f);
        inner_mx[node] = make_float4(fmaxf(amx.x, bmx.x), fmaxf(amx.y, bmx.y), fmaxf(amx.z, bmx.z), 0.f);
        __threadfence();
        node = parent_of_inner[node];
    }
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
    nodes[4 * i + 0] = make_float4(ca.x, ca.y, ca.z, ha.x);
    nodes[4 * i + 1] = make_float4(ha.y, ha.z, cb.x, cb.y);
    nodes[4 * i + 2] = make_float4(cb.z, hb.x, hb.y, hb.z);
    nodes[4 * i + 3] = make_float4(__int_as_float(c0), __int_as_float(c1), 0.f, 0.f);
}

// a tree over a single primitive: one node whose second child can never be entered
__global__ void k_lbvh_single(const int* __restrict__ bvh_prims, const float4* __restrict__ geo0,
                              const float4* __restrict__ geo1, float4* __restrict__ nodes) {
    int p = bvh_prims[0];
    Aabb b = prim_aabb(geo0[p], geo1[p]);
    int c0 = ~((p << 2) | (__float_as_int(geo1[p].z) & 3));
    float3 c, h;
    centre_half(make_float4(b.mn.x, b.mn.y, b.mn.z, 0.f), make_float4(b.mx.x, b.mx.y, b.mx.z, 0.f), c, h);
    nodes[0] = make_float4(c.x, c.y, c.z, h.x);
    nodes[1] = make_float4(h.y, h.z, 0.f, 0.f);
    nodes[2] = make_float4(0.f, -1.f, -1.f, -1.f);          // negative half-extent: near > far, never entered
    nodes[3] = make_float4(__int_as_float(c0), __int_as_float(c0), 0.f, 0.f);
}

}  // namespace pt
