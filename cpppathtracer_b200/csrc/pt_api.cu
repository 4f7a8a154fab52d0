// pt_api.cu -- host side of libpt_b200.so: scene container, LBVH build driver, wave scheduler and the
// extern "C" entry points declared in include/pt_api.h.  There is no CPU fallback anywhere in this file:
// without a CUDA device pt_scene_create fails with PT_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pt_api.h"
#include "pt_lbvh.cuh"
#include "pt_math.cuh"
#include "pt_wavefront.cuh"
#include "pt_display.cuh"

using namespace pt;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(PT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                         std::to_string(__LINE__) + ")");                                \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Every entry point that touches a scene runs on the device the scene was created on, whatever the calling thread's
// current device is (one process driving several GPUs, a host application that switches devices), and restores it.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define PT_ON_SCENE_DEVICE(s) DeviceGuard pt_device_guard_((s) ? (s)->device : 0)

// page-locked staging buffer that only grows: the scene description is re-uploaded by every pt_build, and a pageable
// source makes each cudaMemcpyAsync a synchronous staged copy (3.3 -> 2 ms per rebuild at 128 k primitives)
template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaHostAlloc((void**)&p, count * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

struct HostPrim {
    int type;
    float c[3];
    float radius, y_pos, height;
    int mat_slot;
};

struct MatKey {
    pt_material m;
    bool operator<(const MatKey& o) const { return memcmp(&m, &o.m, sizeof(pt_material)) < 0; }
};

struct Texture { cudaArray_t arr = nullptr; cudaTextureObject_t obj = 0; };

}  // namespace

struct pt_scene {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // host description
    std::vector<HostPrim> prims;
    std::vector<pt_material> materials;      // distinct materials, slot order
    std::map<MatKey, int> material_slots;
    bool dirty = true, built = false;
    std::vector<uint32_t> rank_cache;        // reference leaf rank per primitive, from the last build after an add
    bool rank_valid = false;
    bool layout_valid = false;               // staging buffers + device index lists match the primitive list and its order
    size_t st_n = 0;
    std::vector<int> moved;                  // primitives edited by pt_update_prim since the last build
    std::vector<float> plat_y_host; std::vector<int> plat_ids_host;
    // options
    uint32_t opt_wave_paths = 128u << 20;  // 64 spp per wave at 1080p: 26 GB of hit queues; 5.57 vs 5.40 Gpaths/s with 32 Mi (16 spp), 5.23 with 16 Mi
                                           // (profiles/ab/wave_sweep_v9.log); ensure_queues halves the wave until the queues fit in free memory
    int opt_stage_timing = 0;
    int opt_tie_order = 1;
    int opt_sky_quarter = 1;
    int opt_count_steps = 0;
    int opt_ref_leaf_boxes = 0;
    int opt_inner_run = 6;         // r02 sweep on one box (profiles/ab/r2_inner_run.log): 6 is +0.8 % on C2 and +1.5 % on C3 over 4, -0.6 % on C1
    int opt_sah_treelets = 0;      // 1: the refit passes restructure 7-leaf treelets by dynamic programming (slower build, lower SAH)
    int opt_sah_passes = 4;        // SAH refit passes (tree rotations) after the Karras build; 0: plain LBVH
    int opt_entry_grid = 0;        // finest level the entry grid may use (0, the default: no grid, every walk starts at the root;
                                   // the grid halves the node visits per ray but not the warp's longest walk: DESIGN.md section 3.4)
    int opt_entry_cells_per_prim = 2;   // the finest level has at most this many blocks per primitive
    int opt_split_walk = 0;             // entry-grid library: 1 = bounces run as k_shade_classify + k_walk (rays binned by level class)
    DevBuf<float4> rq_p, rq_d, rq_thr;  // ... and the ray queue between the two (two two-ended regions, single-buffered)
    DevBuf<BounceCtrl> rctrl;
    uint32_t rq_capacity = 0;
    // device scene
    DevBuf<float4> geo0, geo1, nodes, leaf_mn, leaf_mx, leaf_c, inner_mn, inner_mx, test_mn, test_mx;
    DevBuf<MaterialRec> mats;
    DevBuf<float> plat_y;
    DevBuf<int> plat_id, bvh_prims, parent_inner, parent_leaf, flags;
    DevBuf<int2> children;
    DevBuf<uint32_t> codes, codes_sorted, ids, ids_sorted, bounds6;
    DevBuf<int> orig_id;
    DevBuf<uint32_t> sort_hist;
    PinnedBuf<float4> st_g0, st_g1;      // staging of the uploads of pt_build
    PinnedBuf<int> st_orig, st_tree;
    DevBuf<float4> nodes4;               // build variant PT_BVH4
    DevBuf<int4> grid_cells;
    DevBuf<uint4> grid_levels;
    EntryGrid grid{nullptr, nullptr, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0, -1};
    int n_plat = 0, n_tree = 0, has_emitters = 0;
    Texture sky;
    std::map<uint64_t, Texture> textures;
    // frame state
    int width = 0, height = 0;
    DevBuf<float4> accum;
    // wave queues
    DevBuf<float4> path_out;
    DevBuf<float4> hq_p[2], hq_d[2], hq_thr[2];   // hit queues (two two-ended regions each), double-buffered over segments
    DevBuf<float4> q_o, q_d;                      // parity path only: the caller's ray batch
    DevBuf<uint2> hit;                            // parity path only: results by ray slot
    DevBuf<BounceCtrl> ctrl;
    DevBuf<unsigned long long> seg_counter;
    uint32_t capacity = 0;
    // parity scratch
    DevBuf<float> scratch_f[8];
    DevBuf<int32_t> scratch_i[2];
    // display pipeline (pt_dispatch_ray)
    DevBuf<float4> disp_color, disp_normal, disp_mix;
    DevBuf<float> disp_depth;
    DevBuf<uint8_t> disp_bgra;
    int disp_w = 0, disp_h = 0;
    // stats
    pt_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> stage_events;
    int grid_bounce = 0, grid_walk = 0;
    DevBuf<unsigned char> node_height;
    DevBuf<float> sah_dev;
    float sah_host[2] = {0.f, 0.f};   // SAH cost (sum of inner areas / root area) before and after the refit passes
    uint8_t* disp_host = nullptr;      // pinned BGRA frame handed out by pt_dispatch_ray_frame
    size_t disp_host_bytes = 0;
    // pipelined dispatch (pt_dispatch_ray_begin / _end): two frames in flight, the D2H copy of one on its own stream
    // under the trace of the next
    DevBuf<uint8_t> pipe_bgra[2];
    uint8_t* pipe_host[2] = {nullptr, nullptr};
    size_t pipe_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t pipe_filtered[2] = {nullptr, nullptr}, pipe_copied[2] = {nullptr, nullptr};
    uint64_t pipe_begun = 0, pipe_ended = 0;
    int last_depth = 0;
    size_t stage_events_used = 0;
    bool frame_recorded = false;
    DevBuf<float4> read_tmp;
    // pipelined read-back (pt_read_accum_begin / _end): two device staging buffers, the D2H copy on the copy stream
    DevBuf<float4> read_pipe[2];
    cudaEvent_t read_ready[2] = {nullptr, nullptr}, read_done[2] = {nullptr, nullptr};
    uint64_t read_begun = 0, read_ended = 0;
};

namespace {

// MotionalCamera::GetCopy, cuSrc/motional_camera.cu:177-200 (host math, fp32, no contraction)
struct H3 { float x, y, z; };
inline H3 h3(float x, float y, float z) { return H3{x, y, z}; }
inline H3 hsub(H3 a, H3 b) { return h3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline H3 hadd(H3 a, H3 b) { return h3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline H3 hscale(H3 a, float s) { return h3(a.x * s, a.y * s, a.z * s); }
inline float hdot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline H3 hcross(H3 a, H3 b) { return h3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline H3 hnormalize(H3 v) { float inv = 1.0f / sqrtf(hdot(v, v)); return hscale(v, inv); }

struct HostCameraFrame { H3 u, v, w, top_left, horizontal, vertical; float dist; };

HostCameraFrame camera_frame(const pt_camera* c) {
    HostCameraFrame f;
    H3 origin = h3(c->origin[0], c->origin[1], c->origin[2]);
    H3 look_at = h3(c->look_at[0], c->look_at[1], c->look_at[2]);
    H3 vup = h3(c->vup[0], c->vup[1], c->vup[2]);
    float theta = (float)((double)c->view_fov_deg * 3.14159265358979323846 / 180);
    float aspect = (float)c->width / (float)c->height;
    float half_h = tanf(theta / 2);
    float half_w = aspect * half_h;
    f.w = hnormalize(hsub(origin, look_at));
    f.u = hnormalize(hcross(vup, f.w));
    f.v = hcross(f.w, f.u);
    f.dist = sqrtf(hdot(hsub(origin, look_at), hsub(origin, look_at)));
    f.top_left = hsub(hadd(hsub(origin, hscale(f.u, half_w * f.dist)), hscale(f.v, half_h * f.dist)), hscale(f.w, f.dist));
    f.horizontal = hscale(f.u, 2 * half_w * f.dist);
    f.vertical = hscale(f.v, -2 * half_h * f.dist);
    return f;
}

CameraFrame to_device_frame(const pt_camera* c) {
    HostCameraFrame h = camera_frame(c);
    CameraFrame d;
    d.origin = make_float3(c->origin[0], c->origin[1], c->origin[2]);
    d.u = make_float3(h.u.x, h.u.y, h.u.z);
    d.v = make_float3(h.v.x, h.v.y, h.v.z);
    d.top_left = make_float3(h.top_left.x, h.top_left.y, h.top_left.z);
    d.horizontal = make_float3(h.horizontal.x, h.horizontal.y, h.horizontal.z);
    d.vertical = make_float3(h.vertical.x, h.vertical.y, h.vertical.z);
    d.lens_radius = c->lens_radius;
    d.width = c->width; d.height = c->height;
    return d;
}

MaterialRec make_material_rec(const pt_material& m) {
    MaterialRec r;
    memset(&r, 0, sizeof(r));
    r.kd[0] = m.kd[0]; r.kd[1] = m.kd[1]; r.kd[2] = m.kd[2];
    r.emit = m.emit_intensity;
    r.ior = m.refractive_index;
    r.reflectivity = m.reflectivity;
    r.type = m.type;
    float alpha = powf(1000.0f, m.smoothness);       // material.cu:43,69,104
    r.inv_alpha = 1.0 / (double)alpha;               // `1.0 / alpha` at material.cu:45,78,105
    r.flags = (alpha >= 1e9f) ? PT_MATF_SPECULAR : 0;
    r.tex = m.have_tex ? (unsigned long long)m.tex : 0ull;
    return r;
}

int material_slot(pt_scene* s, const pt_material& m_in) {
    MatKey k;
    memset(&k.m, 0, sizeof(k.m));
    k.m.type = m_in.type; k.m.have_tex = m_in.have_tex ? 1 : 0;
    memcpy(k.m.kd, m_in.kd, sizeof(k.m.kd));
    k.m.refractive_index = m_in.refractive_index; k.m.emit_intensity = m_in.emit_intensity;
    k.m.smoothness = m_in.smoothness; k.m.reflectivity = m_in.reflectivity;
    k.m.tex = m_in.have_tex ? m_in.tex : 0;
    auto it = s->material_slots.find(k);
    if (it != s->material_slots.end()) return it->second;
    int slot = (int)s->materials.size();
    s->materials.push_back(k.m);
    s->material_slots[k] = slot;
    return slot;
}

// The reference's leaf visiting order, needed only to resolve exactly tied hits the way it does.
// Restates the *ordering* produced by SceneBVH::Divide (cuSrc/bvh.cu:31-90: sort the range by box centre
// along the longest axis, split at the median index) and the right-son-first pop order of TraceRay
// (bvh.cu:201-202): a leaf's rank is its position counted from the END of the final permutation.
struct RefOrder {
    const std::vector<HostPrim>& prims;
    std::vector<float> mn, mx;   // 3 floats per prim, Object::GetAABBMin/Max (object.cu:134-170)
    std::vector<float> pmn, pmx; // 3 floats per prim: box of the reference node whose split made the prim a leaf
    std::vector<int> order;
    explicit RefOrder(const std::vector<HostPrim>& p) : prims(p) {
        const float tol = PT_BOUNCE_TMIN * 5.f;
        mn.resize(3 * p.size()); mx.resize(3 * p.size()); order.resize(p.size());
        pmn.assign(3 * p.size(), 0.f); pmx.assign(3 * p.size(), 0.f);
        for (size_t i = 0; i < p.size(); ++i) {
            const HostPrim& o = p[i];
            float ar = o.radius >= 0 ? o.radius : -o.radius;
            order[i] = (int)i;
            if (o.type == PT_PRIM_SPHERE) {
                for (int k = 0; k < 3; ++k) { mn[3 * i + k] = o.c[k] - ar; mx[3 * i + k] = o.c[k] + ar; }
            } else if (o.type == PT_PRIM_PLATFORM) {
                mn[3 * i] = -PT_TMAX * 5; mn[3 * i + 1] = o.y_pos - tol; mn[3 * i + 2] = -PT_TMAX * 5;
                mx[3 * i] = PT_TMAX * 5;  mx[3 * i + 1] = o.y_pos + tol; mx[3 * i + 2] = PT_TMAX * 5;
            } else {
                mn[3 * i] = o.c[0] - ar; mn[3 * i + 1] = o.c[1] - o.height / 2 - tol; mn[3 * i + 2] = o.c[2] - ar;
                mx[3 * i] = o.c[0] + ar; mx[3 * i + 1] = o.c[1] + o.height / 2 + tol; mx[3 * i + 2] = o.c[2] + ar;
            }
        }
    }
    // depth: levels below the root; the two halves of the top levels are independent (disjoint index ranges) and
    // run on their own threads -- the ordering is the reference's, only its wall time changes (100 k primitives:
    // 110 ms single-threaded, which was the whole cost of pt_build at that size)
    void split(int l, int r, int depth = 0) {
        if (r - l < 2) return;
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) { lo[k] = mn[3 * order[l] + k]; hi[k] = mx[3 * order[l] + k]; }
        for (int i = l + 1; i < r; ++i)
            for (int k = 0; k < 3; ++k) {
                float a = mn[3 * order[i] + k], b = mx[3 * order[i] + k];
                lo[k] = lo[k] < a ? lo[k] : a;
                hi[k] = hi[k] > b ? hi[k] : b;
            }
        float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
        int axis = (sx >= sy && sx >= sz) ? 0 : (sy >= sz ? 1 : 2);
        const float* MN = mn.data(); const float* MX = mx.data();
        std::sort(order.begin() + l, order.begin() + r, [MN, MX, axis](int a, int b) {
            return (MN[3 * a + axis] + MX[3 * a + axis]) / 2 < (MN[3 * b + axis] + MX[3 * b + axis]) / 2;
        });
        int mid = (l + r) / 2;
        // TraceRay tests a leaf's object whenever the leaf is popped, i.e. whenever its PARENT's box was hit
        // (bvh.cu:175-180,200-202): remember that box for the leaves this split creates
        if (mid - l == 1) for (int k = 0; k < 3; ++k) { pmn[3 * order[l] + k] = lo[k]; pmx[3 * order[l] + k] = hi[k]; }
        if (r - mid == 1) for (int k = 0; k < 3; ++k) { pmn[3 * order[mid] + k] = lo[k]; pmx[3 * order[mid] + k] = hi[k]; }
        if (depth < 3 && r - l >= 4096) {
            std::thread left([this, l, mid, depth] { split(l, mid, depth + 1); });
            split(mid, r, depth + 1);
            left.join();
        } else {
            split(l, mid, depth + 1);
            split(mid, r, depth + 1);
        }
    }
    std::vector<uint32_t> ranks() {
        split(0, (int)order.size());
        std::vector<uint32_t> rk(order.size());
        const size_t n = order.size();
        for (size_t pos = 0; pos < n; ++pos) rk[order[pos]] = (uint32_t)(n - 1 - pos);
        return rk;
    }
};

// Sizes, uploads and fills the entry grid (struct EntryGrid) for the tree that was just packed.  The gridded box is
// computed on the host from the primitives (no device round trip): it only has to CONTAIN every leaf box, so it is
// padded by 2^-10 of the scene's size plus the boxes' own inflation.
int build_entry_grid(pt_scene* s, int nt) {
    s->grid.lmax = -1;
    if (s->opt_entry_grid <= 0 || s->opt_ref_leaf_boxes || nt < 2) return PT_OK;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, amax = 0.0;
    for (const HostPrim& p : s->prims) {
        if (p.type != PT_PRIM_SPHERE && p.type != PT_PRIM_CYLINDER) continue;
        const double ar = std::fabs((double)p.radius);
        const double hy = p.type == PT_PRIM_CYLINDER ? std::fabs((double)p.height) / 2 + 1e-4 : ar;
        const double e[3] = {ar, hy, ar};
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], (double)p.c[k] - e[k]); hi[k] = std::max(hi[k], (double)p.c[k] + e[k]);
            amax = std::max(amax, std::fabs((double)p.c[k]) + e[k]);
        }
    }
    double size = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    if (!(size > 0.0) || !(size < 1e30) || !(amax < 1e30)) return PT_OK;      // degenerate or non-finite scene: no grid
    const double pad = size / 1024.0 + amax * 2e-6 + 1e-6;
    for (int k = 0; k < 3; ++k) { lo[k] -= pad; hi[k] += pad; }
    size += 2 * pad;
    int kexp = 0;
    (void)std::frexp(size, &kexp);                 // size = m 2^kexp, 0.5 <= m < 1  =>  2^kexp > size
    if (kexp < -100 || kexp > 100) return PT_OK;
    std::vector<uint4> levels;
    uint64_t total = 0;
    const uint64_t cap_level = (uint64_t)s->opt_entry_cells_per_prim * (uint64_t)nt + 64, cap_total = 4u << 20;
    for (int l = 0; l <= s->opt_entry_grid && l <= 20; ++l) {
        const double cell = std::ldexp(1.0, kexp - l);
        uint64_t n[3];
        for (int k = 0; k < 3; ++k) n[k] = (uint64_t)std::max(1.0, std::ceil((hi[k] - lo[k]) / cell));
        const uint64_t cells = n[0] * n[1] * n[2];
        if (l > 0 && (cells > cap_level || total + cells > cap_total)) break;
        levels.push_back(make_uint4((uint32_t)total, (uint32_t)n[0], (uint32_t)n[1], (uint32_t)n[2]));
        total += cells;
    }
    const int lmax = (int)levels.size() - 1;
    if (lmax < 1) return PT_OK;
    CK(s->grid_levels.reserve(levels.size())); CK(s->grid_cells.reserve(total));
    CK(cudaMemcpyAsync(s->grid_levels.p, levels.data(), levels.size() * sizeof(uint4), cudaMemcpyHostToDevice, s->stream));
    EntryGridBuild b;
    b.cells = s->grid_cells.p; b.levels = s->grid_levels.p; b.org = make_float3((float)lo[0], (float)lo[1], (float)lo[2]);
    b.k = kexp; b.lmax = lmax; b.n_cells = (uint32_t)total;
    k_entry_grid<<<(unsigned)((total + 127) / 128), 128, 0, s->stream>>>(b, s->nodes.p);
    s->stats.launches += 1;
    s->grid.cells = s->grid_cells.p; s->grid.levels = s->grid_levels.p; s->grid.org = b.org;
    // (float) conversions round to nearest: the padding is three orders of magnitude above that
    s->grid.hi = make_float3((float)hi[0], (float)hi[1], (float)hi[2]);
    s->grid.k = kexp; s->grid.lmax = lmax;
    s->stats.entry_levels = (uint32_t)lmax; s->stats.entry_cells = (uint32_t)total;
    return PT_OK;
}

// Stable LSD radix sort of the (Morton code, leaf slot) pairs in codes / ids (ids = 0..n-1), 8 bits per pass (pt_lbvh.cuh); the result
// lands in codes_sorted / ids_sorted.  All four buffers must hold n entries.
int sort_code_pairs(pt_scene* s, int n) {
    cudaStream_t st = s->stream;
    if (n <= PT_SORT_SMALL) {
        k_sort_small<<<1, 256, 0, st>>>(n, s->codes.p, s->codes_sorted.p, s->ids_sorted.p);
        s->stats.launches += 1;
    } else {
        const int tiles = (n + PT_SORT_TILE - 1) / PT_SORT_TILE;
        CK(s->sort_hist.reserve((size_t)256 * tiles));
        uint32_t* kin = s->codes.p; uint32_t* vin = s->ids.p; uint32_t* kout = s->codes_sorted.p; uint32_t* vout = s->ids_sorted.p;
        for (int shift = 0; shift < 32; shift += 8) {
            k_sort_hist<<<tiles, 256, 0, st>>>(n, shift, kin, s->sort_hist.p);
            k_sort_scan<<<1, 1024, 0, st>>>(256 * tiles, s->sort_hist.p);
            k_sort_scatter<<<tiles, 256, 0, st>>>(n, shift, kin, vin, s->sort_hist.p, kout, vout);
            std::swap(kin, kout); std::swap(vin, vout);
        }
        // four passes: the pairs are back in the buffers they started from -- which now ARE the sorted ones
        std::swap(s->codes, s->codes_sorted); std::swap(s->ids, s->ids_sorted);
        s->stats.launches += 12;
    }
    CK(cudaGetLastError());
    return PT_OK;
}

int upload_and_build(pt_scene* s) {
    const int n = (int)s->prims.size();
    cudaStream_t st = s->stream;
    // device order = the reference's leaf visiting order, so that exact ties resolve by index (pt_wavefront.cuh)
    std::vector<uint32_t> rank(std::max(n, 1));
    std::vector<float4> tb_mn, tb_mx;      // PT_OPT_REF_LEAF_BOXES: per device primitive, the reference parent's box
    // The reference sorts its objects ONCE, in BuildBVH; SceneBVH::UpdateObject only refits the boxes of the tree it has
    // (bvh.cu:122-157), so its leaf visiting order -- all that the tie rule needs -- survives every object update.
    // pt_update_prim therefore keeps the order of the last build that followed an add; only that build pays for the
    // host-side restatement of the reference's sort (35 ms at 128 k primitives, 43x the device LBVH).
    if (s->opt_tie_order && !s->opt_ref_leaf_boxes && n >= 2 && s->rank_valid && (int)s->rank_cache.size() == n) {
        rank = s->rank_cache;
    } else if ((s->opt_tie_order || s->opt_ref_leaf_boxes) && n >= 2) {
        RefOrder ro(s->prims);
        rank = ro.ranks();
        s->rank_cache = rank; s->rank_valid = true;
        if (!s->opt_tie_order) for (int i = 0; i < n; ++i) rank[i] = (uint32_t)i;
        if (s->opt_ref_leaf_boxes) {
            tb_mn.resize(n); tb_mx.resize(n);
            for (int i = 0; i < n; ++i) {
                tb_mn[rank[i]] = make_float4(ro.pmn[3 * i], ro.pmn[3 * i + 1], ro.pmn[3 * i + 2], 0.f);
                tb_mx[rank[i]] = make_float4(ro.pmx[3 * i], ro.pmx[3 * i + 1], ro.pmx[3 * i + 2], 0.f);
            }
        }
    } else for (int i = 0; i < n; ++i) rank[i] = (uint32_t)i;
    const size_t n1 = (size_t)std::max(n, 1);
    CK(s->st_g0.reserve(n1)); CK(s->st_g1.reserve(n1)); CK(s->st_orig.reserve(n1)); CK(s->st_tree.reserve(n1));
    CK(cudaStreamSynchronize(st));          // the staging buffers may still feed the previous build's copies
    float4* const g0 = s->st_g0.p; float4* const g1 = s->st_g1.p; int* const orig = s->st_orig.p;
    int* const tree = s->st_tree.p;
    auto stage_prim = [&](int i) {
        const HostPrim& p = s->prims[i];
        const pt_material& m = s->materials[p.mat_slot];
        if (m.emit_intensity != 0.f) s->has_emitters = 1;
        uint32_t w = ((uint32_t)bin_of_material_type(m.type) << 28) | (uint32_t)p.mat_slot;
        float wf, tf; int ty = p.type;
        memcpy(&wf, &w, 4); memcpy(&tf, &ty, 4);
        const int d = (int)rank[i];
        g0[d] = make_float4(p.c[0], p.c[1], p.c[2], p.radius);
        g1[d] = make_float4(p.height, p.y_pos, tf, wf);
        orig[d] = i;
    };
    // After updates only (pt_update_prim: no add, no material change, same order) the staging buffers and the index
    // lists on the device are those of the last build: restage the moved primitives, re-upload the two geometry arrays.
    const bool moved_only = s->layout_valid && (int)s->st_n == n;
    if (moved_only) {
        for (int i : s->moved) stage_prim(i);
        for (size_t k = 0; k < s->plat_ids_host.size(); ++k) s->plat_y_host[k] = s->prims[orig[s->plat_ids_host[k]]].y_pos;
    } else {
        s->has_emitters = 0;
        for (int i = 0; i < n; ++i) stage_prim(i);
        s->plat_y_host.clear(); s->plat_ids_host.clear();
        int n_tree_prims = 0;
        for (int d = 0; d < n; ++d) {
            const HostPrim& p = s->prims[orig[d]];
            if (p.type == PT_PRIM_PLATFORM) { s->plat_y_host.push_back(p.y_pos); s->plat_ids_host.push_back(d); }
            else if (p.type == PT_PRIM_SPHERE || p.type == PT_PRIM_CYLINDER) tree[n_tree_prims++] = d;
        }
        s->n_tree = n_tree_prims;
        s->st_n = (size_t)n;
    }
    s->moved.clear();
    const std::vector<float>& py = s->plat_y_host; const std::vector<int>& pid = s->plat_ids_host;

    CK(s->geo0.reserve(n1)); CK(s->geo1.reserve(n1));
    CK(s->orig_id.reserve(n1));
    CK(s->plat_y.reserve(std::max<size_t>(py.size(), 1))); CK(s->plat_id.reserve(std::max<size_t>(pid.size(), 1)));
    CK(cudaMemcpyAsync(s->geo0.p, g0, n1 * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->geo1.p, g1, n1 * sizeof(float4), cudaMemcpyHostToDevice, st));
    if (!py.empty()) CK(cudaMemcpyAsync(s->plat_y.p, py.data(), py.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    std::vector<MaterialRec> recs;
    if (!moved_only) {
        recs.resize(std::max<size_t>(s->materials.size(), 1));
        for (size_t i = 0; i < s->materials.size(); ++i) recs[i] = make_material_rec(s->materials[i]);
        CK(s->mats.reserve(recs.size()));
        CK(cudaMemcpyAsync(s->mats.p, recs.data(), recs.size() * sizeof(MaterialRec), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->orig_id.p, orig, n1 * sizeof(int), cudaMemcpyHostToDevice, st));
        if (!pid.empty()) CK(cudaMemcpyAsync(s->plat_id.p, pid.data(), pid.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    const bool use_tb = !tb_mn.empty();
    if (use_tb) {
        CK(s->test_mn.reserve(n)); CK(s->test_mx.reserve(n));
        CK(cudaMemcpyAsync(s->test_mn.p, tb_mn.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->test_mx.p, tb_mx.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
    }
    s->n_plat = (int)py.size();
    s->grid.lmax = -1; s->grid.cells = nullptr; s->grid.levels = nullptr;
    s->stats.entry_levels = 0; s->stats.entry_cells = 0;
    const int nt = s->n_tree;
    cudaEvent_t eb0, eb1;
    CK(cudaEventCreate(&eb0)); CK(cudaEventCreate(&eb1));
    CK(cudaEventRecord(eb0, st));
    if (nt > 0) {
        CK(s->bvh_prims.reserve(nt));
        if (!moved_only) CK(cudaMemcpyAsync(s->bvh_prims.p, tree, nt * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(s->nodes.reserve(4 * (size_t)std::max(nt - 1, 1)));
        if (nt == 1) {
            k_lbvh_single<<<1, 1, 0, st>>>(s->bvh_prims.p, s->geo0.p, s->geo1.p, s->nodes.p);
            s->stats.launches += 1;
#if PT_BVH4
            CK(s->nodes4.reserve(8));
            k_lbvh_collapse4<<<1, 32, 0, st>>>(2, s->nodes.p, s->nodes4.p);
            s->stats.launches += 1;
#endif
        } else {
            CK(s->leaf_mn.reserve(nt)); CK(s->leaf_mx.reserve(nt)); CK(s->leaf_c.reserve(nt)); CK(s->inner_mn.reserve(nt)); CK(s->inner_mx.reserve(nt));
            CK(s->codes.reserve(nt)); CK(s->codes_sorted.reserve(nt)); CK(s->ids.reserve(nt)); CK(s->ids_sorted.reserve(nt));
            CK(s->bounds6.reserve(8)); CK(s->children.reserve(nt)); CK(s->parent_inner.reserve(nt));
            CK(s->parent_leaf.reserve(nt)); CK(s->flags.reserve(nt));
            const uint32_t init6[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
            CK(cudaMemcpyAsync(s->bounds6.p, init6, sizeof(init6), cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(s->flags.p, 0, nt * sizeof(int), st));
            const int tb = 256, gb = (nt + tb - 1) / tb;
            k_lbvh_bounds<<<gb, tb, 0, st>>>(nt, s->bvh_prims.p, s->geo0.p, s->geo1.p, use_tb ? s->test_mn.p : nullptr,
                                             use_tb ? s->test_mx.p : nullptr, s->leaf_mn.p, s->leaf_mx.p, s->leaf_c.p, s->bounds6.p);
            k_lbvh_morton<<<gb, tb, 0, st>>>(nt, s->leaf_c.p, s->bounds6.p, s->codes.p, s->ids.p);
            { int rc = sort_code_pairs(s, nt); if (rc) return rc; }
            k_lbvh_hierarchy<<<gb, tb, 0, st>>>(nt, s->codes_sorted.p, s->children.p, s->parent_inner.p, s->parent_leaf.p);
            CK(s->node_height.reserve(nt));
            LbvhTree lt;
            lt.sorted_ids = s->ids_sorted.p; lt.leaf_mn = s->leaf_mn.p; lt.leaf_mx = s->leaf_mx.p;
            lt.children = s->children.p; lt.parent_of_inner = s->parent_inner.p; lt.parent_of_leaf = s->parent_leaf.p;
            lt.inner_mn = s->inner_mn.p; lt.inner_mx = s->inner_mx.p; lt.height = s->node_height.p;
            // pass 0 fits the boxes of the Karras tree; every further pass is a SAH refit (tree rotations + fit)
            for (int pass = 0; pass <= s->opt_sah_passes; ++pass) {
                if (pass > 0) CK(cudaMemsetAsync(s->flags.p, 0, nt * sizeof(int), st));
                k_lbvh_fit<<<gb, tb, 0, st>>>(nt, lt, s->flags.p, pass > 0 ? (s->opt_sah_treelets ? 2 : 1) : 0);
                if (pass == 0 || pass == s->opt_sah_passes) {
                    CK(s->sah_dev.reserve(2));
                    if (pass == 0) CK(cudaMemsetAsync(s->sah_dev.p, 0, 2 * sizeof(float), st));
                    k_lbvh_sah<<<std::min(gb, 256), tb, 0, st>>>(nt, s->inner_mn.p, s->inner_mx.p, s->sah_dev.p + (pass == 0 ? 0 : 1));
                }
            }
            CK(cudaMemcpyAsync(s->sah_host, s->sah_dev.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
            s->stats.launches += (uint64_t)s->opt_sah_passes + 2;
            k_lbvh_pack<<<gb, tb, 0, st>>>(nt, s->ids_sorted.p, s->bvh_prims.p, s->geo1.p, s->leaf_mn.p, s->leaf_mx.p,
                                           s->children.p, s->inner_mn.p, s->inner_mx.p, s->nodes.p);
            s->stats.launches += 5;
#if PT_BVH4
            CK(s->nodes4.reserve(8 * (size_t)std::max(nt - 1, 1)));
            k_lbvh_collapse4<<<gb, tb, 0, st>>>(nt, s->nodes.p, s->nodes4.p);
            s->stats.launches += 1;
#endif
            int rc = build_entry_grid(s, nt);
            if (rc) return rc;
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(eb1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&s->stats.build_ms, eb0, eb1));
    cudaEventDestroy(eb0); cudaEventDestroy(eb1);
    s->stats.bvh_nodes = nt > 1 ? (uint32_t)(nt - 1) : (uint32_t)nt;
    s->stats.bvh_sah_lbvh = nt > 1 ? s->sah_host[0] : 0.f;
    s->stats.bvh_sah = nt > 1 ? (s->opt_sah_passes > 0 ? s->sah_host[1] : s->sah_host[0]) : 0.f;
    s->dirty = false; s->built = true;
    s->layout_valid = !use_tb;          // (PT_OPT_REF_LEAF_BOXES recomputes everything on every build)
    return PT_OK;
}

SceneDev scene_dev(pt_scene* s) {
    SceneDev d;
    d.grid = s->grid;
    d.nodes = s->nodes.p; d.nodes4 = s->nodes4.p; d.geo0 = s->geo0.p; d.geo1 = s->geo1.p; d.mats = s->mats.p;
    d.plat_y = s->plat_y.p; d.plat_id = s->plat_id.p; d.orig_id = s->orig_id.p;
    d.n_plat = s->n_plat; d.n_tree_prims = s->n_tree; d.sky = s->sky.obj; d.has_emitters = s->has_emitters;
    return d;
}

int ensure_sky(pt_scene* s) {
    if (s->sky.obj) return PT_OK;
    const uint8_t black[4] = {0, 0, 0, 255};
    return pt_set_sky_rgba8(s, black, 1, 1);
}

int ensure_frame(pt_scene* s, int w, int h) {
    if (w <= 0 || h <= 0 || (uint64_t)w * h > (1ull << 27)) return fail(PT_ERR_INVALID, "bad frame size");
    if (w != s->width || h != s->height) {
        s->accum.release();
        CK(s->accum.reserve((size_t)w * h));
        CK(cudaMemsetAsync(s->accum.p, 0, (size_t)w * h * sizeof(float4), s->stream));
        s->width = w; s->height = h;
    }
    return PT_OK;
}

cudaError_t try_reserve_wave(pt_scene* s, uint32_t cap) {
    if (cap <= s->capacity) return cudaSuccess;
    // released first so that growing never holds two copies
    s->path_out.release();
    for (int k = 0; k < 2; ++k) { s->hq_p[k].release(); s->hq_d[k].release(); s->hq_thr[k].release(); }
    s->capacity = 0;
    cudaError_t e = s->path_out.reserve(cap);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        // two regions of `cap` records, each shared by two shader bins growing from its ends (HitQueues)
        e = s->hq_p[k].reserve(2 * (size_t)cap);
        if (e == cudaSuccess) e = s->hq_d[k].reserve(2 * (size_t)cap);
        if (e == cudaSuccess) e = s->hq_thr[k].reserve(2 * (size_t)cap);
    }
    if (e == cudaSuccess) { s->capacity = cap; return e; }
    s->path_out.release();
    for (int k = 0; k < 2; ++k) { s->hq_p[k].release(); s->hq_d[k].release(); s->hq_thr[k].release(); }
    cudaGetLastError();   // clear the sticky allocation error before the caller retries with a smaller wave
    return e;
}

// Sizes the wave: as many samples per pixel as fit PT_OPT_WAVE_PATHS, halved until the hit queues fit in free
// device memory (the default 128 Mi-path wave reserves 26 GB; another tenant of the GPU must not make the frame fail).
int ensure_queues(pt_scene* s, uint32_t npix, uint32_t& spp_per_wave, int max_depth) {
    while (true) {
        cudaError_t e = try_reserve_wave(s, npix * spp_per_wave);
        if (e == cudaSuccess) break;
        if (e != cudaErrorMemoryAllocation || spp_per_wave == 1)
            return fail(PT_ERR_CUDA, std::string("wave queues: ") + cudaGetErrorString(e));
        spp_per_wave = (spp_per_wave + 1) / 2;
    }
    CK(s->ctrl.reserve((size_t)max_depth + 3));
    CK(s->seg_counter.reserve(4));
    return PT_OK;
}

HitQueues hit_queues(pt_scene* s, int k) {
    HitQueues hq; hq.p = s->hq_p[k].p; hq.d = s->hq_d[k].p; hq.thr = s->hq_thr[k].p; hq.capacity = s->capacity;
    return hq;
}

int launch_config(pt_scene* s) {
    if (s->grid_bounce) return PT_OK;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bounce<SRC_HITS, false>, 32 * PT_BOUNCE_WARPS, PT_BOUNCE_SMEM));
    s->grid_bounce = s->sm_count * std::max(occ, 1);
    return PT_OK;
}

template <int SRC>
void launch_bounce(pt_scene* s, const BounceArgs& a, cudaStream_t st) {
    if (s->opt_count_steps) k_bounce<SRC, true><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, PT_BOUNCE_SMEM, st>>>(a);
    else k_bounce<SRC, false><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, PT_BOUNCE_SMEM, st>>>(a);
}

int render_async(pt_scene* s, const pt_camera* cam, int spp, int max_depth, uint64_t seed, uint32_t first_sample,
                 bool accumulate = true, float4* gbuf_normal = nullptr) {
    if (!s || !cam) return fail(PT_ERR_INVALID, "null argument");
    if (!s->built || s->dirty) return fail(PT_ERR_NOT_BUILT, "pt_build must be called after the scene changed");
    if (spp <= 0 || max_depth <= 0 || max_depth > 64) return fail(PT_ERR_INVALID, "bad spp / max_depth");
    int rc = ensure_frame(s, cam->width, cam->height); if (rc) return rc;
    rc = ensure_sky(s); if (rc) return rc;
    rc = launch_config(s); if (rc) return rc;
    const uint32_t npix = (uint32_t)cam->width * (uint32_t)cam->height;
    uint32_t spp_per_wave = std::max<uint32_t>(1u, s->opt_wave_paths / npix);
    spp_per_wave = std::min<uint32_t>(spp_per_wave, (uint32_t)spp);
    rc = ensure_queues(s, npix, spp_per_wave, max_depth); if (rc) return rc;
    const uint32_t cap = npix * spp_per_wave;
    cudaStream_t st = s->stream;
#if PT_ENTRY_GRID_CODE
    const bool split = s->opt_split_walk && s->grid.lmax >= 1;
    if (split) {
        if (s->rq_capacity < s->capacity) {
            s->rq_p.release(); s->rq_d.release(); s->rq_thr.release();
            CK(s->rq_p.reserve(2 * (size_t)s->capacity)); CK(s->rq_d.reserve(2 * (size_t)s->capacity)); CK(s->rq_thr.reserve(2 * (size_t)s->capacity));
            s->rq_capacity = s->capacity;
        }
        CK(s->rctrl.reserve((size_t)max_depth + 2));
        if (!s->grid_walk) {
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_walk<false>, 32 * PT_BOUNCE_WARPS, 0));
            s->grid_walk = s->sm_count * std::max(occ, 1);
        }
    }
#endif
    const CameraFrame frame = to_device_frame(cam);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    BounceArgs args;
    memset(&args, 0, sizeof(args));
    args.sc = scene_dev(s); args.key = key; args.npix = npix; args.max_depth = max_depth; args.inner_run = s->opt_inner_run;
    args.cam_tiled = (cam->width % 8 == 0 && cam->height % 4 == 0) ? 1 : 0;
    args.cam = frame; args.path_out = s->path_out.p; args.gbuf_normal = gbuf_normal; args.step_counters = s->seg_counter.p + 1;
    const bool timing = s->opt_stage_timing != 0;
    size_t ev_i = 0;
    auto stage_event = [&]() -> cudaError_t {
        if (!timing) return cudaSuccess;
        if (ev_i >= s->stage_events.size()) { cudaEvent_t e; cudaError_t r = cudaEventCreate(&e); if (r) return r; s->stage_events.push_back(e); }
        return cudaEventRecord(s->stage_events[ev_i++], st);
    };
    CK(cudaMemsetAsync(s->seg_counter.p, 0, 4 * sizeof(unsigned long long), st));
    CK(cudaEventRecord(s->ev0, st));
    uint64_t launches = 0;
    BounceCtrl* const ctrl = s->ctrl.p;
    BounceCtrl* const ctrl_cam = s->ctrl.p + max_depth + 1;     // work counter of the camera launch
    for (uint32_t s0 = 0; s0 < (uint32_t)spp; s0 += spp_per_wave) {
        const uint32_t ns = std::min<uint32_t>(spp_per_wave, (uint32_t)spp - s0);
        const uint32_t n_paths = npix * ns;
        CK(cudaMemsetAsync(s->ctrl.p, 0, ((size_t)max_depth + 2) * sizeof(BounceCtrl), st));
#if PT_ENTRY_GRID_CODE
        if (split) CK(cudaMemsetAsync(s->rctrl.p, 0, ((size_t)max_depth + 2) * sizeof(BounceCtrl), st));
#endif
        args.first_sample = first_sample + s0; args.n_paths = n_paths;
        // segment 0: camera rays -> hit queue 0
        CK(stage_event());
        // without emitters the hits of the last allowed segment end their paths black: the launch that traces that
        // segment writes the zeros itself and the shade-only launch is skipped (the display pipeline's G-buffer wants
        // the first-segment normals, so a depth-1 dispatch keeps it)
        const bool drop_last = !s->has_emitters && !(gbuf_normal != nullptr && max_depth == 1);
        args.segment = 0; args.trace = 1; args.ctrl_in = ctrl_cam; args.ctrl_out = ctrl + 0; args.hq_out = hit_queues(s, 0);
        args.drop_last_hits = (drop_last && max_depth == 1) ? 1 : 0;
        launch_bounce<SRC_CAMERA>(s, args, st);
        // segment k: shade the hits of segment k-1, trace the survivors; k == max_depth only shades
        for (int k = 1; k <= max_depth; ++k) {
            CK(stage_event());
            if (k == max_depth && drop_last) continue;
            args.drop_last_hits = (drop_last && k == max_depth - 1) ? 1 : 0;
            args.segment = k; args.trace = k < max_depth ? 1 : 0;
            args.ctrl_in = ctrl + (k - 1); args.ctrl_out = ctrl + k;
            args.hq_in = hit_queues(s, (k - 1) & 1); args.hq_out = hit_queues(s, k & 1);
#if PT_ENTRY_GRID_CODE
            if (split && args.trace) {
                // two launches: shade + classify by entry-grid level, then walk one level class per chunk
                args.rctrl = s->rctrl.p + k;
                args.rq.p = s->rq_p.p; args.rq.d = s->rq_d.p; args.rq.thr = s->rq_thr.p; args.rq.capacity = s->rq_capacity;
                if (s->opt_count_steps) {
                    k_shade_classify<true><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, 0, st>>>(args);
                    k_walk<true><<<s->grid_walk, 32 * PT_BOUNCE_WARPS, 0, st>>>(args);
                } else {
                    k_shade_classify<false><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, 0, st>>>(args);
                    k_walk<false><<<s->grid_walk, 32 * PT_BOUNCE_WARPS, 0, st>>>(args);
                }
                ++launches;
                continue;
            }
#endif
            launch_bounce<SRC_HITS>(s, args, st);
        }
        CK(stage_event());
        launches += (uint64_t)max_depth + (drop_last ? 0 : 1);
        const int ac_grid = (int)std::min<uint32_t>((npix + 255u) / 256u, (uint32_t)s->sm_count * 8u);
        if (accumulate) k_accumulate<<<ac_grid, 256, 0, st>>>(npix, ns, s->path_out.p, s->accum.p);
        k_wave_stats<<<1, 1, 0, st>>>(max_depth, s->ctrl.p, s->seg_counter.p);
        launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s->ev1, st));
    s->frame_recorded = true; s->last_depth = max_depth; s->stage_events_used = ev_i;
    s->stats.launches += launches;
    s->stats.paths = (uint64_t)npix * (uint64_t)spp;
    s->stats.wave_paths = cap;
    s->stats.segments = 0;
    s->stats.waves = (uint32_t)(((uint32_t)spp + spp_per_wave - 1) / spp_per_wave);
    s->stats.traverse_launches = s->stats.waves * (uint32_t)max_depth;   // launches that trace (camera + max_depth-1 bounce)
    return PT_OK;
}

int finish_frame(pt_scene* s, int max_depth_hint) {
    (void)max_depth_hint;
    if (!s->frame_recorded) { CK(cudaStreamSynchronize(s->stream)); return PT_OK; }
    unsigned long long seg4[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(seg4, s->seg_counter.p, sizeof(seg4), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaEventElapsedTime(&s->stats.render_ms, s->ev0, s->ev1));
    s->stats.segments = seg4[0]; s->stats.node_steps = seg4[1]; s->stats.prim_tests = seg4[2]; s->stats.hit_records = seg4[3];
    if (s->opt_stage_timing && s->stage_events.size() >= 2) {
        // per wave: one event before each of the max_depth+1 k_bounce launches and one after the last.
        // traverse_ms = launches that trace (segments 0 .. max_depth-1; they also shade the previous segment's hits),
        // shade_ms = the final shade-only launch
        float trace_ms = 0.f, tail_ms = 0.f;
        const size_t per_wave = (size_t)s->last_depth + 2;
        for (size_t w0 = 0; w0 + per_wave <= s->stage_events_used; w0 += per_wave)
            for (int k = 0; k <= s->last_depth; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, s->stage_events[w0 + k], s->stage_events[w0 + k + 1]);
                if (k < s->last_depth) trace_ms += ms; else tail_ms += ms;
            }
        s->stats.traverse_ms = trace_ms; s->stats.shade_ms = tail_ms;
    }
    return PT_OK;
}

template <class T>
int to_device(DevBuf<T>& buf, const T* host, size_t n, cudaStream_t st) {
    CK(buf.reserve(std::max<size_t>(n, 1)));
    if (n) CK(cudaMemcpyAsync(buf.p, host, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return PT_OK;
}

// ---- small kernels behind the parity entry points (same device functions as the render stages) ----
__global__ void k_pack_rays(int n, const float* __restrict__ rays8, float4* __restrict__ ray_o, float4* __restrict__ ray_d,
                            float2* __restrict__ range, BounceCtrl* __restrict__ ctrl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { ctrl->in_count.v = (uint32_t)n; ctrl->rays.v = (uint32_t)n; }
    if (i >= n) return;
    ray_o[i] = make_float4(rays8[8 * i], rays8[8 * i + 1], rays8[8 * i + 2], __uint_as_float((uint32_t)i));
    ray_d[i] = make_float4(rays8[8 * i + 3], rays8[8 * i + 4], rays8[8 * i + 5], 0.f);
    range[i] = make_float2(rays8[8 * i + 6], rays8[8 * i + 7]);
}
__global__ void k_unpack_hits(int n, SceneDev sc, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                              const uint2* __restrict__ hit, int32_t* __restrict__ prim, float* __restrict__ t,
                              float* __restrict__ normal3, float* __restrict__ pos3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 h = hit[i];
    float3 pos = f3(0, 0, 0), nrm = f3(0, 0, 0);
    if (h.y == PT_MISS_CODE) prim[i] = -1;
    else {
        prim[i] = (int32_t)sc.orig_id[h.y >> 3];
        float4 o = ray_o[i], d = ray_d[i];
        hit_attributes(sc.geo0[h.y >> 3], (int)(h.y & 7u), f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), __uint_as_float(h.x), pos, nrm);
    }
    t[i] = __uint_as_float(h.x);
    normal3[3 * i] = nrm.x; normal3[3 * i + 1] = nrm.y; normal3[3 * i + 2] = nrm.z;
    pos3[3 * i] = pos.x; pos3[3 * i + 1] = pos.y; pos3[3 * i + 2] = pos.z;
}
__global__ void k_shade_batch(int n, const MaterialRec* __restrict__ mats, const int32_t* __restrict__ mat_idx,
                              const float* __restrict__ nrm3, const float* __restrict__ in3, const float* __restrict__ u3,
                              float* __restrict__ bounce3, float* __restrict__ atten3, float* __restrict__ rad3,
                              int32_t* __restrict__ ndraws) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MaterialRec m = mats[mat_idx[i]];
    Scatter s;
    shade_bin(bin_of_material_type(m.type), m, f3(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2]),
              f3(in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]), make_float4(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2], 0.f), s);
    bounce3[3 * i] = s.bounce.x; bounce3[3 * i + 1] = s.bounce.y; bounce3[3 * i + 2] = s.bounce.z;
    atten3[3 * i] = s.atten.x; atten3[3 * i + 1] = s.atten.y; atten3[3 * i + 2] = s.atten.z;
    rad3[3 * i] = s.radiance.x; rad3[3 * i + 1] = s.radiance.y; rad3[3 * i + 2] = s.radiance.z;
    ndraws[i] = s.draws;
}
__global__ void k_raygen_batch(int n, CameraFrame cam, const int32_t* __restrict__ xy, const float* __restrict__ u3,
                               float* __restrict__ rays8) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 o, d;
    camera_ray(cam, xy[2 * i], xy[2 * i + 1], make_float4(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2], 0.f), o, d);
    rays8[8 * i] = o.x; rays8[8 * i + 1] = o.y; rays8[8 * i + 2] = o.z;
    rays8[8 * i + 3] = d.x; rays8[8 * i + 4] = d.y; rays8[8 * i + 5] = d.z;
    rays8[8 * i + 6] = 0.f; rays8[8 * i + 7] = PT_TMAX;
}
__global__ void k_sky_batch(int n, cudaTextureObject_t sky, const float* __restrict__ dir3, float* __restrict__ rgb3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = sky_radiance(sky, f3(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]));
    rgb3[3 * i] = c.x; rgb3[3 * i + 1] = c.y; rgb3[3 * i + 2] = c.z;
}
__global__ void k_philox_batch(int n, uint2 key, const uint32_t* __restrict__ pss, float* __restrict__ u4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 u = philox_uniform4(key, pss[3 * i], pss[3 * i + 1], pss[3 * i + 2]);
    u4[4 * i] = u.x; u4[4 * i + 1] = u.y; u4[4 * i + 2] = u.z; u4[4 * i + 3] = u.w;
}
__global__ void k_read_accum(uint32_t npix, const float4* __restrict__ accum, float4* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float4 a = accum[p];
    float n = a.w > 0.f ? a.w : 1.f;
    out[p] = make_float4(a.x / n, a.y / n, a.z / n, a.w);
}

int create_texture(const uint8_t* rgba, int w, int h, size_t row_bytes, Texture& out, cudaStream_t st) {
    if (!rgba || w <= 0 || h <= 0) return fail(PT_ERR_INVALID, "bad texture");
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    Texture t;
    CK(cudaMallocArray(&t.arr, &desc, w, h));
    // the array is first cleared so that texels the reference leaves uninitialised (textures.cu:32) read as 0
    std::vector<uint8_t> staged((size_t)w * h * 4, 0);
    for (int y = 0; y < h; ++y) memcpy(&staged[(size_t)y * w * 4], rgba + (size_t)y * w * 4, row_bytes);
    CK(cudaMemcpy2DToArrayAsync(t.arr, 0, 0, staged.data(), (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray; rd.res.array.array = t.arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));       // textures.cu:46-50
    td.addressMode[0] = cudaAddressModeMirror; td.addressMode[1] = cudaAddressModeMirror;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    CK(cudaCreateTextureObject(&t.obj, &rd, &td, nullptr));
    out = t;
    return PT_OK;
}
void destroy_texture(Texture& t) {
    if (t.obj) cudaDestroyTextureObject(t.obj);
    if (t.arr) cudaFreeArray(t.arr);
    t.obj = 0; t.arr = nullptr;
}

// Frees everything the display path holds on the host side or on the copy stream, and leaves every handle null so
// that the next dispatch allocates afresh.  Called by pt_scene_destroy (and only there: the buffers handed to the
// caller stay valid for the scene's lifetime unless a dispatch replaces them).
void release_display_host(pt_scene* s) {
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); s->copy_stream = nullptr; }
    for (int k = 0; k < 2; ++k) {
        if (s->pipe_host[k]) cudaFreeHost(s->pipe_host[k]);
        s->pipe_host[k] = nullptr;
        s->pipe_bgra[k].release();
        if (s->pipe_filtered[k]) cudaEventDestroy(s->pipe_filtered[k]);
        if (s->pipe_copied[k]) cudaEventDestroy(s->pipe_copied[k]);
        s->pipe_filtered[k] = s->pipe_copied[k] = nullptr;
    }
    s->pipe_bytes = 0; s->pipe_begun = s->pipe_ended = 0;
    for (int k = 0; k < 2; ++k) {
        s->read_pipe[k].release();
        if (s->read_ready[k]) cudaEventDestroy(s->read_ready[k]);
        if (s->read_done[k]) cudaEventDestroy(s->read_done[k]);
        s->read_ready[k] = s->read_done[k] = nullptr;
    }
    s->read_begun = s->read_ended = 0;
    if (s->disp_host) cudaFreeHost(s->disp_host);
    s->disp_host = nullptr; s->disp_host_bytes = 0;
}

// One iteration of the reference's PipelineLoop (cuSrc/path_tracer.cu:256-306) minus the host thread:
// one sample per pixel -> Denoising -> Mix -> BGRA8 -> D2H.  sample_idx is the reference's 1-based
// cur_sample_idx_ (motional_camera.cu:195): 1 restarts the running mean, like MotionalCamera::Refresh.
// trace one sample and run the fused filter + mix + pack into `bgra_dev` (device), all on the scene's stream
int display_sample(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx, uint8_t* bgra_dev) {
    const size_t npix = (size_t)cam->width * cam->height;
    if (cam->width != s->disp_w || cam->height != s->disp_h) {
        s->disp_mix.release(); s->disp_normal.release(); s->disp_bgra.release();
        CK(s->disp_mix.reserve(npix)); CK(s->disp_normal.reserve(npix)); CK(s->disp_bgra.reserve(npix * 4));
        CK(cudaMemsetAsync(s->disp_mix.p, 0, npix * sizeof(float4), s->stream));
        s->disp_w = cam->width; s->disp_h = cam->height;
    }
    int rc = render_async(s, cam, 1, max_depth, seed, sample_idx - 1, /*accumulate=*/false, s->disp_normal.p);
    if (rc) return rc;
    dim3 tb(32, 8), gb((cam->width + 31) / 32, (cam->height + 7) / 8);
    k_denoise_mix<<<gb, tb, 0, s->stream>>>(cam->width, cam->height, s->path_out.p, s->disp_normal.p, s->disp_mix.p,
                                            bgra_dev ? bgra_dev : s->disp_bgra.p, (int)sample_idx);
    s->stats.launches += 1;
    CK(cudaGetLastError());
    return PT_OK;
}

int display_dispatch(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx,
                     uint8_t* bgra_out, size_t nbytes) {
    if (!s || !cam) return fail(PT_ERR_INVALID, "null argument");
    if (sample_idx < 1) return fail(PT_ERR_INVALID, "sample_idx is 1-based");
    const size_t npix = (size_t)cam->width * cam->height;
    if (bgra_out != nullptr && nbytes != npix * 4) return fail(PT_ERR_INVALID, "nbytes must be 4*width*height");
    int rc = display_sample(s, cam, max_depth, seed, sample_idx, nullptr);
    if (rc) return rc;
    if (bgra_out == nullptr) {
        // library-owned pinned frame (the reference's output_buffer_, path_tracer.cu:303-304): one DMA, no staging
        if (s->disp_host_bytes < npix * 4) {
            // the previous frame may still be the target of an in-flight copy only if the caller broke the contract
            // (the buffer is valid until the next dispatch); the stream is idle here (finish_frame synchronised it)
            if (s->disp_host) cudaFreeHost(s->disp_host);
            s->disp_host = nullptr; s->disp_host_bytes = 0;
            CK(cudaHostAlloc((void**)&s->disp_host, npix * 4, cudaHostAllocDefault));
            s->disp_host_bytes = npix * 4;
        }
        bgra_out = s->disp_host;
    }
    CK(cudaMemcpyAsync(bgra_out, s->disp_bgra.p, npix * 4, cudaMemcpyDeviceToHost, s->stream));
    return finish_frame(s, max_depth);
}

// Pipelined dispatch.  The reference's DispatchRay is asynchronous (a task queue drained by a render thread that
// calls back, path_tracer.cu:256-319); here the asynchrony is two calls: _begin enqueues sample + filter + mix on the
// scene's stream and the D2H copy of the frame on a copy stream, _end waits for the OLDEST frame begun and hands out
// its page-locked buffer.  With one frame always in flight (begin(i+1) before end(i)) the copy and the host's
// callback run under the trace of the next sample.
int display_begin(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx) {
    if (!s || !cam) return fail(PT_ERR_INVALID, "null argument");
    if (sample_idx < 1) return fail(PT_ERR_INVALID, "sample_idx is 1-based");
    if (s->pipe_begun - s->pipe_ended >= 2) return fail(PT_ERR_INVALID, "two dispatches are already in flight: call pt_dispatch_ray_end first");
    const size_t bytes = (size_t)cam->width * cam->height * 4;
    if (s->pipe_bytes != bytes) {
        if (s->pipe_begun != s->pipe_ended) return fail(PT_ERR_INVALID, "frame size changed with a dispatch in flight");
        s->pipe_bytes = 0;          // a failure below must not leave the old size paired with freed buffers
        for (int k = 0; k < 2; ++k) {
            if (s->pipe_host[k]) cudaFreeHost(s->pipe_host[k]);
            s->pipe_host[k] = nullptr;
            CK(s->pipe_bgra[k].reserve(bytes));
            CK(cudaHostAlloc((void**)&s->pipe_host[k], bytes, cudaHostAllocDefault));
            if (!s->pipe_filtered[k]) CK(cudaEventCreateWithFlags(&s->pipe_filtered[k], cudaEventDisableTiming));
            if (!s->pipe_copied[k]) CK(cudaEventCreateWithFlags(&s->pipe_copied[k], cudaEventDisableTiming));
        }
        if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        s->pipe_bytes = bytes;
    }
    const int k = (int)(s->pipe_begun & 1);
    // the filter of this dispatch overwrites device frame k: its previous copy must have left the device
    if (s->pipe_begun >= 2) CK(cudaStreamWaitEvent(s->stream, s->pipe_copied[k], 0));
    int rc = display_sample(s, cam, max_depth, seed, sample_idx, s->pipe_bgra[k].p);
    if (rc) return rc;
    CK(cudaEventRecord(s->pipe_filtered[k], s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, s->pipe_filtered[k], 0));
    CK(cudaMemcpyAsync(s->pipe_host[k], s->pipe_bgra[k].p, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->pipe_copied[k], s->copy_stream));
    s->frame_recorded = false;       // no per-frame statistics in this mode (they need a stream synchronisation)
    ++s->pipe_begun;
    return PT_OK;
}
int display_end(pt_scene* s, const uint8_t** bgra, size_t* nbytes) {
    if (!s || !bgra) return fail(PT_ERR_INVALID, "null argument");
    if (s->pipe_begun == s->pipe_ended) return fail(PT_ERR_INVALID, "no dispatch in flight");
    const int k = (int)(s->pipe_ended & 1);
    CK(cudaEventSynchronize(s->pipe_copied[k]));
    *bgra = s->pipe_host[k];
    if (nbytes) *nbytes = s->pipe_bytes;
    ++s->pipe_ended;
    return PT_OK;
}

}  // namespace

extern "C" {

const char* pt_last_error(void) { return g_err.c_str(); }
int pt_set_error_text(const char* msg) { g_err = msg ? msg : ""; return 0; }   // for the other translation units of the library
const char* pt_version(void) {
#ifdef PT_VARIANT_FMA
    return "cpppathtracer_b200 0.2 (sm_100a wavefront, fma-contracted arithmetic)";
#elif PT_ENTRY_GRID_CODE
    return "cpppathtracer_b200 0.2 (sm_100a wavefront, entry grid)";
#else
    return "cpppathtracer_b200 0.2 (sm_100a wavefront)";
#endif
}

int pt_scene_create(pt_scene** out) {
    if (!out) return fail(PT_ERR_INVALID, "null out");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PT_ERR_NO_DEVICE, std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    pt_scene* s = new pt_scene();
    int sms = 0;
    e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e != cudaSuccess) {
        if (s->ev0) cudaEventDestroy(s->ev0);
        if (s->ev1) cudaEventDestroy(s->ev1);
        delete s;
        return fail(PT_ERR_CUDA, std::string("pt_scene_create: ") + cudaGetErrorString(e));
    }
    s->sm_count = sms;
    *out = s;
    return PT_OK;
}

void pt_scene_destroy(pt_scene* s) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return;
    cudaStreamSynchronize(s->stream);
    destroy_texture(s->sky);
    for (auto& kv : s->textures) destroy_texture(kv.second);
    s->geo0.release(); s->geo1.release(); s->nodes.release(); s->leaf_mn.release(); s->leaf_mx.release();
    s->leaf_c.release(); s->test_mn.release(); s->test_mx.release();
    s->inner_mn.release(); s->inner_mx.release(); s->mats.release(); s->plat_y.release(); s->plat_id.release();
    s->bvh_prims.release(); s->parent_inner.release(); s->parent_leaf.release(); s->flags.release(); s->children.release();
    s->codes.release(); s->codes_sorted.release(); s->ids.release(); s->ids_sorted.release(); s->bounds6.release();
    s->orig_id.release(); s->sort_hist.release(); s->accum.release(); s->read_tmp.release(); s->path_out.release(); s->hit.release();
    s->ctrl.release(); s->q_o.release(); s->q_d.release(); s->seg_counter.release();
    s->node_height.release(); s->sah_dev.release(); s->grid_cells.release(); s->grid_levels.release();
    s->st_g0.release(); s->st_g1.release(); s->st_orig.release(); s->st_tree.release(); s->nodes4.release();
    s->rq_p.release(); s->rq_d.release(); s->rq_thr.release(); s->rctrl.release();
    for (int k = 0; k < 2; ++k) { s->hq_p[k].release(); s->hq_d[k].release(); s->hq_thr[k].release(); s->scratch_i[k].release(); }
    for (auto& b : s->scratch_f) b.release();
    s->disp_color.release(); s->disp_normal.release(); s->disp_mix.release(); s->disp_depth.release(); s->disp_bgra.release();
    // display path: the scene-owned page-locked frames (the reference's output_buffer_, path_tracer.cu:303-304, lives as
    // long as its PathTracer) and the copy stream / events of the pipelined dispatch
    release_display_host(s);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (auto e : s->stage_events) cudaEventDestroy(e);
    delete s;
}

int pt_set_option(pt_scene* s, int option, int64_t value) {
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    switch (option) {
        case PT_OPT_WAVE_PATHS:
            if (value < 1024 || value > (1ll << 28)) return fail(PT_ERR_INVALID, "wave_paths out of range");
            s->opt_wave_paths = (uint32_t)value; return PT_OK;
        case PT_OPT_STAGE_TIMING: s->opt_stage_timing = value != 0; return PT_OK;
        case PT_OPT_TIE_ORDER: s->opt_tie_order = value != 0; s->dirty = true; s->rank_valid = false; s->layout_valid = false; return PT_OK;
        case PT_OPT_SKY_QUARTER_ROWS: s->opt_sky_quarter = value != 0; return PT_OK;
        case PT_OPT_COUNT_STEPS: s->opt_count_steps = value != 0; return PT_OK;
        case PT_OPT_REF_LEAF_BOXES: s->opt_ref_leaf_boxes = value != 0; s->dirty = true; s->rank_valid = false; s->layout_valid = false; return PT_OK;
        case PT_OPT_SAH_TREELETS: s->opt_sah_treelets = value != 0; s->dirty = true; return PT_OK;
        case PT_OPT_SAH_PASSES:
            if (value < 0 || value > 16) return fail(PT_ERR_INVALID, "PT_OPT_SAH_PASSES: 0 .. 16");
            s->opt_sah_passes = (int)value; s->dirty = true; return PT_OK;
        case PT_OPT_ENTRY_GRID:
            if (value < 0 || value > 20) return fail(PT_ERR_INVALID, "PT_OPT_ENTRY_GRID: 0 (off) .. 20 levels");
#if !PT_ENTRY_GRID_CODE
            if (value != 0) return fail(PT_ERR_INVALID, "this build carries no entry-grid code in its kernels: use libpt_b200_grid.so");
#endif
            s->opt_entry_grid = (int)value; s->dirty = true; return PT_OK;
        case PT_OPT_ENTRY_CELLS_PER_PRIM:
            if (value < 1 || value > 64) return fail(PT_ERR_INVALID, "PT_OPT_ENTRY_CELLS_PER_PRIM: 1 .. 64");
            s->opt_entry_cells_per_prim = (int)value; s->dirty = true; return PT_OK;
        case PT_OPT_SPLIT_WALK:
#if !PT_ENTRY_GRID_CODE
            if (value != 0) return fail(PT_ERR_INVALID, "this build carries no entry-grid code in its kernels: use libpt_b200_grid.so");
#endif
            s->opt_split_walk = value != 0; return PT_OK;
        case PT_OPT_TRAV_INNER_RUN:
            if (value < 1 || value > 1024) return fail(PT_ERR_INVALID, "inner_run must be 1..1024");
            s->opt_inner_run = (int)value; return PT_OK;
        default: return fail(PT_ERR_INVALID, "unknown option");
    }
}

int pt_set_stream(pt_scene* s, void* cuda_stream) {
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    s->stream = (cudaStream_t)cuda_stream;
    return PT_OK;
}

static int add_prim(pt_scene* s, int type, const float* c, float radius, float y_pos, float height) {
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    if (s->prims.size() >= (1u << 28)) return fail(PT_ERR_INVALID, "too many primitives");
    HostPrim p;
    p.type = type;
    p.c[0] = c ? c[0] : 0.f; p.c[1] = c ? c[1] : 0.f; p.c[2] = c ? c[2] : 0.f;
    p.radius = radius; p.y_pos = y_pos; p.height = height;
    pt_material zero; memset(&zero, 0, sizeof(zero));   // the caller memsets Object to 0 (video_renderer.cpp:75,85)
    p.mat_slot = material_slot(s, zero);
    s->prims.push_back(p);
    s->dirty = true; s->rank_valid = false; s->layout_valid = false;
    return (int)s->prims.size() - 1;
}
int pt_add_sphere(pt_scene* s, const float center[3], float radius) {
    if (!center) return fail(PT_ERR_INVALID, "null center");
    return add_prim(s, PT_PRIM_SPHERE, center, radius, 0.f, 0.f);
}
int pt_add_platform(pt_scene* s, float y_pos) { return add_prim(s, PT_PRIM_PLATFORM, nullptr, 0.f, y_pos, 0.f); }
int pt_add_cylinder(pt_scene* s, const float center[3], float radius, float height) {
    if (!center) return fail(PT_ERR_INVALID, "null center");
    return add_prim(s, PT_PRIM_CYLINDER, center, radius, 0.f, height);
}
int pt_set_material(pt_scene* s, int prim, const pt_material* m) {
    if (!s || !m) return fail(PT_ERR_INVALID, "null argument");
    if (prim < 0 || prim >= (int)s->prims.size()) return fail(PT_ERR_INVALID, "primitive id out of range");
    s->prims[prim].mat_slot = material_slot(s, *m);
    s->dirty = true; s->layout_valid = false;
    return PT_OK;
}
int pt_update_prim(pt_scene* s, int prim, const float center[3], float radius, float y_pos, float height) {
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    if (prim < 0 || prim >= (int)s->prims.size()) return fail(PT_ERR_INVALID, "primitive id out of range");
    HostPrim& p = s->prims[prim];
    if (center) { p.c[0] = center[0]; p.c[1] = center[1]; p.c[2] = center[2]; }
    p.radius = radius; p.y_pos = y_pos; p.height = height;
    s->dirty = true; s->moved.push_back(prim);
    return PT_OK;
}

int pt_set_sky_rgba8(pt_scene* s, const uint8_t* rgba, int width, int height) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    CK(cudaStreamSynchronize(s->stream));
    destroy_texture(s->sky);
    size_t row_bytes = s->opt_sky_quarter ? (size_t)width : (size_t)width * 4;
    return create_texture(rgba, width, height, row_bytes, s->sky, s->stream);
}

int pt_texture_create_rgba8(pt_scene* s, const uint8_t* rgba, int width, int height, uint64_t* handle) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !handle) return fail(PT_ERR_INVALID, "null argument");
    Texture t;
    int rc = create_texture(rgba, width, height, (size_t)width * 4, t, s->stream);
    if (rc) return rc;
    s->textures[(uint64_t)t.obj] = t;
    *handle = (uint64_t)t.obj;
    return PT_OK;
}
int pt_texture_destroy(pt_scene* s, uint64_t handle) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    auto it = s->textures.find(handle);
    if (it == s->textures.end()) return fail(PT_ERR_INVALID, "unknown texture");
    cudaStreamSynchronize(s->stream);
    destroy_texture(it->second);
    s->textures.erase(it);
    return PT_OK;
}

int pt_build(pt_scene* s) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    return upload_and_build(s);
}

int pt_render_frame_async(pt_scene* s, const pt_camera* cam, int spp, int max_depth, uint64_t seed, uint32_t first_sample) {
    PT_ON_SCENE_DEVICE(s);
    return render_async(s, cam, spp, max_depth, seed, first_sample);
}
int pt_synchronize(pt_scene* s) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    return finish_frame(s, 0);
}
int pt_render_frame(pt_scene* s, const pt_camera* cam, int spp, int max_depth, uint64_t seed, uint32_t first_sample) {
    PT_ON_SCENE_DEVICE(s);
    int rc = render_async(s, cam, spp, max_depth, seed, first_sample);
    if (rc) return rc;
    rc = finish_frame(s, max_depth);
    if (rc) return rc;
    return PT_OK;
}

int pt_ensure_frame(pt_scene* s, int width, int height) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    return ensure_frame(s, width, height);
}

int pt_reset_accum(pt_scene* s) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    if (s->accum.p) CK(cudaMemsetAsync(s->accum.p, 0, (size_t)s->width * s->height * sizeof(float4), s->stream));
    return PT_OK;
}

int pt_read_accum(pt_scene* s, float* rgba, size_t nfloats) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !rgba) return fail(PT_ERR_INVALID, "null argument");
    const size_t npix = (size_t)s->width * s->height;
    if (npix == 0 || nfloats != 4 * npix) return fail(PT_ERR_INVALID, "nfloats must be 4*width*height of the last frame");
    CK(s->read_tmp.reserve(npix));
    k_read_accum<<<(unsigned)((npix + 255) / 256), 256, 0, s->stream>>>((uint32_t)npix, s->accum.p, s->read_tmp.p);
    s->stats.launches += 1;
    CK(cudaMemcpyAsync(rgba, s->read_tmp.p, npix * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return PT_OK;
}

// Read-back in two halves: _begin normalises the SUM buffer into a device staging buffer on the scene's stream and enqueues the
// D2H copy on the copy stream; _end waits for the OLDEST copy begun.  With one read-back always in flight the copy of frame i
// (33 MB at 1080p: 1.5 ms over PCIe) runs under the render of frame i + 1, and the scene's stream never waits for the host.
int pt_read_accum_begin(pt_scene* s, float* rgba, size_t nfloats) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !rgba) return fail(PT_ERR_INVALID, "null argument");
    const size_t npix = (size_t)s->width * s->height;
    if (npix == 0 || nfloats != 4 * npix) return fail(PT_ERR_INVALID, "nfloats must be 4*width*height of the last frame");
    if (s->read_begun - s->read_ended >= 2) return fail(PT_ERR_INVALID, "two read-backs are already in flight: call pt_read_accum_end first");
    const int k = (int)(s->read_begun & 1);
    if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    if (!s->read_ready[k]) CK(cudaEventCreateWithFlags(&s->read_ready[k], cudaEventDisableTiming));
    if (!s->read_done[k]) CK(cudaEventCreateWithFlags(&s->read_done[k], cudaEventDisableTiming));
    CK(s->read_pipe[k].reserve(npix));
    // staging buffer k was last read by the copy begun two calls ago, which _end has waited for (at most two in flight)
    k_read_accum<<<(unsigned)((npix + 255) / 256), 256, 0, s->stream>>>((uint32_t)npix, s->accum.p, s->read_pipe[k].p);
    s->stats.launches += 1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(s->read_ready[k], s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, s->read_ready[k], 0));
    CK(cudaMemcpyAsync(rgba, s->read_pipe[k].p, npix * sizeof(float4), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->read_done[k], s->copy_stream));
    ++s->read_begun;
    return PT_OK;
}
int pt_read_accum_end(pt_scene* s) {
    PT_ON_SCENE_DEVICE(s);
    if (!s) return fail(PT_ERR_INVALID, "null scene");
    if (s->read_begun == s->read_ended) return fail(PT_ERR_INVALID, "no read-back in flight");
    CK(cudaEventSynchronize(s->read_done[s->read_ended & 1]));
    ++s->read_ended;
    return PT_OK;
}

int pt_accum_device_ptr(pt_scene* s, float** dptr, size_t* nfloats) {
    if (!s || !dptr || !nfloats) return fail(PT_ERR_INVALID, "null argument");
    if (!s->accum.p) return fail(PT_ERR_INVALID, "no frame rendered yet");
    *dptr = (float*)s->accum.p; *nfloats = 4 * (size_t)s->width * s->height;
    return PT_OK;
}

int pt_get_stats(pt_scene* s, pt_stats* out) {
    if (!s || !out) return fail(PT_ERR_INVALID, "null argument");
    *out = s->stats;
    return PT_OK;
}

int pt_trace_rays(pt_scene* s, int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !rays8 || n < 0) return fail(PT_ERR_INVALID, "bad argument");
    if (!s->built || s->dirty) return fail(PT_ERR_NOT_BUILT, "pt_build must be called after the scene changed");
    if (n == 0) return PT_OK;
    int rc = ensure_sky(s); if (rc) return rc;
    rc = launch_config(s); if (rc) return rc;
    cudaStream_t st = s->stream;
    rc = to_device(s->scratch_f[0], rays8, (size_t)8 * n, st); if (rc) return rc;
    CK(s->q_o.reserve(n)); CK(s->q_d.reserve(n)); CK(s->hit.reserve(n)); CK(s->ctrl.reserve(4)); CK(s->seg_counter.reserve(4));
    CK(s->scratch_f[1].reserve((size_t)2 * n)); CK(s->scratch_f[2].reserve(n)); CK(s->scratch_f[3].reserve((size_t)3 * n));
    CK(s->scratch_f[4].reserve((size_t)3 * n)); CK(s->scratch_i[0].reserve(n));
    CK(cudaMemsetAsync(s->ctrl.p, 0, 2 * sizeof(BounceCtrl), st));
    const int g = (n + 255) / 256;
    k_pack_rays<<<g, 256, 0, st>>>(n, s->scratch_f[0].p, s->q_o.p, s->q_d.p, (float2*)s->scratch_f[1].p, s->ctrl.p);
    // the SAME bounce kernel as the render path, fed from the caller's batch (SRC_RAYS) instead of a shader
    BounceArgs args;
    memset(&args, 0, sizeof(args));
    args.sc = scene_dev(s); args.segment = 0; args.max_depth = 1; args.trace = 1; args.inner_run = s->opt_inner_run;
    args.ctrl_in = s->ctrl.p; args.ctrl_out = s->ctrl.p + 1;
    args.ray_o = s->q_o.p; args.ray_d = s->q_d.p; args.ray_range = (const float2*)s->scratch_f[1].p; args.slot_hit = s->hit.p;
    args.n_rays = (uint32_t)n; args.step_counters = s->seg_counter.p + 1;
    CK(cudaEventRecord(s->ev0, st));
    if (s->opt_count_steps) {
        CK(s->scratch_i[1].reserve(n));
        CK(cudaMemsetAsync(s->seg_counter.p, 0, 4 * sizeof(unsigned long long), st));
        args.ray_steps = (uint32_t*)s->scratch_i[1].p;
        k_bounce<SRC_RAYS, true><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, PT_BOUNCE_SMEM, st>>>(args);
    } else {
        k_bounce<SRC_RAYS, false><<<s->grid_bounce, 32 * PT_BOUNCE_WARPS, PT_BOUNCE_SMEM, st>>>(args);
    }
    CK(cudaEventRecord(s->ev1, st));
    s->frame_recorded = false;
    k_unpack_hits<<<g, 256, 0, st>>>(n, scene_dev(s), s->q_o.p, s->q_d.p, s->hit.p, s->scratch_i[0].p,
                                     s->scratch_f[2].p, s->scratch_f[3].p, s->scratch_f[4].p);
    s->stats.launches += 3;
    CK(cudaGetLastError());
    if (prim) CK(cudaMemcpyAsync(prim, s->scratch_i[0].p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (t) CK(cudaMemcpyAsync(t, s->scratch_f[2].p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (normal3) CK(cudaMemcpyAsync(normal3, s->scratch_f[3].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    if (hitpos3) CK(cudaMemcpyAsync(hitpos3, s->scratch_f[4].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&s->stats.traverse_ms, s->ev0, s->ev1));      // device time of the closest-hit launch alone
    return PT_OK;
}

int pt_last_trace_steps(pt_scene* s, int n, uint32_t* steps) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !steps || n <= 0) return fail(PT_ERR_INVALID, "bad argument");
    if (!s->opt_count_steps || s->scratch_i[1].n < (size_t)n) return fail(PT_ERR_INVALID, "needs PT_OPT_COUNT_STEPS and a previous pt_trace_rays of >= n rays");
    CK(cudaMemcpyAsync(steps, s->scratch_i[1].p, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return PT_OK;
}

int pt_shade_batch(pt_scene* s, int n, const pt_material* mats, int n_mats, const int32_t* mat_idx, const float* pos3,
                   const float* normal3, const float* in_dir3, const float* u3, float* bounce3, float* atten3,
                   float* radiance3, int32_t* n_draws) {
    PT_ON_SCENE_DEVICE(s);
    (void)pos3;
    if (!s || !mats || !mat_idx || !normal3 || !in_dir3 || !u3 || n < 0 || n_mats <= 0) return fail(PT_ERR_INVALID, "bad argument");
    if (n == 0) return PT_OK;
    for (int i = 0; i < n; ++i) if (mat_idx[i] < 0 || mat_idx[i] >= n_mats) return fail(PT_ERR_INVALID, "mat_idx out of range");
    cudaStream_t st = s->stream;
    std::vector<MaterialRec> recs(n_mats);
    for (int i = 0; i < n_mats; ++i) recs[i] = make_material_rec(mats[i]);
    DevBuf<MaterialRec> dm;
    int rc = to_device(dm, recs.data(), recs.size(), st); if (rc) return rc;
    rc = to_device(s->scratch_i[0], mat_idx, n, st); if (rc) { dm.release(); return rc; }
    rc = to_device(s->scratch_f[0], normal3, (size_t)3 * n, st); if (rc) { dm.release(); return rc; }
    rc = to_device(s->scratch_f[1], in_dir3, (size_t)3 * n, st); if (rc) { dm.release(); return rc; }
    rc = to_device(s->scratch_f[2], u3, (size_t)3 * n, st); if (rc) { dm.release(); return rc; }
    cudaError_t e = cudaSuccess;
    for (int k = 3; k < 6 && e == cudaSuccess; ++k) e = s->scratch_f[k].reserve((size_t)3 * n);
    if (e == cudaSuccess) e = s->scratch_i[1].reserve(n);
    if (e == cudaSuccess) {
        k_shade_batch<<<(n + 127) / 128, 128, 0, st>>>(n, dm.p, s->scratch_i[0].p, s->scratch_f[0].p, s->scratch_f[1].p,
                                                       s->scratch_f[2].p, s->scratch_f[3].p, s->scratch_f[4].p,
                                                       s->scratch_f[5].p, s->scratch_i[1].p);
        s->stats.launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && bounce3) e = cudaMemcpyAsync(bounce3, s->scratch_f[3].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && atten3) e = cudaMemcpyAsync(atten3, s->scratch_f[4].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && radiance3) e = cudaMemcpyAsync(radiance3, s->scratch_f[5].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_draws) e = cudaMemcpyAsync(n_draws, s->scratch_i[1].p, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dm.release();
    CK(e);
    return PT_OK;
}

int pt_raygen_batch(pt_scene* s, const pt_camera* cam, int n, const int32_t* xy, const float* u3, float* rays8) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !cam || !xy || !u3 || !rays8 || n < 0) return fail(PT_ERR_INVALID, "bad argument");
    if (n == 0) return PT_OK;
    cudaStream_t st = s->stream;
    int rc = to_device(s->scratch_i[0], xy, (size_t)2 * n, st); if (rc) return rc;
    rc = to_device(s->scratch_f[0], u3, (size_t)3 * n, st); if (rc) return rc;
    CK(s->scratch_f[1].reserve((size_t)8 * n));
    k_raygen_batch<<<(n + 127) / 128, 128, 0, st>>>(n, to_device_frame(cam), s->scratch_i[0].p, s->scratch_f[0].p, s->scratch_f[1].p);
    s->stats.launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rays8, s->scratch_f[1].p, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PT_OK;
}

int pt_sky_lookup(pt_scene* s, int n, const float* dir3, float* rgb3) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !dir3 || !rgb3 || n < 0) return fail(PT_ERR_INVALID, "bad argument");
    if (n == 0) return PT_OK;
    int rc = ensure_sky(s); if (rc) return rc;
    cudaStream_t st = s->stream;
    rc = to_device(s->scratch_f[0], dir3, (size_t)3 * n, st); if (rc) return rc;
    CK(s->scratch_f[1].reserve((size_t)3 * n));
    k_sky_batch<<<(n + 127) / 128, 128, 0, st>>>(n, s->sky.obj, s->scratch_f[0].p, s->scratch_f[1].p);
    s->stats.launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb3, s->scratch_f[1].p, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PT_OK;
}

int pt_philox_uniforms(pt_scene* s, int n, uint64_t seed, const uint32_t* pss, float* u4) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !pss || !u4 || n < 0) return fail(PT_ERR_INVALID, "bad argument");
    if (n == 0) return PT_OK;
    cudaStream_t st = s->stream;
    DevBuf<uint32_t> d;
    int rc = to_device(d, pss, (size_t)3 * n, st); if (rc) return rc;
    cudaError_t e = s->scratch_f[0].reserve((size_t)4 * n);
    if (e == cudaSuccess) {
        k_philox_batch<<<(n + 127) / 128, 128, 0, st>>>(n, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), d.p, s->scratch_f[0].p);
        s->stats.launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(u4, s->scratch_f[0].p, (size_t)n * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    d.release();
    CK(e);
    return PT_OK;
}

int pt_sort_code_pairs(pt_scene* s, int n, const uint32_t* codes, uint32_t* codes_sorted, uint32_t* slots_sorted) {
    PT_ON_SCENE_DEVICE(s);
    if (!s || !codes || n < 0) return fail(PT_ERR_INVALID, "bad argument");
    if (n == 0) return PT_OK;
    cudaStream_t st = s->stream;
    CK(s->codes.reserve(n)); CK(s->codes_sorted.reserve(n)); CK(s->ids.reserve(n)); CK(s->ids_sorted.reserve(n));
    std::vector<uint32_t> iota((size_t)n);
    for (int i = 0; i < n; ++i) iota[i] = (uint32_t)i;
    CK(cudaMemcpyAsync(s->codes.p, codes, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->ids.p, iota.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    int rc = sort_code_pairs(s, n); if (rc) return rc;
    if (codes_sorted) CK(cudaMemcpyAsync(codes_sorted, s->codes_sorted.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (slots_sorted) CK(cudaMemcpyAsync(slots_sorted, s->ids_sorted.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    s->dirty = true;          // the build's scratch buffers were used
    return PT_OK;
}

int pt_camera_basis(const pt_camera* cam, float out19[19]) {
    if (!cam || !out19) return fail(PT_ERR_INVALID, "null argument");
    HostCameraFrame f = camera_frame(cam);
    const H3* v[6] = {&f.u, &f.v, &f.w, &f.top_left, &f.horizontal, &f.vertical};
    for (int k = 0; k < 6; ++k) { out19[3 * k] = v[k]->x; out19[3 * k + 1] = v[k]->y; out19[3 * k + 2] = v[k]->z; }
    out19[18] = f.dist;
    return PT_OK;
}

int pt_camera_control(pt_camera* cam, int op, float arg, float move_speed) {
    if (!cam || op < 0 || op > 10) return fail(PT_ERR_INVALID, "bad camera control");
    H3 o = h3(cam->origin[0], cam->origin[1], cam->origin[2]), a = h3(cam->look_at[0], cam->look_at[1], cam->look_at[2]);
    const H3 vup = h3(cam->vup[0], cam->vup[1], cam->vup[2]);
    auto neg = [](H3 v) { return h3(-v.x, -v.y, -v.z); };
    if (op <= 3) {
        // left = -normalize(cross(vup, w)), back = -normalize(cross(left, vup))   (motional_camera.cu:77-78,93-95)
        const H3 w = hnormalize(hsub(o, a));
        const H3 left = neg(hnormalize(hcross(vup, w)));
        const H3 dir = op >= 2 ? neg(hnormalize(hcross(left, vup))) : left;
        const H3 step = hscale(dir, arg * move_speed);
        if (op == 0 || op == 3) { o = hadd(o, step); a = hadd(a, step); } else { o = hsub(o, step); a = hsub(a, step); }
    } else if (op <= 5) {
        const H3 step = hscale(vup, arg * move_speed);
        if (op == 4) { o = hadd(o, step); a = hadd(a, step); } else { o = hsub(o, step); a = hsub(a, step); }
    } else if (op <= 9) {
        a = hadd(o, hnormalize(hsub(a, o)));
        const H3 w = hnormalize(hsub(a, o));
        const H3 left = hnormalize(hcross(vup, w));
        const H3 axis = op <= 7 ? hnormalize(hcross(w, left)) : left;
        a = (op == 6 || op == 8) ? hadd(a, hscale(axis, arg)) : hsub(a, hscale(axis, arg));
        a = hadd(o, hnormalize(hsub(a, o)));
    } else {
        cam->view_fov_deg = (float)((double)cam->view_fov_deg + (double)arg * 3.14159265358979323846 / 180.0f);
    }
    cam->origin[0] = o.x; cam->origin[1] = o.y; cam->origin[2] = o.z;
    cam->look_at[0] = a.x; cam->look_at[1] = a.y; cam->look_at[2] = a.z;
    return PT_OK;
}

int pt_dispatch_ray(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx,
                    uint8_t* bgra_out, size_t nbytes) {
    PT_ON_SCENE_DEVICE(s);
    if (!bgra_out) return fail(PT_ERR_INVALID, "null argument");
    return display_dispatch(s, cam, max_depth, seed, sample_idx, bgra_out, nbytes);
}

int pt_dispatch_ray_begin(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx) {
    PT_ON_SCENE_DEVICE(s);
    return display_begin(s, cam, max_depth, seed, sample_idx);
}
int pt_dispatch_ray_end(pt_scene* s, const uint8_t** bgra, size_t* nbytes) {
    PT_ON_SCENE_DEVICE(s);
    return display_end(s, bgra, nbytes);
}

int pt_dispatch_ray_frame(pt_scene* s, const pt_camera* cam, int max_depth, uint64_t seed, uint32_t sample_idx,
                          const uint8_t** bgra, size_t* nbytes) {
    PT_ON_SCENE_DEVICE(s);
    if (!bgra) return fail(PT_ERR_INVALID, "null argument");
    int rc = display_dispatch(s, cam, max_depth, seed, sample_idx, nullptr, 0);
    if (rc) return rc;
    *bgra = s->disp_host;
    if (nbytes) *nbytes = (size_t)cam->width * cam->height * 4;
    return PT_OK;
}

}  // extern "C"
