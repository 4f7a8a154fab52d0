/*
 * ref_cuda_harness.cu -- TEST / BASELINE INFRASTRUCTURE.  C-ABI around the reference's CUDA code
 * rebuilt for sm_100 (cuSrc/{object,material,bvh,motional_camera}.cu unmodified; cuSrc/path_tracer.cu
 * with the one-line rvalue fix at :258 applied by the Makefile into oracle/_ref/ and #included below
 * so that its file-local PathTracerParams / SamplePixel / Denoising / Mix are reachable).
 * cuSrc/textures.cu needs OpenCV and is replaced by the three functions below.
 * Used by bench.py (reported baseline "reference CUDA on sm_100") and by -m gpu parity tests.
 */
#include <vector>
#include "path_tracer_patched.cu"

Logger logger("/dev/null");

static cudaTextureObject_t g_sky_tex = 0;
static cudaArray_t g_sky_arr = nullptr;

cudaTextureObject_t PocaTextureUtils::AddTexByFile(std::string, cudaTextureAddressMode, cudaTextureFilterMode) { return g_sky_tex; }
void PocaTextureUtils::DestroyTexture(cudaTextureObject_t) {}
__device__ float4 PocaTextureUtils::GetTexture2D(cudaTextureObject_t tex_obj, float u, float v) {
    return tex2D<float4>(tex_obj, u, v);                      /* cuSrc/textures.cu:68-71 */
}

static std::vector<Object*> g_objs;   /* never freed, like the reference (bvh.cu:27) */
static SceneBVHGPUHandle g_root = nullptr;
static bool g_built = false;

struct refcuda_material { int32_t type; int32_t have_tex; float kd[3]; float refractive_index, emit_intensity, smoothness, reflectivity; };

static inline float int_as_float_h(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

__global__ void k_ref_trace(SceneBVHGPUHandle root, int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray ray;
    ray.origin = make_float3(rays8[8 * i], rays8[8 * i + 1], rays8[8 * i + 2]);
    ray.dir = make_float3(rays8[8 * i + 3], rays8[8 * i + 4], rays8[8 * i + 5]);
    ray.tmin = rays8[8 * i + 6]; ray.tmax = rays8[8 * i + 7];
    IntersectionAttributes attr; attr.normal = make_float3(0.f); attr.hit_pos = make_float3(0.f);
    Object obj;
    bool hit = root->TraceRay(ray, attr, obj);
    if (hit) {
        prim[i] = obj.type_ == PrimitiveType::Platform ? __float_as_int(obj.height_) : __float_as_int(obj.y_pos_);
        Ray r2 = ray; IntersectionAttributes a2 = attr;
        obj.IntersectionTest(r2, a2);
        t[i] = r2.tmax;
    } else { prim[i] = -1; t[i] = ray.tmax; attr.normal = make_float3(0.f); attr.hit_pos = make_float3(0.f); }
    normal3[3 * i] = attr.normal.x; normal3[3 * i + 1] = attr.normal.y; normal3[3 * i + 2] = attr.normal.z;
    hitpos3[3 * i] = attr.hit_pos.x; hitpos3[3 * i + 1] = attr.hit_pos.y; hitpos3[3 * i + 2] = attr.hit_pos.z;
}

__global__ void k_ref_sum(int n, const float3* src, float* sum3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum3[3 * i] += src[i].x; sum3[3 * i + 1] += src[i].y; sum3[3 * i + 2] += src[i].z;
}

extern "C" {

int refcuda_load_scene(int n, const int32_t* type, const float* center3, const float* radius, const float* y_pos,
                       const float* height, const int32_t* mat_idx, const refcuda_material* mats) {
    if (g_built || !g_objs.empty()) { SceneBVH::ReleaseBVH(); g_built = false; }
    g_objs.clear();
    for (int i = 0; i < n; ++i) {
        Object* o = new Object();
        memset(o, 0, sizeof(Object));
        o->type_ = (PrimitiveType::Enum)type[i];
        const refcuda_material& m = mats[mat_idx[i]];
        o->material_.type_ = (MaterialType::Enum)m.type; o->material_.have_tex_ = false;
        o->material_.kd_ = make_float3(m.kd[0], m.kd[1], m.kd[2]);
        o->material_.refractive_index_ = m.refractive_index; o->material_.emit_intensity_ = m.emit_intensity;
        o->material_.smoothness_ = m.smoothness; o->material_.reflectivity_ = m.reflectivity;
        o->center_ = make_float3(center3[3 * i], center3[3 * i + 1], center3[3 * i + 2]);
        o->radius_ = radius[i]; o->y_pos_ = y_pos[i]; o->height_ = height[i];
        if (o->type_ == PrimitiveType::Platform) o->height_ = int_as_float_h(i); else o->y_pos_ = int_as_float_h(i);
        g_objs.push_back(o);
        SceneBVH::AddObject(o);
    }
    g_root = SceneBVH::BuildBVH();
    g_built = true;
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int refcuda_set_sky(const uint8_t* rgba, int w, int h, int quarter_rows) {
    if (g_sky_tex) { cudaDestroyTextureObject(g_sky_tex); cudaFreeArray(g_sky_arr); g_sky_tex = 0; }
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&g_sky_arr, &desc, w, h) != cudaSuccess) return -1;
    std::vector<uint8_t> zero((size_t)w * h * 4, 0);
    cudaMemcpy2DToArray(g_sky_arr, 0, 0, zero.data(), w * 4, w * 4, h, cudaMemcpyHostToDevice);
    /* cuSrc/textures.cu:32: width*sizeof(uint8_t) bytes per row */
    if (cudaMemcpy2DToArray(g_sky_arr, 0, 0, rgba, w * 4, quarter_rows ? w : w * 4, h, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = g_sky_arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeMirror; td.addressMode[1] = cudaAddressModeMirror;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    return cudaCreateTextureObject(&g_sky_tex, &rd, &td, nullptr) == cudaSuccess ? 0 : -3;
}

int refcuda_trace(int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3) {
    float *d_r, *d_t, *d_n, *d_h; int32_t* d_p;
    cudaMalloc(&d_r, (size_t)n * 32); cudaMalloc(&d_t, (size_t)n * 4); cudaMalloc(&d_n, (size_t)n * 12);
    cudaMalloc(&d_h, (size_t)n * 12); cudaMalloc(&d_p, (size_t)n * 4);
    cudaMemcpy(d_r, rays8, (size_t)n * 32, cudaMemcpyHostToDevice);
    k_ref_trace<<<(n + 127) / 128, 128>>>(g_root, n, d_r, d_p, d_t, d_n, d_h);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(prim, d_p, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(t, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(normal3, d_n, (size_t)n * 12, cudaMemcpyDeviceToHost); cudaMemcpy(hitpos3, d_h, (size_t)n * 12, cudaMemcpyDeviceToHost);
    cudaFree(d_r); cudaFree(d_t); cudaFree(d_n); cudaFree(d_h); cudaFree(d_p);
    return e == cudaSuccess ? 0 : -1;
}

/* Runs the reference's per-sample pipeline `spp` times (PipelineLoop, path_tracer.cu:256-306, minus the
 * host thread / callback).  kernel_ms: CUDA-event time of the SamplePixel launches only; service_ms: wall
 * time of the whole per-sample loop with its 3 launches, 3 cudaDeviceSynchronize and the D2H copy.
 * traced_pixels: the grid is (W/16, H/16) blocks of 16x16 (path_tracer.cu:276-278) -> rows beyond
 * 16*(H/16) are never traced.  sum_rgb (optional, W*H*3): per-pixel sum of the raw render target. */
int refcuda_render(int w, int h, const float* origin, const float* look_at, float fov, float lens, int spp, int warmup,
                   unsigned long long seed, float* sum_rgb, float* kernel_ms, float* service_ms, long long* traced_pixels) {
    MotionalCamera cam(w, h, make_float3(origin[0], origin[1], origin[2]), make_float3(look_at[0], look_at[1], look_at[2]));
    cam.SetViewFov(fov); cam.lens_radius_ = lens;
    size_t np = (size_t)w * h;
    uint8_t* d_out; float* d_depth; float3 *d_normal, *d_rt, *d_dn, *d_mix; curandState* d_rng; float* d_sum = nullptr;
    cudaMalloc(&d_out, np * 4); cudaMalloc(&d_depth, np * 4); cudaMalloc(&d_normal, np * sizeof(float4));
    cudaMalloc(&d_rt, np * sizeof(float4)); cudaMalloc(&d_dn, np * sizeof(float3)); cudaMalloc(&d_mix, np * sizeof(float3));
    cudaMalloc(&d_rng, np * sizeof(curandState));
    cudaMemset(d_mix, 0, np * sizeof(float3)); cudaMemset(d_rt, 0, np * sizeof(float4));
    if (sum_rgb) { cudaMalloc(&d_sum, np * 12); cudaMemset(d_sum, 0, np * 12); }
    std::vector<uint8_t> host_out(np * 4);
    InitCuRand<<<dim3(w / 16), dim3(16)>>>(d_rng, seed, w, h);       /* path_tracer.cu:103-107 */
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    dim3 tpb(16, 16), nb(w / 16, h / 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float k_ms = 0.f; double s_ms = 0.0;
    for (int s = -warmup; s < spp; ++s) {
        auto t0 = std::chrono::steady_clock::now();
        MotionalCamera cma = cam.GetCopy();
        PathTracerParams params;
        params.sky_tex_obj = g_sky_tex; params.bvh_root = g_root; params.max_recursion_depth = 8;
        params.camera = cma; params.depth_info_buffer = d_depth; params.normal_info_buffer = d_normal;
        params.render_target = d_rt; params.d_rng_states = d_rng;
        cudaEventRecord(e0);
        SamplePixel<<<nb, tpb>>>(params);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) return -2;
        Denoising<<<nb, tpb>>>(d_rt, d_normal, d_depth, d_dn, 1.f);
        cudaDeviceSynchronize();
        Mix<<<nb, tpb>>>(d_dn, d_mix, d_out, (int)cma.cur_sample_idx_);
        cudaDeviceSynchronize();
        cudaMemcpy(host_out.data(), d_out, np * 4, cudaMemcpyDeviceToHost);
        auto t1 = std::chrono::steady_clock::now();
        if (s >= 0) {
            float ms; cudaEventElapsedTime(&ms, e0, e1); k_ms += ms;
            s_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
            if (d_sum) k_ref_sum<<<(unsigned)((np + 255) / 256), 256>>>((int)np, d_rt, d_sum);
        }
    }
    cudaDeviceSynchronize();
    if (sum_rgb) { cudaMemcpy(sum_rgb, d_sum, np * 12, cudaMemcpyDeviceToHost); cudaFree(d_sum); }
    *kernel_ms = k_ms; *service_ms = (float)s_ms; *traced_pixels = (long long)(w / 16) * 16 * (long long)(h / 16) * 16;
    cudaFree(d_out); cudaFree(d_depth); cudaFree(d_normal); cudaFree(d_rt); cudaFree(d_dn); cudaFree(d_mix); cudaFree(d_rng);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
/* The reference's own Denoising + Mix kernels (path_tracer.cu:177-254) on caller-supplied buffers.
 * w, h must be multiples of 16 (the kernels derive the image size from the grid, :186-187). */
int refcuda_denoise_mix(int w, int h, const float* color3, const float* normal3, float* mix3, uint8_t* bgra, int idx) {
    if (w % 16 || h % 16) return -9;
    size_t np = (size_t)w * h;
    float3 *d_c, *d_n, *d_dn, *d_mix; float* d_depth; uint8_t* d_out;
    cudaMalloc(&d_c, np * 12); cudaMalloc(&d_n, np * 12); cudaMalloc(&d_dn, np * 12); cudaMalloc(&d_mix, np * 12);
    cudaMalloc(&d_depth, np * 4); cudaMalloc(&d_out, np * 4);
    cudaMemcpy(d_c, color3, np * 12, cudaMemcpyHostToDevice); cudaMemcpy(d_n, normal3, np * 12, cudaMemcpyHostToDevice);
    cudaMemcpy(d_mix, mix3, np * 12, cudaMemcpyHostToDevice); cudaMemcpy(d_out, bgra, np * 4, cudaMemcpyHostToDevice);
    std::vector<float> depth(np, DEFAULT_RAY_TMAX);
    cudaMemcpy(d_depth, depth.data(), np * 4, cudaMemcpyHostToDevice);
    dim3 tpb(16, 16), nb(w / 16, h / 16);
    Denoising<<<nb, tpb>>>(d_c, d_n, d_depth, d_dn, 1.f);
    Mix<<<nb, tpb>>>(d_dn, d_mix, d_out, idx);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(mix3, d_mix, np * 12, cudaMemcpyDeviceToHost); cudaMemcpy(bgra, d_out, np * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_c); cudaFree(d_n); cudaFree(d_dn); cudaFree(d_mix); cudaFree(d_depth); cudaFree(d_out);
    return e == cudaSuccess ? 0 : -1;
}
}
