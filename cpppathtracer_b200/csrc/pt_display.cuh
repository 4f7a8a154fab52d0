// pt_display.cuh -- the reference's per-sample display pipeline, fused: Denoising (cuSrc/path_tracer.cu:177-239)
// + Mix (cuSrc/path_tracer.cu:241-254) in ONE kernel (the reference runs two launches with a
// cudaDeviceSynchronize after each and a float3 temp buffer in between, path_tracer.cu:285-294).
#pragma once
#include "pt_math.cuh"

namespace pt {

// color[p]  : radiance of this sample (SamplePixel's render_target, path_tracer.cu:172)
// normal[p] : first-segment normal, or -dir on a miss (path_tracer.cu:152,159-163,174)
// depth     : always DEFAULT_RAY_TMAX -- TraceRay takes the ray by value so payload.ray.tmax never changes
//             (bvh.h:22, path_tracer.cu:162); its weight is exp(0) = 1 and drops out
// mix[p]    : running mean of the clamped filtered samples; bgra: 255.99 * mix as B,G,R (+ opaque alpha;
//             the reference never writes that byte)
// The 5x5 neighbourhood is addressed as v*width+u and only rejected when that offset leaves the image
// (path_tracer.cu:212-216), so taps wrap across row ends; reproduced.
// exp(x) for x <= 0 in double, relative error < 1e-11: range reduction by ln 2 (two-part constant) and a degree-9
// Taylor polynomial on |r| <= 0.347 (next term r^10/10! < 7e-12).  The reference's weights are
// `(float)min(exp((double)-d2 / M_PI), 1.0)` (path_tracer.cu:224,228): the library exp is ~90 FP64 operations, and 50
// of them per pixel made this kernel as long as the whole path trace of the sample (0.52 ms at 1080p).  After the
// rounding to float this version differs from the correctly rounded value for about 2 weights in 10^4, by one float
// ulp -- far below what an 8-bit channel can show (tests/test_gpu_display.py pins the bytes to the reference's own
// kernels; the division by M_PI is a multiplication by its reciprocal, a 1e-16 effect).  Results below half the smallest float denormal are returned as 0, which is what the conversion gives.
__device__ __forceinline__ double exp_nonpositive(double x) {
    if (x < -104.0) return 0.0;
    const double kf = rint(x * 1.4426950408889634);
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = 1.0 / 362880.0;
    p = fma(p, r, 1.0 / 40320.0); p = fma(p, r, 1.0 / 5040.0); p = fma(p, r, 1.0 / 720.0); p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0); p = fma(p, r, 1.0 / 6.0); p = fma(p, r, 0.5); p = fma(p, r, 1.0); p = fma(p, r, 1.0);
    return __longlong_as_double(__double_as_longlong(p) + ((long long)(int)kf << 52));   // 2^k, k >= -151: stays normal
}

__global__ void __launch_bounds__(256) k_denoise_mix(int width, int height, const float4* __restrict__ color,
                                                     const float4* __restrict__ normal, float4* __restrict__ mix,
                                                     uint8_t* __restrict__ bgra, int cur_sample_idx) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const float kern[5][5] = {{1.f, 4.f, 7.f, 4.f, 1.f}, {4.f, 16.f, 26.f, 16.f, 4.f}, {7.f, 26.f, 41.f, 26.f, 7.f},
                              {4.f, 16.f, 26.f, 16.f, 4.f}, {1.f, 4.f, 7.f, 4.f, 1.f}};
    const int offset = y * width + x;
    const float4 c4 = color[offset], n4 = normal[offset];
    const float3 cval = f3(c4.x, c4.y, c4.z), nval = f3(n4.x, n4.y, n4.z);
    float3 sum = f3(0.f, 0.f, 0.f);
    float cum_w = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int u = x + (i - 2), v = y + (j - 2);
            const int cur = v * width + u;
            float c_w = 0.f, n_w = 0.f, p_w = 0.f;
            float3 ctmp = f3(0.f, 0.f, 0.f);
            if (!(cur < 0 || cur >= width * height)) {
                const float4 t4 = color[cur], m4 = normal[cur];
                ctmp = f3(t4.x, t4.y, t4.z);
                float3 t = cval - ctmp;
                float dist2 = dot(t, t);
                c_w = (float)fmin(exp_nonpositive(-(double)dist2 * (1.0 / PT_PI_D)), 1.0);   // exp(-(dist2)/M_PI): double
                t = nval - f3(m4.x, m4.y, m4.z);
                dist2 = fmaxf(dot(t, t), 0.0f);
                n_w = (float)fmin(exp_nonpositive(-(double)dist2 * (1.0 / PT_PI_D)), 1.0);
                p_w = 1.0f;                                                   // equal depths: exp(-0/pi)
            }
            const float weight = c_w * n_w * p_w;
            sum = sum + (weight * kern[i][j]) * ctmp;
            cum_w += weight * kern[i][j];
        }
    }
    const float3 den = sum / cum_w;
    // Mix: lerp(mix, clamp(src, 0, 1), 1/idx) = a + t*(b-a)  (helper_math.h:1154-1157,1168-1171)
    const float t = 1.f / (float)cur_sample_idx;
    const float4 m0 = mix[offset];
    const float3 a = f3(m0.x, m0.y, m0.z);
    const float3 b = f3(fmaxf(0.f, fminf(den.x, 1.f)), fmaxf(0.f, fminf(den.y, 1.f)), fmaxf(0.f, fminf(den.z, 1.f)));
    // idx == 1 is the reference's reset: lerp(old, b, 1) must not propagate a NaN/garbage `old`
    const float3 m = (cur_sample_idx <= 1) ? b : a + t * (b - a);
    mix[offset] = make_float4(m.x, m.y, m.z, 0.f);
    uchar4 px;
    px.x = (uint8_t)(255.99f * m.z); px.y = (uint8_t)(255.99f * m.y); px.z = (uint8_t)(255.99f * m.x); px.w = 255;
    reinterpret_cast<uchar4*>(bgra)[offset] = px;
}

}  // namespace pt
