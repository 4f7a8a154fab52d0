/*
 * host_shim.h -- TEST INFRASTRUCTURE.  Force-included (-include) when the reference's OWN
 * tracing sources (cuSrc/{object,material,bvh,motional_camera}.cu under /root/reference)
 * are compiled UNMODIFIED as host C++ into oracle/_ref/libpt_ref_host.so.
 * Nothing from the reference is copied into this repo; this file only supplies what the
 * Windows/CUDA-only build environment would have supplied:
 *   - headers the reference forgot (logger.hpp:63-69 uses va_list, timer.hpp uses chrono)
 *   - glibc random() clash          (include/ray_tracing_math.hpp:30)
 *   - rvalue unary minus on float3  (Common/helper_math.h:274 takes a non-const ref;
 *                                    used on temporaries at cuSrc/motional_camera.cu:78)
 *   - a host stand-in for cuRAND    (include/ray_tracing_math.hpp:4,82-104)  -> pluggable stream
 *   - a host stand-in for tex2D     (cuSrc/textures.cu:68-71)
 */
#pragma once
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <cstdint>
#include <ctime>
#include <cmath>
#include <chrono>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>

#define random poca_random

inline float3 operator-(const float3& a) { return make_float3(-a.x, -a.y, -a.z); }

/* ---- cuRAND stand-in: the harness decides where the uniforms come from ---- */
#define CURAND_KERNEL_H_
struct curandState {
    const float* fixed;   /* if non-null: replay these uniforms in draw order */
    int          n_fixed;
    int          cursor;  /* draws consumed from the current source           */
    float        u[4];    /* else: the 4 uniforms of the current Philox block  */
};
typedef struct curandState curandState_t;

inline void curand_init(unsigned long long, unsigned long long, unsigned long long, curandState* s) {
    s->fixed = nullptr; s->n_fixed = 0; s->cursor = 0;
    s->u[0] = s->u[1] = s->u[2] = s->u[3] = 0.5f;
}
inline float curand_uniform(curandState* s) {
    int i = s->cursor++;
    if (s->fixed) return s->fixed[i < s->n_fixed ? i : s->n_fixed - 1];
    return s->u[i & 3];
}

/* ---- texture unit stand-in (defined in ref_host_harness.cpp) ---- */
template <class T> T tex2D(cudaTextureObject_t tex, float u, float v);
template <> float4 tex2D<float4>(cudaTextureObject_t tex, float u, float v);
