// Facade of the reference's include/ray_tracing_common.h + the float3 helpers its caller uses
// (make_float3 and friends come from Common/helper_math.h / cuda_runtime.h in the reference).
// Header-only, host-only: no CUDA toolkit is needed to compile a caller against libpt_b200.so.
#pragma once
#include <stdint.h>
#include <cmath>

#include "../../include/pt_api.h"

#ifndef __VECTOR_TYPES_H__   // do not clash with cuda_runtime.h when the caller also includes it
struct float3 { float x, y, z; };
inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
#endif
inline float3 make_float3(float s) { return make_float3(s, s, s); }

typedef uint32_t uint;

#define DEFAULT_RAY_TMAX 1e30f   // include/ray_tracing_common.h:11
#define BOUNCE_RAY_TMIN 2e-5f    // include/ray_tracing_common.h:12

// include/ray_tracing_common.h:37-40
struct DispatchRayArgs {
    void* cbParam;
    void (*Callback)(uint8_t* data, int width, int height, void* cbParam);
};

namespace pt_facade {
// fp32 vector helpers with the operand order of Common/helper_math.h (host fallbacks, :78-81,1264,1325,1436)
inline float3 add(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 scale(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 neg(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 normalize(float3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return scale(v, inv); }
}  // namespace pt_facade
