// pt_math.cuh -- fp32 vector math, Philox stream and the exact primitive / material arithmetic of the
// render path.  This translation unit is compiled with --fmad=false: every `a*b+c` below is two
// IEEE roundings, exactly as the reference's expressions evaluate when compiled without contraction
// (the CPU oracles are built with -ffp-contract=off), so hit distances, normals and scatter
// directions are bit-reproducible against them up to libm differences in sin/cos/pow/asin/atan.
// Where fusing is harmless (BVH slab tests: conservative culling only) fmaf() is written explicitly.
//
// Reference lines cited are relative to the reference repository root.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pt {

#define PT_TMAX 1e30f          // DEFAULT_RAY_TMAX  include/ray_tracing_common.h:11
#define PT_BOUNCE_TMIN 2e-5f   // BOUNCE_RAY_TMIN   include/ray_tracing_common.h:12
#define PT_PI_D 3.14159265358979323846

// ---- float3 algebra with the operand order of Common/helper_math.h -------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }   // helper_math.h:1013
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // :1264
__device__ __forceinline__ float3 cross(float3 a, float3 b) {                                               // :1436
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// normalize = v * rsqrtf(dot(v,v)) (:1325); evaluated as the IEEE 1/sqrt so host and device agree bit for bit
__device__ __forceinline__ float3 normalize(float3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
__device__ __forceinline__ float3 reflect(float3 i, float3 n) { return i - (2.0f * n) * dot(n, i); }        // :1427

// ---- Philox4x32-10, keyed (seed) / counter (pixel, sample, slot, 0) -------------------------------
// slot 0 = camera ray, slot 1+k = shader of segment k.  Uniforms follow curand_uniform: (0,1].
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
    return __uint2float_rn(x) * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}
__device__ __forceinline__ float4 philox_uniform4(uint2 key, uint32_t pixel, uint32_t sample, uint32_t slot) {
    uint4 r = philox4x32_10(make_uint4(pixel, sample, slot, 0u), key);
    return make_float4(u32_to_uniform(r.x), u32_to_uniform(r.y), u32_to_uniform(r.z), u32_to_uniform(r.w));
}

// ---- device-side scene records --------------------------------------------------------------------
// geo0 = (cx, cy, cz, radius)   geo1 = (height, y_pos, bits(type), bits(material slot))
struct CameraFrame {            // what MotionalCamera::GetCopy hands to the kernels (motional_camera.cu:177-200)
    float3 origin, u, v, top_left, horizontal, vertical;
    float lens_radius;
    int width, height;
};
struct MaterialRec {            // 48 B; one per distinct material
    float kd[3];
    float emit;
    float ior;
    float reflectivity;
    int   type;
    int   flags;                // bit0: lobe degenerates to the axis (alpha >= 1e9 => z rounds to 1.0f)
    double inv_alpha;           // 1.0 / (double)powf(1000.0f, smoothness)   (material.cu:43,45,69,104-105)
    unsigned long long tex;     // kd texture or 0
};
#define PT_MATF_SPECULAR 1

// hit part codes kept in the low 3 bits of the hit record
enum { PART_SPHERE_NEAR = 0, PART_SPHERE_FAR = 1, PART_PLATFORM = 2, PART_CYL_TOP = 3, PART_CYL_BOTTOM = 4,
       PART_CYL_SIDE = 5 };

// ---- intersectors: t only; the normal is rebuilt from (t, part) in the shade stage ----------------
// Each returns true and lowers tmax when the reference's test would (strict <, > against tmax, tmin).

// SphereIntersectionTest, cuSrc/object.cu:10-35
__device__ __forceinline__ bool hit_sphere(float4 g0, float3 o, float3 d, float tmin, float& tmax, int& part) {
    float3 ac = o - f3(g0.x, g0.y, g0.z);
    float a = dot(d, d);
    float b = dot(ac, d);
    float c = dot(ac, ac) - g0.w * g0.w;
    float disc = b * b - a * c;
    if (disc > 0.f) {
        float sq = sqrtf(disc);
        float t = (-b - sq) / a;
        if (t < tmax && t > tmin) { tmax = t; part = PART_SPHERE_NEAR; return true; }
        t = (-b + sq) / a;
        if (t < tmax && t > tmin) { tmax = t; part = PART_SPHERE_FAR; return true; }
    }
    return false;
}
// PlatformIntersectionTest, cuSrc/object.cu:37-48
__device__ __forceinline__ bool hit_platform(float y_pos, float3 o, float3 d, float tmin, float& tmax) {
    if ((o.y < y_pos && d.y > 0.f) || (o.y > y_pos && d.y < 0.f)) {
        float t = (y_pos - o.y) / d.y;
        if (t < tmax && t > tmin) { tmax = t; return true; }
    }
    return false;
}
// CylinderIntersectionTest, cuSrc/object.cu:50-112 (caps first, then both side roots, each may overwrite)
__device__ __forceinline__ bool hit_cylinder(float4 g0, float height, float3 o, float3 d, float tmin, float& tmax,
                                             int& part) {
    bool ret = false;
    const float radius = g0.w;
    float upper = g0.y + height / 2;
    if ((o.y < upper && d.y > 0.f) || (o.y > upper && d.y < 0.f)) {
        float t = (upper - o.y) / d.y;
        float3 p = o + t * d;
        if (t < tmax && t > tmin && sqrtf((p.x - g0.x) * (p.x - g0.x) + (p.z - g0.z) * (p.z - g0.z)) < radius) {
            tmax = t; part = PART_CYL_TOP; ret = true;
        }
    }
    float lower = g0.y - height / 2;
    if ((o.y < lower && d.y > 0.f) || (o.y > lower && d.y < 0.f)) {
        float t = (lower - o.y) / d.y;
        float3 p = o + t * d;
        if (t < tmax && t > tmin && sqrtf((p.x - g0.x) * (p.x - g0.x) + (p.z - g0.z) * (p.z - g0.z)) < radius) {
            tmax = t; part = PART_CYL_BOTTOM; ret = true;
        }
    }
    float cx = o.x - g0.x, cz = o.z - g0.z;
    float a = d.x * d.x + d.z * d.z;
    float b = cx * d.x + cz * d.z;
    float c = cx * cx + cz * cz - radius * radius;
    float disc = b * b - a * c;
    if (disc > 0.f) {
        float sq = sqrtf(disc);
        float t = (-b - sq) / a;
        float py = o.y + t * d.y;
        if (t < tmax && t > tmin && py > lower && py < upper) { tmax = t; part = PART_CYL_SIDE; ret = true; }
        t = (-b + sq) / a;
        py = o.y + t * d.y;
        if (t < tmax && t > tmin && py > lower && py < upper) { tmax = t; part = PART_CYL_SIDE; ret = true; }
    }
    return ret;
}

// attr.hit_pos / attr.normal of the winning test (object.cu:21-23,29-30,42-43,61-62,75-76,96-98,105-107)
__device__ __forceinline__ void hit_normal(float4 g0, int part, float3 pos, float3 d, float3& normal) {
    switch (part) {
        case PART_SPHERE_NEAR: normal = (pos - f3(g0.x, g0.y, g0.z)) / g0.w; break;   // signed radius
        case PART_SPHERE_FAR: normal = normalize(pos - f3(g0.x, g0.y, g0.z)); break;
        case PART_CYL_SIDE: normal = normalize(f3(pos.x - g0.x, 0.f, pos.z - g0.z)); break;
        default: normal = normalize(f3(0.f, -d.y, 0.f)); break;                       // platform and caps
    }
}
__device__ __forceinline__ void hit_attributes(float4 g0, int part, float3 o, float3 d, float t, float3& pos,
                                               float3& normal) {
    pos = o + t * d;
    hit_normal(g0, part, pos, d, normal);
}

// ---- shading ----------------------------------------------------------------------------------------
// to_world, include/ray_tracing_math.hpp:51-63
__device__ __forceinline__ float3 to_world(float3 a, float3 N) {
    float3 C;
    if (fabsf(N.x) > fabsf(N.y)) {
        float inv = 1.0f / sqrtf(N.x * N.x + N.z * N.z);
        C = f3(N.z * inv, 0.0f, -N.x * inv);
    } else {
        float inv = 1.0f / sqrtf(N.y * N.y + N.z * N.z);
        C = f3(0.f, N.z * inv, -N.y * inv);
    }
    float3 B = cross(C, N);
    return (a.x * B + a.y * C) + a.z * N;
}
// schlick, include/ray_tracing_math.hpp:65-69: pow(float,int) is a DOUBLE pow in the reference's host
// semantics; x^5 by three double multiplications stays within 2 ulp(double) of it, i.e. the same float
__device__ __forceinline__ float schlick(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 *= r0;
    double x = (double)(1.0f - cosine);
    double x2 = x * x;
    return (float)((double)r0 + (double)(1.0f - r0) * (x2 * x2 * x));
}
// refract, include/ray_tracing_math.hpp:71-80 (`1.0 - ...` is a double subtraction)
__device__ __forceinline__ bool refract(float3 v, float3 n, float ni_over_nt, float3& refracted) {
    float3 uv = normalize(v);
    float dt = dot(uv, n);
    float disc = (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1.0f - dt * dt)));
    if (disc > 0.f) {
        refracted = normalize(ni_over_nt * (uv - n * dt) - n * sqrtf(disc));
        return true;
    }
    return false;
}
// the lobe sample every shader builds (material.cu:24-27,45-48,78-80,85-87,105-107)
//   z = pow(x1, 1.0/e) evaluated in double, r = sqrt(1-z*z), phi = 2*M_PI*x2 (double product)
__device__ __forceinline__ float3 lobe_cosine(float x1, float x2) {
    // exponent 1/2: (float)pow((double)x1, 0.5) == sqrtf(x1) (double sqrt then float rounding is innocuous)
    float z = sqrtf(x1);
    float r = sqrtf(1.0f - z * z);
    float phi = (float)(2 * PT_PI_D * (double)x2);
    float s, c;
    sincosf(phi, &s, &c);
    return f3(r * c, r * s, z);
}
__device__ __forceinline__ float3 lobe_glossy(float x1, float x2, const MaterialRec& m) {
    if (m.flags & PT_MATF_SPECULAR) {
        // alpha >= 1e9: x1^(1/alpha) >= 1 - 2^-25 for every x1 >= 2^-33, so z == 1.0f and r == 0 exactly
        float phi = (float)(2 * PT_PI_D * (double)x2);
        float s, c;
        sincosf(phi, &s, &c);
        return f3(0.0f * c, 0.0f * s, 1.0f);
    }
    float z = (float)pow((double)x1, m.inv_alpha);
    float r = sqrtf(1.0f - z * z);
    float phi = (float)(2 * PT_PI_D * (double)x2);
    float s, c;
    sincosf(phi, &s, &c);
    return f3(r * c, r * s, z);
}

struct Scatter { float3 bounce, atten, radiance; int draws; };

__device__ __forceinline__ float3 material_kd(const MaterialRec& m) {   // Material::GetKd, material.cu:11-18
    if (m.tex) {
        float4 t = tex2D<float4>((cudaTextureObject_t)m.tex, 0.f, 0.f);
        return f3(t.x, t.y, t.z);
    }
    return f3(m.kd[0], m.kd[1], m.kd[2]);
}
__device__ __forceinline__ float3 material_emission(const MaterialRec& m) {   // emit_intensity_ * kd_ (always the colour field)
    return m.emit * f3(m.kd[0], m.kd[1], m.kd[2]);
}

// DiffuseHitShader, material.cu:20-38
__device__ __forceinline__ void shade_diffuse(const MaterialRec& m, float3 normal, float4 u, Scatter& o) {
    o.draws = 2;
    o.bounce = to_world(lobe_cosine(u.x, u.y), normal);
    o.atten = dot(normal, o.bounce) > 0.0f ? material_kd(m) : f3(0.f, 0.f, 0.f);
    o.radiance = material_emission(m);
}
// MirrorHitShader, material.cu:40-64 -- the dispatcher runs it for MaterialType::Metal
__device__ __forceinline__ void shade_glossy(const MaterialRec& m, float3 normal, float3 in_dir, float4 u, Scatter& o) {
    o.draws = 2;
    float3 wo = to_world(lobe_glossy(u.x, u.y, m), reflect(in_dir, normal));
    o.atten = dot(normal, wo) > 0.0f ? material_kd(m) : f3(0.f, 0.f, 0.f);
    o.bounce = wo;
    o.radiance = material_emission(m);
}
// MetalHitShader, material.cu:66-99 -- the dispatcher runs it for MaterialType::Mirror
__device__ __forceinline__ void shade_mix(const MaterialRec& m, float3 normal, float3 in_dir, float4 u, Scatter& o) {
    o.draws = 3;
    if (u.z < m.reflectivity) o.bounce = to_world(lobe_glossy(u.x, u.y, m), reflect(in_dir, normal));
    else o.bounce = to_world(lobe_cosine(u.x, u.y), normal);
    o.atten = dot(o.bounce, normal) < 0.f ? f3(0.f, 0.f, 0.f) : material_kd(m);
    o.radiance = material_emission(m);
}
// GlassHitShader, material.cu:101-143
__device__ __forceinline__ void shade_glass(const MaterialRec& m, float3 normal, float3 in_dir, float4 u, Scatter& o) {
    o.draws = 3;
    float3 local = lobe_glossy(u.x, u.y, m);
    float3 outward, refracted = f3(0.f, 0.f, 0.f);
    float ni_over_nt, reflect_prob, cosine;
    in_dir = normalize(in_dir);
    float din = dot(in_dir, normal);
    if (din > 0.f) {
        outward = -normal;
        ni_over_nt = m.ior;
        cosine = sqrtf(1.0f - m.ior * m.ior * (1.0f - din * din));
    } else {
        outward = normal;
        ni_over_nt = 1.f / m.ior;
        cosine = -din;
    }
    if (refract(in_dir, outward, ni_over_nt, refracted)) reflect_prob = schlick(cosine, m.ior);
    else reflect_prob = 1.0f;
    if (u.z < reflect_prob) o.bounce = to_world(local, reflect(in_dir, normal));
    else o.bounce = to_world(local, refracted);
    o.atten = material_kd(m);
    o.radiance = material_emission(m);
}
// shader bins (the Metal<->Mirror swap of material.cu:151-156 is applied when the bin is chosen)
// Shader bins.  Bin 0 names the outcome "left the scene": such a ray is shaded by the sky where it ends and never
// enters a queue, so the hit queues hold NUM_BINS - 1 regions (bin b at region b - 1) and bin 0 stays empty.
enum { BIN_MISS = 0, BIN_DIFFUSE = 1, BIN_GLOSSY = 2, BIN_MIX = 3, BIN_GLASS = 4, NUM_BINS = 5 };
__host__ __device__ __forceinline__ int bin_of_material_type(int type) {
    switch (type) {
        case 1: return BIN_GLOSSY;   // MaterialType::Metal  -> MirrorHitShader
        case 2: return BIN_MIX;      // MaterialType::Mirror -> MetalHitShader
        case 3: return BIN_GLASS;
        default: return BIN_DIFFUSE; // Diffuse, Test, anything else (material.cu:160-161)
    }
}
__device__ __forceinline__ void shade_bin(int bin, const MaterialRec& m, float3 normal, float3 in_dir, float4 u,
                                          Scatter& o) {
    switch (bin) {
        case BIN_GLOSSY: shade_glossy(m, normal, in_dir, u, o); break;
        case BIN_MIX: shade_mix(m, normal, in_dir, u, o); break;
        case BIN_GLASS: shade_glass(m, normal, in_dir, u, o); break;
        default: shade_diffuse(m, normal, u, o); break;
    }
}

// Miss, cuSrc/path_tracer.cu:117-122: the divisions by M_PI are DOUBLE divisions in the reference.  They are
// evaluated here as double multiplications by 1/pi: the double results differ by <= 1 ulp(double), i.e. the float
// they round to is the same except within 2^-29 of a rounding tie (and feeds an 8-bit-weight texture filter).
__device__ __forceinline__ float3 sky_radiance(cudaTextureObject_t sky, float3 dir) {
    float3 d = normalize(dir);
    const double inv_pi = 0.31830988618379067154;
    float v = (float)((double)asinf(d.z) * inv_pi + 0.5);
    float u = (float)((double)(atanf(d.y / d.x) / 2) * inv_pi);
    float4 t = tex2D<float4>(sky, u, v);
    return f3(t.x, t.y, t.z);
}

// MotionalCamera::RayGen, cuSrc/motional_camera.cu:202-213.  U = (rd.x, rd.y, rd.z) draws.
__device__ __forceinline__ void camera_ray(const CameraFrame& c, int x, int y, float4 U, float3& o, float3& d) {
    float3 rd = c.lens_radius * f3(U.x, U.y, U.z);
    float3 offset = c.u * rd.x + c.v * rd.y;
    float dx = (float)x / (float)c.width;
    float dy = (float)y / (float)c.height;
    o = c.origin + offset;
    d = normalize((((c.top_left + dx * c.horizontal) + dy * c.vertical) - c.origin) - offset);
}

}  // namespace pt
