/*
 * pt_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement ("port" oracle) of the reference's render
 * hot path.  C-style C++ (std::sort is the only library algorithm, kept because the reference's BVH
 * build is defined in terms of it).  Build:  g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math.
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this file is
 * pinned against the reference's own sources compiled for the host (oracle/_ref/libpt_ref_host.so,
 * built by oracle/ref_build/Makefile) -- bit-exact on trace / shade / raygen / render in
 * tests/test_oracle.py -- and against tests/golden/ (npz files) generated from that library by
 * tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library.  The product (cpppathtracer_b200/) never does.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * All arithmetic is fp32 evaluated left to right without contraction, except where the reference's
 * C++ promotes to double (noted at each site).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <omp.h>

#include "pt_philox.h"

namespace {

/* ---------- helper_math.h float3 semantics (Common/helper_math.h) ---------- */
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }      /* :  float3*float and float*float3 */
inline V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }       /* :1013-1016 */
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   /* :1264-1267 */
inline float length(V3 a) { return sqrtf(dot(a, a)); }                        /* :1307-1310 */
inline V3 normalize(V3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return scale(v, inv); } /* :1325-1329, host rsqrtf :78-81 */
inline V3 cross(V3 a, V3 b) {                                                  /* :1436-1439 */
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V3 reflect(V3 i, V3 n) { return sub(i, scale(scale(n, 2.0f), dot(n, i))); } /* :1427-1430: i - 2.0f*n*dot(n,i) */

#define RT_TMAX 1e30f       /* DEFAULT_RAY_TMAX, include/ray_tracing_common.h:11 */
#define RT_BOUNCE_TMIN 2e-5f /* BOUNCE_RAY_TMIN,  include/ray_tracing_common.h:12 */
#define RT_PI 3.14159265358979323846 /* M_PI (double), include/ray_tracing_math.hpp:12-14 */
#define RMAX(a, b) ((a) > (b) ? (a) : (b)) /* include/ray_tracing_math.hpp:16-22 */
#define RMIN(a, b) ((a) < (b) ? (a) : (b))

enum { SPHERE = 0, PLATFORM = 1, CYLINDER = 2 };                  /* include/object.h:7-15 */
enum { DIFFUSE = 0, METAL = 1, MIRROR = 2, GLASS = 3, TESTMAT = 4 }; /* include/material.h:5-15 */

struct Mat { int32_t type; int32_t have_tex; float kd[3]; float ior, emit, smooth, refl; };
struct Prim { int32_t type; Mat mat; V3 c; float radius, y_pos, height; int32_t id; };
struct Ray { V3 o, d; float tmin, tmax; };
struct Hit { V3 normal, pos; };

/* ---------- intersectors ---------- */
/* SphereIntersectionTest, cuSrc/object.cu:10-35 */
bool hit_sphere(const Prim& s, Ray& ray, Hit& h) {
    V3 ac = sub(ray.o, s.c);
    float a = dot(ray.d, ray.d);
    float b = dot(ac, ray.d);
    float c = dot(ac, ac) - s.radius * s.radius;
    float disc = b * b - a * c;
    if (disc > 0) {
        float t = (-b - sqrtf(disc)) / a;
        if (t < ray.tmax && t > ray.tmin) {
            ray.tmax = t;
            h.pos = add(ray.o, scale(ray.d, t));
            h.normal = divs(sub(h.pos, s.c), s.radius);   /* signed radius: inward normal for r<0 */
            return true;
        }
        t = (-b + sqrtf(disc)) / a;
        if (t < ray.tmax && t > ray.tmin) {
            ray.tmax = t;
            h.pos = add(ray.o, scale(ray.d, t));
            h.normal = normalize(sub(h.pos, s.c));
            return true;
        }
    }
    return false;
}
/* PlatformIntersectionTest, cuSrc/object.cu:37-48 */
bool hit_platform(const Prim& s, Ray& ray, Hit& h) {
    if ((ray.o.y < s.y_pos && ray.d.y > 0.f) || (ray.o.y > s.y_pos && ray.d.y < 0.f)) {
        float t = (s.y_pos - ray.o.y) / ray.d.y;
        if (t < ray.tmax && t > ray.tmin) {
            ray.tmax = t;
            h.pos = add(ray.o, scale(ray.d, t));
            h.normal = normalize(v3(0, -ray.d.y, 0));
            return true;
        }
    }
    return false;
}
/* CylinderIntersectionTest, cuSrc/object.cu:50-112 */
bool hit_cylinder(const Prim& s, Ray& ray, Hit& h) {
    bool ret = false;
    float upper = s.c.y + s.height / 2;
    if ((ray.o.y < upper && ray.d.y > 0.f) || (ray.o.y > upper && ray.d.y < 0.f)) {
        float t = (upper - ray.o.y) / ray.d.y;
        V3 p = add(ray.o, scale(ray.d, t));
        if (t < ray.tmax && t > ray.tmin &&
            sqrtf((p.x - s.c.x) * (p.x - s.c.x) + (p.z - s.c.z) * (p.z - s.c.z)) < s.radius) {
            ray.tmax = t; h.pos = p; h.normal = normalize(v3(0, -ray.d.y, 0)); ret = true;
        }
    }
    float lower = s.c.y - s.height / 2;
    if ((ray.o.y < lower && ray.d.y > 0.f) || (ray.o.y > lower && ray.d.y < 0.f)) {
        float t = (lower - ray.o.y) / ray.d.y;
        V3 p = add(ray.o, scale(ray.d, t));
        if (t < ray.tmax && t > ray.tmin &&
            sqrtf((p.x - s.c.x) * (p.x - s.c.x) + (p.z - s.c.z) * (p.z - s.c.z)) < s.radius) {
            ray.tmax = t; h.pos = p; h.normal = normalize(v3(0, -ray.d.y, 0)); ret = true;
        }
    }
    float dx = ray.d.x, dz = ray.d.z, r = s.radius;
    float cx = ray.o.x - s.c.x, cz = ray.o.z - s.c.z;
    float a = dx * dx + dz * dz;
    float b = cx * dx + cz * dz;
    float c = cx * cx + cz * cz - r * r;
    float disc = b * b - a * c;
    if (disc > 0.f) {
        float t = (-b - sqrtf(disc)) / a;
        V3 p = add(ray.o, scale(ray.d, t));
        if (t < ray.tmax && t > ray.tmin && p.y > lower && p.y < upper) {
            ray.tmax = t; h.pos = p; h.normal = normalize(v3(p.x - s.c.x, 0.f, p.z - s.c.z)); ret = true;
        }
        t = (-b + sqrtf(disc)) / a;
        p = add(ray.o, scale(ray.d, t));
        if (t < ray.tmax && t > ray.tmin && p.y > lower && p.y < upper) {
            ray.tmax = t; h.pos = p; h.normal = normalize(v3(p.x - s.c.x, 0.f, p.z - s.c.z)); ret = true;
        }
    }
    return ret;
}
/* Object::IntersectionTest, cuSrc/object.cu:114-128 */
bool hit_prim(const Prim& p, Ray& ray, Hit& h) {
    switch (p.type) {
        case SPHERE: return hit_sphere(p, ray, h);
        case PLATFORM: return hit_platform(p, ray, h);
        case CYLINDER: return hit_cylinder(p, ray, h);
        default: return false;
    }
}
/* Object::GetAABBMax / GetAABBMin, cuSrc/object.cu:134-170 (ABS macro: a >= 0 ? a : -a) */
V3 aabb_max(const Prim& p) {
    float tol = RT_BOUNCE_TMIN * 5.f;
    float ar = p.radius >= 0 ? p.radius : -p.radius;
    switch (p.type) {
        case SPHERE: return add(p.c, v3(ar, ar, ar));
        case PLATFORM: return v3(RT_TMAX * 5, p.y_pos + tol, RT_TMAX * 5);
        case CYLINDER: return v3(p.c.x + ar, p.c.y + p.height / 2 + tol, p.c.z + ar);
        default: return v3(0, 0, 0);
    }
}
V3 aabb_min(const Prim& p) {
    float tol = RT_BOUNCE_TMIN * 5.f;
    float ar = p.radius >= 0 ? p.radius : -p.radius;
    switch (p.type) {
        case SPHERE: return sub(p.c, v3(ar, ar, ar));
        case PLATFORM: return v3(-RT_TMAX * 5, p.y_pos - tol, -RT_TMAX * 5);
        case CYLINDER: return v3(p.c.x - ar, p.c.y - p.height / 2 - tol, p.c.z - ar);
        default: return v3(0, 0, 0);
    }
}

/* ---------- BVH: SceneBVH::Divide / BuildBVH, cuSrc/bvh.cu:31-120 ---------- */
struct Node { V3 mn, mx; bool leaf; int l, r; int prim; };
struct Scene {
    std::vector<Prim> prims;   /* insertion order = prim id */
    std::vector<Node> nodes;   /* pre-order, root = 0 */
};

int divide(Scene& sc, std::vector<int>& order, int l, int r) {
    if (l >= r) return -1;
    int idx = (int)sc.nodes.size();
    sc.nodes.push_back(Node());
    const Prim& pl = sc.prims[order[l]];
    V3 lmn = aabb_min(pl), lmx = aabb_max(pl);
    if (l == r - 1) {
        Node& n = sc.nodes[idx];
        n.l = n.r = -1; n.mn = lmn; n.mx = lmx; n.leaf = true; n.prim = order[l];
        return idx;
    }
    float minx = lmn.x, miny = lmn.y, minz = lmn.z, maxx = lmx.x, maxy = lmx.y, maxz = lmx.z;
    for (int i = l + 1; i < r; ++i) {
        V3 mn = aabb_min(sc.prims[order[i]]), mx = aabb_max(sc.prims[order[i]]);
        minx = RMIN(minx, mn.x); miny = RMIN(miny, mn.y); minz = RMIN(minz, mn.z);
        maxx = RMAX(maxx, mx.x); maxy = RMAX(maxy, mx.y); maxz = RMAX(maxz, mx.z);
    }
    float sx = maxx - minx, sy = maxy - miny, sz = maxz - minz;
    const std::vector<Prim>& P = sc.prims;
    if (sx >= sy && sx >= sz)
        std::sort(order.begin() + l, order.begin() + r, [&P](int a, int b) {
            return (aabb_min(P[a]).x + aabb_max(P[a]).x) / 2 < (aabb_min(P[b]).x + aabb_max(P[b]).x) / 2; });
    else if (sy >= sz)
        std::sort(order.begin() + l, order.begin() + r, [&P](int a, int b) {
            return (aabb_min(P[a]).y + aabb_max(P[a]).y) / 2 < (aabb_min(P[b]).y + aabb_max(P[b]).y) / 2; });
    else
        std::sort(order.begin() + l, order.begin() + r, [&P](int a, int b) {
            return (aabb_min(P[a]).z + aabb_max(P[a]).z) / 2 < (aabb_min(P[b]).z + aabb_max(P[b]).z) / 2; });
    int mid = (l + r) / 2;
    int ls = divide(sc, order, l, mid);
    int rs = divide(sc, order, mid, r);
    Node& n = sc.nodes[idx];
    n.l = ls; n.r = rs; n.mn = v3(minx, miny, minz); n.mx = v3(maxx, maxy, maxz); n.leaf = false; n.prim = -1;
    return idx;
}

/* SceneBVH::TraceRay, cuSrc/bvh.cu:167-205.  Returns the prim index or -1; ray.tmax ends at the hit t. */
int trace(const Scene& sc, Ray ray, Hit& h, float* t_out) {
    int stack[512];
    int top = 0;
    stack[top++] = 0;
    int best = -1;
    while (top > 0) {
        int ni = stack[--top];
        if (ni == -1) continue;
        const Node& n = sc.nodes[ni];
        if (n.leaf) {
            if (hit_prim(sc.prims[n.prim], ray, h)) best = n.prim;
        }
        float tmin = -RT_TMAX * 2, tmax = RT_TMAX * 2;
        if (ray.d.x != 0.f) {
            float t0 = (n.mn.x - ray.o.x) / ray.d.x, t1 = (n.mx.x - ray.o.x) / ray.d.x;
            tmin = RMAX(tmin, RMIN(t0, t1)); tmax = RMIN(tmax, RMAX(t0, t1));
        }
        if (ray.d.y != 0.f) {
            float t0 = (n.mn.y - ray.o.y) / ray.d.y, t1 = (n.mx.y - ray.o.y) / ray.d.y;
            tmin = RMAX(tmin, RMIN(t0, t1)); tmax = RMIN(tmax, RMAX(t0, t1));
        }
        if (ray.d.z != 0.f) {
            float t0 = (n.mn.z - ray.o.z) / ray.d.z, t1 = (n.mx.z - ray.o.z) / ray.d.z;
            tmin = RMAX(tmin, RMIN(t0, t1)); tmax = RMIN(tmax, RMAX(t0, t1));
        }
        if (tmin > tmax || tmin > ray.tmax || tmax < ray.tmin) continue;
        stack[top++] = n.l;
        stack[top++] = n.r;
    }
    if (t_out) *t_out = ray.tmax;
    return best;
}

/* ---------- shading ---------- */
/* to_world, include/ray_tracing_math.hpp:51-63 */
V3 to_world(V3 a, V3 N) {
    V3 B, C;
    if (fabsf(N.x) > fabsf(N.y)) {
        float inv = 1.0f / sqrtf(N.x * N.x + N.z * N.z);
        C = v3(N.z * inv, 0.0f, -N.x * inv);
    } else {
        float inv = 1.0f / sqrtf(N.y * N.y + N.z * N.z);
        C = v3(0.f, N.z * inv, -N.y * inv);
    }
    B = cross(C, N);
    return add(add(scale(B, a.x), scale(C, a.y)), scale(N, a.z));
}
/* schlick, include/ray_tracing_math.hpp:65-69 -- pow(float,int) promotes to double on the host */
float schlick(float cosine, float ref_idx) {
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 *= r0;
    return (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cosine), 5.0));
}
/* refract, include/ray_tracing_math.hpp:71-80 -- `1.0 - ...` is a double subtraction */
bool refract(V3 v, V3 n, float ni_over_nt, V3* out) {
    V3 uv = normalize(v);
    float dt = dot(uv, n);
    float disc = (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1 - dt * dt)));
    if (disc > 0) {
        *out = normalize(sub(scale(sub(uv, scale(n, dt)), ni_over_nt), scale(n, sqrtf(disc))));
        return true;
    }
    return false;
}
/* the lobe sample shared by all four shaders (material.cu:24-27,45-48,78-80,85-87,105-107):
 * z = pow(x1, 1.0/e) in DOUBLE, r = sqrt(1-z*z), phi = 2*M_PI*x2 (double product) */
V3 lobe(float x1, float x2, double inv_exponent) {
    float z = (float)pow((double)x1, inv_exponent);
    float r = sqrtf(1.0f - z * z);
    float phi = (float)(2 * RT_PI * (double)x2);
    return v3(r * cosf(phi), r * sinf(phi), z);
}
struct Scatter { V3 bounce, atten, radiance; int draws; };
inline V3 kd_of(const Mat& m) { return v3(m.kd[0], m.kd[1], m.kd[2]); }  /* GetKd, material.cu:11-18 (no textures bound) */

/* DiffuseHitShader, material.cu:20-38 */
void shade_diffuse(const Mat& m, V3 normal, const float* u, Scatter& o) {
    float x1 = u[0], x2 = u[1]; o.draws = 2;
    V3 local = lobe(x1, x2, 1.0 / 2);
    o.bounce = to_world(local, normal);
    o.atten = dot(normal, o.bounce) > 0.0f ? kd_of(m) : v3(0, 0, 0);
    o.radiance = scale(kd_of(m), m.emit);
}
/* MirrorHitShader, material.cu:40-64 -- dispatched for MaterialType::Metal (material.cu:151-153) */
void shade_mirror_fn(const Mat& m, V3 normal, V3 in_dir, const float* u, Scatter& o) {
    float alpha = powf(1000.0f, m.smooth);
    float x1 = u[0], x2 = u[1]; o.draws = 2;
    V3 local = lobe(x1, x2, 1.0 / (double)alpha);
    V3 wo = to_world(local, reflect(in_dir, normal));
    o.atten = dot(normal, wo) > 0.0f ? kd_of(m) : v3(0, 0, 0);
    o.bounce = wo;
    o.radiance = scale(kd_of(m), m.emit);
}
/* MetalHitShader, material.cu:66-99 -- dispatched for MaterialType::Mirror (material.cu:154-156) */
void shade_metal_fn(const Mat& m, V3 normal, V3 in_dir, const float* u, Scatter& o) {
    float x1 = u[0], x2 = u[1];
    float alpha = powf(1000.0f, m.smooth);
    o.draws = 3;
    if (u[2] < m.refl) {
        V3 local = lobe(x1, x2, 1.0 / (double)alpha);
        o.bounce = to_world(local, reflect(in_dir, normal));
    } else {
        V3 local = lobe(x1, x2, 1.0 / 2.0);
        o.bounce = to_world(local, normal);
    }
    o.atten = dot(o.bounce, normal) < 0 ? v3(0, 0, 0) : kd_of(m);
    o.radiance = scale(kd_of(m), m.emit);
}
/* GlassHitShader, material.cu:101-143 */
void shade_glass(const Mat& m, V3 normal, V3 in_dir, const float* u, Scatter& o) {
    float x1 = u[0], x2 = u[1]; o.draws = 3;
    float alpha = powf(1000.0f, m.smooth);
    V3 local = lobe(x1, x2, 1.0 / (double)alpha);
    V3 outward, refracted = v3(0, 0, 0);
    float ni_over_nt, reflect_prob, cosine;
    in_dir = normalize(in_dir);
    if (dot(in_dir, normal) > 0) {
        outward = neg(normal);
        ni_over_nt = m.ior;
        cosine = dot(in_dir, normal);
        cosine = sqrtf(1 - m.ior * m.ior * (1 - cosine * cosine));
    } else {
        outward = normal;
        ni_over_nt = 1.f / m.ior;
        cosine = -dot(in_dir, normal);
    }
    if (refract(in_dir, outward, ni_over_nt, &refracted)) reflect_prob = schlick(cosine, m.ior);
    else reflect_prob = 1.0f;
    if (u[2] < reflect_prob) o.bounce = to_world(local, reflect(in_dir, normal));
    else o.bounce = to_world(local, refracted);
    o.atten = kd_of(m);
    o.radiance = scale(kd_of(m), m.emit);
}
/* Material::EvalAttenuationAndCreateRay, material.cu:145-163 (note the Metal<->Mirror swap) */
void shade(const Mat& m, V3 normal, V3 in_dir, const float* u, Scatter& o) {
    switch (m.type) {
        case DIFFUSE: shade_diffuse(m, normal, u, o); break;
        case METAL: shade_mirror_fn(m, normal, in_dir, u, o); break;
        case MIRROR: shade_metal_fn(m, normal, in_dir, u, o); break;
        case GLASS: shade_glass(m, normal, in_dir, u, o); break;
        default: shade_diffuse(m, normal, u, o);
    }
}

/* ---------- sky: Miss (path_tracer.cu:117-122) + texture unit (textures.cu:46-50,68-71) ---------- */
struct Sky { int w, h; std::vector<uint8_t> rgba; };
inline int mirror_index(int i, int n) {
    if (i < 0) i = -1 - i;
    if (i >= n) i = 2 * n - 1 - i;
    if (i < 0) i = 0;
    if (i >= n) i = n - 1;
    return i;
}
inline float mirror_coord(float x) {
    float f = floorf(x), fr = x - f;
    return (((long long)f) & 1LL) ? 1.0f - fr : fr;
}
V3 tex_lookup(const Sky& t, float u, float v) {
    if (t.w == 0) return v3(0, 0, 0);
    float x = mirror_coord(u) * (float)t.w - 0.5f, y = mirror_coord(v) * (float)t.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = floorf((x - fx) * 256.0f + 0.5f) * (1.0f / 256.0f);  /* 8 fractional weight bits */
    float b = floorf((y - fy) * 256.0f + 0.5f) * (1.0f / 256.0f);
    int i0 = mirror_index((int)fx, t.w), i1 = mirror_index((int)fx + 1, t.w);
    int j0 = mirror_index((int)fy, t.h), j1 = mirror_index((int)fy + 1, t.h);
    float w00 = (1.f - a) * (1.f - b), w10 = a * (1.f - b), w01 = (1.f - a) * b, w11 = a * b;
    float c[3];
    for (int k = 0; k < 3; ++k) {
        float t00 = t.rgba[4 * ((size_t)j0 * t.w + i0) + k] * (1.0f / 255.0f);
        float t10 = t.rgba[4 * ((size_t)j0 * t.w + i1) + k] * (1.0f / 255.0f);
        float t01 = t.rgba[4 * ((size_t)j1 * t.w + i0) + k] * (1.0f / 255.0f);
        float t11 = t.rgba[4 * ((size_t)j1 * t.w + i1) + k] * (1.0f / 255.0f);
        c[k] = w00 * t00 + w10 * t10 + w01 * t01 + w11 * t11;
    }
    return v3(c[0], c[1], c[2]);
}
V3 miss(const Sky& sky, V3 dir) {
    V3 d = normalize(dir);
    float v = (float)((double)asinf(d.z) / RT_PI + 0.5);
    float u = (float)((double)(atanf(d.y / d.x) / 2) / RT_PI);
    return tex_lookup(sky, u, v);
}

/* ---------- camera: MotionalCamera::GetCopy / RayGen, cuSrc/motional_camera.cu:177-213 ---------- */
struct Cam {
    int w, h; V3 origin, look_at, vup; float fov, lens_radius, dist;
    V3 u, v, wv, top_left, horizontal, vertical;
};
void cam_setup(Cam& c) {
    float theta = (float)((double)c.fov * RT_PI / 180);           /* view_fov_ * M_PI / 180 : double, stored float */
    float aspect = (float)c.w / (float)c.h;
    float half_h = tanf(theta / 2);
    float half_w = aspect * half_h;
    c.wv = normalize(sub(c.origin, c.look_at));
    c.u = normalize(cross(c.vup, c.wv));
    c.v = cross(c.wv, c.u);
    c.dist = length(sub(c.origin, c.look_at));
    c.top_left = sub(add(sub(c.origin, scale(c.u, half_w * c.dist)), scale(c.v, half_h * c.dist)), scale(c.wv, c.dist));
    c.horizontal = scale(c.u, 2 * half_w * c.dist);
    c.vertical = scale(c.v, -2 * half_h * c.dist);
}
Ray raygen(const Cam& c, int x, int y, const float* U /* rd.x, rd.y, rd.z */) {
    Ray r;
    V3 rd = v3(c.lens_radius * U[0], c.lens_radius * U[1], c.lens_radius * U[2]);
    V3 offset = add(scale(c.u, rd.x), scale(c.v, rd.y));
    float dx = (float)x / (float)c.w, dy = (float)y / (float)c.h;
    r.o = add(c.origin, offset);
    r.d = normalize(sub(sub(add(add(c.top_left, scale(c.horizontal, dx)), scale(c.vertical, dy)), c.origin), offset));
    r.tmin = 0.f; r.tmax = RT_TMAX;
    return r;
}

Scene g_scene;
Sky g_sky = {0, 0, std::vector<uint8_t>()};
Cam g_cam;

}  // namespace

extern "C" {

struct orc_stats { unsigned long long paths, segments, misses, depth_exhausted; };

void orc_reset(void) { g_scene.prims.clear(); g_scene.nodes.clear(); }

int orc_add_object(int prim_type, const float* center, float radius, float y_pos, float height, const Mat* m) {
    Prim p; memset(&p, 0, sizeof(p));
    p.type = prim_type; p.mat = *m; p.c = v3(center[0], center[1], center[2]);
    p.radius = radius; p.y_pos = y_pos; p.height = height; p.id = (int)g_scene.prims.size();
    g_scene.prims.push_back(p);
    return p.id;
}
int orc_build(void) {
    std::vector<int> order(g_scene.prims.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    g_scene.nodes.clear();
    divide(g_scene, order, 0, (int)order.size());
    return (int)g_scene.nodes.size();
}
void orc_leaf_order(int32_t* rank_of_prim) {
    std::vector<int> st; st.push_back(0); int rank = 0;
    while (!st.empty()) {
        int n = st.back(); st.pop_back();
        if (n == -1) continue;
        if (g_scene.nodes[n].leaf) rank_of_prim[g_scene.nodes[n].prim] = rank++;
        st.push_back(g_scene.nodes[n].l); st.push_back(g_scene.nodes[n].r);
    }
}
void orc_set_sky(const uint8_t* rgba, int w, int h, int quarter_rows) {
    g_sky.w = w; g_sky.h = h; g_sky.rgba.assign((size_t)w * h * 4, 0);
    size_t row_bytes = quarter_rows ? (size_t)w : (size_t)w * 4;   /* textures.cu:32 copies `width` bytes per row */
    for (int y = 0; y < h; ++y) memcpy(&g_sky.rgba[(size_t)y * w * 4], rgba + (size_t)y * w * 4, row_bytes);
}
void orc_trace(int n, const float* rays8, int32_t* prim, float* t, float* normal3, float* hitpos3) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        Ray ray; ray.o = v3(rays8[8 * i], rays8[8 * i + 1], rays8[8 * i + 2]);
        ray.d = v3(rays8[8 * i + 3], rays8[8 * i + 4], rays8[8 * i + 5]);
        ray.tmin = rays8[8 * i + 6]; ray.tmax = rays8[8 * i + 7];
        Hit h; h.normal = v3(0, 0, 0); h.pos = v3(0, 0, 0);
        float tt;
        int p = trace(g_scene, ray, h, &tt);
        prim[i] = p; t[i] = tt;
        if (p < 0) { h.normal = v3(0, 0, 0); h.pos = v3(0, 0, 0); }
        normal3[3 * i] = h.normal.x; normal3[3 * i + 1] = h.normal.y; normal3[3 * i + 2] = h.normal.z;
        hitpos3[3 * i] = h.pos.x; hitpos3[3 * i + 1] = h.pos.y; hitpos3[3 * i + 2] = h.pos.z;
    }
}
void orc_shade(int n, const Mat* mats, const int32_t* mat_idx, const float* pos3, const float* nrm3,
               const float* in3, const float* u3, float* bounce3, float* atten3, float* rad3, int32_t* ndraws) {
    (void)pos3;
    for (int i = 0; i < n; ++i) {
        Scatter s;
        shade(mats[mat_idx[i]], v3(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2]),
              v3(in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]), u3 + 3 * i, s);
        bounce3[3 * i] = s.bounce.x; bounce3[3 * i + 1] = s.bounce.y; bounce3[3 * i + 2] = s.bounce.z;
        atten3[3 * i] = s.atten.x; atten3[3 * i + 1] = s.atten.y; atten3[3 * i + 2] = s.atten.z;
        rad3[3 * i] = s.radiance.x; rad3[3 * i + 1] = s.radiance.y; rad3[3 * i + 2] = s.radiance.z;
        ndraws[i] = s.draws;
    }
}
void orc_camera_setup(int w, int h, const float* origin, const float* look_at, float fov_deg, float lens_radius,
                      float* out19) {
    g_cam.w = w; g_cam.h = h; g_cam.origin = v3(origin[0], origin[1], origin[2]);
    g_cam.look_at = v3(look_at[0], look_at[1], look_at[2]); g_cam.vup = v3(0.f, 1.f, 0.f);
    g_cam.fov = fov_deg; g_cam.lens_radius = lens_radius;
    cam_setup(g_cam);
    if (out19) {
        const V3* v[6] = {&g_cam.u, &g_cam.v, &g_cam.wv, &g_cam.top_left, &g_cam.horizontal, &g_cam.vertical};
        for (int k = 0; k < 6; ++k) { out19[3 * k] = v[k]->x; out19[3 * k + 1] = v[k]->y; out19[3 * k + 2] = v[k]->z; }
        out19[18] = g_cam.dist;
    }
}
/* MotionalCamera::MoveEye* / RotateAround* / ScaleFov, cuSrc/motional_camera.cu:76-168 (move_speed_ = 50) */
void orc_camera_op(float* pose7, int op, float arg) {
    V3 o = v3(pose7[0], pose7[1], pose7[2]), a = v3(pose7[3], pose7[4], pose7[5]);
    V3 vup = v3(0.f, 1.f, 0.f);
    float fov = pose7[6];
    const float speed = 50.f;
    if (op <= 3) {
        V3 w = normalize(sub(o, a));
        V3 left = neg(normalize(cross(vup, w)));
        V3 dir = left;
        if (op >= 2) dir = neg(normalize(cross(left, vup)));        /* back */
        V3 step = scale(dir, arg * speed);
        if (op == 0 || op == 3) { o = add(o, step); a = add(a, step); }
        else { o = sub(o, step); a = sub(a, step); }
    } else if (op == 4) { V3 s = scale(vup, arg * speed); o = add(o, s); a = add(a, s); }
    else if (op == 5) { V3 s = scale(vup, arg * speed); o = sub(o, s); a = sub(a, s); }
    else if (op <= 9) {
        a = add(o, normalize(sub(a, o)));
        V3 w = normalize(sub(a, o));
        V3 left = normalize(cross(vup, w));
        if (op == 6 || op == 7) {
            V3 up = normalize(cross(w, left));
            a = (op == 6) ? add(a, scale(up, arg)) : sub(a, scale(up, arg));
        } else {
            a = (op == 8) ? add(a, scale(left, arg)) : sub(a, scale(left, arg));
        }
        a = add(o, normalize(sub(a, o)));
    } else if (op == 10) {
        fov = (float)((double)fov + (double)arg * RT_PI / 180.0f);   /* view_fov_ += d * M_PI / 180.0f */
    }
    pose7[0] = o.x; pose7[1] = o.y; pose7[2] = o.z; pose7[3] = a.x; pose7[4] = a.y; pose7[5] = a.z; pose7[6] = fov;
}
void orc_raygen(int n, const int32_t* xy, const float* u3, float* rays8) {
    for (int i = 0; i < n; ++i) {
        Ray r = raygen(g_cam, xy[2 * i], xy[2 * i + 1], u3 + 3 * i);
        rays8[8 * i] = r.o.x; rays8[8 * i + 1] = r.o.y; rays8[8 * i + 2] = r.o.z;
        rays8[8 * i + 3] = r.d.x; rays8[8 * i + 4] = r.d.y; rays8[8 * i + 5] = r.d.z;
        rays8[8 * i + 6] = r.tmin; rays8[8 * i + 7] = r.tmax;
    }
}
void orc_sky_lookup(int n, const float* dir3, float* rgb3) {
    for (int i = 0; i < n; ++i) {
        V3 c = miss(g_sky, v3(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]));
        rgb3[3 * i] = c.x; rgb3[3 * i + 1] = c.y; rgb3[3 * i + 2] = c.z;
    }
}
/* SamplePixel, cuSrc/path_tracer.cu:124-175, for samples [first, first+spp) of every pixel; per-pixel
 * fp32 sum in sample order.  All rows rendered (the reference's W/16 x H/16 grid drops trailing rows). */
void orc_render(int spp, int first_sample, int max_depth, unsigned long long seed, int nthreads,
                float* sum_rgb, float* normal_rgb, orc_stats* stats) {
    const int W = g_cam.w, H = g_cam.h;
    unsigned long long segs = 0, misses = 0, exhausted = 0;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : segs, misses, exhausted)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint32_t pixel = (uint32_t)(y * W + x);
            V3 sum = v3(0, 0, 0), first_normal = v3(0, 0, 0);
            for (int s = first_sample; s < first_sample + spp; ++s) {
                float U[4];
                pt_philox_uniform4(seed, pixel, (uint32_t)s, 0u, U);
                Ray ray = raygen(g_cam, x, y, U);
                V3 radiance = v3(0, 0, 0), atten = v3(1, 1, 1), normals = v3(0, 0, 0);
                int depth = 0; bool first = false, missed = false;
                while (depth < max_depth) {
                    Hit h; float t;
                    ++segs;
                    int p = trace(g_scene, ray, h, &t);
                    V3 seg_rad, seg_att, hit_pos, bounce;
                    if (p >= 0) {
                        pt_philox_uniform4(seed, pixel, (uint32_t)s, 1u + (uint32_t)depth, U);
                        Scatter sc; shade(g_scene.prims[p].mat, h.normal, ray.d, U, sc);
                        seg_rad = sc.radiance; seg_att = sc.atten; hit_pos = h.pos; bounce = sc.bounce;
                        radiance = add(radiance, mul(atten, seg_rad));
                        atten = mul(atten, seg_att);
                        if (!first) { normals = add(normals, h.normal); first = true; }
                        ray.o = hit_pos; ray.d = normalize(bounce); ray.tmin = RT_BOUNCE_TMIN; ray.tmax = RT_TMAX;
                        depth++;
                    } else {
                        /* Miss sets recursion_depth = 32 (path_tracer.cu:121): the loop ends after this segment;
                         * payload.attenuation keeps its previous value but is never used again */
                        seg_rad = miss(g_sky, ray.d);
                        radiance = add(radiance, mul(atten, seg_rad));
                        if (!first) { normals = add(normals, neg(ray.d)); first = true; }
                        missed = true; ++misses;
                        break;
                    }
                }
                if (!missed) ++exhausted;
                sum = add(sum, radiance);
                first_normal = normals;
            }
            sum_rgb[3 * (size_t)pixel] = sum.x; sum_rgb[3 * (size_t)pixel + 1] = sum.y; sum_rgb[3 * (size_t)pixel + 2] = sum.z;
            if (normal_rgb) {
                normal_rgb[3 * (size_t)pixel] = first_normal.x; normal_rgb[3 * (size_t)pixel + 1] = first_normal.y;
                normal_rgb[3 * (size_t)pixel + 2] = first_normal.z;
            }
        }
    if (stats) { stats->paths = (unsigned long long)W * H * spp; stats->segments = segs; stats->misses = misses;
                 stats->depth_exhausted = exhausted; }
}
/* Debug / test aid: the ray every pixel's path (sample `sample`) shoots as its segment number `seg`
 * (0 = camera ray), or valid=0 when the path ended before that.  Same walk as orc_render. */
void orc_segment_rays(int seg, int sample, int max_depth, unsigned long long seed, float* rays8, int32_t* valid) {
    const int W = g_cam.w, H = g_cam.h;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint32_t pixel = (uint32_t)(y * W + x);
            float U[4];
            pt_philox_uniform4(seed, pixel, (uint32_t)sample, 0u, U);
            Ray ray = raygen(g_cam, x, y, U);
            int depth = 0; bool ok = true;
            while (depth < seg) {
                if (depth >= max_depth) { ok = false; break; }
                Hit h; float t;
                int p = trace(g_scene, ray, h, &t);
                if (p < 0) { ok = false; break; }
                pt_philox_uniform4(seed, pixel, (uint32_t)sample, 1u + (uint32_t)depth, U);
                Scatter sc; shade(g_scene.prims[p].mat, h.normal, ray.d, U, sc);
                ray.o = h.pos; ray.d = normalize(sc.bounce); ray.tmin = RT_BOUNCE_TMIN; ray.tmax = RT_TMAX;
                depth++;
            }
            valid[pixel] = ok ? 1 : 0;
            float* r = rays8 + 8 * (size_t)pixel;
            r[0] = ray.o.x; r[1] = ray.o.y; r[2] = ray.o.z; r[3] = ray.d.x; r[4] = ray.d.y; r[5] = ray.d.z;
            r[6] = ray.tmin; r[7] = ray.tmax;
        }
}
/* Denoising (cuSrc/path_tracer.cu:177-239) + Mix (cuSrc/path_tracer.cu:241-254) for one sample.
 * color3 / normal3: SamplePixel's render_target and normal buffer (W*H*3); depth is constant (see
 * pt_display.cuh) so p_w = exp(-0/pi) = 1.  mix3 is the running mean (in/out), idx the 1-based
 * cur_sample_idx_.  `exp(-(dist2) / M_PI)` is a double exp; taps wrap across row ends exactly like the
 * reference's `cur_off = v*width+u` bound check (:212-216).  bgra[4p+3] is never written by the reference;
 * this restatement leaves it untouched too. */
void orc_denoise_mix(int W, int H, const float* color3, const float* normal3, float* mix3, uint8_t* bgra, int idx) {
    static const float kern[5][5] = {{1.f, 4.f, 7.f, 4.f, 1.f}, {4.f, 16.f, 26.f, 16.f, 4.f}, {7.f, 26.f, 41.f, 26.f, 7.f},
                                     {4.f, 16.f, 26.f, 16.f, 4.f}, {1.f, 4.f, 7.f, 4.f, 1.f}};
    std::vector<float> den((size_t)W * H * 3);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int off = y * W + x;
            V3 cval = v3(color3[3 * off], color3[3 * off + 1], color3[3 * off + 2]);
            V3 nval = v3(normal3[3 * off], normal3[3 * off + 1], normal3[3 * off + 2]);
            V3 sum = v3(0, 0, 0); float cum_w = 0.0f;
            for (int i = 0; i < 5; ++i)
                for (int j = 0; j < 5; ++j) {
                    int u = x + (i - 2), v = y + (j - 2);
                    int cur = v * W + u;
                    float c_w, n_w, p_w; V3 ctmp;
                    if (cur < 0 || cur >= W * H) { c_w = n_w = p_w = 0; ctmp = v3(0, 0, 0); }
                    else {
                        ctmp = v3(color3[3 * cur], color3[3 * cur + 1], color3[3 * cur + 2]);
                        V3 t = sub(cval, ctmp);
                        float dist2 = dot(t, t);
                        c_w = (float)fmin(exp(-(double)dist2 / RT_PI), 1.0);
                        t = sub(nval, v3(normal3[3 * cur], normal3[3 * cur + 1], normal3[3 * cur + 2]));
                        dist2 = fmaxf(dot(t, t), 0.0f);
                        n_w = (float)fmin(exp(-(double)dist2 / RT_PI), 1.0);
                        p_w = 1.0f;
                    }
                    float weight = c_w * n_w * p_w;
                    sum = add(sum, scale(ctmp, weight * kern[i][j]));
                    cum_w += weight * kern[i][j];
                }
            V3 d = divs(sum, cum_w);
            den[3 * off] = d.x; den[3 * off + 1] = d.y; den[3 * off + 2] = d.z;
        }
    const float t = 1.f / (float)idx;
    for (int off = 0; off < W * H; ++off)
        for (int k = 0; k < 3; ++k) {
            float a = mix3[3 * off + k];
            float b = fmaxf(0.f, fminf(den[3 * off + k], 1.f));                 /* clamp, helper_math.h:1168-1171 */
            float m = a + t * (b - a);                                        /* lerp,  helper_math.h:1146-1149 */
            mix3[3 * off + k] = m;
            bgra[4 * off + (2 - k)] = (uint8_t)(255.99f * m);
        }
}
int orc_sizeof(int) { return 0; }
int orc_max_threads(void) { return omp_get_max_threads(); }
}
