/*
 * pt_philox.h -- TEST INFRASTRUCTURE (oracle side).  Not linked into the product.
 *
 * Counter-based Philox4x32-10 stream shared by the two CPU oracles
 * (oracle/ref_build: the reference's own sources compiled for the host;
 *  oracle/pt_oracle.c: the plain-C restatement).
 *
 * The reference draws its uniforms from a per-pixel cuRAND XORWOW state seeded
 * with the host clock (include/ray_tracing_math.hpp:82-104, cuSrc/path_tracer.cu:36-42),
 * so its output is not reproducible.  north_star replaces that with a stateless
 * stream keyed by (pixel, sample, bounce); the algorithm below is the published
 * Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 "philox4x32_R(10,..)"),
 * pinned by the Random123 known-answer vectors in tests/test_oracle_philox.py.
 *
 * Stream layout (must match cpppathtracer_b200/csrc/pt_rng.cuh):
 *   key     = (seed_lo, seed_hi)
 *   counter = (pixel_index, sample_index, slot, 0)
 *   slot 0        : MotionalCamera::RayGen   (3 draws, motional_camera.cu:204)
 *   slot 1 + k    : shader of path segment k (2 or 3 draws, material.cu:23,44,68,77,103,133)
 *   uniform       : u = x * 2^-32 + 2^-33  in (0,1]   (curand_uniform convention)
 */
#ifndef PT_PHILOX_H
#define PT_PHILOX_H
#include <stdint.h>

#define PT_PHILOX_M0 0xD2511F53u
#define PT_PHILOX_M1 0xCD9E8D57u
#define PT_PHILOX_W0 0x9E3779B9u
#define PT_PHILOX_W1 0xBB67AE85u

static inline void pt_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PT_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PT_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PT_PHILOX_W0; k1 += PT_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* (0,1] mapping: x*2^-32 is exact in fp32 after the u32->f32 conversion, one rounding in the add. */
static inline float pt_u32_to_uniform(uint32_t x) {
    return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}

static inline void pt_philox_uniform4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, float u[4]) {
    uint32_t ctr[4] = {pixel, sample, slot, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    pt_philox4x32_10(ctr, key, o);
    for (int i = 0; i < 4; ++i) u[i] = pt_u32_to_uniform(o[i]);
}
#endif
