/*
 * cuda_shim.h -- TEST / BASELINE INFRASTRUCTURE.  Force-included when the reference's CUDA sources are
 * rebuilt UNMODIFIED by nvcc for sm_100 into oracle/_ref/libpt_ref_cuda.so (the "reference's own CUDA
 * kernels rebuilt for sm_100" baseline of BASELINE.md and the device-side level-1 oracle).
 * Supplies only what MSVC / Windows headers supplied implicitly (see host_shim.h for the list).
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
inline __host__ __device__ float3 operator-(const float3& a) { return make_float3(-a.x, -a.y, -a.z); }
