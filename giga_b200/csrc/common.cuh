// Shared helpers for the giga_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace giga {

constexpr int G = 40;          // TSDF / plane resolution
constexpr int G2 = G * G;      // 1600
constexpr int G3 = G * G * G;  // 64000
constexpr int C = 32;          // plane feature channels / decoder hidden size

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace giga
