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

// Programmatic dependent launch (PDL): kernels on the fast path are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel may become resident -- and run its prologue (barrier
// init, TMEM allocation, weight staging) -- while its predecessor in the stream is still draining.  pdl_launch() lets the
// successor start early; pdl_wait() blocks until the predecessor grid has completed and its writes are visible, and must
// precede the first access to anything a previous kernel wrote (or still reads).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// Packed fp32 pair arithmetic (sm_100: FFMA2 / FADD2).  Each component is an IEEE fp32 fma / add, so the
// results are bit-identical to the scalar instructions; the pair form halves the issue slots of FMA-bound loops.
__device__ __forceinline__ void fma2(float2& d, const float2 a, const float s) {   // d += a * (s, s)
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const float2 ss = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&ss)));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void add2(float2& d, const float2 a) {                   // d += a
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(dd) : "l"(*reinterpret_cast<const unsigned long long*>(&a)));
  d = *reinterpret_cast<float2*>(&dd);
}

// ReLU / max that PROPAGATE NaN (fmaxf returns the non-NaN operand and would swallow an fp16-range overflow: a value beyond 65504 becomes the
// operand pair (inf, -inf), the next contraction turns it into NaN, and that NaN must reach the outputs -- giga_ctx_overflow_count)
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float relu_nan(float x) { return max_nan(x, 0.f); }

}  // namespace giga
