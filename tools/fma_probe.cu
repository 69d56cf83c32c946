// FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a: can the packed form raise the fp32 FMA rate, alone or mixed with scalar FFMA?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fma_probe tools/fma_probe.cu && /tmp/fma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

// MODE 0: 32 scalar FFMA per iteration; 1: 16 FFMA2 (= 32 FMAs); 2: 8 FFMA2 + 16 FFMA (= 32 FMAs); 3: 16 FFMA2 + 16 FFMA (= 48 FMAs)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
  float s[16];
  float2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { s[i] = threadIdx.x * 0.001f + i; v[i] = make_float2(s[i], s[i] + 1.f); }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a), "f"(b));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { fma2(v[i], a2, b2); }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        fma2(v[i], a2, b2);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[2 * i]) : "f"(a), "f"(b));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[2 * i + 1]) : "f"(a), "f"(b));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        fma2(v[i], a2, b2);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a), "f"(b));
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += s[i] + v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int fmas_per_iter, int blocks_per_sm) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * blocks_per_sm, iters = 20000;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * 256);
  probe<MODE><<<blocks, 256>>>(out, 100, 1.0001f, 0.5f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)blocks * 256 * iters * fmas_per_iter;
  const double per_clk_sm = fmas / (ms * 1e-3) / (khz * 1e3) / sms;
  printf("%-34s %d CTAs/SM x 256 thr: %7.3f ms  %6.1f TFLOP/s  %6.1f FMA/clk/SM (at the %d MHz attribute clock)\n", name, blocks_per_sm, ms,
         2.0 * fmas / (ms * 1e-3) / 1e12, per_clk_sm, khz / 1000);
  cudaFree(out);
}

int main() {
  for (int bps : {2, 4, 8}) {
    run<0>("FFMA  (32 scalar)", 32, bps);
    run<1>("FFMA2 (16 packed)", 32, bps);
    run<2>("8 FFMA2 + 16 FFMA", 32, bps);
    run<3>("16 FFMA2 + 16 FFMA", 48, bps);
  }
  return 0;
}
