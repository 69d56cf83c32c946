// tcgen05 kind::f16 probe (not product): (1) cycles per 128xNx16 fp16 MMA with the SWIZZLE_NONE K-major operand layout the
// kernels use (16-byte k-chunks = 8 halfs), (2) accuracy of the 3-product fp16 split (x = xh + xl, both fp16;
// x*y ~ xh*yh + xl*yh + xh*yl, fp32 accumulate) against fp64 -- the candidate replacement for 3xTF32 (same 11-bit
// significands, twice the K per MMA and half the operand bytes).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tc_f16_probe tools/tc_f16_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../giga_b200/csrc/tc.cuh"
using namespace giga;

__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {   // fp16 A/B (format 0), fp32 accumulate, K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}

template <int N, bool F16>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int nwarps, int chains, int a_stride_rows, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 48 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = F16 ? 0x3c003c00u : 0x3f800000u;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (tid == 0) tc::mbar_init(&bar, nwarps);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 32 * 1024;
  const uint32_t ks_a = a_stride_rows * 16, ks_b = N * 16;
  const uint32_t idesc = F16 ? make_idesc_f16(128, N) : tc::make_idesc_tf32(128, N);
  long long t0 = clock64();
  if (warp < nwarps) {
    if (tc::elect_one()) {
      for (int it = 0; it < iters; ++it)
        for (int c = 0; c < chains; ++c) {
          const uint32_t d = tmem + (((warp * chains + c) * N) & 511);
          const uint64_t ad = tc::make_desc(a0 + ((it * 7 + c) % 9) * 16, ks_a, 128), bd = tc::make_desc(b0, ks_b, 128);
          if (F16) mma_f16(d, ad, bd, idesc, it > 0 ? 1u : 0u);
          else tc::mma_tf32(d, ad, bd, idesc, it > 0 ? 1u : 0u);
        }
      tc::mma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N, bool F16>
void run(int nwarps, int chains, int a_rows) {
  long long* d; cudaMalloc(&d, 8 * 1024);
  const int iters = 200;
  cudaFuncSetAttribute(rate_kernel<N, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  rate_kernel<N, F16><<<148, 128, 48 * 1024>>>(iters, nwarps, chains, a_rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per = avg / (iters * chains * nwarps);
  printf("%s N=%3d warps=%d chains/warp=%d : %7.1f cycles / MMA = %6.2f cycles per unit of K  %s\n", F16 ? "f16 (K=16)" : "tf32 (K=8)", N, nwarps, chains,
         per, per / (F16 ? 16 : 8), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

// ---- accuracy: D[128 x 32] = A[128 x K] . B[32 x K]^T, K = 64, split operands in smem ------------------------------
constexpr int AK = 64;
__global__ void __launch_bounds__(128) acc_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int scale_lo) {
  // smem: A hi/lo: [k/8][128 rows][8 halfs], B hi/lo: [k/8][32 rows][8 halfs]
  __shared__ __align__(128) __half sAh[AK / 8][128][8], sAl[AK / 8][128][8], sBh[AK / 8][32][8], sBl[AK / 8][32][8];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const float ls = scale_lo ? 2048.f : 1.f;
  for (int k = 0; k < AK; ++k) {
    const float v = A[tid * AK + k];
    const __half h = __float2half_rn(v);
    sAh[k / 8][tid][k % 8] = h;
    sAl[k / 8][tid][k % 8] = __float2half_rn((v - __half2float(h)) * ls);
    if (tid < 32) {
      const float w = B[tid * AK + k];
      const __half wh = __float2half_rn(w);
      sBh[k / 8][tid][k % 8] = wh;
      sBl[k / 8][tid][k % 8] = __float2half_rn((w - __half2float(wh)) * ls);
    }
  }
  if (warp == 0) tc::tmem_alloc(&slot, 128);
  if (tid == 0) tc::mbar_init(&bar, 1);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 32);
      for (int ks = 0; ks < AK / 16; ++ks) {
        const uint64_t ah = tc::make_desc(tc::smem_u32(&sAh[2 * ks][0][0]), 128 * 16, 128), al = tc::make_desc(tc::smem_u32(&sAl[2 * ks][0][0]), 128 * 16, 128);
        const uint64_t bh = tc::make_desc(tc::smem_u32(&sBh[2 * ks][0][0]), 32 * 16, 128), bl = tc::make_desc(tc::smem_u32(&sBl[2 * ks][0][0]), 32 * 16, 128);
        mma_f16(tmem, ah, bh, idesc, ks > 0);
        mma_f16(tmem + 32, al, bh, idesc, ks > 0);
        mma_f16(tmem + 64, ah, bl, idesc, ks > 0);
      }
      tc::mma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float v[32], a[32], b[32];
  const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
  tc::tmem_ld32(row, v); tc::tmem_ld32(row + 32, a); tc::tmem_ld32(row + 64, b);
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j] + (a[j] + b[j]) / ls;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

void accuracy(float amp_a, float amp_b, int scale_lo) {
  static float hA[128 * AK], hB[32 * AK], hD[128 * 32];
  for (auto& v : hA) v = amp_a * (2.f * rand() / RAND_MAX - 1.f);
  for (auto& v : hB) v = amp_b * (2.f * rand() / RAND_MAX - 1.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  acc_kernel<<<1, 128>>>(dA, dB, dD, scale_lo);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
  double worst = 0, worst32 = 0, scale = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 32; ++j) {
      double r = 0; float r32 = 0;
      for (int k = 0; k < AK; ++k) { r += (double)hA[i * AK + k] * hB[j * AK + k]; r32 = fmaf(hA[i * AK + k], hB[j * AK + k], r32); }
      worst = fmax(worst, fabs(hD[i * 32 + j] - r)); worst32 = fmax(worst32, fabs(r32 - r)); scale = fmax(scale, fabs(r));
    }
  printf("3xFP16 split K=%d amp_a=%g amp_b=%g lo_scale=%d: max |err| %.3e (fp32 fma chain: %.3e), max |D| %.3e -> rel %.2e  %s\n", AK, amp_a, amp_b,
         scale_lo, worst, worst32, scale, worst / scale, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  for (int w : {1, 2, 3}) { run<32, false>(w, 4, 343); run<32, true>(w, 4, 343); }
  run<64, false>(2, 4, 343); run<64, true>(2, 4, 343);
  run<128, false>(2, 2, 343); run<128, true>(2, 2, 343);
  run<160, true>(2, 1, 343); run<256, true>(1, 2, 343);
  accuracy(1.f, 1.f, 0); accuracy(1.f, 1.f, 1);
  accuracy(20.f, 0.1f, 0); accuracy(20.f, 0.1f, 1);
  accuracy(0.01f, 0.1f, 0); accuracy(0.01f, 0.1f, 1);
  accuracy(1e-3f, 1e-2f, 0); accuracy(1e-3f, 1e-2f, 1);
  return 0;
}
