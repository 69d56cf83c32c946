// tcgen05 throughput probe (not product): cycles per 128xNx8 tf32 MMA (SWIZZLE_NONE K-major operands in smem)
// as a function of N, number of issuing warps and number of independent accumulators.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tc_rate tools/tc_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../giga_b200/csrc/tc.cuh"
using namespace giga;

__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }

// each issuing warp w issues `iters` rounds of `chains` MMAs (accumulators w*chains + c), then commits
template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int nwarps, int chains, int a_stride_rows, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 48 * 1024 / 4; e += 128) reinterpret_cast<float*>(smem)[e] = 1.0f;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (tid == 0) tc::mbar_init(&bar, nwarps);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 32 * 1024;
  const uint32_t ks_a = a_stride_rows * 16, ks_b = N * 16;
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, N);
  long long t0 = clock64();
  if (warp < nwarps) {
    if (tc::elect_one()) {
      for (int it = 0; it < iters; ++it)
        for (int c = 0; c < chains; ++c) {
          const uint32_t d = tmem + ((warp * chains + c) * N) % (512 - N + 1 > 0 ? 512 : 512);
          tc::mma_tf32(tmem + (((warp * chains + c) * N) & 511) , tc::make_desc(a0 + ((it * 7 + c) % 9) * 16, ks_a, 128), tc::make_desc(b0, ks_b, 128), IDESC, it > 0 ? 1u : 0u);
          (void)d;
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

template <int N>
void run(int nwarps, int chains, int a_rows, int ctas_per_sm) {
  long long* d; cudaMalloc(&d, 8 * 1024);
  const int iters = 200;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  rate_kernel<N><<<148 * ctas_per_sm, 128, 48 * 1024>>>(iters, nwarps, chains, a_rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024]; cudaMemcpy(h, d, 8 * 148 * ctas_per_sm, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148 * ctas_per_sm; ++i) avg += h[i]; avg /= 148 * ctas_per_sm;
  const double per = avg / (iters * chains * nwarps);
  printf("N=%3d warps=%d chains/warp=%d a_rows=%d ctas/sm=%d : %8.1f cycles / MMA (CTA-level), %8.1f SM-level  %s\n", N, nwarps, chains, a_rows,
         ctas_per_sm, per, per / ctas_per_sm, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  // TMEM alloc of 512 columns per CTA: only 1 CTA/SM can be resident at a time for this probe (ctas/sm>1 serialise)
  for (int chains : {1, 2, 4, 8}) run<32>(1, chains, 343, 1);
  for (int w : {2, 3, 4}) run<32>(w, 4, 343, 1);
  run<64>(1, 4, 343, 1); run<64>(2, 4, 343, 1);
  run<128>(1, 4, 343, 1); run<128>(2, 2, 343, 1);
  run<256>(1, 2, 343, 1);
  printf("-- conv-like issue patterns\n");
  run<32>(3, 2, 343, 1);   // 40^2 layers today: 3 issuing warps x 2 chains (mt)
  run<32>(3, 1, 343, 1);
  run<64>(3, 1, 175, 1);   // 20^2 / 10^2 layers today: 3 warps x 1 chain, N=64
  run<64>(3, 2, 175, 1);
  run<64>(2, 2, 343, 1);   // N-concat D1 for NTILE=32, NT=2
  run<128>(2, 1, 175, 1);
  run<128>(1, 1, 175, 1);
  run<128>(1, 2, 175, 1);
  return 0;
}
