// Hop-latency probe (not product): where do the cycles of one decoder "layer round trip" go?
//   compute threads: [t0] tcgen05.ld -> [t1] wait::ld -> ALU -> tcgen05.st -> [t2] wait::st + fence -> [t3] mbarrier.arrive
//   issuer lane:     wakes from try_wait [t4] -> 6 TS MMAs issued [t5] -> commit
//   compute threads: wake from try_wait on the commit barrier [t6]
// All stamps are clock64 of the same SM; the kernel prints averages of the deltas over many rounds for NW = 4 or 8 compute warps.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tc_hop_probe tools/tc_hop_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../giga_b200/csrc/tc.cuh"
using namespace giga;

__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void wait_plain(uint64_t* bar, uint32_t parity) {   // try_wait without a suspend-time hint
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_test(uint64_t* bar, uint32_t parity) {    // pure spin on test_wait
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
}

// WAITMODE 0: tc::mbar_wait (suspend hint), 1: plain try_wait loop, 2: test_wait spin
template <int NW, int WAITMODE>
__global__ void __launch_bounds__(32 * NW + 32) hop_kernel(int rounds, long long* out) {
  __shared__ __align__(128) uint8_t sB[8192];
  __shared__ uint64_t acc_full, a_ready;
  __shared__ uint32_t slot;
  __shared__ long long stamp[8];   // written by thread 0 (compute) and the issuer lane
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 8192 / 4; e += 32 * NW + 32) reinterpret_cast<uint32_t*>(sB)[e] = 0x2c002c00u;
  if (warp == NW) tc::tmem_alloc(&slot, 64);
  if (tid == 0) { tc::mbar_init(&acc_full, 1); tc::mbar_init(&a_ready, 32 * NW); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    if (WAITMODE == 0) tc::mbar_wait(bar, parity);
    else if (WAITMODE == 1) wait_plain(bar, parity);
    else wait_test(bar, parity);
  };
  long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (warp == NW) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(128, 32);
      const uint64_t bh = tc::make_desc(tc::smem_u32(sB), 512, 128);
      for (int r = 0; r < rounds; ++r) {
        wait(&a_ready, (uint32_t)(r & 1));
        const long long t4 = clock64();
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mma_ts(tmem, tmem + 32 + ks * 16, bh + ks * 64, idesc, ks > 0 ? 1u : 0u);
          mma_ts(tmem, tmem + 32 + ks * 16 + 8, bh + ks * 64, idesc, 1u);
          mma_ts(tmem, tmem + 32 + ks * 16, bh + 128 + ks * 64, idesc, 1u);
        }
        tc::mma_commit(&acc_full);
        const long long t5 = clock64();
        stamp[4] = t4; stamp[5] = t5;
      }
    }
    __syncwarp();
  } else {
    const int half = (warp >> 2) & 1;   // NW = 8: two threads per row, 16 columns each
    const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (NW == 8 ? 16 * half : 0);
    float keep = 0.f;
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      tmem_st8(row + 32, z); tmem_st8(row + 40, z);
      if (NW == 4) { tmem_st8(row + 48, z); tmem_st8(row + 56, z); }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc::fence_before_sync();
      mbar_arrive_(&a_ready);
    }
    for (int r = 0; r < rounds; ++r) {
      wait(&acc_full, (uint32_t)(r & 1));
      const long long t6 = clock64();
      tc::fence_after_sync();
      const long long t0 = clock64();
      float v[16];
      tc::tmem_ld16(row, v);
      const long long t1 = clock64();
      uint32_t h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x0 = fmaxf(fmaf(v[2 * j], 0.03125f, 0.01f), 0.f), x1 = fmaxf(fmaf(v[2 * j + 1], 0.03125f, 0.01f), 0.f);
        const __half2 hh = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
        keep += x0;
      }
      if (NW == 4) {   // one thread per row: both column halves
        float w[16];
        tc::tmem_ld16(row + 16, w);
        uint32_t h2[8], l2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x0 = fmaxf(fmaf(w[2 * j], 0.03125f, 0.01f), 0.f), x1 = fmaxf(fmaf(w[2 * j + 1], 0.03125f, 0.01f), 0.f);
          const __half2 hh = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          h2[j] = *reinterpret_cast<const uint32_t*>(&hh);
          l2[j] = *reinterpret_cast<const uint32_t*>(&ll);
          keep += x0;
        }
        tmem_st8(row + 48, h2); tmem_st8(row + 56, l2);
      }
      tmem_st8(row + 32, h); tmem_st8(row + 40, l);
      const long long t2 = clock64();
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc::fence_before_sync();
      const long long t3 = clock64();
      if (r + 1 < rounds) mbar_arrive_(&a_ready);
      if (tid == 0 && r > 0) {
        acc[0] += t1 - t0;            // tcgen05.ld + wait
        acc[1] += t2 - t1;            // ALU + st issue
        acc[2] += t3 - t2;            // wait::st + fence
        acc[3] += t6 - stamp[5];      // commit issued -> compute thread awake (MMA execution + commit + wake)
        acc[4] += stamp[5] - stamp[4];   // issuer: awake -> MMAs + commit issued
        acc[6] += t3 - t6;            // compute thread's own phase
      }
      if (tid == 0) stamp[3] = t3;
    }
    if (keep == 123.456f) out[100] = (long long)keep;
    if (tid == 0)
      for (int i = 0; i < 8; ++i) out[i] = acc[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == NW) tc::tmem_dealloc(tmem, 64);
}

template <int NW, int WM>
void run() {
  long long* d; cudaMalloc(&d, 8 * 256);
  cudaMemset(d, 0, 8 * 256);
  const int rounds = 400;
  const long long t_dummy = 0; (void)t_dummy;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  hop_kernel<NW, WM><<<1, 32 * NW + 32>>>(rounds, d);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const double n = rounds - 1;
  const double own = h[6] / n, total = ms * 1e-3 * 1.965e9 / rounds;
  printf("warps=%d wait=%s : round %.0f cyc | compute: ld %.0f, ALU+st %.0f, wait::st+fence %.0f (own phase %.0f) | issuer awake->commit issued %.0f | "
         "commit issued->compute awake %.0f | arrive->issuer awake (rest) %.0f  %s\n",
         NW, WM == 0 ? "try_wait+hint" : WM == 1 ? "try_wait" : "test_wait spin", total, h[0] / n, h[1] / n, h[2] / n, own, h[4] / n, h[3] / n,
         total - own - h[4] / n - h[3] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<4, 0>(); run<4, 1>(); run<4, 2>();
  run<8, 0>(); run<8, 1>(); run<8, 2>();
  return 0;
}
