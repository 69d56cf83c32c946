// tcgen05 "TS" probe (not product): A operand in TENSOR MEMORY (written by tcgen05.st), B in shared memory.
//   (1) layout check: D = A . B^T with A[128 x 32] fp16 stored as TMEM lane = row, 32-bit column j = (k = 2j | k = 2j+1 << 16);
//   (2) issue / throughput: cycles per 128 x N x 16 MMA, TS vs SS, dependent accumulate chains vs independent, 1..4 issuers;
//   (3) round-trip latency of one decoder "layer round": tcgen05.ld -> ALU -> tcgen05.st (A hi/lo) -> MMA x6 -> commit -> wait,
//       with 1 / 3 software-interleaved chains per epilogue thread, TS vs SS (STS + fence.proxy.async).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tc_ts_probe tools/tc_ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../giga_b200/csrc/tc.cuh"
using namespace giga;

__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- (1) layout / correctness
// A [128][32] halfs (row-major, global), B [32 n][32 k] halfs (row-major), D [128][32] floats
__global__ void __launch_bounds__(128) ts_layout_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
  __shared__ __align__(128) uint8_t sB[4 * 512];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&slot, 64);
  if (tid == 0) tc::mbar_init(&bar, 1);
  for (int e = tid; e < 32 * 32; e += 128) {
    const int n = e / 32, k = e % 32;
    *reinterpret_cast<__half*>(sB + (k / 8) * 512 + n * 16 + (k % 8) * 2) = B[n * 32 + k];
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t r[16];
  for (int j = 0; j < 16; ++j) {
    const uint32_t lo = __half_as_ushort(A[tid * 32 + 2 * j]), hi = __half_as_ushort(A[tid * 32 + 2 * j + 1]);
    r[j] = lo | (hi << 16);
  }
  tmem_st16(row + 0, r);
  tmem_wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(128, 32);
      for (int ks = 0; ks < 2; ++ks)
        mma_f16_ts(tmem + 32, tmem + ks * 8, tc::make_desc(tc::smem_u32(sB) + ks * 2 * 512, 512, 128), idesc, ks > 0 ? 1u : 0u);
      tc::mma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float v[32];
  tc::tmem_ld32(row + 32, v);
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------- (2) issue / throughput
// MODE 0: SS (A smem), 1: TS (A tmem).  Each issuing warp runs `iters` rounds; a round = for every chain of the warp KSTEPS
// accumulating MMAs into that chain's accumulator (first one of the round overwrites).  DEP = 1: the KSTEPS MMAs of a chain are issued
// back to back (dependent); DEP = 0: k-step-major order (chains interleaved).
template <int N, int MODE, int CHAINS, int KSTEPS, int DEP>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int nwarps, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 64 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (tid == 0) tc::mbar_init(&bar, nwarps);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  {   // defined A contents in TMEM columns [448, 512)
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = 0x3c003c00u;
    const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 448; c < 512; c += 16) tmem_st16(row + c, r);
    tmem_wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 40 * 1024;
  const uint32_t ks_a = 129 * 16, ks_b = N * 16;
  const uint32_t idesc = tc::make_idesc_f16(128, N);
  long long t0 = clock64();
  if (warp < nwarps) {
    if (tc::elect_one()) {
      uint64_t ad[KSTEPS], bd[KSTEPS];
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        ad[k] = tc::make_desc(a0 + (k % 4) * 2 * ks_a, ks_a, 128);
        bd[k] = tc::make_desc(b0 + (k % 2) * 2 * ks_b, ks_b, 128);
      }
      const uint32_t dbase = tmem + ((warp * CHAINS * N) % 448);
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {
        if (DEP) {
#pragma unroll
          for (int c = 0; c < CHAINS; ++c)
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (MODE == 0) tc::mma_f16(dbase + c * N, ad[k], bd[k], idesc, k > 0 ? 1u : 0u);
              else mma_f16_ts(dbase + c * N, tmem + 448 + (k % 8) * 8, bd[k], idesc, k > 0 ? 1u : 0u);
            }
        } else {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
              if (MODE == 0) tc::mma_f16(dbase + c * N, ad[k], bd[k], idesc, k > 0 ? 1u : 0u);
              else mma_f16_ts(dbase + c * N, tmem + 448 + (k % 8) * 8, bd[k], idesc, k > 0 ? 1u : 0u);
            }
        }
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

template <int N, int MODE, int CHAINS, int KSTEPS, int DEP>
void run_rate(int nwarps) {
  long long* d; cudaMalloc(&d, 8 * 1024);
  const int iters = 100;
  cudaFuncSetAttribute(rate_kernel<N, MODE, CHAINS, KSTEPS, DEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rate_kernel<N, MODE, CHAINS, KSTEPS, DEP><<<148, 128, 64 * 1024>>>(iters, nwarps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per = avg / ((double)iters * CHAINS * KSTEPS * nwarps);
  printf("rate %s N=%3d issuers=%d chains/issuer=%d ksteps=%2d %s : %7.1f cycles / MMA (per SM)  %s\n", MODE ? "TS" : "SS", N, nwarps, CHAINS, KSTEPS,
         DEP ? "chain-major(dependent)" : "kstep-major(interleaved)", per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

// ---------------------------------------------------------------- (3) layer round trip
// 128 epilogue threads (warps 0-3) + issuer warp 4.  NCH chains per epilogue thread, software-interleaved.  A round of a chain:
//   epilogue: wait acc_full[c]; tcgen05.ld 32 cols; v = relu(v * s + b); split hi/lo; write A (TS: tcgen05.st 2 x 16 cols; SS: 8 x 16 B STS + proxy fence);
//             arrive a_ready[c] (128)
//   issuer:   wait a_ready[c]; 6 MMAs (3 products x 2 k-steps, N = 32) into acc[c]; commit acc_full[c]
template <int MODE, int NCH, int ALU>
__global__ void __launch_bounds__(160) round_kernel(int rounds, long long* out, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t acc_full[4], a_ready[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 64 * 1024 / 4; e += 160) reinterpret_cast<uint32_t*>(smem)[e] = 0x2c002c00u;   // 2^-4
  if (warp == 4) tc::tmem_alloc(&slot, 512);
  if (tid == 0)
    for (int c = 0; c < 4; ++c) { tc::mbar_init(&acc_full[c], 1); tc::mbar_init(&a_ready[c], 128); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  constexpr int KS_A = 128 * 16 + 16;
  const uint32_t b_hi = tc::smem_u32(smem), b_lo = b_hi + 2048;                    // [k/8][n 32][8] halfs: 4 x 512 B each
  const uint32_t idesc = tc::make_idesc_f16(128, 32);
  // TMEM: chain c: acc cols [c*64, +32), A hi cols [c*64+32, +16), A lo [c*64+48, +16)
  long long t0 = clock64();
  if (warp == 4) {
    if (tc::elect_one()) {
      for (int r = 0; r < rounds; ++r)
        for (int c = 0; c < NCH; ++c) {
          tc::mbar_wait(&a_ready[c], (uint32_t)(r & 1));
          tc::fence_after_sync();
          const uint32_t d = tmem + c * 64;
          if (MODE == 1) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t bh = tc::make_desc(b_hi + ks * 1024, 512, 128), bl = tc::make_desc(b_lo + ks * 1024, 512, 128);
              mma_f16_ts(d, d + 32 + ks * 8, bh, idesc, ks > 0 ? 1u : 0u);
              mma_f16_ts(d, d + 48 + ks * 8, bh, idesc, 1u);
              mma_f16_ts(d, d + 32 + ks * 8, bl, idesc, 1u);
            }
          } else {
            const uint32_t a_hi = tc::smem_u32(smem) + 8192 + c * 2 * 4 * KS_A, a_lo = a_hi + 4 * KS_A;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t bh = tc::make_desc(b_hi + ks * 1024, 512, 128), bl = tc::make_desc(b_lo + ks * 1024, 512, 128);
              const uint64_t ah = tc::make_desc(a_hi + ks * 2 * KS_A, KS_A, 128), al = tc::make_desc(a_lo + ks * 2 * KS_A, KS_A, 128);
              tc::mma_f16(d, ah, bh, idesc, ks > 0 ? 1u : 0u);
              tc::mma_f16(d, al, bh, idesc, 1u);
              tc::mma_f16(d, ah, bl, idesc, 1u);
            }
          }
          tc::mma_commit(&acc_full[c]);
        }
    }
    __syncwarp();
  } else {
    const uint32_t row = tmem + ((uint32_t)(warp * 32) << 16);
    float keep = 0.f;
    {   // round 0: initial A operands
      uint32_t r16[16];
      for (int j = 0; j < 16; ++j) r16[j] = 0x2c002c00u;
      for (int c = 0; c < NCH; ++c) {
        if (MODE == 1) { tmem_st16(row + c * 64 + 32, r16); tmem_st16(row + c * 64 + 48, r16); tmem_wait_st(); }
        else tc::fence_smem_to_async();
        tc::fence_before_sync();
        mbar_arrive_(&a_ready[c]);
      }
    }
    for (int r = 0; r < rounds; ++r)
      for (int c = 0; c < NCH; ++c) {
        tc::mbar_wait(&acc_full[c], (uint32_t)(r & 1));
        tc::fence_after_sync();
        float v[32];
        tc::tmem_ld32(row + c * 64, v);
        uint32_t h[16], l[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x0 = v[2 * j], x1 = v[2 * j + 1];
          if (ALU) { x0 = fmaxf(fmaf(x0, 0.03125f, 0.01f), 0.f); x1 = fmaxf(fmaf(x1, 0.03125f, 0.01f), 0.f); }
          const __half2 hh = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          h[j] = *reinterpret_cast<const uint32_t*>(&hh);
          l[j] = *reinterpret_cast<const uint32_t*>(&ll);
          keep += x0;
        }
        if (r + 1 < rounds) {
          if (MODE == 1) {
            tmem_st16(row + c * 64 + 32, h);
            tmem_st16(row + c * 64 + 48, l);
            tmem_wait_st();
          } else {
            uint8_t* a_hi = smem + 8192 + c * 2 * 4 * KS_A;
            uint8_t* a_lo = a_hi + 4 * KS_A;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              *reinterpret_cast<uint4*>(a_hi + kc * KS_A + tid * 16) = make_uint4(h[4 * kc], h[4 * kc + 1], h[4 * kc + 2], h[4 * kc + 3]);
              *reinterpret_cast<uint4*>(a_lo + kc * KS_A + tid * 16) = make_uint4(l[4 * kc], l[4 * kc + 1], l[4 * kc + 2], l[4 * kc + 3]);
            }
            tc::fence_smem_to_async();
          }
          tc::fence_before_sync();
          mbar_arrive_(&a_ready[c]);
        }
      }
    if (keep == 123.456f) sink[tid] = keep;
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

template <int MODE, int NCH, int ALU>
void run_round() {
  long long* d; cudaMalloc(&d, 8 * 1024);
  float* sink; cudaMalloc(&sink, 1024);
  const int rounds = 200;
  cudaFuncSetAttribute(round_kernel<MODE, NCH, ALU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  round_kernel<MODE, NCH, ALU><<<148, 160, 64 * 1024>>>(rounds, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("round %s chains/thread=%d alu=%d : %7.1f cycles per round of all chains, %7.1f per chain-layer  %s\n", MODE ? "TS" : "SS", NCH, ALU, avg / rounds,
         avg / rounds / NCH, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  // ---- (1)
  {
    __half hA[128 * 32], hB[32 * 32];
    srand(1);
    for (auto& x : hA) x = __float2half((rand() % 2001 - 1000) / 500.f);
    for (auto& x : hB) x = __float2half((rand() % 2001 - 1000) / 500.f);
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    ts_layout_kernel<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    float hD[128 * 32];
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double s = 0;
        for (int k = 0; k < 32; ++k) s += (double)__half2float(hA[m * 32 + k]) * __half2float(hB[n * 32 + k]);
        maxerr = fmax(maxerr, fabs(s - hD[m * 32 + n]));
      }
    printf("TS layout (lane=row, col j = k 2j | 2j+1, k-step advance 8 cols): max |err| = %.3g  %s  %s\n", maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH",
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (!(maxerr < 1e-3)) {   // diagnostic: B = identity -> D[m][n] should equal A[m][n]
      for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) hB[n * 32 + k] = __float2half(n == k ? 1.f : 0.f);
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) hA[m * 32 + k] = __float2half((float)(k + 1 + 64 * (m & 1)));
      cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
      ts_layout_kernel<<<1, 128>>>(dA, dB, dD);
      cudaDeviceSynchronize();
      cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
      for (int m : {0, 1, 33}) { printf("  D[%d][:] =", m); for (int n = 0; n < 32; ++n) printf(" %g", hD[m * 32 + n]); printf("\n"); }
    }
  }
  // ---- (2)
  run_rate<32, 0, 1, 6, 1>(1); run_rate<32, 1, 1, 6, 1>(1);
  run_rate<32, 0, 3, 6, 1>(1); run_rate<32, 1, 3, 6, 1>(1);
  run_rate<32, 0, 3, 6, 0>(1); run_rate<32, 1, 3, 6, 0>(1);
  run_rate<32, 1, 6, 6, 1>(1); run_rate<32, 1, 6, 6, 0>(1);
  run_rate<32, 1, 3, 6, 1>(2); run_rate<32, 1, 3, 6, 0>(2);
  run_rate<32, 1, 3, 6, 1>(4); run_rate<32, 1, 3, 6, 0>(4);
  run_rate<32, 0, 3, 6, 1>(2); run_rate<32, 0, 3, 6, 1>(4);
  run_rate<96, 0, 1, 18, 1>(1); run_rate<96, 0, 1, 18, 1>(2); run_rate<96, 1, 1, 18, 1>(1); run_rate<96, 1, 1, 18, 1>(2);
  run_rate<96, 0, 2, 18, 0>(1); run_rate<96, 0, 2, 18, 0>(2);
  run_rate<160, 0, 1, 18, 1>(1); run_rate<160, 1, 1, 18, 1>(1);
  run_rate<64, 1, 3, 6, 1>(2); run_rate<64, 0, 3, 6, 1>(2);
  // ---- (3)
  run_round<0, 1, 1>(); run_round<1, 1, 1>();
  run_round<0, 3, 1>(); run_round<1, 3, 1>();
  run_round<1, 1, 0>(); run_round<1, 3, 0>();
  run_round<1, 2, 1>(); run_round<1, 4, 1>();
  return 0;
}
