// Thin inline-PTX layer over the sm_100a tensor-core path: tcgen05.mma (kind::tf32) with TMEM
// accumulators, shared-memory matrix descriptors, mbarrier completion.  Layout/descriptor facts
// below were validated on a B200 with tools/tc_probe.cu (profiles/r01_tc_probe.txt).
//
// Operand layout used everywhere (canonical K-major, SWIZZLE_NONE): an operand with R rows (M or N)
// and K columns is stored as 16-byte "k-chunks" of E consecutive k for one row (E = 4 for tf32/fp32 words, 8 for fp16):
//     byte address(r, k) = (k/E) * KSTRIDE + r*16 + (k%E)*(16/E)
// so an 8-row x 16-byte core matrix is 128 contiguous bytes, SBO (next 8 rows) = 128 B and
// LBO (next k-chunk) = KSTRIDE (any multiple of 16 B; R*16 + 16 keeps thread-issued stores
// bank-conflict free).  One MMA consumes two k-chunks (K = 8 tf32 or K = 16 fp16); advancing by one MMA's K
// advances the start address by 2*KSTRIDE.  Advancing the start by 16*s bytes shifts the operand by s rows --
// the "flattened shift" the implicit-GEMM convolution uses for its 9 taps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace giga {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (what tcgen05.mma reads through)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// kind::f16: fp16 A/B (format code 0), fp32 accumulate, both operands K-major; one MMA consumes K = 16 = two 16-byte
// k-chunks of 8 halfs per row -- the SAME shared-memory bytes per MMA as kind::tf32 (K = 8), i.e. twice the K per
// operand fetch (tools/tc_f16_probe.cu: identical cycles per MMA, half the cycles per unit of K)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Bounded wait: a wrong descriptor must trap, never hang the GPU.  Plain try_wait (hardware-default suspend window): an explicit
// suspend-time hint was measured 2x slower per hand-off (tools/tc_hop_probe.cu: 1679 vs 933 cycles per layer round trip).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 24) && !ok; ++i) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
  if (!ok) __trap();
}
// Same, for roles that wait LONG and off the critical path (producers waiting for a free buffer): back off between polls so the
// spinning warp does not take issue slots / mbarrier-unit bandwidth from the warps doing the work.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned sleep_ns = 256) {
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 22) && !ok; ++i) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!ok) __nanosleep(sleep_ns);
  }
  if (!ok) __trap();
}
// this thread's TMEM lane (its warp's quadrant), 32 consecutive fp32 columns starting at taddr
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-column variant (register-lean epilogues)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// One lane of a fully converged warp.  Issuing tcgen05.mma / bulk copies under this predicate (instead of
// under a thread-index test) lets the compiler keep descriptors in uniform registers: back-to-back
// UTCHMMA without the per-instruction ELECT/BRA.U.ANY lane loop (profiles/r01d: 3-4x faster issue).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// 3xTF32 operand split: hi keeps the top 19 bits (exactly representable in tf32), lo = v - hi is exact in fp32
__host__ __device__ __forceinline__ float tf32_hi(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(v) & 0xffffe000u);
#else
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xffffe000u;
  memcpy(&v, &u, 4);
  return v;
#endif
}

}  // namespace tc
}  // namespace giga
