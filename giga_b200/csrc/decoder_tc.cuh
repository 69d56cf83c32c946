// Tensor-core implicit decoder (tcgen05 / TMEM, 3xFP16 operand splitting -> fp32-level accuracy).
//
// Same function as decode_points_kernel (decoder.cuh) -- tri-plane gather + LocalDecoder heads
// (conv_onet/models/decoder.py:117-176, layers.py:39-47, models/__init__.py:119-123) -- but every
// dense contraction runs on the 5th-gen tensor cores:
//   fc_c  : [128 pts x 96 feat] . [96 x 160]   (all five blocks' fc_c at once, one plane (K=32) at a time)
//   fc_0/1: [128 pts x 32]      . [32 x 32]    x 10 per head (the ResnetBlockFC chain)
// A CTA owns 128 query points of one scene; thread t <-> point t <-> TMEM lane t.
//   * weights: staged by cp.async.bulk (UBLKCP) with mbarrier completion, prefetched one stage ahead;
//   * gather: warp-cooperative (lane = channel, 12 coalesced 128 B texel reads per point); the
//     96 features of the warp's 32 points stay in REGISTERS (F[3][32] per lane) and are re-split
//     into the shared-memory A operand (hi/lo fp16 pair) for every head -- gathered once per tile;
//   * MMAs are issued by one elected lane of warp 0; accumulators live in TMEM: columns [0,160) = fc_c outputs of the
//     five blocks, [160,256) = three partial accumulators (hi*hi, lo*hi, hi*lo) of the current 32x32 layer;
//   * epilogues (bias, residual, ReLU, hi/lo split, write next A operand) run thread-per-point
//     straight out of TMEM (tcgen05.ld 32x32b.x32), fc_p (K=3) and fc_out (N<=4) stay on CUDA cores;
//   * 3xFP16: x*y ~= xh*yh + xl*yh + xh*yl with xh = fp16_rn(x), xl = fp16_rn(x - xh) (x - xh is exact in fp32); fp16 has
//     tf32's 11-bit significand, so this is the 3xTF32 scheme at twice the K per MMA and half the operand bytes
//     (tools/tc_f16_probe.cu: 2e-7 relative vs fp64).  fp16's narrow exponent is handled by a per-head power-of-two
//     pre-scale 2^s of the weights chosen on the host (max |w| 2^s in [512, 1024)): the lo halves of all but
//     negligible weights stay normal fp16 numbers; accumulators come out scaled by 2^s and the epilogues multiply by
//     2^-s (exact).  Activations are O(1..50): their lo halves underflow gracefully below 2^-3 (absolute error
//     <= 2^-25, fp32-class relative to the activation scale).  Activations are clamped to fp16's +-65504.
#pragma once
#include "common.cuh"
#include "decoder.cuh"
#include "tc.cuh"
#include "unet_tall.cuh"   // tc::bulk_g2s / mbar_arrive_expect_tx

namespace giga {

// ---- packed per-head parameter blob for the tensor-core path (floats) ----
constexpr int TW_FCP = 0;                         // Wt[3][32] + b[32]                      (128)
constexpr int TW_FCC = 128;                       // [plane 3][hi,lo] fp16 operands, halfs at (k/8)*1280 + n*8 + k%8, n = blk*32+j
constexpr int TW_FCC_SLICE = 2560;                //   one (plane, hi|lo) operand: N=160 x K=32 halfs = 10240 B (in floats)
constexpr int TW_BC = TW_FCC + 6 * TW_FCC_SLICE;  // 15488: fc_c biases [5][32]
constexpr int TW_BLK = TW_BC + 160;               // 15648: per block W0hi W0lo W1hi W1lo (2048 B each, halfs at (k/8)*256 + n*8 + k%8), b0[32], b1[32]
constexpr int TW_BLK_SIZE = 4 * 512 + 64;         // 2112
constexpr int TW_OUT = TW_BLK + 5 * TW_BLK_SIZE;  // 26208: Wt[32][4] + b[4]
constexpr int TW_INV = TW_OUT + 132;              // 2^-s: undoes the head's weight pre-scale
constexpr int TW_HEAD = TW_INV + 4;               // 26344 floats per head

constexpr int TD_PTS = 128;
constexpr int TD_KS_A = TD_PTS * 16 + 16;         // A k-chunk stride (bytes), +16 keeps the gather stores conflict free
constexpr int TD_A_BYTES = 4 * TD_KS_A;           // one A operand (K=32 = 4 k-chunks of 8 halfs): 8256 B
constexpr int TD_W_BYTES = 2 * TW_FCC_SLICE * 4;  // staged weights: fc_c plane slice hi+lo = 20480 B (chain stage uses 8.4 KB of it)
constexpr int TD_OFF_ALO = TD_A_BYTES;
constexpr int TD_OFF_W = 2 * TD_A_BYTES;          // 16512
constexpr int TD_OFF_TINFO = TD_OFF_W + TD_W_BYTES;          // 36992
constexpr int TD_OFF_WB0 = TD_OFF_TINFO + TD_PTS * 24 * 4;   // 49280: dedicated chain-weight buffer 0 (buffer 1 aliases sW)
constexpr int TD_WB_BYTES = TW_BLK_SIZE * 4;                 // 8448
constexpr int TD_OFF_BAR = TD_OFF_WB0 + TD_WB_BYTES;         // 57728
constexpr int TD_SMEM_BYTES = TD_OFF_BAR + 48;               // 57776  (TMEM keeps it at 2 CTAs / SM)
constexpr int TD_TMEM_COLS = 256;
constexpr uint32_t TD_KS_WC = 160 * 16;           // fc_c B operand k-chunk stride (N=160)
constexpr uint32_t TD_KS_W = 32 * 16;             // 32x32 B operand k-chunk stride

// one of the three 3xFP16 products of a 32x32 layer (kind 0: hi*hi, 1: lo*hi, 2: hi*lo) -> accumulator d_tmem + 32*kind
// (independent accumulators: dependent MMAs into one TMEM tile cost ~100 cycles each).  Issued by three different warps
// concurrently: a single thread only issues one small MMA every ~85 cycles.  K = 32 = 2 MMAs of K = 16.
__device__ __forceinline__ void issue_k32_one(int kind, uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                              uint32_t b_kstride, uint32_t idesc) {
  const uint32_t a = kind == 1 ? a_lo : a_hi, b = kind == 2 ? b_lo : b_hi;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    tc::mma_f16(d_tmem + 32 * kind, tc::make_desc(a + ks * 2 * TD_KS_A, TD_KS_A, 128), tc::make_desc(b + ks * 2 * b_kstride, b_kstride, 128),
                idesc, ks > 0 ? 1u : 0u);
}

// sum of the three partial accumulators of a 32x32 layer for this thread's row
__device__ __forceinline__ void tmem_ld32_sum3(uint32_t taddr, float* v) {
  float a[32];
  tc::tmem_ld32(taddr, v);
  tc::tmem_ld32(taddr + 32, a);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] += a[j];
  tc::tmem_ld32(taddr + 64, a);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] += a[j];
}

// issue the K=32 contraction  D[128 x N] (+)= (Ahi+Alo)[128x32] . (Bhi+Blo)[N x 32]^T  as 2 k-steps x 3 MMAs
__device__ __forceinline__ void issue_k32_x3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             uint32_t b_kstride, uint32_t idesc, bool accumulate_first) {
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint64_t ah = tc::make_desc(a_hi + ks * 2 * TD_KS_A, TD_KS_A, 128);
    const uint64_t al = tc::make_desc(a_lo + ks * 2 * TD_KS_A, TD_KS_A, 128);
    const uint64_t bh = tc::make_desc(b_hi + ks * 2 * b_kstride, b_kstride, 128);
    const uint64_t bl = tc::make_desc(b_lo + ks * 2 * b_kstride, b_kstride, 128);
    tc::mma_f16(d_tmem, ah, bh, idesc, acc);
    tc::mma_f16(d_tmem, al, bh, idesc, 1u);
    tc::mma_f16(d_tmem, ah, bl, idesc, 1u);
    acc = 1u;
  }
}

// unscaled fp16 split of an activation (clamped to fp16's range): hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split_act2(float v0, float v1, uint32_t& h, uint32_t& l) {
  v0 = fminf(fmaxf(v0, -H_MAX), H_MAX);
  v1 = fminf(fmaxf(v1, -H_MAX), H_MAX);
  const __half2 hh = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  h = *reinterpret_cast<const uint32_t*>(&hh);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}

// thread t writes row t of the A operand (hi and lo) from 32 fp32 values
__device__ __forceinline__ void store_a_row(uint8_t* a_hi, uint8_t* a_lo, int row, const float* v) {
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_act2(v[8 * kc + 2 * i], v[8 * kc + 2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(a_hi + kc * TD_KS_A + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_lo + kc * TD_KS_A + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// grid ((ceil(N/128) + ceil(N2/128)) * B), block 128, dynamic smem TD_SMEM_BYTES, 2 CTAs/SM (2 x 256 TMEM columns).
// Two jobs can share one launch (forward(): the grasp heads at p and the TSDF head at p_tsdf).  The grid is ONE-dimensional
// and ordered by cost: CTAs [0, tiles1*B) evaluate `heads` at pts for all scenes, the rest `heads2` at pts2 -- CTAs are
// dispatched in blockIdx order, so every long (3-head) tile starts before any short one and the kernel's tail consists of
// short tiles (with a (tile, scene) grid the last scenes' long tiles ran alone at the end: 158 -> 12x us).
__global__ void __launch_bounds__(TD_PTS, 2)
decode_points_tc_kernel(const float* __restrict__ planes,  // [3][B][40][40][32]
                        const float* __restrict__ pts,     // [B][N][3]
                        const float* __restrict__ tw,      // [4][TW_HEAD]
                        int B, int N, unsigned heads,
                        float* __restrict__ qual, float* __restrict__ rot, float* __restrict__ width,
                        float* __restrict__ occ,
                        const float* __restrict__ pts2, int N2, unsigned heads2, int tiles1,   // second job (N2 = 0: none)
                        unsigned long long* __restrict__ tl) {   // tl: optional debug timeline
  extern __shared__ __align__(128) uint8_t smem_tc[];
  uint8_t* smem = smem_tc;
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + TD_OFF_ALO;
  uint8_t* sW = smem + TD_OFF_W;
  float* tinfo = reinterpret_cast<float*>(smem + TD_OFF_TINFO);
  uint8_t* sWB0 = smem + TD_OFF_WB0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TD_OFF_BAR);       // MMA completion
  uint64_t* wbar = bar + 1;                                             // [2] chain-weight buffers landed
  uint64_t* pbar = bar + 3;                                             // fc_c plane slice landed
  uint64_t* bar3 = bar + 4;                                             // 32x32 layer: three issuing warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TD_OFF_BAR + 40);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tile = blockIdx.x, b;
  if (tile < tiles1 * B) { b = tile / tiles1; tile -= b * tiles1; }
  else {
    const int tiles2 = (N2 + TD_PTS - 1) / TD_PTS;
    tile -= tiles1 * B;
    b = tile / tiles2; tile -= b * tiles2;
    pts = pts2; N = N2; heads = heads2;
  }
  const int n0 = tile * TD_PTS;
  const int n = n0 + tid;
  const bool valid = n < N;
  const int nc = valid ? n : N - 1;

  unsigned long long* tlc = tl ? tl + (size_t)blockIdx.x * 32 : nullptr;
  int tslot = 0;
  auto stamp = [&]() {
    if (tlc && tid == 0 && tslot < 31) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      tlc[tslot] = t;
    }
    ++tslot;
  };
  stamp();   // 0: start
  if (warp == 0) tc::tmem_alloc(tmem_slot, TD_TMEM_COLS);
  if (tid == 0) {
    tc::mbar_init(bar, 1); tc::mbar_init(&wbar[0], 1); tc::mbar_init(&wbar[1], 1); tc::mbar_init(pbar, 1); tc::mbar_init(bar3, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  pdl_wait();   // plane features (and points) of the preceding kernels are complete; outputs are only written below
  // ---- gather: the warp's 32 points, lane = channel; features stay in registers ----
  float F[3][32];
  {
    float* tw_ = tinfo + warp * 32 * 24;
    {
      const int nq = min(n0 + warp * 32 + lane, N - 1);
      TexInfo t;
      point_taps(pts + ((size_t)b * N + nq) * 3, t);
      int* ti = reinterpret_cast<int*>(tw_) + lane * 24;
      float* tf = tw_ + lane * 24 + 12;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ti[pl * 4 + q] = t.off[pl][q];
          tf[pl * 4 + q] = t.w[pl][q];
        }
    }
    __syncwarp();
    const float* pb[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) pb[pl] = planes + ((size_t)pl * B + b) * (G2 * C) + lane;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int4* oi = reinterpret_cast<const int4*>(tw_ + q * 24);
      const float4* wf = reinterpret_cast<const float4*>(tw_ + q * 24 + 12);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const int4 o = oi[pl];
        const float4 w = wf[pl];
        const float v0 = __ldg(pb[pl] + o.x), v1 = __ldg(pb[pl] + o.y), v2 = __ldg(pb[pl] + o.z), v3 = __ldg(pb[pl] + o.w);
        F[pl][q] = v0 * w.x + v1 * w.y + v2 * w.z + v3 * w.w;
      }
    }
  }

  const float* pp = pts + ((size_t)b * N + nc) * 3;
  const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  stamp();   // 1: gather done
  pdl_launch();   // after the gather: the successor's CTAs take shared memory out of the L1 carve-out the gather reads through

  tc::fence_before_sync();
  __syncthreads();   // TMEM address + mbarrier init visible
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_hi = tc::smem_u32(sAhi), a_lo = tc::smem_u32(sAlo), w_s = tc::smem_u32(sW);
  constexpr uint32_t IDESC_160 = tc::make_idesc_f16(128, 160);
  constexpr uint32_t IDESC_32 = tc::make_idesc_f16(128, 32);
  uint32_t phase = 0, phase3 = 0;

  // Weight staging is asynchronous (cp.async.bulk + mbarrier, issued by one elected lane of warp 0):
  //   chain block i+1 is prefetched while block i computes (buffers: dedicated WB0 / the start of sW),
  //   block 0 of a head is prefetched during its fc_c plane rounds, and the next head's first plane slice
  //   during block 4.  Only plane slices 1 and 2 of each head are fetched on demand.
  uint32_t wuse[2] = {0u, 0u}, puse = 0u;          // completed-use counters -> mbarrier parities
  bool plane0_prefetched = false;
  auto issue_bulk = [&](void* dst, const float* src, uint32_t bytes, uint64_t* mb) {
    if (warp == 0) {
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(mb, bytes);
        tc::bulk_g2s(dst, src, bytes, mb);
      }
      __syncwarp();
    }
  };

#pragma unroll 1
  for (int head = 0; head < 4; ++head) {
    if (!(heads & (1u << head))) continue;
    const float* W = tw + (size_t)head * TW_HEAD;
    const float winv = __ldg(W + TW_INV);   // 2^-s of this head's weight pre-scale
    int next_head = -1;
    for (int hh = head + 1; hh < 4; ++hh)
      if (heads & (1u << hh)) { next_head = hh; break; }

    // ---- fc_c for all five blocks: C[128 x 160] = F[128 x 96] . Wc^T, one plane (K = 32) per round ----
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      // (head, plane) weight slice hi|lo, already in operand layout; sW is free: its last readers' MMAs completed
      if (!(pl == 0 && plane0_prefetched)) issue_bulk(sW, W + TW_FCC + pl * 2 * TW_FCC_SLICE, TD_W_BYTES, pbar);
      if (pl == 0) issue_bulk(sWB0, W + TW_BLK, TD_WB_BYTES, &wbar[0]);   // chain block 0 -> dedicated buffer
      // A operand from the register-resident features: element (row = warp*32+q, k = lane)
      uint8_t* ah = sAhi + (lane >> 3) * TD_KS_A + (lane & 7) * 2 + (warp * 32) * 16;
      uint8_t* al = sAlo + (lane >> 3) * TD_KS_A + (lane & 7) * 2 + (warp * 32) * 16;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float v = fminf(fmaxf(F[pl][q], -H_MAX), H_MAX);
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(ah + q * 16) = h;
        *reinterpret_cast<__half*>(al + q * 16) = __float2half_rn(v - __half2float(h));
      }
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      if (warp == 0) {
        tc::mbar_wait(pbar, puse & 1u);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          issue_k32_x3(tmem, a_hi, a_lo, w_s, w_s + TW_FCC_SLICE * 4, TD_KS_WC, IDESC_160, pl > 0);
          tc::mma_commit(bar);
        }
        __syncwarp();
      }
      ++puse;
      tc::mbar_wait(bar, phase);   // MMAs done: A / W buffers reusable, C columns readable
      phase ^= 1u;
      tc::fence_after_sync();
      stamp();   // per head: 3 plane rounds
    }
    plane0_prefetched = false;

    // ---- fc_p on CUDA cores ----
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      h[j] = __ldg(W + TW_FCP + 96 + j) + __ldg(W + TW_FCP + j) * px + __ldg(W + TW_FCP + 32 + j) * py +
             __ldg(W + TW_FCP + 64 + j) * pz;

#pragma unroll 1
    for (int blk = 0; blk < 5; ++blk) {
      const int wb = blk & 1;                                   // 0: WB0, 1: start of sW
      const uint8_t* wbuf = wb ? sW : sWB0;
      const uint32_t w_b = tc::smem_u32(wbuf);
      // prefetch: next block's weights into the other buffer; during the last block the next head's first plane slice
      if (blk < 4) issue_bulk(wb ? (void*)sWB0 : (void*)sW, W + TW_BLK + (blk + 1) * TW_BLK_SIZE, TD_WB_BYTES, &wbar[wb ^ 1]);
      else if (next_head >= 0) {
        issue_bulk(sW, tw + (size_t)next_head * TW_HEAD + TW_FCC, TD_W_BYTES, pbar);
        plane0_prefetched = true;
      }
      float v[32];
      tc::tmem_ld32(tmem_row + blk * 32, v);           // fc_c[blk] output for this point
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        h[j] += fmaf(v[j], winv, __ldg(W + TW_BC + blk * 32 + j));  // net = net + fc_c[blk](c)
        v[j] = fmaxf(h[j], 0.f);
      }
      store_a_row(sAhi, sAlo, tid, v);
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::mbar_wait(&wbar[wb], wuse[wb] & 1u);         // this block's weights have landed (all threads: they read b0/b1)
      ++wuse[wb];
      if (warp < 3) {
        tc::fence_after_sync();
        if (tc::elect_one()) {
          issue_k32_one(warp, tmem + 160, a_hi, a_lo, w_b, w_b + 2048, TD_KS_W, IDESC_32);   // fc_0
          tc::mma_commit(bar3);
        }
        __syncwarp();
      }
      tc::mbar_wait(bar3, phase3);
      phase3 ^= 1u;
      tc::fence_after_sync();
      const float* bs = reinterpret_cast<const float*>(wbuf) + 2048;   // b0[32], b1[32]
      tmem_ld32_sum3(tmem_row + 160, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(fmaf(v[j], winv, bs[j]), 0.f);
      store_a_row(sAhi, sAlo, tid, v);
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      if (warp < 3) {
        tc::fence_after_sync();
        if (tc::elect_one()) {
          issue_k32_one(warp, tmem + 160, a_hi, a_lo, w_b + 4096, w_b + 6144, TD_KS_W, IDESC_32);   // fc_1
          tc::mma_commit(bar3);
        }
        __syncwarp();
      }
      tc::mbar_wait(bar3, phase3);
      phase3 ^= 1u;
      tc::fence_after_sync();
      tmem_ld32_sum3(tmem_row + 160, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] += fmaf(v[j], winv, bs[32 + j]);     // x + fc_1(relu(fc_0(relu(x))))
      tc::fence_before_sync();
      __syncthreads();   // everyone has read b0/b1 and T before this weight buffer / T are overwritten
      stamp();   // per head: 5 block rounds
    }

    // ---- fc_out(relu(net)) + head epilogue on CUDA cores ----
    float o[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) o[m] = __ldg(W + TW_OUT + 128 + m);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float r = fmaxf(h[k], 0.f);
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + TW_OUT + k * 4));
      o[0] = fmaf(w.x, r, o[0]); o[1] = fmaf(w.y, r, o[1]);
      o[2] = fmaf(w.z, r, o[2]); o[3] = fmaf(w.w, r, o[3]);
    }
    if (valid) {
      const size_t idx = (size_t)b * N + n;
      const bool raw = (heads & 16u) != 0;
      if (head == 0) {
        qual[idx] = raw ? o[0] : 1.f / (1.f + expf(-o[0]));
      } else if (head == 1) {
        const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
        const float d = raw ? 1.f : fmaxf(nrm, 1e-12f);
        st4(rot + idx * 4, make_float4(o[0] / d, o[1] / d, o[2] / d, o[3] / d));
      } else if (head == 2) {
        width[idx] = o[0];
      } else {
        occ[idx] = o[0];
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TD_TMEM_COLS);
}

}  // namespace giga
