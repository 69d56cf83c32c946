// Tensor-core U-Net on "tall" pre-split activations: bulk-async (TMA-class) staging + tcgen05 + TMEM.
//
// TALL layout of an activation with C channels at resolution HW (all 3B plane images stacked):
//   rows: 0 = zero pad; image i at rows 1+i*(HW+1) .. i*(HW+1)+HW; one zero row after every image;
//   cols: 0 and HW+1 are zero pad (WP = HW+2);  flattened position f = row*WP + col;
//   memory: [hi|lo][C/4 k-chunks][PS positions][4 channels] fp32, position f stored at MARGIN+f;
//   hi = rn_tf32(v), lo = v - hi (exact), so hi+lo == v bit-exactly and each 16-byte element is one
//   row of one k-chunk of a K-major SWIZZLE_NONE UMMA operand (tc.cuh).
// Consequences:
//   * the operand window a CTA needs (its 128*NT output positions plus a one-row/one-column halo, for
//     one k-chunk) is ONE contiguous byte range in global memory -> a single cp.async.bulk
//     (UBLKCP, completion on an mbarrier) per (hi|lo, k-chunk); zero padding is stored, not computed;
//   * the 9 taps of a 3x3 conv are the same staged window advanced by (dy*WP+dx) rows = a start-address
//     offset in the matrix descriptor (hardware-validated: tools/tc_probe.cu T5/T6);
//   * producers write the split in their epilogue (the values are in registers anyway).
// Accuracy: 3xTF32 (hi*hi + lo*hi + hi*lo).  Tensor-core accumulation truncates, so error grows with
// the accumulate chain; the hi*hi products therefore go to one of two alternating TMEM accumulator
// sets that are drained into fp32 REGISTERS every G chunks (<= 18 MMAs per chain, round-to-nearest
// adds on the CUDA cores), while the ~2^-12-scaled correction products accumulate in two further sets.
//
// CTA = 160 threads: warps 0-3 own TMEM lanes 0..127 (drain + epilogue, thread t <-> position f0+t),
// warp 4 lane 0 is the control thread (bulk copies + MMA issue).  No __syncthreads in the main loop;
// everything is mbarrier-driven: full[s] (bytes landed) / empty[s] (tcgen05.commit) per stage,
// acc_full[set] (commit) / acc_empty[set] (128 drainer arrivals) per accumulator set.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace giga {

constexpr int TALL_MARGIN = 640;   // zero positions before/after the image stack (>= WP+1 + 2*128 + slack)

__host__ __device__ constexpr int tall_wp(int hw) { return hw + 2; }
__host__ __device__ constexpr int tall_rows(int hw, int n_img) { return 1 + n_img * (hw + 1); }
__host__ __device__ constexpr long tall_positions(int hw, int n_img) { return (long)tall_rows(hw, n_img) * tall_wp(hw); }
// plane stride (positions) of a TALL tensor sized for n_img images
__host__ __device__ constexpr long tall_ps(int hw, int n_img) { return 2 * TALL_MARGIN + ((tall_positions(hw, n_img) + 255) / 256) * 256; }
__host__ __device__ constexpr long tall_floats(int hw, int n_img, int ch) { return 2L * (ch / 4) * tall_ps(hw, n_img) * 4; }

__device__ __forceinline__ float rn_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l) {
  h.x = rn_tf32(v.x); l.x = v.x - h.x;
  h.y = rn_tf32(v.y); l.y = v.y - h.y;
  h.z = rn_tf32(v.z); l.z = v.z - h.z;
  h.w = rn_tf32(v.w); l.w = v.w - h.w;
}
__device__ __forceinline__ long tall_pos(int hw, int img, int y, int x) { return (long)(1 + img * (hw + 1) + y) * (hw + 2) + x + 1; }

// ---- layout conversion / pooling (thread = one valid pixel x one k-chunk) --------------------------
// grid: ceil(n_img*HW*HW*(C/4) / 256), block 256
template <int HW, int C4>   // C4 = channels / 4
__global__ void __launch_bounds__(256) nchw_to_tall_kernel(const float* __restrict__ src, float* __restrict__ dst, long ps, int n_img) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long total = (long)n_img * C4 * HW * HW;
  if (t >= total) return;
  const int pix = (int)(t % (HW * HW));
  const int kc = (int)((t / (HW * HW)) % C4);
  const int img = (int)(t / ((long)HW * HW * C4));
  const float* p = src + ((size_t)img * (4 * C4) + 4 * kc) * (HW * HW) + pix;
  float4 v = make_float4(__ldg(p), __ldg(p + HW * HW), __ldg(p + 2 * HW * HW), __ldg(p + 3 * HW * HW)), h, l;
  split4(v, h, l);
  const long pos = TALL_MARGIN + tall_pos(HW, img, pix / HW, pix % HW);
  st4(dst + ((size_t)kc * ps + pos) * 4, h);
  st4(dst + ((size_t)(C4 + kc) * ps + pos) * 4, l);
}

template <int HW, int C4>
__global__ void __launch_bounds__(256) tall_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long ps, int n_img) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long total = (long)n_img * C4 * HW * HW;
  if (t >= total) return;
  const int pix = (int)(t % (HW * HW));
  const int kc = (int)((t / (HW * HW)) % C4);
  const int img = (int)(t / ((long)HW * HW * C4));
  const long pos = TALL_MARGIN + tall_pos(HW, img, pix / HW, pix % HW);
  const float4 h = ld4(src + ((size_t)kc * ps + pos) * 4), l = ld4(src + ((size_t)(C4 + kc) * ps + pos) * 4);
  float* p = dst + ((size_t)img * (4 * C4) + 4 * kc) * (HW * HW) + pix;
  p[0] = h.x + l.x; p[HW * HW] = h.y + l.y; p[2 * HW * HW] = h.z + l.z; p[3 * HW * HW] = h.w + l.w;
}

// MaxPool2d(2,2) (unet.py:64): TALL(2*HW) -> TALL(HW); hi+lo reconstructs the fp32 value exactly
template <int HW, int C4>   // HW = OUTPUT resolution
__global__ void __launch_bounds__(256) pool_tall_kernel(const float* __restrict__ src, long ps_in, float* __restrict__ dst, long ps_out, int n_img) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long total = (long)n_img * C4 * HW * HW;
  if (t >= total) return;
  const int pix = (int)(t % (HW * HW));
  const int kc = (int)((t / (HW * HW)) % C4);
  const int img = (int)(t / ((long)HW * HW * C4));
  const int y = pix / HW, x = pix % HW;
  float4 m;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long pos = TALL_MARGIN + tall_pos(2 * HW, img, 2 * y + (q >> 1), 2 * x + (q & 1));
    const float4 h = ld4(src + ((size_t)kc * ps_in + pos) * 4), l = ld4(src + ((size_t)(C4 + kc) * ps_in + pos) * 4);
    const float4 v = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
    m = q == 0 ? v : make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
  }
  float4 h, l;
  split4(m, h, l);
  const long pos = TALL_MARGIN + tall_pos(HW, img, y, x);
  st4(dst + ((size_t)kc * ps_out + pos) * 4, h);
  st4(dst + ((size_t)(C4 + kc) * ps_out + pos) * 4, l);
}

// ---- mbarrier / bulk-copy helpers (beyond tc.cuh) ----------------------------------------------------
namespace tc {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk async copy (TMA engine, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
}  // namespace tc

// MODE 0: 3x3 conv + bias + ReLU, same resolution (unet.py conv3x3 + F.relu)
// MODE 1: ConvTranspose2d(k2,s2) + bias (unet.py:25-31): blockIdx.y = (a,b) output parity, output at 2x resolution
template <int HW_, int CIN0_, int CIN1_, int COUT_, int NTILE_, int NT_, int MODE_, bool FUSE_FINAL_>
struct TallCfg {
  static constexpr int HW = HW_, CIN0 = CIN0_, CIN1 = CIN1_, CIN = CIN0_ + CIN1_, COUT = COUT_;
  static constexpr int NTILE = NTILE_, NT = NT_, MODE = MODE_;
  static constexpr bool FUSE_FINAL = FUSE_FINAL_;
  static constexpr int WP = HW + 2, HP1 = HW + 1;
  static constexpr int NTAPS = MODE == 0 ? 9 : 1;
  static constexpr int NC = CIN / 8;                         // 8-channel chunks (one K=8 MMA step per tap)
  static constexpr int G = MODE == 0 ? 1 : NC;               // chunks per accumulator group (drain period)
  static constexpr int NG = NC / G;
  static constexpr int MT = NT * 128;
  static constexpr int HALO = MODE == 0 ? WP + 1 : 0;
  static constexpr int ROWS_WIN = MT + 2 * HALO;
  static constexpr int KS_A = ROWS_WIN * 16;                 // bytes per staged k-chunk
  static constexpr int A_BYTES = 4 * KS_A;                   // hi kc0, hi kc1, lo kc0, lo kc1
  static constexpr int B_BYTES = NTAPS * 2 * 2 * NTILE * 16; // [tap][kc][hi|lo][n][4]
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NNT = MODE == 0 ? COUT / NTILE : 4;   // grid.y
  static constexpr int ACC = NT * NTILE;                     // columns per accumulator set
  static constexpr int TMEM_NEED = 4 * ACC;                  // X, Y (hi*hi, alternating), Z1 (lo*hi), Z2 (hi*lo); conv_final reuses X
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : 256;
  static constexpr int OFF_BAR = 2 * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_BAR + 96;
  static constexpr int SMEM_BYTES = OFF_BIAS + 64 * 4;
  static constexpr int FIN_KS = 128 * 16 + 16;
  static constexpr int FIN_A_BYTES = 8 * FIN_KS;
  static constexpr int NTHREADS = 160;
  static_assert(CIN0 % 8 == 0 && CIN1 % 8 == 0 && NTILE % 32 == 0 && NTILE <= 64 && NC % G == 0 && NC >= 2, "tiling");
  static_assert(MODE == 1 || COUT % NTILE == 0, "N tiling");
  static_assert(MODE == 0 || NTILE == COUT, "transpose conv: one (a,b) parity class per N tile");
  static_assert(TMEM_NEED <= 256, "TMEM budget (2 CTAs/SM)");
  static_assert(!FUSE_FINAL || (MODE == 0 && COUT == 32 && NTILE == 32 && 2 * FIN_A_BYTES + 8192 <= 2 * STAGE_BYTES), "fused conv_final");
  static_assert(HALO + MT + 128 <= TALL_MARGIN, "margin");
  static int num_ctas(int n_img) { return ceil_div((int)ceil_div((long)tall_positions(HW, n_img), 128L), NT); }
  // packed weights: [ntile][chunk][tap][kc 2][hi|lo][n NTILE][4]
  static constexpr long weight_floats() { return (long)NNT * NC * (B_BYTES / 4); }
};

// grid (num_ctas, NNT), block 160, dynamic smem SMEM_BYTES
template <class K>
__global__ void __launch_bounds__(160, 2)
conv_tall_kernel(const float* __restrict__ src0, long ps0,   // TALL [2][CIN0/4][ps0][4]
                 const float* __restrict__ src1, long ps1,   // TALL [2][CIN1/4][ps1][4] or null
                 const float* __restrict__ wt,               // packed weights (see TallCfg)
                 const float* __restrict__ bias,             // [COUT]
                 float* __restrict__ out, long pso,          // TALL [2][COUT/4][pso][4] at HW (MODE 0) or 2*HW (MODE 1)
                 const float* __restrict__ fin_w, const float* __restrict__ fin_b, float* __restrict__ fin_out,  // FUSE_FINAL
                 int n_img, unsigned long long* __restrict__ tl) {   // tl: optional per-CTA timeline (32 x u64), debug only
  constexpr int HW = K::HW, WP = K::WP, HP1 = K::HP1, NTILE = K::NTILE, NT = K::NT, ACC = K::ACC;
  extern __shared__ __align__(128) uint8_t smem_tall[];
  uint8_t* smem = smem_tall;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);  // [2]
  uint64_t* empty = full + 2;                                        // [2]
  uint64_t* acc_full = full + 4;                                     // [2]
  uint64_t* acc_empty = full + 6;                                    // [2]
  uint64_t* fin_bar = full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + K::OFF_BAR + 80);
  float* sbias = reinterpret_cast<float*>(smem + K::OFF_BIAS);   // this CTA's NTILE biases (first touch costs an L2 round trip)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int f0 = blockIdx.x * K::MT;
  const int nt_idx = blockIdx.y;
  const int total_rows = tall_rows(HW, n_img);
  unsigned long long* tlc = tl ? tl + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 : nullptr;
  auto stamp = [&](int slot) {
    if (tlc) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      tlc[slot] = t;
    }
  };
  if (tid == 0) stamp(0);

  if (warp == 4) tc::tmem_alloc(tmem_slot, K::TMEM_COLS);
  if (tid < NTILE) sbias[tid] = __ldg(bias + (K::MODE == 0 ? blockIdx.y * NTILE : 0) + tid);
  if (tid == 0) {
    tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
    tc::mbar_init(&empty[0], 1); tc::mbar_init(&empty[1], 1);
    tc::mbar_init(&acc_full[0], 1); tc::mbar_init(&acc_full[1], 1);
    tc::mbar_init(&acc_empty[0], 128); tc::mbar_init(&acc_empty[1], 128);
    tc::mbar_init(fin_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ===================== control warp: bulk copies + MMA issue by one elected lane =====================
    // (the whole warp runs the loop so that addresses/descriptors stay warp-uniform -> uniform registers)
    constexpr uint32_t IDESC = tc::make_idesc_tf32(128, NTILE);
    const float* wt_cta = wt + (size_t)nt_idx * K::NC * (K::B_BYTES / 4);
    auto loads = [&](int c) {
      const int s = c & 1;
      uint8_t* stA = smem + s * K::STAGE_BYTES;
      if (c >= 2) tc::mbar_wait(&empty[s], (uint32_t)(((c - 2) >> 1) & 1));   // MMAs of chunk c-2 have consumed the stage
      const int ch0 = c * 8;
      const float* sp; long ps; int kch, kcn;
      if (ch0 < K::CIN0) { sp = src0; ps = ps0; kch = ch0 / 4; kcn = K::CIN0 / 4; }
      else { sp = src1; ps = ps1; kch = (ch0 - K::CIN0) / 4; kcn = K::CIN1 / 4; }
      const long pos0 = TALL_MARGIN + f0 - K::HALO;
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)K::STAGE_BYTES);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
            tc::bulk_g2s(stA + (hl * 2 + kc) * K::KS_A, sp + ((size_t)(hl * kcn + kch + kc) * ps + pos0) * 4, K::KS_A, &full[s]);
        tc::bulk_g2s(stA + K::A_BYTES, wt_cta + (size_t)c * (K::B_BYTES / 4), K::B_BYTES, &full[s]);
      }
      __syncwarp();
    };
    if ((tid & 31) == 0) stamp(1);
    loads(0);
    loads(1);
#pragma unroll 1
    for (int c = 0; c < K::NC; ++c) {
      const int s = c & 1, g = c / K::G, set = g & 1;
      const bool group_first = (c % K::G) == 0, group_last = (c % K::G) == K::G - 1;
      tc::mbar_wait(&full[s], (uint32_t)((c >> 1) & 1));
      if ((tid & 31) == 0 && c < 4) stamp(2 + 2 * c);
      if (group_first && g >= 2) tc::mbar_wait(&acc_empty[set], (uint32_t)(((g - 2) >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t a_hi = tc::smem_u32(smem + s * K::STAGE_BYTES), a_lo = a_hi + 2 * K::KS_A, b0 = a_hi + K::A_BYTES;
      if (tc::elect_one()) {
        // Dependent MMAs into one TMEM accumulator serialise on its ~100-cycle read-modify-write latency
        // (measured: profiles/r01_conv_timeline.txt), far longer than a 128xNTILEx8 MMA executes (NTILE/2
        // cycles).  So consecutive MMAs go round-robin over 3*NT independent accumulator chains.
#pragma unroll
        for (int tap = 0; tap < K::NTAPS; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint32_t bb = b0 + tap * (4 * NTILE * 16);          // [kc][hi|lo][n][4]: k-chunk stride 2*NTILE*16
          const uint64_t bh = tc::make_desc(bb, 2 * NTILE * 16, 128);
          const uint64_t bl = tc::make_desc(bb + NTILE * 16, 2 * NTILE * 16, 128);
#pragma unroll
          for (int mt = 0; mt < NT; ++mt) {   // hi*hi -> the drained set
            const uint32_t aoff = (uint32_t)(mt * 128 + (K::MODE == 0 ? dy * WP + dx : 0)) * 16;
            tc::mma_tf32(tmem + set * ACC + mt * NTILE, tc::make_desc(a_hi + aoff, K::KS_A, 128), bh, IDESC, (group_first && tap == 0) ? 0u : 1u);
          }
#pragma unroll
          for (int mt = 0; mt < NT; ++mt) {   // lo*hi
            const uint32_t aoff = (uint32_t)(mt * 128 + (K::MODE == 0 ? dy * WP + dx : 0)) * 16;
            tc::mma_tf32(tmem + 2 * ACC + mt * NTILE, tc::make_desc(a_lo + aoff, K::KS_A, 128), bh, IDESC, (c == 0 && tap == 0) ? 0u : 1u);
          }
#pragma unroll
          for (int mt = 0; mt < NT; ++mt) {   // hi*lo
            const uint32_t aoff = (uint32_t)(mt * 128 + (K::MODE == 0 ? dy * WP + dx : 0)) * 16;
            tc::mma_tf32(tmem + 3 * ACC + mt * NTILE, tc::make_desc(a_hi + aoff, K::KS_A, 128), bl, IDESC, (c == 0 && tap == 0) ? 0u : 1u);
          }
        }
        tc::mma_commit(&empty[s]);
        if (group_last) tc::mma_commit(&acc_full[set]);
      }
      __syncwarp();
      if ((tid & 31) == 0 && c < 4) stamp(3 + 2 * c);
      if (c >= 1 && c + 1 < K::NC) loads(c + 1);
    }
    if ((tid & 31) == 0) stamp(10);
  } else {
    // =============================== drain + epilogue warps (thread t <-> TMEM lane t) ===============================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    float acc[NT][NTILE];
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
      for (int j = 0; j < NTILE; ++j) acc[mt][j] = 0.f;
#pragma unroll 1
    for (int g = 0; g < K::NG; ++g) {
      const int set = g & 1;
      tc::mbar_wait(&acc_full[set], (uint32_t)((g >> 1) & 1));
      tc::fence_after_sync();
      if (tid == 0 && g < 4) stamp(12 + g);
#pragma unroll
      for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int n0 = 0; n0 < NTILE; n0 += 32) {
          float v[32];
          tc::tmem_ld32(tmem_row + set * ACC + mt * NTILE + n0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[mt][n0 + j] += v[j];
        }
      tc::fence_before_sync();
      tc::mbar_arrive(&acc_empty[set]);
    }
    if (tid == 0) stamp(20);
    // the last group's commit covers every MMA, including the correction set
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
      for (int n0 = 0; n0 < NTILE; n0 += 32) {
        float v[32], w[32];
        tc::tmem_ld32(tmem_row + 2 * ACC + mt * NTILE + n0, v);
        tc::tmem_ld32(tmem_row + 3 * ACC + mt * NTILE + n0, w);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[mt][n0 + j] += v[j] + w[j];
      }

#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      const int f = f0 + mt * 128 + tid;
      const int r = f / WP, cc = f - r * WP;
      const int rr = r - 1, img = rr / HP1, y = rr - img * HP1, x = cc - 1;
      const bool valid = cc >= 1 && cc <= HW && r >= 1 && r < total_rows && y < HW;
      if constexpr (K::MODE == 0 && !K::FUSE_FINAL) {
        if (valid) {
          const long pos = TALL_MARGIN + f;
#pragma unroll
          for (int kc = 0; kc < NTILE / 4; ++kc) {
            const int co = nt_idx * NTILE + 4 * kc;
            const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
            const float4 v = make_float4(fmaxf(acc[mt][4 * kc + 0] + bv.x, 0.f), fmaxf(acc[mt][4 * kc + 1] + bv.y, 0.f),
                                         fmaxf(acc[mt][4 * kc + 2] + bv.z, 0.f), fmaxf(acc[mt][4 * kc + 3] + bv.w, 0.f));
            float4 h, l;
            split4(v, h, l);
            st4(out + ((size_t)(co / 4) * pso + pos) * 4, h);
            st4(out + ((size_t)(K::COUT / 4 + co / 4) * pso + pos) * 4, l);
          }
        }
      } else if constexpr (K::MODE == 1) {
        if (valid) {
          const int a = nt_idx >> 1, b = nt_idx & 1;
          const long pos = TALL_MARGIN + tall_pos(2 * HW, img, 2 * y + a, 2 * x + b);
#pragma unroll
          for (int kc = 0; kc < NTILE / 4; ++kc) {
            const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
            const float4 v = make_float4(acc[mt][4 * kc + 0] + bv.x, acc[mt][4 * kc + 1] + bv.y, acc[mt][4 * kc + 2] + bv.z, acc[mt][4 * kc + 3] + bv.w);
            float4 h, l;
            split4(v, h, l);
            st4(out + ((size_t)kc * pso + pos) * 4, h);
            st4(out + ((size_t)(K::COUT / 4 + kc) * pso + pos) * 4, l);
          }
        }
      } else {
        // conv_final (1x1, no activation) fused: relu(conv) -> A operand -> 32x32 3xTF32 MMA -> +bias -> channels-last
        uint8_t* fa_hi = smem;
        uint8_t* fa_lo = smem + K::FIN_A_BYTES;
        uint8_t* fw = smem + 2 * K::FIN_A_BYTES;   // [hi|lo][kc 8][n 32][4]
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
          const float4 v = make_float4(fmaxf(acc[mt][4 * kc + 0] + bv.x, 0.f), fmaxf(acc[mt][4 * kc + 1] + bv.y, 0.f),
                                       fmaxf(acc[mt][4 * kc + 2] + bv.z, 0.f), fmaxf(acc[mt][4 * kc + 3] + bv.w, 0.f));
          float4 h, l;
          split4(v, h, l);
          *reinterpret_cast<float4*>(fa_hi + kc * K::FIN_KS + tid * 16) = h;
          *reinterpret_cast<float4*>(fa_lo + kc * K::FIN_KS + tid * 16) = l;
        }
        if (mt == 0) {
          const float4* wsrc = reinterpret_cast<const float4*>(fin_w);
          float4* wdst = reinterpret_cast<float4*>(fw);
          for (int e = tid; e < 512; e += 128) wdst[e] = __ldg(wsrc + e);
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        tc::named_bar_sync(1, 128);
        if (warp == 0) {
          tc::fence_after_sync();
          const uint32_t ah0 = tc::smem_u32(fa_hi), al0 = tc::smem_u32(fa_lo), w0 = tc::smem_u32(fw);
          constexpr uint32_t ID32 = tc::make_idesc_tf32(128, 32);
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = tc::make_desc(ah0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
              const uint64_t al = tc::make_desc(al0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
              const uint64_t bh = tc::make_desc(w0 + ks * 2 * 512, 512, 128);
              const uint64_t bl = tc::make_desc(w0 + 4096 + ks * 2 * 512, 512, 128);
              tc::mma_tf32(tmem, ah, bh, ID32, ks > 0 ? 1u : 0u);
              tc::mma_tf32(tmem, al, bh, ID32, 1u);
              tc::mma_tf32(tmem, ah, bl, ID32, 1u);
            }
            tc::mma_commit(fin_bar);
          }
          __syncwarp();
        }
        tc::mbar_wait(fin_bar, (uint32_t)(mt & 1));
        tc::fence_after_sync();
        float v[32];
        tc::tmem_ld32(tmem_row, v);
        if (valid) {
          float* op = fin_out + ((size_t)img * (HW * HW) + y * HW + x) * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            st4(op + j, make_float4(v[j] + __ldg(fin_b + j), v[j + 1] + __ldg(fin_b + j + 1), v[j + 2] + __ldg(fin_b + j + 2),
                                    v[j + 3] + __ldg(fin_b + j + 3)));
        }
        tc::fence_before_sync();
        tc::named_bar_sync(1, 128);   // fa_* / final accumulator free for the next M tile
      }
    }
  }
  if (tid == 0) stamp(21);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, K::TMEM_COLS);
  if (tid == 0) {
    stamp(22);
    if (tlc) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tlc[31] = smid;
    }
  }
}


// =====================================================================================================
// Persistent variant: one CTA per SM walks a strided list of work items (M-tile group x N tile) so that
//   * the bulk-copy ring never drains between tiles (first-load latency paid once per CTA, not per tile),
//   * the epilogue of tile t (drain warps) overlaps the MMAs of tile t+1 (control warp runs up to two
//     chunks ahead through the alternating accumulator sets X/Y; the correction sets are double-buffered
//     across tiles),
//   * for the 40^2 layers the whole layer's weights (<= 147 KB) stay resident in shared memory (WRES).
// Same arithmetic / operand layouts / accuracy scheme as conv_tall_kernel above.
template <class K_, int S_, bool WRES_, bool NCONCAT_ = false>
struct PersistCfg {
  using K = K_;
  static constexpr int S = S_;
  static constexpr bool WRES = WRES_;
  // NCONCAT: hi*hi and hi*lo share the A operand, so they are ONE MMA against the row-concatenated B operand
  // [Bhi;Blo] (N = 2*NTILE): A is fetched once for both products (the A fetch is 80 % of an N=32 MMA's
  // shared-memory traffic, which is what bounds these MMAs).  Two issue warps: D1 = Ahi.[Bhi;Blo]^T, D2 = Alo.Bhi^T.
  static constexpr bool NCONCAT = NCONCAT_;
  static constexpr int NMMAW = NCONCAT ? 2 : 3;                                  // MMA-issue warps
  static constexpr int ACC1 = NCONCAT ? 2 * K_::ACC : K_::ACC;                   // columns of a drained set (X or Y)
  static constexpr int ZCOLS = NCONCAT ? K_::ACC : 2 * K_::ACC;                  // columns of one tile's correction accumulators
  static constexpr int OFF_Z = 2 * ACC1;
  static constexpr int OFF_FINACC = OFF_Z + 2 * ZCOLS;
  static constexpr int W_BYTES = WRES ? K::NC * K::B_BYTES : 0;                 // resident weights (one N tile)
  static constexpr int STAGE_BYTES = K::A_BYTES + (WRES ? 0 : K::B_BYTES);
  static constexpr int OFF_STAGES = W_BYTES;
  static constexpr int OFF_FIN = OFF_STAGES + S * STAGE_BYTES;                   // conv_final scratch (FUSE_FINAL)
  static constexpr int FIN_BYTES = K::FUSE_FINAL ? 2 * K::FIN_A_BYTES + 8192 : 0;
  static constexpr int OFF_BIAS = OFF_FIN + FIN_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 64 * 4;
  static constexpr int NBAR = 2 * S + 10;                                        // full[S] empty[S] acc_full[2] acc_empty[2] z_full[2] z_empty[2] wbar fin
  static constexpr int NTHREADS = 32 * (5 + NMMAW);                              // 4 drain warps + loader warp + MMA-issue warps
  static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr int ACC = K::ACC;
  static constexpr int TMEM_COLS = 512;                                          // X Y | Z1a Z2a | Z1b Z2b | final  (6*ACC + 32 <= 512)
  static_assert(OFF_FINACC + 32 <= 512, "TMEM");
  static_assert(!NCONCAT || K::NTILE == 32, "N-concat doubles the MMA's N; kept for the NTILE=32 layers");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(!WRES || K::NNT == 1, "resident weights cover one N tile");
};

// A single thread issues a 128xNTILEx8 tcgen05.mma only every ~85 cycles (measured, profiles/r01i: one issuing
// warp per SM ran at 93 cycles/MMA) while the MMA itself is shared-memory-bound at ~40 cycles, so three
// warps issue concurrently: warp 5 the hi*hi products (drained sets X/Y), warps 6/7 the two correction
// products (sets Z1/Z2).  Warp 4 only moves data (bulk copies), warps 0-3 drain and run the epilogue.
// grid (min(#items, #SMs)), block 256, dynamic smem P::SMEM_BYTES, 1 CTA / SM
template <class P>
__global__ void __launch_bounds__(P::NTHREADS, 1)
conv_tall_persistent_kernel(const float* __restrict__ src0, long ps0, const float* __restrict__ src1, long ps1,
                            const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out, long pso,
                            const float* __restrict__ fin_w, const float* __restrict__ fin_b, float* __restrict__ fin_out,
                            int n_img, int n_groups, unsigned long long* __restrict__ tl) {   // tl: optional stall accounting (debug)
  using K = typename P::K;
  constexpr int HW = K::HW, WP = K::WP, HP1 = K::HP1, NTILE = K::NTILE, NT = K::NT, ACC = K::ACC, S = P::S, NC = K::NC;
  extern __shared__ __align__(128) uint8_t smem_pt[];
  uint8_t* smem = smem_pt;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
  uint64_t* empty = full + S;
  uint64_t* acc_full = empty + S;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* z_full = acc_empty + 2;
  uint64_t* z_empty = z_full + 2;
  uint64_t* wbar = z_empty + 2;
  uint64_t* fin_bar = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::OFF_BAR + P::NBAR * 8);
  float* sbias = reinterpret_cast<float*>(smem + P::OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int total_rows = tall_rows(HW, n_img);
  const int n_items = n_groups * K::NNT;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items blockIdx.x + j*gridDim.x

  if (warp == 4) tc::tmem_alloc(tmem_slot, P::TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], P::NMMAW); }   // every MMA warp releases a stage
    tc::mbar_init(&acc_full[0], 1); tc::mbar_init(&acc_full[1], 1);
    tc::mbar_init(&acc_empty[0], 128); tc::mbar_init(&acc_empty[1], 128);
    tc::mbar_init(&z_full[0], P::NMMAW - 1); tc::mbar_init(&z_full[1], P::NMMAW - 1);
    tc::mbar_init(&z_empty[0], 128); tc::mbar_init(&z_empty[1], 128);
    tc::mbar_init(wbar, 1); tc::mbar_init(fin_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (K::FUSE_FINAL && tid < 128) {   // conv_final weights: staged once per CTA
    const float4* wsrc = reinterpret_cast<const float4*>(fin_w);
    float4* wdst = reinterpret_cast<float4*>(smem + P::OFF_FIN + 2 * K::FIN_A_BYTES);
    for (int e = tid; e < 512; e += 128) wdst[e] = __ldg(wsrc + e);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int total_chunks = my_items * NC;
  // debug stall accounting: cycles spent in each kind of wait, per role (slots: 32 u64 per CTA)
  unsigned long long* tlc = tl ? tl + (size_t)blockIdx.x * 32 : nullptr;
  long long w_a = 0, w_b = 0, w_c = 0;
  const long long t_begin = clock64();
  auto twait = [&](uint64_t* bar, uint32_t parity, long long& accum) {
    const long long t = clock64();
    tc::mbar_wait(bar, parity);
    accum += clock64() - t;
  };

  if (warp == 4) {
    // ============================ loader warp: bulk copies only ============================
    if (P::WRES && my_items > 0) {   // the layer's weights, resident for the whole kernel (single N tile)
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(wbar, (uint32_t)P::W_BYTES);
        for (int c = 0; c < NC; ++c)
          tc::bulk_g2s(smem + c * K::B_BYTES, wt + (size_t)c * (K::B_BYTES / 4), K::B_BYTES, wbar);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int i = 0; i < total_chunks; ++i) {   // i = flat chunk index of this CTA
      const int s = i % S, j = i / NC, c = i - j * NC;
      const int item = (int)blockIdx.x + j * (int)gridDim.x;
      const int grp = item / K::NNT, nt = item - grp * K::NNT;
      if (i >= S) twait(&empty[s], (uint32_t)(((i / S) - 1) & 1), w_a);   // all three MMA warps are done with the stage
      uint8_t* stA = smem + P::OFF_STAGES + s * P::STAGE_BYTES;
      const int ch0 = c * 8;
      const float* sp; long ps; int kch, kcn;
      if (ch0 < K::CIN0) { sp = src0; ps = ps0; kch = ch0 / 4; kcn = K::CIN0 / 4; }
      else { sp = src1; ps = ps1; kch = (ch0 - K::CIN0) / 4; kcn = K::CIN1 / 4; }
      const long pos0 = TALL_MARGIN + (long)grp * K::MT - K::HALO;
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)P::STAGE_BYTES);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
            tc::bulk_g2s(stA + (hl * 2 + kc) * K::KS_A, sp + ((size_t)(hl * kcn + kch + kc) * ps + pos0) * 4, K::KS_A, &full[s]);
        if (!P::WRES) tc::bulk_g2s(stA + K::A_BYTES, wt + ((size_t)nt * NC + c) * (K::B_BYTES / 4), K::B_BYTES, &full[s]);
      }
      __syncwarp();
    }
    if (tlc && (tid & 31) == 0) { tlc[0] = (unsigned long long)w_a; tlc[1] = (unsigned long long)(clock64() - t_begin); }
  } else if (warp >= 5) {
    // ============================ MMA-issue warps ============================
    // 3-product form: 5 = hi*hi, 6 = lo*hi, 7 = hi*lo.   N-concat form: 5 = Ahi.[Bhi;Blo]^T, 6 = Alo.Bhi^T.
    const int kind = warp - 5;
    constexpr int NMMA = (P::NCONCAT ? 2 : 1) * NTILE;             // N of warp 5's MMAs
    const uint32_t idesc = (P::NCONCAT && kind == 0) ? tc::make_idesc_tf32(128, NMMA) : tc::make_idesc_tf32(128, NTILE);
    if (P::WRES && my_items > 0) tc::mbar_wait(wbar, 0u);
#pragma unroll 1
    for (int i = 0; i < total_chunks; ++i) {
      const int s = i % S, j = i / NC, c = i - j * NC, set = i & 1, zp = j & 1;
      twait(&full[s], (uint32_t)((i / S) & 1), w_a);
      if (kind == 0) {
        if (i >= 2) twait(&acc_empty[set], (uint32_t)(((i - 2) >> 1) & 1), w_b);
      } else if (c == 0 && j >= 2) {
        twait(&z_empty[zp], (uint32_t)(((j - 2) >> 1) & 1), w_b);
      }
      const long long t_issue = clock64();
      tc::fence_after_sync();
      const uint32_t a_hi = tc::smem_u32(smem + P::OFF_STAGES + s * P::STAGE_BYTES), a_lo = a_hi + 2 * K::KS_A;
      const uint32_t b0 = P::WRES ? tc::smem_u32(smem + c * K::B_BYTES) : a_hi + K::A_BYTES;
      const uint32_t a_base = kind == 1 ? a_lo : a_hi;
      const uint32_t b_off = (!P::NCONCAT && kind == 2) ? NTILE * 16 : 0;          // [kc][hi|lo][n][4]: lo rows follow the hi rows
      const uint32_t d_stride = (P::NCONCAT && kind == 0) ? 2 * NTILE : NTILE;     // accumulator columns per M tile
      const uint32_t d_base = kind == 0 ? tmem + set * P::ACC1 : tmem + P::OFF_Z + zp * P::ZCOLS + (kind - 1) * ACC;
      const bool fresh = kind == 0 ? true : (c == 0);   // first MMA of the chain overwrites the accumulator
      if (tc::elect_one()) {
#pragma unroll
        for (int tap = 0; tap < K::NTAPS; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint64_t bd = tc::make_desc(b0 + tap * (4 * NTILE * 16) + b_off, 2 * NTILE * 16, 128);
#pragma unroll
          for (int mt = 0; mt < NT; ++mt) {
            const uint32_t aoff = (uint32_t)(mt * 128 + (K::MODE == 0 ? dy * WP + dx : 0)) * 16;
            tc::mma_tf32(d_base + mt * d_stride, tc::make_desc(a_base + aoff, K::KS_A, 128), bd, idesc, (fresh && tap == 0) ? 0u : 1u);
          }
        }
        tc::mma_commit(&empty[s]);
        if (kind == 0) tc::mma_commit(&acc_full[set]);
        else if (c == NC - 1) tc::mma_commit(&z_full[zp]);
      }
      __syncwarp();
      w_c += clock64() - t_issue;
    }
    if (tlc && (tid & 31) == 0) {
      tlc[4 + 4 * kind] = (unsigned long long)w_a; tlc[5 + 4 * kind] = (unsigned long long)w_b;
      tlc[6 + 4 * kind] = (unsigned long long)w_c; tlc[7 + 4 * kind] = (unsigned long long)(clock64() - t_begin);
    }
  } else {
    // ============================ drain + epilogue warps ============================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t fin_use = 0;
#pragma unroll 1
    for (int j = 0; j < my_items; ++j) {
      const int item = (int)blockIdx.x + j * (int)gridDim.x;
      const int grp = item / K::NNT, nt_idx = item - grp * K::NNT;
      const int f0 = grp * K::MT, zp = j & 1;
      if (tid < NTILE) sbias[tid] = __ldg(bias + (K::MODE == 0 ? nt_idx * NTILE : 0) + tid);
      float acc[NT][NTILE];
#pragma unroll
      for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int q = 0; q < NTILE; ++q) acc[mt][q] = 0.f;
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
        const int i = j * NC + c, set = i & 1;
        twait(&acc_full[set], (uint32_t)((i >> 1) & 1), w_a);
        tc::fence_after_sync();
        const long long t_d = clock64();
#pragma unroll
        for (int mt = 0; mt < NT; ++mt)
#pragma unroll
          for (int n0 = 0; n0 < NTILE; n0 += 32) {
            float v[32];
            if constexpr (P::NCONCAT) {   // columns [mt*2N, +N) = hi*hi, [mt*2N+N, +N) = hi*lo
              float w[32];
              tc::tmem_ld32(tmem_row + set * P::ACC1 + mt * 2 * NTILE + n0, v);
              tc::tmem_ld32(tmem_row + set * P::ACC1 + mt * 2 * NTILE + NTILE + n0, w);
#pragma unroll
              for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += v[q] + w[q];
            } else {
              tc::tmem_ld32(tmem_row + set * P::ACC1 + mt * NTILE + n0, v);
#pragma unroll
              for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += v[q];
            }
          }
        tc::fence_before_sync();
        tc::mbar_arrive(&acc_empty[set]);
        w_b += clock64() - t_d;
      }
      const long long t_e = clock64();
      // the two correction sets of this tile (issued by warps 6 and 7)
      twait(&z_full[zp], (uint32_t)((j >> 1) & 1), w_a);
      tc::fence_after_sync();
#pragma unroll
      for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int n0 = 0; n0 < NTILE; n0 += 32) {
          float v[32];
          tc::tmem_ld32(tmem_row + P::OFF_Z + zp * P::ZCOLS + mt * NTILE + n0, v);
#pragma unroll
          for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += v[q];
          if constexpr (!P::NCONCAT) {
            tc::tmem_ld32(tmem_row + P::OFF_Z + zp * P::ZCOLS + ACC + mt * NTILE + n0, v);
#pragma unroll
            for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += v[q];
          }
        }
      tc::fence_before_sync();
      tc::mbar_arrive(&z_empty[zp]);
      tc::named_bar_sync(1, 128);   // sbias visible to all epilogue threads (and previous tile's readers are done)

#pragma unroll
      for (int mt = 0; mt < NT; ++mt) {
        const int f = f0 + mt * 128 + tid;
        const int r = f / WP, cc = f - r * WP;
        const int rr = r - 1, img = rr / HP1, y = rr - img * HP1, x = cc - 1;
        const bool valid = cc >= 1 && cc <= HW && r >= 1 && r < total_rows && y < HW;
        if constexpr (K::MODE == 0 && !K::FUSE_FINAL) {
          if (valid) {
            const long pos = TALL_MARGIN + f;
#pragma unroll
            for (int kc = 0; kc < NTILE / 4; ++kc) {
              const int co = nt_idx * NTILE + 4 * kc;
              const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
              const float4 v = make_float4(fmaxf(acc[mt][4 * kc + 0] + bv.x, 0.f), fmaxf(acc[mt][4 * kc + 1] + bv.y, 0.f),
                                           fmaxf(acc[mt][4 * kc + 2] + bv.z, 0.f), fmaxf(acc[mt][4 * kc + 3] + bv.w, 0.f));
              float4 h, l;
              split4(v, h, l);
              st4(out + ((size_t)(co / 4) * pso + pos) * 4, h);
              st4(out + ((size_t)(K::COUT / 4 + co / 4) * pso + pos) * 4, l);
            }
          }
        } else if constexpr (K::MODE == 1) {
          if (valid) {
            const int a = nt_idx >> 1, b = nt_idx & 1;
            const long pos = TALL_MARGIN + tall_pos(2 * HW, img, 2 * y + a, 2 * x + b);
#pragma unroll
            for (int kc = 0; kc < NTILE / 4; ++kc) {
              const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
              const float4 v = make_float4(acc[mt][4 * kc + 0] + bv.x, acc[mt][4 * kc + 1] + bv.y, acc[mt][4 * kc + 2] + bv.z, acc[mt][4 * kc + 3] + bv.w);
              float4 h, l;
              split4(v, h, l);
              st4(out + ((size_t)kc * pso + pos) * 4, h);
              st4(out + ((size_t)(K::COUT / 4 + kc) * pso + pos) * 4, l);
            }
          }
        } else {
          uint8_t* fa_hi = smem + P::OFF_FIN;
          uint8_t* fa_lo = fa_hi + K::FIN_A_BYTES;
          uint8_t* fw = fa_lo + K::FIN_A_BYTES;
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            const float4 bv = *reinterpret_cast<const float4*>(sbias + 4 * kc);
            const float4 v = make_float4(fmaxf(acc[mt][4 * kc + 0] + bv.x, 0.f), fmaxf(acc[mt][4 * kc + 1] + bv.y, 0.f),
                                         fmaxf(acc[mt][4 * kc + 2] + bv.z, 0.f), fmaxf(acc[mt][4 * kc + 3] + bv.w, 0.f));
            float4 h, l;
            split4(v, h, l);
            *reinterpret_cast<float4*>(fa_hi + kc * K::FIN_KS + tid * 16) = h;
            *reinterpret_cast<float4*>(fa_lo + kc * K::FIN_KS + tid * 16) = l;
          }
          tc::fence_smem_to_async();
          tc::fence_before_sync();
          tc::named_bar_sync(1, 128);
          if (warp == 0) {
            tc::fence_after_sync();
            const uint32_t ah0 = tc::smem_u32(fa_hi), al0 = tc::smem_u32(fa_lo), w0 = tc::smem_u32(fw);
            constexpr uint32_t ID32 = tc::make_idesc_tf32(128, 32);
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ah = tc::make_desc(ah0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
                const uint64_t al = tc::make_desc(al0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
                const uint64_t bh = tc::make_desc(w0 + ks * 2 * 512, 512, 128);
                const uint64_t bl = tc::make_desc(w0 + 4096 + ks * 2 * 512, 512, 128);
                tc::mma_tf32(tmem + P::OFF_FINACC, ah, bh, ID32, ks > 0 ? 1u : 0u);
                tc::mma_tf32(tmem + P::OFF_FINACC, al, bh, ID32, 1u);
                tc::mma_tf32(tmem + P::OFF_FINACC, ah, bl, ID32, 1u);
              }
              tc::mma_commit(fin_bar);
            }
            __syncwarp();
          }
          tc::mbar_wait(fin_bar, fin_use & 1u);
          ++fin_use;
          tc::fence_after_sync();
          float v[32];
          tc::tmem_ld32(tmem_row + P::OFF_FINACC, v);
          if (valid) {
            float* op = fin_out + ((size_t)img * (HW * HW) + y * HW + x) * 32;
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              st4(op + q, make_float4(v[q] + __ldg(fin_b + q), v[q + 1] + __ldg(fin_b + q + 1), v[q + 2] + __ldg(fin_b + q + 2),
                                      v[q + 3] + __ldg(fin_b + q + 3)));
          }
          tc::fence_before_sync();
          tc::named_bar_sync(1, 128);   // scratch / final accumulator free for the next M tile
        }
      }
      tc::named_bar_sync(1, 128);   // everyone is done with sbias before the next item overwrites it
      w_c += clock64() - t_e;
    }
    if (tlc && tid == 0) {
      tlc[16] = (unsigned long long)w_a; tlc[17] = (unsigned long long)w_b; tlc[18] = (unsigned long long)w_c;
      tlc[19] = (unsigned long long)(clock64() - t_begin); tlc[20] = (unsigned long long)my_items;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, P::TMEM_COLS);
}

}  // namespace giga
