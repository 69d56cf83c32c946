// Tensor-core U-Net on "tall" pre-split activations: bulk-async (TMA-class) staging + tcgen05 + TMEM.
//
// TALL layout of an activation with C channels at resolution HW (all 3B plane images stacked):
//   rows: 0 = zero pad; image i at rows 1+i*(HW+1) .. i*(HW+1)+HW; one zero row after every image;
//   cols: 0 and HW+1 are zero pad (WP = HW+2);  flattened position f = row*WP + col;
//   memory: [hi|lo][C/8 k-chunks][PS positions][8 channels] fp16, position f stored at MARGIN+f, i.e. 16-byte
//   elements (the code addresses them as 4 floats);  hi = fp16_rn(v), lo = fp16_rn((v - hi) * 2^11): v - hi is exact
//   in fp32 and the 2^11 scale keeps lo a NORMAL fp16 over the whole fp16 range of v, so hi + lo*2^-11 carries 22
//   significant bits (fp32-class) at every magnitude.  Each 16-byte element is one row of one k-chunk of a K-major
//   SWIZZLE_NONE UMMA operand (tc.cuh).
// Consequences:
//   * the operand window a CTA needs (its 128*NT output positions plus a one-row/one-column halo, for
//     one k-chunk) is ONE contiguous byte range in global memory -> a single cp.async.bulk
//     (UBLKCP, completion on an mbarrier) per (hi|lo, k-chunk); zero padding is stored, not computed;
//   * the 9 taps of a 3x3 conv are the same staged window advanced by (dy*WP+dx) rows = a start-address
//     offset in the matrix descriptor (hardware-validated: tools/tc_probe.cu T5/T6);
//   * producers write the split in their epilogue (the values are in registers anyway).
// Accuracy: 3xFP16 operand splitting, x*y ~= xh*yh + xl*yh + xh*yl (fp16 has tf32's 11-bit significand, so this is the
// 3xTF32 scheme at twice the K per MMA and half the operand bytes: tools/tc_f16_probe.cu, profiles/r01s_tc_f16_probe.txt;
// measured 2e-7 relative vs fp64).  Tensor-core accumulation truncates, so error grows with the accumulate chain; the
// hi*hi products therefore go to one of two alternating TMEM accumulator sets that are drained into fp32 REGISTERS after
// every 16-channel chunk (<= 9 MMAs per chain, round-to-nearest adds on the CUDA cores), while the 2^11-scaled
// correction products accumulate in separate sets and are folded in with one FMA (x 2^-11) at the end of the tile.
// Range: activations and weights must stay inside fp16's +-65504 (GIGA's are O(10)).  Nothing is clamped: a larger value becomes the pair
// (inf, -inf), the next contraction makes it NaN, ReLU / max-pool propagate NaN (max.NaN), so the plane features and every head output
// computed from them read NaN -- counted by giga_ctx_overflow_count().
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc.cuh"

namespace giga {

constexpr int TALL_MARGIN = 640;   // zero positions before/after the image stack (>= WP+1 + 2*128 + slack)
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;
constexpr float H_MAX = 65504.f;

__host__ __device__ constexpr int tall_wp(int hw) { return hw + 2; }
__host__ __device__ constexpr int tall_rows(int hw, int n_img) { return 1 + n_img * (hw + 1); }
__host__ __device__ constexpr long tall_positions(int hw, int n_img) { return (long)tall_rows(hw, n_img) * tall_wp(hw); }
// plane stride (positions) of a TALL tensor sized for n_img images
__host__ __device__ constexpr long tall_ps(int hw, int n_img) { return 2 * TALL_MARGIN + ((tall_positions(hw, n_img) + 255) / 256) * 256; }
// storage in floats (= 4-byte words): [hi|lo][ch/8][ps] 16-byte elements
__host__ __device__ constexpr long tall_floats(int hw, int n_img, int ch) { return 2L * (ch / 8) * tall_ps(hw, n_img) * 4; }

// 8 fp32 values -> one 16-byte hi element and one 16-byte (scaled) lo element
__device__ __forceinline__ void split8(const float* v, uint4& h, uint4& l) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn((v[2 * i] - hf.x) * LO_SCALE, (v[2 * i + 1] - hf.y) * LO_SCALE);
    hw[i] = *reinterpret_cast<const uint32_t*>(&hh);
    lw[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
// the value a (hi, lo) pair stands for
__device__ __forceinline__ void join8(const uint4& h, const uint4& l, float* v) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    v[2 * i] = fmaf(lf.x, LO_INV, hf.x);
    v[2 * i + 1] = fmaf(lf.y, LO_INV, hf.y);
  }
}
__device__ __forceinline__ uint4 ldu4(const float* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stu4(float* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ long tall_pos(int hw, int img, int y, int x) { return (long)(1 + img * (hw + 1) + y) * (hw + 2) + x + 1; }

// ---- layout conversion / pooling (thread = one valid pixel x one 8-channel k-chunk) --------------------------
// grid: ceil(n_img*HW*HW*C8 / 256), block 256
template <int HW, int C8>
__global__ void __launch_bounds__(256) tall_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, long ps, int n_img) {
  pdl_launch();
  pdl_wait();
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long total = (long)n_img * C8 * HW * HW;
  if (t >= total) return;
  const int pix = (int)(t % (HW * HW));
  const int kc = (int)((t / (HW * HW)) % C8);
  const int img = (int)(t / ((long)HW * HW * C8));
  const long pos = TALL_MARGIN + tall_pos(HW, img, pix / HW, pix % HW);
  float v[8];
  join8(ldu4(src + ((size_t)kc * ps + pos) * 4), ldu4(src + ((size_t)(C8 + kc) * ps + pos) * 4), v);
  float* p = dst + ((size_t)img * (8 * C8) + 8 * kc) * (HW * HW) + pix;
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j * HW * HW] = v[j];
}

// MaxPool2d(2,2) (unet.py:64): TALL(2*HW) -> TALL(HW) on the values the (hi, lo) pairs stand for
template <int HW, int C8>   // HW = OUTPUT resolution
__global__ void __launch_bounds__(256) pool_tall_kernel(const float* __restrict__ src, long ps_in, float* __restrict__ dst, long ps_out, int n_img) {
  pdl_launch();
  pdl_wait();
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long total = (long)n_img * C8 * HW * HW;
  if (t >= total) return;
  const int pix = (int)(t % (HW * HW));
  const int kc = (int)((t / (HW * HW)) % C8);
  const int img = (int)(t / ((long)HW * HW * C8));
  const int y = pix / HW, x = pix % HW;
  float m[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long pos = TALL_MARGIN + tall_pos(2 * HW, img, 2 * y + (q >> 1), 2 * x + (q & 1));
    float v[8];
    join8(ldu4(src + ((size_t)kc * ps_in + pos) * 4), ldu4(src + ((size_t)(C8 + kc) * ps_in + pos) * 4), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = q == 0 ? v[j] : max_nan(m[j], v[j]);
  }
  uint4 h, l;
  split8(m, h, l);
  const long pos = TALL_MARGIN + tall_pos(HW, img, y, x);
  stu4(dst + ((size_t)kc * ps_out + pos) * 4, h);
  stu4(dst + ((size_t)(C8 + kc) * ps_out + pos) * 4, l);
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
// One arrival per WARP (the barrier's count is in warps): 32 per-thread arrivals are 32 serialised shared-memory atomics on one word;
// tools/tc_hop_probe.cu measured ~300 cycles per hand-off for 256 arriving threads.  Each lane orders its own prior accesses with its
// fence, the warp barrier orders the lanes before the elected lane's release.
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
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
  static constexpr int NC = CIN / 16;                        // 16-channel chunks (one K=16 fp16 MMA step per tap)
  static constexpr int G = MODE == 0 ? 1 : NC;               // chunks per accumulator group (drain period)
  static constexpr int NG = NC / G;
  static constexpr int MT = NT * 128;
  static constexpr int HALO = MODE == 0 ? WP + 1 : 0;
  static constexpr int ROWS_WIN = MT + 2 * HALO;
  static constexpr int KS_A = ROWS_WIN * 16;                 // bytes per staged k-chunk
  static constexpr int A_BYTES = 4 * KS_A;                   // hi kc0, hi kc1, lo kc0, lo kc1
  static constexpr int B_BYTES = NTAPS * 2 * 2 * NTILE * 16; // [tap][kc][hi|lo][n][8 halfs]
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NNT = MODE == 0 ? COUT / NTILE : 4;   // grid.y
  static constexpr int ACC = NT * NTILE;                     // columns per accumulator set
  static constexpr int TMEM_NEED = 4 * ACC;                  // X, Y (hi*hi, alternating), Z1 (lo*hi), Z2 (hi*lo); conv_final reuses X
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : 256;
  static constexpr int OFF_BAR = 2 * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_BAR + 96;
  static constexpr int SMEM_BYTES = OFF_BIAS + 64 * 4;
  static constexpr int FIN_KS = 128 * 16 + 16;
  static constexpr int FIN_A_BYTES = 4 * FIN_KS;             // conv_final A operand: K = 32 = 4 k-chunks of 8 halfs
  static constexpr int NTHREADS = 160;
  static_assert(CIN0 % 16 == 0 && CIN1 % 16 == 0 && NTILE % 32 == 0 && NTILE <= 64 && NC % G == 0 && NC >= 2, "tiling");
  static_assert(MODE == 1 || COUT % NTILE == 0, "N tiling");
  static_assert(MODE == 0 || NTILE == COUT, "transpose conv: one (a,b) parity class per N tile");
  static_assert(TMEM_NEED <= 256, "TMEM budget (2 CTAs/SM)");
  static_assert(!FUSE_FINAL || (MODE == 0 && COUT == 32 && NTILE == 32 && 2 * FIN_A_BYTES + 4096 <= 2 * STAGE_BYTES), "fused conv_final");
  static_assert(HALO + MT + 128 <= TALL_MARGIN, "margin");
  static int num_ctas(int n_img) { return ceil_div((int)ceil_div((long)tall_positions(HW, n_img), 128L), NT); }
  // packed weights: [ntile][chunk][tap][kc 2][hi|lo][n NTILE][8 halfs]
  static constexpr long weight_floats() { return (long)NNT * NC * (B_BYTES / 4); }
};

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
  static constexpr int FIN_BYTES = K::FUSE_FINAL ? 2 * K::FIN_A_BYTES + 4096 : 0;
  static constexpr int OFF_BIAS = OFF_FIN + FIN_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 64 * 4;
  static constexpr int NBAR = 2 * S + 19;                                        // full[S] empty[S] acc_full[2] acc_empty[2] z_full[2] z_empty[2] wbar fin pub item[8]
  static constexpr int NTHREADS = 32 * (6 + NMMAW);                              // 4 drain warps + loader warp + MMA-issue warps + publisher warp
  static constexpr int PUB_WARP = 5 + NMMAW;
  static constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16 + 32;                // + TMEM slot + item ring
  static constexpr int ACC = K::ACC;
  static constexpr int TMEM_COLS = 512;                                          // X Y | Z1a Z2a | Z1b Z2b | final  (6*ACC + 32 <= 512)
  static_assert(OFF_FINACC + 64 <= 512, "TMEM");   // conv_final: hi*hi columns + scaled-correction columns
  static_assert(!NCONCAT || K::NTILE == 32, "N-concat doubles the MMA's N; kept for the NTILE=32 layers");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(!WRES || K::NNT == 1, "resident weights cover one N tile");
};

// Tile-level dependencies between consecutive persistent layers (same resolution): instead of waiting for the WHOLE
// previous layer (griddepcontrol.wait), a consumer layer's loader waits until the producer has finished exactly the
// position groups its operand window covers.  The producer's epilogue publishes a group with
// __threadfence + barrier + atomicAdd (release); the consumer polls with ld.acquire.gpu and orders its bulk copies
// (async proxy) behind the acquire with fence.proxy.async.  A consumer CTA can only become resident on an SM whose
// producer CTA has exited (both hold all 512 TMEM columns), and every producer CTA is resident before the consumer is
// launched, so producers never wait for consumers: no deadlock.  Consumers walk the item list from the other end
// (`reverse`), so the SMs that got one item fewer in the producer get one more in the consumer: the per-layer tail
// (item quantisation over 148 SMs) and the consumer's prologue disappear into the producer's tail.
// The flags are zeroed once per encode by nchw_to_tall_kernel (which every layer of the step transitively follows).
struct LayerDep {
  unsigned* ready_out;        // producer side: [n_groups] counters of this layer's finished (group, N tile) items, or null
  const unsigned* ready_in;   // consumer side: the previous layer's counters, or null (then: griddepcontrol.wait)
  int in_mt;                  // producer's positions per group
  int in_groups;              // producer's number of groups
  unsigned in_target;         // producer's N tiles per group
  int in_half_res;            // 1: the producer is the transpose conv one level down (its groups index positions at HW/2)
  int reverse;                // walk the item list from the last SM
  int early_trigger;          // grid == #SMs: the successor cannot become resident before one of this grid's CTAs exits
  unsigned* work;             // dynamic item scheduling: global counter of this layer's next unclaimed item (null: static strided list)
  unsigned long long* dbg;    // debug (GIGA_LAYER_TIMES): [min CTA start, max CTA start, min CTA end, max CTA end] in ns (globaltimer)
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded poll: a broken dependency must trap, never hang the GPU
__device__ __forceinline__ void wait_groups(const unsigned* ready, int g_lo, int g_hi, unsigned target) {
  for (int g = g_lo; g <= g_hi; ++g) {
    bool ok = false;
    for (int i = 0; i < (1 << 22) && !ok; ++i) {
      ok = ld_acquire_gpu(ready + g) >= target;
      if (!ok) __nanosleep(40);
    }
    if (!ok) __trap();
  }
  asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes acquired above -> visible to the bulk copies below
}

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
                            int n_img, int n_groups, const LayerDep dep,
                            unsigned long long* __restrict__ tl) {   // tl: optional stall accounting (debug)
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
  uint64_t* pub_bar = fin_bar + 1;   // 4 epilogue-warp arrivals per item -> publisher warp (LayerDep)
  uint64_t* item_bar = pub_bar + 1;  // [8] the loader has published item j (ring slot j & 7)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::OFF_BAR + P::NBAR * 8);
  // Items (M-tile group x N tile) are handed out by the loader warp: from a global counter when dep.work is set (the CTAs that
  // become resident first simply take more items -- no per-layer tail from the static 148-way split), else the strided list.
  // The other roles learn the item (or -1 = no more) from an 8-slot ring; the loader is never more than 5 items ahead.
  volatile int* ring = reinterpret_cast<volatile int*>(smem + P::OFF_BAR + P::NBAR * 8 + 16);
  float* sbias = reinterpret_cast<float*>(smem + P::OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int total_rows = tall_rows(HW, n_img);
  const int n_items = n_groups * K::NNT;
  const int bid = dep.reverse ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;

  // A layer that waits for its whole predecessor releases its own successor only after that wait (below): the successor
  // may then rely on everything older being complete.  A tile-dependent consumer releases its successor at once.
  if (dep.ready_in || dep.early_trigger) pdl_launch();
  if (dep.dbg && tid == 0) {
    const unsigned long long t = globaltimer_ns();
    atomicMin(dep.dbg + 0, t);
    atomicMax(dep.dbg + 1, t);
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, P::TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], P::NMMAW); }   // every MMA warp releases a stage
    tc::mbar_init(&acc_full[0], 1); tc::mbar_init(&acc_full[1], 1);
    tc::mbar_init(&acc_empty[0], 4); tc::mbar_init(&acc_empty[1], 4);   // drain side: one arrival per warp
    tc::mbar_init(&z_full[0], P::NMMAW - 1); tc::mbar_init(&z_full[1], P::NMMAW - 1);
    tc::mbar_init(&z_empty[0], 4); tc::mbar_init(&z_empty[1], 4);
    tc::mbar_init(wbar, 1); tc::mbar_init(fin_bar, 1); tc::mbar_init(pub_bar, 4);
    for (int i = 0; i < 8; ++i) tc::mbar_init(&item_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (K::FUSE_FINAL && tid < 128) {   // conv_final weights: staged once per CTA
    const float4* wsrc = reinterpret_cast<const float4*>(fin_w);
    float4* wdst = reinterpret_cast<float4*>(smem + P::OFF_FIN + 2 * K::FIN_A_BYTES);
    for (int e = tid; e < 256; e += 128) wdst[e] = __ldg(wsrc + e);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 4 && P::WRES) {   // the layer's weights, resident for the whole kernel (single N tile): constants,
    if (tc::elect_one()) {                       // so their copy overlaps the wait for the previous layer
      tc::mbar_arrive_expect_tx(wbar, (uint32_t)P::W_BYTES);
      for (int c = 0; c < NC; ++c)
        tc::bulk_g2s(smem + c * K::B_BYTES, wt + (size_t)c * (K::B_BYTES / 4), K::B_BYTES, wbar);
    }
    __syncwarp();
  }
  if (!dep.ready_in) {
    pdl_wait();     // activations of the previous layer(s) are complete and visible (weights / biases are constants)
    if (!dep.early_trigger) pdl_launch();   // small grids: release the successor only now (it may land on an SM without a CTA of ours)
  }
  auto next_item = [&](int j) -> int {   // every role but the loader: wait for the loader's verdict on item j
    tc::mbar_wait(&item_bar[j & 7], (uint32_t)((j >> 3) & 1));
    return ring[j & 7];
  };
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
#pragma unroll 1
    for (int j = 0;; ++j) {
      int item = -1;
      if ((tid & 31) == 0) {
        item = dep.work ? (int)atomicAdd(dep.work, 1u) : bid + j * (int)gridDim.x;
        if (item >= n_items) item = -1;
        ring[j & 7] = item;
        tc::mbar_arrive(&item_bar[j & 7]);   // release: the ring entry is visible to whoever sees the phase complete
      }
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item < 0) break;
      const int grp = item / K::NNT, nt = item - grp * K::NNT;
      if (dep.ready_in) {   // the producer groups this item's operand window [f0 - HALO, f0 + MT + HALO) covers
        long lo = (long)grp * K::MT - K::HALO, hi = (long)grp * K::MT + K::MT + K::HALO - 1;
        if (lo < 0) lo = 0;
        if (dep.in_half_res) {   // position at HW -> the HW/2 position whose transpose-conv outputs land there (monotone, clamped into the images)
          auto down = [&](long f) -> long {
            long r = f / WP - 1;
            if (r < 0) r = 0;
            long img = r / HP1;
            if (img > n_img - 1) img = n_img - 1;
            long y = r - img * HP1, x = f % WP - 1;
            y = y < 0 ? 0 : (y > HW - 1 ? HW - 1 : y);
            x = x < 0 ? 0 : (x > HW - 1 ? HW - 1 : x);
            return (1 + img * (HW / 2 + 1) + y / 2) * (HW / 2 + 2) + x / 2 + 1;
          };
          lo = down(lo);
          hi = down(hi);
        }
        const int g_lo = (int)(lo / dep.in_mt);
        int g_hi = (int)(hi / dep.in_mt);
        if (g_hi > dep.in_groups - 1) g_hi = dep.in_groups - 1;
        if (tc::elect_one()) wait_groups(dep.ready_in, g_lo, g_hi, dep.in_target);
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {   // i = flat chunk index of this CTA
      const int i = j * NC + c, s = i % S;
      if (i >= S) twait(&empty[s], (uint32_t)(((i / S) - 1) & 1), w_a);   // all three MMA warps are done with the stage
      uint8_t* stA = smem + P::OFF_STAGES + s * P::STAGE_BYTES;
      const int ch0 = c * 16;
      const float* sp; long ps; int kch, kcn;
      if (ch0 < K::CIN0) { sp = src0; ps = ps0; kch = ch0 / 8; kcn = K::CIN0 / 8; }
      else { sp = src1; ps = ps1; kch = (ch0 - K::CIN0) / 8; kcn = K::CIN1 / 8; }
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
    }
    if (tlc && (tid & 31) == 0) { tlc[0] = (unsigned long long)w_a; tlc[1] = (unsigned long long)(clock64() - t_begin); }
  } else if (warp == P::PUB_WARP) {
    // ============================ publisher warp (tile-level dependencies) ============================
    // The device-wide release fence costs ~1 us; done here it delays nobody in this CTA (done by an epilogue thread it
    // stalled the drain warps' chunk barriers and made the layer 2.5 us slower: profiles/r01w_layer_times.txt).
    if (dep.ready_out) {
#pragma unroll 1
      for (int j = 0;; ++j) {
        const int item = next_item(j);
        if (item < 0) break;
        tc::mbar_wait(pub_bar, (uint32_t)(j & 1));   // all 128 epilogue threads have stored the item (arrive = release, wait = acquire)
        if (tc::elect_one()) {
          __threadfence();                           // cumulative: their stores are ordered before the counter device-wide
          atomicAdd(dep.ready_out + item / K::NNT, 1u);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 5) {
    // ============================ MMA-issue warps ============================
    // 3-product form: 5 = hi*hi, 6 = lo*hi, 7 = hi*lo.   N-concat form: 5 = Ahi.[Bhi;Blo]^T, 6 = Alo.Bhi^T.
    const int kind = warp - 5;
    constexpr int NMMA = (P::NCONCAT ? 2 : 1) * NTILE;             // N of warp 5's MMAs
    const uint32_t idesc = (P::NCONCAT && kind == 0) ? tc::make_idesc_f16(128, NMMA) : tc::make_idesc_f16(128, NTILE);
    if (P::WRES) tc::mbar_wait(wbar, 0u);
#pragma unroll 1
    for (int j = 0;; ++j) {
      if (next_item(j) < 0) break;
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
      const int i = j * NC + c, s = i % S, set = i & 1, zp = j & 1;
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
            tc::mma_f16(d_base + mt * d_stride, tc::make_desc(a_base + aoff, K::KS_A, 128), bd, idesc, (fresh && tap == 0) ? 0u : 1u);
          }
        }
        tc::mma_commit(&empty[s]);
        if (kind == 0) tc::mma_commit(&acc_full[set]);
        else if (c == NC - 1) tc::mma_commit(&z_full[zp]);
      }
      __syncwarp();
      w_c += clock64() - t_issue;
      }
    }
    if (tlc && (tid & 31) == 0) {
      tlc[4 + 4 * kind] = (unsigned long long)w_a; tlc[5 + 4 * kind] = (unsigned long long)w_b;
      tlc[6 + 4 * kind] = (unsigned long long)w_c; tlc[7 + 4 * kind] = (unsigned long long)(clock64() - t_begin);
    }
  } else {
    // ============================ drain + epilogue warps ============================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t fin_use = 0;
    int n_done = 0;
#pragma unroll 1
    for (int j = 0;; ++j) {
      const int item = next_item(j);
      if (item < 0) break;
      ++n_done;
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
              for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += fmaf(w[q], LO_INV, v[q]);
            } else {
              tc::tmem_ld32(tmem_row + set * P::ACC1 + mt * NTILE + n0, v);
#pragma unroll
              for (int q = 0; q < 32; ++q) acc[mt][n0 + q] += v[q];
            }
          }
        tc::fence_before_sync();
        tc::warp_arrive(&acc_empty[set]);
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
          for (int q = 0; q < 32; ++q) acc[mt][n0 + q] = fmaf(v[q], LO_INV, acc[mt][n0 + q]);
          if constexpr (!P::NCONCAT) {
            tc::tmem_ld32(tmem_row + P::OFF_Z + zp * P::ZCOLS + ACC + mt * NTILE + n0, v);
#pragma unroll
            for (int q = 0; q < 32; ++q) acc[mt][n0 + q] = fmaf(v[q], LO_INV, acc[mt][n0 + q]);
          }
        }
      tc::fence_before_sync();
      tc::warp_arrive(&z_empty[zp]);
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
            for (int kc = 0; kc < NTILE / 8; ++kc) {
              const int co = nt_idx * NTILE + 8 * kc;
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = relu_nan(acc[mt][8 * kc + q] + sbias[8 * kc + q]);
              uint4 h, l;
              split8(v, h, l);
              stu4(out + ((size_t)(co / 8) * pso + pos) * 4, h);
              stu4(out + ((size_t)(K::COUT / 8 + co / 8) * pso + pos) * 4, l);
            }
          }
        } else if constexpr (K::MODE == 1) {
          if (valid) {
            const int a = nt_idx >> 1, b = nt_idx & 1;
            const long pos = TALL_MARGIN + tall_pos(2 * HW, img, 2 * y + a, 2 * x + b);
#pragma unroll
            for (int kc = 0; kc < NTILE / 8; ++kc) {
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = acc[mt][8 * kc + q] + sbias[8 * kc + q];
              uint4 h, l;
              split8(v, h, l);
              stu4(out + ((size_t)kc * pso + pos) * 4, h);
              stu4(out + ((size_t)(K::COUT / 8 + kc) * pso + pos) * 4, l);
            }
          }
        } else {
          uint8_t* fa_hi = smem + P::OFF_FIN;
          uint8_t* fa_lo = fa_hi + K::FIN_A_BYTES;
          uint8_t* fw = fa_lo + K::FIN_A_BYTES;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = relu_nan(acc[mt][8 * kc + q] + sbias[8 * kc + q]);
            uint4 h, l;
            split8(v, h, l);
            *reinterpret_cast<uint4*>(fa_hi + kc * K::FIN_KS + tid * 16) = h;
            *reinterpret_cast<uint4*>(fa_lo + kc * K::FIN_KS + tid * 16) = l;
          }
          tc::fence_smem_to_async();
          tc::fence_before_sync();
          tc::named_bar_sync(1, 128);
          if (warp == 0) {
            tc::fence_after_sync();
            const uint32_t ah0 = tc::smem_u32(fa_hi), al0 = tc::smem_u32(fa_lo), w0 = tc::smem_u32(fw);
            constexpr uint32_t ID32 = tc::make_idesc_f16(128, 32);
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {   // K = 32 = 2 MMAs of K = 16; weights [hi|lo][kc 4][n 32][8 halfs]
                const uint64_t ah = tc::make_desc(ah0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
                const uint64_t al = tc::make_desc(al0 + ks * 2 * K::FIN_KS, K::FIN_KS, 128);
                const uint64_t bh = tc::make_desc(w0 + ks * 2 * 512, 512, 128);
                const uint64_t bl = tc::make_desc(w0 + 2048 + ks * 2 * 512, 512, 128);
                tc::mma_f16(tmem + P::OFF_FINACC, ah, bh, ID32, ks > 0 ? 1u : 0u);
                tc::mma_f16(tmem + P::OFF_FINACC + 32, al, bh, ID32, ks > 0 ? 1u : 0u);   // 2^11-scaled corrections
                tc::mma_f16(tmem + P::OFF_FINACC + 32, ah, bl, ID32, 1u);
              }
              tc::mma_commit(fin_bar);
            }
            __syncwarp();
          }
          tc::mbar_wait(fin_bar, fin_use & 1u);
          ++fin_use;
          tc::fence_after_sync();
          float v[32], z[32];
          tc::tmem_ld32(tmem_row + P::OFF_FINACC, v);
          tc::tmem_ld32(tmem_row + P::OFF_FINACC + 32, z);
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaf(z[q], LO_INV, v[q]);
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
      if (dep.ready_out) tc::warp_arrive(pub_bar);   // this warp's stores of the item are done -> publisher warp
      tc::named_bar_sync(1, 128);   // everyone is done with sbias before the next item overwrites it
      w_c += clock64() - t_e;
    }
    if (tlc && tid == 0) {
      tlc[16] = (unsigned long long)w_a; tlc[17] = (unsigned long long)w_b; tlc[18] = (unsigned long long)w_c;
      tlc[19] = (unsigned long long)(clock64() - t_begin); tlc[20] = (unsigned long long)n_done;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, P::TMEM_COLS);
  if (dep.dbg && tid == 0) {
    const unsigned long long t = globaltimer_ns();
    atomicMin(dep.dbg + 2, t);
    atomicMax(dep.dbg + 3, t);
  }
}

}  // namespace giga
