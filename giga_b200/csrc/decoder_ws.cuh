// Warp-specialised tensor-core implicit decoder (tcgen05 / TMEM, 3xFP16 operand splitting -> fp32-level accuracy).
//
// Function (reference, relative to src/vgn/ConvONets): tri-plane bilinear gather (conv_onet/models/decoder.py:117-122, common.py:238-261,
// ATen grid_sampler bilinear/border/align_corners) + LocalDecoder heads (decoder.py:133-176, layers.py:39-47) + head epilogues
// (conv_onet/models/__init__.py:119-123).
//
// Structure: ONE persistent CTA per SM, 320 threads; an item = one tile of 128 query points of one scene x the NH heads of its job:
//   warps 0-3   compute warpgroup: thread t <-> point t <-> TMEM lane t.  The tile's NH heads run as NH independent chains,
//               software-interleaved per thread: while the tensor core works on chain c the thread is in the epilogue of chain c+1
//               (the layer round trip of a single chain -- tcgen05.ld -> ALU -> tcgen05.st -> MMA -> commit -- is ~750 cycles, with
//               three chains interleaved ~440 cycles per chain-layer: tools/tc_ts_probe.cu);
//   warps 4-7   gather warpgroup: gathers the NEXT item's tri-plane features into the other of two shared-memory feature operands
//               while the compute warpgroup is busy with the current item (the gather is L2-latency bound, the chains ALU/tensor bound);
//   warp 8      MMA issue (one elected lane);
//   warp 9      work distribution (items handed out from a global counter: heavy 3-head items first) and weight streaming
//               (cp.async.bulk + mbarrier; fc_c slice single-buffered, chain weights double-buffered per block).
// Per chain c two 32-column TMEM tiles X[c], Y[c] alternate between accumulator and A operand.  Block b:
//     E1: h += Y (= fc_c_b + fc_1_{b-1}) + bias;   A = split(relu(h)) -> written over Y      MMA fc_0_b:  X <- A(Y) . W0^T
//     E2: A = split(relu(X + b0))                  -> written over X                          MMA fc_1_b:  Y += A(X) . W1^T
//     and right after the fc_0 MMAs of all chains were issued, fc_c_{b+1} of ALL the tile's heads: Y[0..NH) <- feat . Wc_{b+1}^T, one
//     N = 32 NH batch that runs on the tensor pipe while the threads are in E2 (fc_1 then accumulates on top) -- off the critical path.
//   * the chain layers' A operand lives in TENSOR MEMORY (tcgen05.st by the epilogue thread that owns the row; "TS" MMA form): no
//     shared-memory store, no proxy fence, and the MMA runs at its N/2-cycle floor (16 cycles for N = 32) instead of the 40 cycles the
//     shared-memory A fetch costs; the tensor pipe executes one thread's MMAs in issue order, which is what makes the X/Y reuse safe;
//   * the 96 tri-plane features are gathered ONCE per tile (lane = channel, coalesced 128 B texel reads), split and stored as the K-major
//     shared-memory A operand of the fc_c contractions of all heads and blocks;
//   * 3xFP16: x*y ~= xh*yh + xl*yh + xh*yl, all three products accumulate in ONE TMEM accumulator; per-head power-of-two weight pre-scale
//     2^s (host) keeps the weights' lo halves normal, epilogues multiply by 2^-s (exact);
//   * range: an activation or feature beyond fp16's 65504 turns into (inf, -inf) -> NaN accumulators; ReLU is max.NaN, so the NaN reaches the
//     head output and is counted in `overflow` (giga_ctx_overflow_count) -- loud, never a silent clamp.
#pragma once
#include "common.cuh"
#include "decoder.cuh"
#include "tc.cuh"
#include "unet_tall.cuh"   // tc::bulk_g2s / mbar_arrive / mbar_arrive_expect_tx

namespace giga {

// ---- head constants (floats, per head; staged in shared memory once per CTA) ----
constexpr int HC_FCP = 0;        // Wt[3][32], then b_p[32]                          (128)
constexpr int HC_OUT = 128;      // Wt[32][4] (zero padded), then b_out[4]           (132)
constexpr int HC_B1 = 260;       // fc_1 bias of the LAST block [32]
constexpr int HC_INV = 292;      // 2^-s
constexpr int HC_SIZE = 296;

// ---- streamed weight blobs, one per job type: [block 5][ fcc | chain ] ----
//   fcc   : [hi|lo][k-chunk 12][n = head*32 + j][8 halfs]                                        12288 * NH bytes
//   chain : W0 [head][hi|lo][k-chunk 4][n 32][8 halfs], W1 same, b0 [head][32] f32, bin [head][32] f32   (8192 + 256) * NH bytes
//           bin = the bias E1 of this block adds: bc_b (+ b1_{b-1} for b > 0)
__host__ __device__ constexpr int wd_fcc_bytes(int nh) { return 12288 * nh; }
__host__ __device__ constexpr int wd_chain_bytes(int nh) { return 8448 * nh; }
__host__ __device__ constexpr int wd_block_bytes(int nh) { return wd_fcc_bytes(nh) + wd_chain_bytes(nh); }

constexpr int WD_PTS = 128;                            // points per warpgroup tile = MMA M
constexpr int WD_KS_F = WD_PTS * 16 + 16;              // feature operand k-chunk stride (bytes); +16: conflict-free gather stores
constexpr int WD_FEAT_BYTES = 2 * 12 * WD_KS_F;        // hi + lo, K = 96: 49,536
constexpr int WD_OFF_FEAT = 0;                         // [2 buffers]: item k uses buffer k & 1
constexpr int WD_OFF_FCC = 2 * WD_FEAT_BYTES;          // 99,072
constexpr int WD_OFF_CH = WD_OFF_FCC + wd_fcc_bytes(3);          // 135,936: [2 stages]
constexpr int WD_OFF_TINFO = WD_OFF_CH + 2 * wd_chain_bytes(3);  // 186,624: [128 points][24 words] tap table of the tile being gathered
constexpr int WD_OFF_HC = WD_OFF_TINFO + WD_PTS * 24 * 4;        // 198,912: head constants of all four heads (4,736 B)
constexpr int WD_OFF_PTS = WD_OFF_HC + 4 * HC_SIZE * 4;          // 203,648: [2 buffers][128] float4 query points of the gathered tiles
constexpr int WD_OFF_FIN = WD_OFF_PTS + 2 * WD_PTS * 16;         // 207,744: [3 chains][128] float4 partial fc_out sums of the second column half
constexpr int WD_OFF_BAR = WD_OFF_FIN + 3 * WD_PTS * 16;         // 213,888
constexpr int WD_NBAR = 27;
constexpr int WD_SMEM_BYTES = WD_OFF_BAR + WD_NBAR * 8 + 32;
constexpr int WD_THREADS = 448;
constexpr int WD_TMEM_COLS = 512;                      // two item sets (item k uses set k & 1 at column 256 (k & 1)): X[c] = 32c, Y[c] = 96 + 32c
constexpr int WD_MAX_JOBS = 4;

struct DecJob {
  const float* pts;   // [B][N][3]
  int N;
  int type;           // 0: the three grasp heads (qual, rot, width) as one 3-chain tile; 1 + h: head h alone
  int tiles;          // tiles per scene = ceil(N / 128)
  int item0;          // first global item index of this job (items = B * tiles, scene-major)
};
struct DecArgs {
  DecJob job[WD_MAX_JOBS];
  int njobs, n_items, B;
  unsigned raw;                       // GIGA_HEAD_RAW: skip sigmoid / normalise
  const float* planes;                // [3][B][40][40][32]
  const float* hc;                    // [4][HC_SIZE]
  const uint8_t* wblob[5];            // per job type
  float *qual, *rot, *width, *occ;
  unsigned* sched;                    // [0] next item, [1] CTAs done, [2] overflow count
  unsigned long long* tl;             // optional debug timeline [grid][32]
  unsigned debug;                     // timing experiments only (results are wrong): 1 = gather skips the texel loads, 2 = no fc_c MMAs
};

namespace wd {
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// ASYNCHRONOUS load of 16 fp32 columns of this thread's lane: the registers are valid only after tmem_ld_wait16(r) -- which names them
// as in/out operands so that the compiler cannot move a use above the wait
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// One arrival per WARP (barrier counts are in warps): 32 per-thread arrivals are 32 serialised shared-memory atomics -- with 256 threads
// arriving the hand-off to the MMA-issue lane took ~300 cycles longer (tools/tc_hop_probe.cu).  Every lane's prior tcgen05 / shared-memory
// writes are ordered before the elected lane's release by its own fence + the warp barrier.
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) tc::mbar_arrive(bar);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(tc::smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
// 16 fp32 values -> 8 hi words + 8 lo words (fp16 pairs): one K = 16 step of the A operand
__device__ __forceinline__ void split16(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
}
// One 16-column half of an epilogue phase.  TYPE 1 (E1): hcol += acc * winv + bias; A = split(relu(hcol)).  TYPE 2 (E2): A = split(relu(acc * winv + bias)).
// The A operand of K-step ks lives in the columns of its own accumulator half: hi at t + 16 ks, lo at t + 16 ks + 8.
template <int TYPE>
__device__ __forceinline__ void phase_half(float* hcol, const uint32_t* acc, const float* bias, float winv, uint32_t t_half) {
  float v[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 q = *reinterpret_cast<const float4*>(bias + 4 * j4);
    const float bq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = 4 * j4 + i;
      if (TYPE == 1) {
        hcol[j] = fmaf(__uint_as_float(acc[j]), winv, hcol[j] + bq[i]);
        v[j] = relu_nan(hcol[j]);
      } else {
        v[j] = relu_nan(fmaf(__uint_as_float(acc[j]), winv, bq[i]));
      }
    }
  }
  uint32_t h[8], l[8];
  split16(v, h, l);
  tmem_st8(t_half, h);
  tmem_st8(t_half + 8, l);
}
__device__ __forceinline__ void job_of_item(const DecArgs& a, int item, int& j, int& b, int& tile) {
  j = 0;
#pragma unroll
  for (int q = 1; q < WD_MAX_JOBS; ++q)
    if (q < a.njobs && item >= a.job[q].item0) j = q;
  const int r = item - a.job[j].item0;
  b = r / a.job[j].tiles;
  tile = r - b * a.job[j].tiles;
}
}  // namespace wd


// barriers (arrival counts; thread groups arrive once per WARP): 0-1 feat_ready[buf] (4 gather warps) | 2-3 feat_free[buf] (1) | 4-6 a_ready[c]
//           (8 compute warps) | 7-12 acc_full[set][c] (1) | 13 wfull_fcc | 14 wempty_fcc (1) | 15-16 wfull_ch[2] | 17-18 wempty_ch[2] (1)
//           | 19-22 item_bar[4] (1) | 23-24 tmem_free[set] (8) | 25 fin_bar (4) | 26 feat_first (12: the CTA's first tile is gathered by
//           all twelve compute + gather warps -- at kernel start every SM gathers at once and nothing else can run yet)
// Items alternate between two TMEM column sets and two acc_full barrier sets: the first contraction of item k + 1 (fc_c of block 0) is issued
// while the compute threads are still in the fc_out epilogue of item k (which reads item k's set), and its completion must not advance a
// barrier the compute threads have not yet waited on for item k's last layer (an mbarrier may not run two phases ahead of a waiter).
constexpr int WB_FEAT_READY = 0, WB_FEAT_FREE = 2, WB_A_READY = 4, WB_ACC_FULL = 7, WB_WFULL_FCC = 13, WB_WEMPTY_FCC = 14, WB_WFULL_CH = 15,
              WB_WEMPTY_CH = 17, WB_ITEM = 19, WB_TMEM_FREE = 23, WB_FIN = 25, WB_FEAT_FIRST = 26;

// ---- gather warpgroup: one tile's 96 tri-plane features -> shared-memory A operand (hi | lo) of the fc_c contractions --------------
// One WARP gathers rows [row0, row0 + nrows) of the tile (nrows <= 32, a multiple of 4).
__device__ __forceinline__ void wd_gather_rows(const DecArgs& args, const DecJob& job, int b, int tile, uint8_t* smem, uint8_t* feat, float4* pout,
                                               int row0, int nrows) {
  const int lane = threadIdx.x & 31;
  const int N = job.N;
  const int n0 = tile * WD_PTS;
  float* tw_ = reinterpret_cast<float*>(smem + WD_OFF_TINFO) + row0 * 24;
  if (lane < nrows) {
    const int nq = min(n0 + row0 + lane, N - 1);   // rows beyond N evaluate the last point (their outputs are not stored)
    TexInfo t;
    const float* pq = job.pts + ((size_t)b * N + nq) * 3;
    point_taps(pq, t);
    pout[row0 + lane] = make_float4(pq[0], pq[1], pq[2], 0.f);   // the compute thread of this row reads it back (fc_p)
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
  for (int pl = 0; pl < 3; ++pl) pb[pl] = args.planes + ((size_t)pl * args.B + b) * (G2 * C) + lane;
  uint8_t* fa = feat + (lane >> 3) * WD_KS_F + (lane & 7) * 2 + row0 * 16;
  // lane = channel: every texel is one coalesced 128 B read.  Batches of 8 points: all 96 texel loads of a batch are issued before the
  // first dependent use (the feature stores below could alias the tap table as far as the compiler knows, so without the explicit
  // batching only one point's loads are in flight).
  constexpr int GB = 4;   // 48 loads in flight per warp (register budget: 448 threads -> 128 registers)
#pragma unroll 1
  for (int q0 = 0; q0 < ((args.debug & 1u) ? 0 : nrows); q0 += GB) {
    float tex[GB][12];
#pragma unroll
    for (int qq = 0; qq < GB; ++qq) {
      const int4* oi = reinterpret_cast<const int4*>(tw_ + (q0 + qq) * 24);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const int4 o = oi[pl];
        tex[qq][pl * 4 + 0] = __ldg(pb[pl] + o.x);
        tex[qq][pl * 4 + 1] = __ldg(pb[pl] + o.y);
        tex[qq][pl * 4 + 2] = __ldg(pb[pl] + o.z);
        tex[qq][pl * 4 + 3] = __ldg(pb[pl] + o.w);
      }
    }
#pragma unroll
    for (int qq = 0; qq < GB; ++qq) {
      const float4* wf = reinterpret_cast<const float4*>(tw_ + (q0 + qq) * 24 + 12);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const float4 w = wf[pl];
        const float f = tex[qq][pl * 4] * w.x + tex[qq][pl * 4 + 1] * w.y + tex[qq][pl * 4 + 2] * w.z + tex[qq][pl * 4 + 3] * w.w;   // ATen order: nw, ne, sw, se
        const __half hh = __float2half_rn(f);
        *reinterpret_cast<__half*>(fa + pl * 4 * WD_KS_F + (q0 + qq) * 16) = hh;
        *reinterpret_cast<__half*>(fa + 12 * WD_KS_F + pl * 4 * WD_KS_F + (q0 + qq) * 16) = __float2half_rn(f - __half2float(hh));
      }
    }
  }
  __syncwarp();   // the tap table is rewritten by this warp for its next item
}

// ---- compute warpgroups: the NH chains of one tile ---------------------------------------------------------------------------------
// TMEM columns: X[c] = 32c, Y[c] = 96 + 32c.  TWO threads per point: warpgroup `half` (0: warps 0-3, 1: warps 4-7; warp w and w + 4 share
// TMEM lane quadrant w and one scheduler) owns hidden columns [16 half, 16 half + 16) of every layer -- the epilogue math is elementwise,
// so the halves are independent until fc_out.  One warp per scheduler left 70 % of the issue slots idle (a phase is a dependent
// ld -> ~90 ALU instructions -> st sequence); two interleave.
template <int NH>
__device__ __forceinline__ void wd_compute_item(const DecArgs& args, const DecJob& job, int b, int tile, uint8_t* smem, int wtid, int half, uint32_t tmem,
                                                uint64_t* bars, uint32_t& pf, uint32_t u0, int k, unsigned long long* tlc, int& tslot) {
  const int set = k & 1;
  uint64_t* a_ready = bars + WB_A_READY;
  uint64_t* acc_full = bars + WB_ACC_FULL + 3 * set;
  uint64_t* wfull_ch = bars + WB_WFULL_CH;
  const int warp4 = wtid >> 5;
  const int N = job.N;
  const int n = tile * WD_PTS + wtid;
  const bool valid = n < N;
  const int head0 = job.type == 0 ? 0 : job.type - 1;
  const float* hcs = reinterpret_cast<const float*>(smem + WD_OFF_HC);
  const int col0 = 16 * half;
  auto stamp = [&]() {
    if (tlc && wtid == 0 && half == 0 && tslot < 21) tlc[tslot] = globaltimer_ns();
    ++tslot;
  };
  stamp();

  // ---- fc_p on CUDA cores; the point was parked in shared memory by the gather thread of this row ----
  if (k == 0) tc::mbar_wait(&bars[WB_FEAT_FIRST], 0u);
  else tc::mbar_wait(&bars[WB_FEAT_READY + (k & 1)], (uint32_t)(((k >> 1) - (1 - (k & 1))) & 1));
  const float4 pxyz = reinterpret_cast<const float4*>(smem + WD_OFF_PTS)[(k & 1) * WD_PTS + wtid];
  const float px = pxyz.x, py = pxyz.y, pz = pxyz.z;
  float h[NH][16];
  float winv[NH];
#pragma unroll
  for (int c = 0; c < NH; ++c) {
    const float* W = hcs + (head0 + c) * HC_SIZE + col0;
    winv[c] = hcs[(head0 + c) * HC_SIZE + HC_INV];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 wx = ld4(W + HC_FCP + 4 * j4), wy = ld4(W + HC_FCP + 32 + 4 * j4), wz = ld4(W + HC_FCP + 64 + 4 * j4), bb = ld4(W + HC_FCP + 96 + 4 * j4);
      h[c][4 * j4 + 0] = bb.x + wx.x * px + wy.x * py + wz.x * pz;
      h[c][4 * j4 + 1] = bb.y + wx.y * px + wy.y * py + wz.y * pz;
      h[c][4 * j4 + 2] = bb.z + wx.z * px + wy.z * py + wz.z * pz;
      h[c][4 * j4 + 3] = bb.w + wx.w * px + wy.w * py + wz.w * pz;
    }
  }

  const uint32_t trow = tmem + ((uint32_t)(warp4 * 32) << 16) + 256 * set + col0;
#pragma unroll 1
  for (int blk = 0; blk < 5; ++blk) {
    const uint32_t u = u0 + blk;
    tc::mbar_wait(&wfull_ch[u & 1], (u >> 1) & 1u);   // this block's biases are in shared memory
    const float* cb = reinterpret_cast<const float*>(smem + WD_OFF_CH + (u & 1) * wd_chain_bytes(3) + NH * 8192) + col0;   // b0 [NH][32], bin [NH][32]
#pragma unroll
    for (int ph = 0; ph < 2 * NH; ++ph) {
      // phase ph: E1 (net += fc_c + fc_1; A <- relu(net), over Y[c]) for ph < NH, E2 (A <- relu(fc_0 + b0), over X[c]) after
      const int c = ph % NH;
      const bool e1 = ph < NH;
      const uint32_t t_cur = trow + (e1 ? 96 : 0) + 32 * c;
      const float* bias = cb + (e1 ? NH * 32 : 0) + c * 32;
      tc::mbar_wait(&acc_full[c], (pf >> (3 * set + c)) & 1u);
      pf ^= 1u << (3 * set + c);
      tc::fence_after_sync();
      uint32_t va[16];
      wd::tmem_ld16_async(t_cur, va);
      wd::tmem_ld_wait16(va);
      if (e1) wd::phase_half<1>(&h[c][0], va, bias, winv[c], t_cur);
      else wd::phase_half<2>(nullptr, va, bias, winv[c], t_cur);
      wd::tmem_wait_st();
      tc::fence_before_sync();
      wd::warp_arrive(&a_ready[c]);
    }
    stamp();
  }

  // ---- last fc_1 (in Y), fc_out(relu(net)) + head epilogue on CUDA cores: each half sums its 16 inputs, half 0 combines and stores ----
  float4* fin = reinterpret_cast<float4*>(smem + WD_OFF_FIN);
  float4 o[NH];
#pragma unroll
  for (int c = 0; c < NH; ++c) {
    const int head = head0 + c;
    const float* W = hcs + head * HC_SIZE;
    tc::mbar_wait(&acc_full[c], (pf >> (3 * set + c)) & 1u);
    pf ^= 1u << (3 * set + c);
    tc::fence_after_sync();
    uint32_t va[16];
    wd::tmem_ld16_async(trow + 96 + 32 * c, va);
    wd::tmem_ld_wait16(va);
    if (c == NH - 1) {   // this thread's last TMEM read of the item: the next item's first MMA may overwrite Y
      tc::fence_before_sync();
      wd::warp_arrive(&bars[WB_TMEM_FREE + set]);
    }
    float v[16];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const float4 b1 = ld4(W + HC_B1 + col0 + 4 * k4);
      v[4 * k4 + 0] = relu_nan(fmaf(__uint_as_float(va[4 * k4 + 0]), winv[c], h[c][4 * k4 + 0] + b1.x));
      v[4 * k4 + 1] = relu_nan(fmaf(__uint_as_float(va[4 * k4 + 1]), winv[c], h[c][4 * k4 + 1] + b1.y));
      v[4 * k4 + 2] = relu_nan(fmaf(__uint_as_float(va[4 * k4 + 2]), winv[c], h[c][4 * k4 + 2] + b1.z));
      v[4 * k4 + 3] = relu_nan(fmaf(__uint_as_float(va[4 * k4 + 3]), winv[c], h[c][4 * k4 + 3] + b1.w));
    }
    o[c] = half == 0 ? ld4(W + HC_OUT + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (head == 1) {   // rot: 4 outputs
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 w = ld4(W + HC_OUT + (col0 + kk) * 4);
        o[c].x = fmaf(w.x, v[kk], o[c].x); o[c].y = fmaf(w.y, v[kk], o[c].y);
        o[c].z = fmaf(w.z, v[kk], o[c].z); o[c].w = fmaf(w.w, v[kk], o[c].w);
      }
    } else {           // one output: Wt[k][0]
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) o[c].x = fmaf(W[HC_OUT + (col0 + kk) * 4], v[kk], o[c].x);
    }
    if (half == 1) fin[c * WD_PTS + wtid] = o[c];
  }
  if (half == 1) {
    wd::warp_arrive(&bars[WB_FIN]);   // release: the partial sums are visible to the first-half threads of these rows
  } else {
    tc::mbar_wait(&bars[WB_FIN], (uint32_t)(k & 1));
#pragma unroll
    for (int c = 0; c < NH; ++c) {
      const int head = head0 + c;
      const float4 q = fin[c * WD_PTS + wtid];
      const float4 r = make_float4(o[c].x + q.x, o[c].y + q.y, o[c].z + q.z, o[c].w + q.w);
      if (valid) {
        if (!(fabsf(r.x) <= 3.0e38f)) atomicAdd(args.sched + 2, 1u);   // NaN / inf: an activation left fp16's range (or the input was not finite)
        const size_t idx = (size_t)b * N + n;
        if (head == 0) {
          args.qual[idx] = args.raw ? r.x : 1.f / (1.f + expf(-r.x));
        } else if (head == 1) {
          const float nrm = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
          const float inv = args.raw ? 1.f : 1.f / fmaxf(nrm, 1e-12f);   // one IEEE division, four multiplies (<= 1 ulp from x / d)
          st4(args.rot + idx * 4, make_float4(r.x * inv, r.y * inv, r.z * inv, r.w * inv));
        } else if (head == 2) {
          args.width[idx] = r.x;
        } else {
          args.occ[idx] = r.x;
        }
      }
    }
  }
  stamp();
}

// ---- the MMA-issue lane's work on one item ---------------------------------------------------------------------------------
template <int NH>
__device__ __forceinline__ void wd_issue_item(uint8_t* smem, uint32_t tmem, uint64_t* bars, uint32_t& pa, uint32_t u0, int k, unsigned debug, long long* iacc) {
  uint64_t* feat_ready = bars + WB_FEAT_READY;
  uint64_t* feat_free = bars + WB_FEAT_FREE;
  uint64_t* a_ready = bars + WB_A_READY;
  uint64_t* acc_full = bars + WB_ACC_FULL + 3 * (k & 1);
  uint64_t* wfull_fcc = bars + WB_WFULL_FCC;
  uint64_t* wempty_fcc = bars + WB_WEMPTY_FCC;
  uint64_t* wfull_ch = bars + WB_WFULL_CH;
  uint64_t* wempty_ch = bars + WB_WEMPTY_CH;
  constexpr uint32_t IDESC_C = tc::make_idesc_f16(128, 32 * NH), IDESC_32 = tc::make_idesc_f16(128, 32);
  constexpr uint32_t KS_WC = 32 * NH * 16;   // fc_c B operand k-chunk stride
  const int buf = k & 1;
  const uint32_t f_hi = tc::smem_u32(smem + WD_OFF_FEAT + buf * WD_FEAT_BYTES), f_lo = f_hi + 12 * WD_KS_F;
  const uint32_t wc_hi = tc::smem_u32(smem + WD_OFF_FCC), wc_lo = wc_hi + 12 * KS_WC;
  const uint32_t X = tmem + 256 * (k & 1), Y = X + 96;
  // descriptors are built once per item; advancing the start address by `bytes` is one add on the low word (addresses < 256 KB)
  const uint64_t d_fhi = tc::make_desc(f_hi, WD_KS_F, 128), d_flo = tc::make_desc(f_lo, WD_KS_F, 128);
  const uint64_t d_wchi = tc::make_desc(wc_hi, KS_WC, 128), d_wclo = tc::make_desc(wc_lo, KS_WC, 128);
  const uint64_t d_w = tc::make_desc(tc::smem_u32(smem + WD_OFF_CH), 512, 128);
  auto issue_fcc = [&](uint32_t u) {   // Y[0..NH) <- feat . Wc^T  (fresh)
    const long long ta = iacc ? clock64() : 0;
    tc::mbar_wait(wfull_fcc, u & 1u);
    tc::fence_after_sync();
    const long long tb = iacc ? clock64() : 0;
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      if ((debug & 2u) && ks > 0) break;
      const uint64_t ah = d_fhi + ((ks * 2 * WD_KS_F) >> 4), al = d_flo + ((ks * 2 * WD_KS_F) >> 4);
      const uint64_t bh = d_wchi + ((ks * 2 * KS_WC) >> 4), bl = d_wclo + ((ks * 2 * KS_WC) >> 4);
      tc::mma_f16(Y, ah, bh, IDESC_C, ks > 0 ? 1u : 0u);
      tc::mma_f16(Y, al, bh, IDESC_C, 1u);
      tc::mma_f16(Y, ah, bl, IDESC_C, 1u);
    }
    tc::mma_commit(wempty_fcc);
    if (iacc) { iacc[0] += tb - ta; iacc[1] += clock64() - tb; }
  };
  // D <- (+)= A . W^T with A at columns `a`: K-step ks has its hi halfs in 8 columns at a + 16 ks, its lo halfs in the next 8  (W at byte offset w_off of the chain stages: hi, lo at +2048; [k-chunk 4][n 32][8 halfs])
  auto issue_layer = [&](uint32_t d, uint32_t a, uint32_t w_off, uint32_t accumulate) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t bh = d_w + ((w_off + ks * 1024) >> 4), bl = d_w + ((w_off + 2048 + ks * 1024) >> 4);
      wd::mma_ts(d, a + ks * 16, bh, IDESC_32, (ks > 0 || accumulate) ? 1u : 0u);
      wd::mma_ts(d, a + ks * 16 + 8, bh, IDESC_32, 1u);
      wd::mma_ts(d, a + ks * 16, bl, IDESC_32, 1u);
    }
  };
  const long long t_item = iacc ? clock64() : 0;
  if (k == 0) tc::mbar_wait(&bars[WB_FEAT_FIRST], 0u);   // features of this item are in shared memory (the first tile has its own barrier)
  else tc::mbar_wait(&feat_ready[buf], (uint32_t)(((k >> 1) - (1 - (k & 1))) & 1));
  if (k > 1) tc::mbar_wait(&bars[WB_TMEM_FREE + (k & 1)], (uint32_t)(((k >> 1) - 1) & 1));   // the compute threads have read the last results of item k - 2 (same set)
  tc::fence_after_sync();
  if (iacc) iacc[6] += clock64() - t_item;
  issue_fcc(u0);
#pragma unroll
  for (int c = 0; c < NH; ++c) tc::mma_commit(&acc_full[c]);
#pragma unroll 1
  for (int blk = 0; blk < 5; ++blk) {
    const uint32_t u = u0 + blk;
    tc::mbar_wait(&wfull_ch[u & 1], (u >> 1) & 1u);
    const uint32_t w0 = (u & 1) * wd_chain_bytes(3), w1 = w0 + NH * 4096;   // byte offsets within the chain stages
#pragma unroll
    for (int c = 0; c < NH; ++c) {   // fc_0: X[c] <- A(Y[c]) . W0^T
      const long long ta = iacc ? clock64() : 0;
      tc::mbar_wait(&a_ready[c], (pa >> c) & 1u);
      pa ^= 1u << c;
      tc::fence_after_sync();
      const long long tb = iacc ? clock64() : 0;
      issue_layer(X + 32 * c, Y + 32 * c, w0 + c * 4096, 0u);
      tc::mma_commit(&acc_full[c]);
      if (iacc) { iacc[2] += tb - ta; iacc[3] += clock64() - tb; }
    }
    if (blk < 4) {   // fc_c of the next block: overwrites Y (its A contents were consumed by the fc_0 MMAs just issued); runs during E2
      issue_fcc(u + 1);
      if (blk == 3) tc::mma_commit(&feat_free[buf]);   // last reader of the feature operand
    }
#pragma unroll
    for (int c = 0; c < NH; ++c) {   // fc_1: Y[c] (+)= A(X[c]) . W1^T
      const long long ta = iacc ? clock64() : 0;
      tc::mbar_wait(&a_ready[c], (pa >> c) & 1u);
      pa ^= 1u << c;
      tc::fence_after_sync();
      const long long tb = iacc ? clock64() : 0;
      issue_layer(Y + 32 * c, X + 32 * c, w1 + c * 4096, blk < 4 ? 1u : 0u);
      tc::mma_commit(&acc_full[c]);
      if (iacc) { iacc[4] += tb - ta; iacc[5] += clock64() - tb; }
    }
    tc::mma_commit(&wempty_ch[u & 1]);
  }
}

// grid = min(#SMs, n_items), block WD_THREADS, dynamic smem WD_SMEM_BYTES, 1 CTA / SM
__global__ void __launch_bounds__(WD_THREADS, 1) decode_points_ws_kernel(const __grid_constant__ DecArgs args) {
  extern __shared__ __align__(128) uint8_t smem_wd[];
  uint8_t* smem = smem_wd;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WD_OFF_BAR);
  uint64_t* item_bar = bars + WB_ITEM;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WD_OFF_BAR + WD_NBAR * 8);
  volatile int* ring = reinterpret_cast<volatile int*>(smem + WD_OFF_BAR + WD_NBAR * 8 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 13) tc::tmem_alloc(tmem_slot, WD_TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars[WB_FEAT_READY + i], 4); tc::mbar_init(&bars[WB_FEAT_FREE + i], 1); }   // counts in warps
    for (int i = 0; i < 3; ++i) tc::mbar_init(&bars[WB_A_READY + i], 8);
    for (int i = 0; i < 6; ++i) tc::mbar_init(&bars[WB_ACC_FULL + i], 1);
    tc::mbar_init(&bars[WB_WFULL_FCC], 1); tc::mbar_init(&bars[WB_WEMPTY_FCC], 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars[WB_WFULL_CH + i], 1); tc::mbar_init(&bars[WB_WEMPTY_CH + i], 1); }
    for (int i = 0; i < 4; ++i) tc::mbar_init(&item_bar[i], 1);
    tc::mbar_init(&bars[WB_TMEM_FREE], 8);
    tc::mbar_init(&bars[WB_TMEM_FREE + 1], 8);
    tc::mbar_init(&bars[WB_FIN], 4);
    tc::mbar_init(&bars[WB_FEAT_FIRST], 12);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < 4 * HC_SIZE / 4; e += WD_THREADS)   // head constants (parameters: not written by the preceding kernels)
    reinterpret_cast<float4*>(smem + WD_OFF_HC)[e] = __ldg(reinterpret_cast<const float4*>(args.hc) + e);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // plane features / points of the preceding kernels are complete; the previous decoder launch has reset the scheduler words

  auto next_item = [&](int k) -> int {
    tc::mbar_wait(&item_bar[k & 3], (uint32_t)((k >> 2) & 1));
    return ring[k & 3];
  };
  unsigned long long* tlc = args.tl ? args.tl + (size_t)blockIdx.x * 32 : nullptr;

  if (warp == 13) {
    // ============================ scheduler + weight loader ============================
    if ((tid & 31) == 0) {
      uint64_t* wfull_fcc = bars + WB_WFULL_FCC;
      uint64_t* wempty_fcc = bars + WB_WEMPTY_FCC;
      uint64_t* wfull_ch = bars + WB_WFULL_CH;
      uint64_t* wempty_ch = bars + WB_WEMPTY_CH;
      int fetched = 0, last = 0;
      auto item_at = [&](int k) -> int {   // fetch (and publish) items up to ordinal k; -1 = no more work
        while (fetched <= k) {
          int it = last < 0 ? -1 : (int)atomicAdd(args.sched, 1u);
          if (it >= args.n_items) it = -1;
          last = it;
          ring[fetched & 3] = it;
          tc::mbar_arrive(&item_bar[fetched & 3]);
          ++fetched;
        }
        return ring[k & 3];
      };
      auto blob_of = [&](int item, int& nh) -> const uint8_t* {
        int j, b, tile;
        wd::job_of_item(args, item, j, b, tile);
        const int type = args.job[j].type;
        nh = type == 0 ? 3 : 1;
        return args.wblob[type];
      };
      auto load_fcc = [&](uint32_t u) -> bool {
        const int item = item_at((int)(u / 5));
        if (item < 0) return false;
        if (u % 5 == 0) item_at((int)(u / 5) + 1);   // the gather warpgroup works one item ahead of the chains
        int nh;
        const uint8_t* blob = blob_of(item, nh);
        if (u >= 1) tc::mbar_wait(wempty_fcc, (u - 1) & 1u);   // tight poll: the single fc_c buffer's refill is on the issue lane's critical path
        tc::mbar_arrive_expect_tx(wfull_fcc, (uint32_t)wd_fcc_bytes(nh));
        tc::bulk_g2s(smem + WD_OFF_FCC, blob + (size_t)(u % 5) * wd_block_bytes(nh), (uint32_t)wd_fcc_bytes(nh), wfull_fcc);
        return true;
      };
      auto load_ch = [&](uint32_t u) -> bool {
        const int item = item_at((int)(u / 5));
        if (item < 0) return false;
        int nh;
        const uint8_t* blob = blob_of(item, nh);
        if (u >= 2) tc::mbar_wait_relaxed(&wempty_ch[u & 1], ((u >> 1) - 1) & 1u);
        tc::mbar_arrive_expect_tx(&wfull_ch[u & 1], (uint32_t)wd_chain_bytes(nh));
        tc::bulk_g2s(smem + WD_OFF_CH + (u & 1) * wd_chain_bytes(3), blob + (size_t)(u % 5) * wd_block_bytes(nh) + wd_fcc_bytes(nh),
                     (uint32_t)wd_chain_bytes(nh), &wfull_ch[u & 1]);
        return true;
      };
      // Buffers become free in this order: fcc(u) once the fc_c batch of use u-1 has completed (it is issued early in block u-2 / at the
      // item start), ch(u+1) once block u-1 has completed.  Uses of items that do not exist are skipped (load_* returns false).
      bool more = load_fcc(0);
      if (more) { load_ch(0); load_ch(1); }
      for (uint32_t u = 1; more; ++u) {
        const bool a = load_fcc(u);
        const bool c = load_ch(u + 1);
        more = a || c;
      }
      // all items handed out: the last CTA to get here re-arms the scheduler for the next launch
      if (atomicAdd(args.sched + 1, 1u) == gridDim.x - 1) {
        args.sched[0] = 0u;
        args.sched[1] = 0u;
      }
    }
    __syncwarp();
  } else if (warp == 12) {
    // ============================ MMA issue ============================
    if (tc::elect_one()) {
      uint32_t pa = 0u;
      long long iacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // debug cycle accounting of the issue lane (tools/decode_timeline.py)
      long long* iacc = tlc ? iacc_ : nullptr;
      const long long t_begin = clock64();
#pragma unroll 1
      for (int k = 0;; ++k) {
        const int item = next_item(k);
        if (item < 0) break;
        int j, b, tile;
        wd::job_of_item(args, item, j, b, tile);
        if (args.job[j].type == 0) wd_issue_item<3>(smem, tmem, bars, pa, (uint32_t)k * 5u, k, args.debug, iacc);
        else wd_issue_item<1>(smem, tmem, bars, pa, (uint32_t)k * 5u, k, args.debug, iacc);
      }
      if (tlc) {
        for (int i = 0; i < 7; ++i) tlc[i == 0 ? 28 : 20 + i] = (unsigned long long)iacc_[i];   // slots 28, 21..26
        tlc[27] = (unsigned long long)(clock64() - t_begin);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ============================ gather warpgroup (one item ahead) ============================
    const int wtid = tid - 256;
#pragma unroll 1
    for (int k = 0;; ++k) {
      tc::mbar_wait_relaxed(&item_bar[k & 3], (uint32_t)((k >> 2) & 1));
      const int item = ring[k & 3];
      if (item < 0) break;
      int j, b, tile;
      wd::job_of_item(args, item, j, b, tile);
      if (k >= 2) tc::mbar_wait_relaxed(&bars[WB_FEAT_FREE + (k & 1)], (uint32_t)(((k >> 1) - 1) & 1));   // the fc_c MMAs of item k-2 have read this buffer
      // the CTA's first tile is gathered by all twelve warps: 16 rows per gather warp, 8 per compute warp
      wd_gather_rows(args, args.job[j], b, tile, smem, smem + WD_OFF_FEAT + (k & 1) * WD_FEAT_BYTES,
                     reinterpret_cast<float4*>(smem + WD_OFF_PTS) + (k & 1) * WD_PTS, k == 0 ? (wtid >> 5) * 16 : (wtid >> 5) * 32, k == 0 ? 16 : 32);
      tc::fence_smem_to_async();
      wd::warp_arrive(&bars[k == 0 ? WB_FEAT_FIRST : WB_FEAT_READY + (k & 1)]);
      if (k == 0) tc::mbar_wait(&bars[WB_FEAT_FIRST], 0u);   // the compute warps use tap-table rows 64..127 for the first tile: do not start the
                                                             // next gather (which owns all 128 rows) before they are done
    }
  } else {
    // ============================ compute warpgroups (two threads per point: column halves) ============================
    const int half = warp >> 2, wtid = tid & 127;
    uint32_t pf = 0u;
    int tslot = 0;
#pragma unroll 1
    for (int k = 0;; ++k) {
      const int item = next_item(k);
      if (item < 0) break;
      int j, b, tile;
      wd::job_of_item(args, item, j, b, tile);
      const DecJob& job = args.job[j];
      if (k == 0) {   // help gathering the CTA's first tile: rows 64 + 8 * warp
        wd_gather_rows(args, job, b, tile, smem, smem + WD_OFF_FEAT, reinterpret_cast<float4*>(smem + WD_OFF_PTS), 64 + 8 * warp, 8);
        tc::fence_smem_to_async();
        wd::warp_arrive(&bars[WB_FEAT_FIRST]);
      }
      if (job.type == 0) wd_compute_item<3>(args, job, b, tile, smem, wtid, half, tmem, bars, pf, (uint32_t)k * 5u, k, tlc, tslot);
      else wd_compute_item<1>(args, job, b, tile, smem, wtid, half, tmem, bars, pf, (uint32_t)k * 5u, k, tlc, tslot);
    }
  }
  pdl_launch();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 13) tc::tmem_dealloc(tmem, WD_TMEM_COLS);
}

}  // namespace giga
