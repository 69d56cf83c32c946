// Tensor-core version of the fused Conv3d(1->32,k3,p1) + ReLU + tri-plane means (encoder/voxels.py:36,57-66,95-108).
//
// Formulation.  The conv has ONE input channel, so the natural K axis of an implicit GEMM is the 27 taps.  The dz taps
// are folded into a 16-byte "voxel element" (one row of one K-major k-chunk, tc.cuh):
//     E[v] = ( h(z-1), h(z), h(z+1), l(z-1), l(z), l(z+1), 1, 0 )   fp16,  h = fp16_rn(x), l = fp16_rn(x - h), 0 outside
// written once per call by tsdf_elements_kernel into a zero-padded volume E[b][42 ixp][42 iyp][40 iz].  The (dx,dy) taps
// are then pure row shifts of that volume (dy: 40 rows, dx: one plane), i.e. start-address offsets of the shared-memory
// matrix descriptor -- the flattened-shift trick of unet_tall.cuh -- and one MMA (K = 16 = two k-chunks) covers two
// (dx,dy) taps x {3 dz} x {hi, lo}.  The B operand is the row concatenation [Wh ; Wl] (N = 64): columns 0-31 receive
// (h + l) * wh, columns 32-63 (h + l) * wl, so f = D[c] + D[32 + c] carries all four split products; the bias rides on
// the constant 1 of the centre tap.  Weights (and bias) are pre-scaled by 2^s on the host so their lo halves stay normal
// fp16 numbers (decoder_tc.cuh); the epilogue multiplies by 2^-s.
//
// CTA = (tile of 3 iy lines = 120 voxel rows + 8 idle rows, scene), marching along ix; 160 threads:
//   warp 4    one cp.async.bulk per ix plane (the tile's rows +- one line, 3.3 KB) into a 7-slot ring, three steps
//             ahead, and the MMA issue: per step 3 planes x 2 MMAs (taps (dy0,dy1) and (-,dy2)) into one of two TMEM
//             accumulator sets
//   warps 0-3 thread t <-> voxel row t <-> TMEM lane t: drain, *2^-s, ReLU, and the three axis sums
//               yz[c][iz][iy] = sum_ix  -> registers (the march axis)
//               xy[c][iy][ix] = sum_iz  -> per-step row sums through shared memory
//               xz[c][iz][ix] = sum_iy  -> per-step sum of the tile's 3 lines, one partial per tile (deterministic:
//                                          (l0 + l1) + l2 per tile, tiles ascending in xz_finish_tc_kernel)
// 14 tiles x B CTAs, 4 CTAs / SM (128 TMEM columns, ~50 KB shared memory, <= 102 registers): one wave at B = 32.
#pragma once
#include "common.cuh"
#include "tc.cuh"
#include "unet_tall.cuh"

namespace giga {

constexpr int CT_LINES = 3;                              // iy lines per tile
constexpr int CT_NT = (G + CT_LINES - 1) / CT_LINES;     // 14 tiles per scene
constexpr int CT_GP = G + 2;                             // 42: padded ix / iy extent of the element volume
constexpr int CT_PLANE_ROWS = CT_GP * G;                 // 1680 rows (elements) per ix plane
constexpr long CT_SCENE_ROWS = (long)CT_GP * CT_PLANE_ROWS;   // 70560
constexpr int CT_TAIL_ROWS = 256;                        // readable slack after the last scene (the last tile's junk rows)
constexpr int CT_WIN_ROWS = 128 + 2 * G;                 // 208 rows staged per plane: the tile's rows +- one line
constexpr int CT_WIN_BYTES = CT_WIN_ROWS * 16;           // 3328
constexpr int CT_SLOT_BYTES = CT_WIN_BYTES;              // 26 x 128 B
constexpr int CT_NSLOT = 7;                              // ring of plane windows
constexpr int CT_AHEAD = 5;                              // planes in flight ahead of the oldest one a step reads (3 steps of lead)
constexpr int CT_B_MMA_BYTES = 2 * 64 * 16;              // one MMA's B operand: [kc 2][n 64][8 halfs]
constexpr int CT_B_BYTES = 6 * CT_B_MMA_BYTES;           // 12288: [dx 3][mma 2]
constexpr int CT_W_WORDS = CT_B_BYTES / 4 + 4;           // packed weights (+ 2^-s)
constexpr int CT_OFF_B = CT_NSLOT * CT_SLOT_BYTES;       // 23296
constexpr int CT_OFF_RED = CT_OFF_B + CT_B_BYTES;        // 35584: red[32][132] floats
constexpr int CT_RS = 132;                               // red row stride: 16-byte aligned rows, 4-bank skew per channel
constexpr int CT_OFF_BAR = CT_OFF_RED + C * CT_RS * 4;   // 52480
constexpr int CT_SMEM_BYTES = CT_OFF_BAR + 20 * 8;       // 52640  (4 CTAs / SM)
constexpr int CT_TMEM_COLS = 128;                        // two accumulator sets of 64 columns
constexpr int CT_THREADS = 160;

__host__ __device__ constexpr long ct_element_words(int B) { return ((long)B * CT_SCENE_ROWS + CT_TAIL_ROWS) * 4; }

// TSDF -> voxel elements (interior only; the padding planes / lines stay zero from the one-time memset).
// grid ceil(B*64000/256), block 256
__global__ void __launch_bounds__(256) tsdf_elements_kernel(const float* __restrict__ x, float* __restrict__ E, int B) {
  pdl_launch();
  pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= B * G3) return;
  const int b = t / G3, v = t - b * G3;
  const int ix = v / G2, iy = (v / G) % G, iz = v % G;
  const float* p = x + t;
  float xv[3];
  xv[0] = iz > 0 ? __ldg(p - 1) : 0.f;
  xv[1] = __ldg(p);
  xv[2] = iz < G - 1 ? __ldg(p + 1) : 0.f;
  __half h[3], l[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float c = fminf(fmaxf(xv[k], -H_MAX), H_MAX);
    h[k] = __float2half_rn(c);
    l[k] = __float2half_rn(c - __half2float(h[k]));
  }
  const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
  const __half2 e0 = __halves2half2(h[0], h[1]), e1 = __halves2half2(h[2], l[0]), e2 = __halves2half2(l[1], l[2]), e3 = __halves2half2(one, zero);
  const long row = (long)b * CT_SCENE_ROWS + ((long)(ix + 1) * CT_GP + (iy + 1)) * G + iz;
  stu4(E + row * 4, make_uint4(*reinterpret_cast<const uint32_t*>(&e0), *reinterpret_cast<const uint32_t*>(&e1),
                               *reinterpret_cast<const uint32_t*>(&e2), *reinterpret_cast<const uint32_t*>(&e3)));
}

// grid (CT_NT, B), block 160, dynamic smem CT_SMEM_BYTES
__global__ void __launch_bounds__(CT_THREADS, 4)
conv_in_tc_kernel(const float* __restrict__ E,        // voxel elements (see above)
                  const float* __restrict__ wt,       // packed B operands [dx 3][mma 2][kc 2][n 64][8 halfs] + 2^-s
                  float* __restrict__ pre,            // [3][B][32][40][40]  (xz, xy, yz), NCHW
                  float* __restrict__ xz_part,        // [B][CT_NT][40 ix][32][40 iz]
                  int B, unsigned long long* __restrict__ tl) {   // tl: optional stall accounting (debug), 32 u64 per CTA
  extern __shared__ __align__(128) uint8_t smem_ct[];
  uint8_t* smem = smem_ct;
  float* red = reinterpret_cast<float*>(smem + CT_OFF_RED);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CT_OFF_BAR);   // [7] plane window landed
  uint64_t* empty = full + CT_NSLOT;                                  // [7] MMAs that read the slot completed
  uint64_t* acc_full = empty + CT_NSLOT;                              // [2]
  uint64_t* acc_empty = acc_full + 2;                                 // [2] 128 drainer arrivals
  uint64_t* wbar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int row0 = G + tile * (CT_LINES * G);          // first row of the tile inside a plane (iyp = 1 + 3*tile, iz = 0)

  pdl_launch();
  if (warp == 4) tc::tmem_alloc(tmem_slot, CT_TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < CT_NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&acc_full[0], 1); tc::mbar_init(&acc_full[1], 1);
    tc::mbar_init(&acc_empty[0], 128); tc::mbar_init(&acc_empty[1], 128);
    tc::mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  unsigned long long* tlc = tl ? tl + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 : nullptr;
  long long w_a = 0, w_b = 0, w_c = 0, w_d = 0, w_e = 0;
  const long long t_begin = clock64();
  auto twait = [&](uint64_t* bar, uint32_t parity, long long& accum) {
    const long long t = clock64();
    tc::mbar_wait(bar, parity);
    accum += clock64() - t;
  };

  if (warp == 4) {
    // ============================ loader + MMA issue ============================
    if (tc::elect_one()) {   // weights: constants, no dependency on the previous kernel
      tc::mbar_arrive_expect_tx(wbar, (uint32_t)CT_B_BYTES);
      tc::bulk_g2s(smem + CT_OFF_B, wt, CT_B_BYTES, wbar);
    }
    __syncwarp();
    pdl_wait();              // the element volume of this call is complete
    const float* Eb = E + ((long)b * CT_SCENE_ROWS + row0 - G) * 4;
    auto load_plane = [&](int p) {               // padded plane index ixp -> slot p % 5
      const int s = p % CT_NSLOT;
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)CT_WIN_BYTES);
        tc::bulk_g2s(smem + s * CT_SLOT_BYTES, Eb + (long)p * CT_PLANE_ROWS * 4, CT_WIN_BYTES, &full[s]);
      }
      __syncwarp();
    };
#pragma unroll 1
    for (int p = 0; p < CT_AHEAD; ++p) load_plane(p);
    constexpr uint32_t IDESC = tc::make_idesc_f16(128, 64);
    constexpr uint32_t LBO = G * 16;             // second k-chunk = the same window one line (40 rows) further
    const uint32_t b0 = tc::smem_u32(smem + CT_OFF_B), a0 = tc::smem_u32(smem);
    tc::mbar_wait(wbar, 0u);
    tc::mbar_wait(&full[0], 0u);
    tc::mbar_wait(&full[1], 0u);
#pragma unroll 1
    for (int ix = 0; ix < G; ++ix) {
      const int set = ix & 1;
      // prefetch plane ix+5 into the slot plane ix-2 used (its last reader was step ix-2): the copy has three steps to land
      if (ix + CT_AHEAD < CT_GP) {
        if (ix >= 2) twait(&empty[(ix - 2) % CT_NSLOT], (uint32_t)(((ix - 2) / CT_NSLOT) & 1), w_a);
        load_plane(ix + CT_AHEAD);
      }
      const int pn = ix + 2;                     // newest plane of this step
      twait(&full[pn % CT_NSLOT], (uint32_t)((pn / CT_NSLOT) & 1), w_b);
      if (ix >= 2) twait(&acc_empty[set], (uint32_t)(((ix - 2) >> 1) & 1), w_c);
      const long long t_i = clock64();
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint32_t a = a0 + (uint32_t)((ix + dx) % CT_NSLOT) * CT_SLOT_BYTES;
          // MMA 0: k-chunks = lines dy 0 and dy 1;  MMA 1: k-chunks = dy 1 (zero weights) and dy 2
          tc::mma_f16(tmem + set * 64, tc::make_desc(a, LBO, 128), tc::make_desc(b0 + (dx * 2 + 0) * CT_B_MMA_BYTES, 64 * 16, 128), IDESC,
                      dx > 0 ? 1u : 0u);
          tc::mma_f16(tmem + set * 64, tc::make_desc(a + LBO, LBO, 128), tc::make_desc(b0 + (dx * 2 + 1) * CT_B_MMA_BYTES, 64 * 16, 128), IDESC, 1u);
        }
        tc::mma_commit(&empty[ix % CT_NSLOT]);   // plane ix is not read again
        tc::mma_commit(&acc_full[set]);
      }
      __syncwarp();
      w_d += clock64() - t_i;
    }
    if (tlc && (tid & 31) == 0) {
      tlc[0] = (unsigned long long)w_a; tlc[1] = (unsigned long long)w_b; tlc[2] = (unsigned long long)w_c; tlc[3] = (unsigned long long)w_d;
      tlc[4] = (unsigned long long)(clock64() - t_begin);
    }
  } else {
    // ============================ drain + reductions ============================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const float inv40 = __ldg(wt + CT_B_BYTES / 4) / 40.0f;       // 2^-s (exact) folded into the final means
    const int line = tid / G, iz = tid - line * G;                 // line 3 = the 8 idle rows
    const int iy = tile * CT_LINES + line;
    const bool valid = line < CT_LINES && iy < G;
    const int nlines = min(CT_LINES, G - tile * CT_LINES);
    float yz[C];
#pragma unroll
    for (int j = 0; j < C; ++j) yz[j] = 0.f;
    if (!valid) {   // idle rows / lines past the volume never contribute: their red entries stay zero
#pragma unroll
      for (int j = 0; j < C; ++j) red[j * CT_RS + tid] = 0.f;
    }
    float* pre_xy = pre + ((size_t)(1 * B + b) * C) * G2;
#pragma unroll 1
    for (int ix = 0; ix < G; ++ix) {
      const int set = ix & 1;
      twait(&acc_full[set], (uint32_t)((ix >> 1) & 1), w_a);
      tc::fence_after_sync();
      long long t0 = clock64();
      tc::named_bar_sync(1, 128);                        // the previous step's readers are done with red
      w_b += clock64() - t0; t0 = clock64();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                   // 16 channels at a time (register budget: 4 CTAs / SM)
        float f[16], w[16];
        tc::tmem_ld16(tmem_row + set * 64 + hf * 16, f);       // (h + l) * wh      (warp-collective: never under `valid`)
        tc::tmem_ld16(tmem_row + set * 64 + 32 + hf * 16, w);  // (h + l) * wl
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {                 // values stay scaled by 2^s (ReLU and the sums commute with it)
            const float r = fmaxf(f[j] + w[j], 0.f);
            yz[hf * 16 + j] += r;
            red[(hf * 16 + j) * CT_RS + tid] = r;
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&acc_empty[set]);
      w_c += clock64() - t0; t0 = clock64();
      tc::named_bar_sync(1, 128);
      w_d += clock64() - t0; t0 = clock64();
      // xy[c][iy][ix] = sum over iz of one line: 96 (c, line) row sums, ascending iz
      if (tid < C * CT_LINES) {
        const int c = tid / CT_LINES, ln = tid - c * CT_LINES;
        if (ln < nlines) {
          const float* r = red + c * CT_RS + ln * G;
          float4 s = ld4(r);                             // four interleaved partial sums (short dependency chains)
#pragma unroll
          for (int k = 4; k < G; k += 4) {
            const float4 q = ld4(r + k);
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
          }
          pre_xy[(c * G + tile * CT_LINES + ln) * G + ix] = ((s.x + s.y) + (s.z + s.w)) * inv40;
        }
      }
      // xz partial[ix][c][iz] = (l0 + l1) + l2 over the tile's lines (idle lines hold zeros), four iz per thread
      float* part = xz_part + (((size_t)b * CT_NT + tile) * G + ix) * (C * G);
      for (int o = tid; o < C * (G / 4); o += 128) {
        const int c = o / (G / 4), z = (o - c * (G / 4)) * 4;
        float4 s = ld4(red + c * CT_RS + z);
#pragma unroll
        for (int ln = 1; ln < CT_LINES; ++ln) {
          const float4 q = ld4(red + c * CT_RS + ln * G + z);
          s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
        }
        st4(part + c * G + z, s);
      }
      w_e += clock64() - t0;
    }
    if (tlc && tid == 0) {
      tlc[8] = (unsigned long long)w_a; tlc[9] = (unsigned long long)w_b; tlc[10] = (unsigned long long)w_c; tlc[11] = (unsigned long long)w_d;
      tlc[12] = (unsigned long long)w_e; tlc[13] = (unsigned long long)(clock64() - t_begin);
    }
    // yz[c][iz][iy]: transpose through smem so that the tile's (up to 3) consecutive iy are written together
    tc::named_bar_sync(1, 128);
    float* stage = red;   // [32][40][3]
    if (line < CT_LINES) {
#pragma unroll
      for (int c = 0; c < C; ++c) stage[(c * G + iz) * CT_LINES + line] = yz[c] * inv40;
    }
    tc::named_bar_sync(1, 128);
    float* pre_yz = pre + ((size_t)(2 * B + b) * C) * G2;
    for (int o = tid; o < C * G * CT_LINES; o += 128) {
      const int cz = o / CT_LINES, r = o - cz * CT_LINES;   // cz = c*40 + iz
      if (r < nlines) pre_yz[cz * G + tile * CT_LINES + r] = stage[o];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, CT_TMEM_COLS);
}

// xz[b][c][iz][ix] = (sum over the CT_NT tile partials, ascending) * 2^-s / 40      grid (32, B), block 256
__global__ void __launch_bounds__(256)
xz_finish_tc_kernel(const float* __restrict__ xz_part, const float* __restrict__ wt, float* __restrict__ pre, int B) {
  __shared__ float tile[G * 41];
  pdl_launch();
  pdl_wait();
  const int c = blockIdx.x, b = blockIdx.y;
  const float inv40 = __ldg(wt + CT_B_BYTES / 4) / 40.0f;
  for (int e = threadIdx.x; e < G2; e += 256) {
    const int ix = e / G, z = e % G;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < CT_NT; ++t) s += xz_part[((((size_t)b * CT_NT + t) * G + ix) * C + c) * G + z];
    tile[z * 41 + ix] = s * inv40;
  }
  __syncthreads();
  float* o = pre + ((size_t)(0 * B + b) * C + c) * G2;
  for (int e = threadIdx.x; e < G2; e += 256) o[e] = tile[(e / G) * 41 + (e % G)];
}

}  // namespace giga
