// Tensor-core version of the fused Conv3d(1->32,k3,p1) + ReLU + tri-plane means (encoder/voxels.py:57-66,95-108).
//
// Formulation.  Take (ix, iy) as a zero-padded 42x42 image plane, the three dz neighbours as the K axis and
// march along iz:   f[c][ix][iy][iz] = b[c] + sum_{dx,dy} sum_{dz} x[ix+dx-1][iy+dy-1][iz+dz-1] * w[c][dx][dy][dz]
// is a 3x3 "2-D convolution with 3 input channels" per iz.  With the flattened padded position
// p = ixp*42 + iyp the 9 (dx,dy) taps are row shifts of ONE shared-memory operand whose row p holds
// (x[..][iz-1], x[..][iz], x[..][iz+1], 1) -- 12 contiguous bytes of the TSDF plus a constant that carries the
// bias through the centre tap -- exactly the descriptor trick of
// unet_tall.cuh; the K=8 MMA's second k-chunk points at a block of zeros.  Per iz step and 128-position tile:
// 9 taps x { Ahi.[Bhi;Blo]^T (N=64), Alo.Bhi^T (N=32) } = 18 small MMAs instead of 128*864 FMAs.
//
// CTA = 3 full ix-rows (126 positions) of one scene, 256 threads, 2 CTAs / SM:
//   warps 0-3  thread t <-> position t <-> TMEM lane t: drain, bias+ReLU, and the three reductions
//              xy[c][iy][ix] = sum_iz  -> registers (the march axis)
//              xz[c][iz][ix] = sum_iy  -> per-step row sums through shared memory (complete: rows are whole)
//              yz[c][iz][iy] = sum_ix  -> per-step column sums over the 3 rows -> one partial per CTA
//                                         (deterministic order; finished by yz_finish_kernel)
//   warps 4-5  build the A operand (tf32 hi/lo split of a sliding iz window held in registers; one new TSDF
//              value per row per step), double buffered
//   warps 6-7  issue the MMAs (one product kind each); accumulators double buffered across steps so the
//              epilogue of step iz overlaps the MMAs of step iz+1.
// Outputs are written straight in the TALL pre-split layout the U-Net consumes (no NCHW round trip).
#pragma once
#include "common.cuh"
#include "tc.cuh"
#include "unet_tall.cuh"

namespace giga {

constexpr int CT_ROWS = 3;                       // ix rows per CTA
constexpr int CT_NG = (G + CT_ROWS - 1) / CT_ROWS;   // 14 row groups per scene
constexpr int CT_WP = G + 2;                     // 42
constexpr int CT_HALO = CT_WP + 1;               // 43
constexpr int CT_WIN = 128 + 2 * CT_HALO;        // 214 staged rows
constexpr int CT_ACH = 216 * 16;                 // bytes of one staged k-chunk (hi or lo), 3456
constexpr int CT_STAGE = 4 * CT_ACH;             // [hi][zeros][lo][zeros]
constexpr int CT_B_BYTES = 9 * 2048;             // [tap][kc 2][hi|lo][n 32][4], kc 1 = zeros
constexpr int CT_OFF_B = 2 * CT_STAGE;                       // 27648
constexpr int CT_OFF_RED = CT_OFF_B + CT_B_BYTES;            // 46080: red[32*3][44]
constexpr int CT_RED_STRIDE = 44;                 // 16-byte aligned rows: the reductions read float4
constexpr int CT_RED_FLOATS = C * CT_ROWS * CT_RED_STRIDE;
constexpr int CT_OFF_XZ = CT_OFF_RED + CT_RED_FLOATS * 4;    // xzacc[32*3][40]
constexpr int CT_XZ_FLOATS = C * CT_ROWS * G;
constexpr int CT_OFF_BIAS = CT_OFF_XZ + CT_XZ_FLOATS * 4;
constexpr int CT_OFF_BAR = CT_OFF_BIAS + 128;
constexpr int CT_SMEM_BYTES = CT_OFF_BAR + 8 * 8 + 16;       // ~77.5 KB -> 2 CTAs / SM
constexpr int CT_TMEM_COLS = 256;                            // 2 sets x (D1 64 + D2 32) = 192
constexpr long CT_WEIGHT_FLOATS = CT_B_BYTES / 4;

// grid (CT_NG, B), block 256, dynamic smem CT_SMEM_BYTES
__global__ void __launch_bounds__(256, 2)
conv_in_tc_kernel(const float* __restrict__ x,        // [B][40][40][40]
                  const float* __restrict__ wt,       // packed B operand (CT_WEIGHT_FLOATS)
                  const float* __restrict__ bias,     // [32]
                  float* __restrict__ pre_tall, long ps,   // TALL [2][8][ps][4]: images plane*B + b
                  float* __restrict__ yz_part,        // [B][CT_NG][40 iz][32][40 iy]
                  int B, unsigned long long* __restrict__ tl) {   // tl: optional stall accounting (debug)
  extern __shared__ __align__(128) uint8_t smem_ci[];
  uint8_t* smem = smem_ci;
  float* red = reinterpret_cast<float*>(smem + CT_OFF_RED);
  float* xzacc = reinterpret_cast<float*>(smem + CT_OFF_XZ);
  float* sbias = reinterpret_cast<float*>(smem + CT_OFF_BIAS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CT_OFF_BAR);   // [2] A operand of a step staged (64 loader arrivals)
  uint64_t* empty = full + 2;                                          // [2] both MMA warps are done with the stage
  uint64_t* acc_full = full + 4;                                       // [2] both products of a step complete
  uint64_t* acc_empty = full + 6;                                      // [2] 128 drainer arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + CT_OFF_BAR + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x, b = blockIdx.y;
  const int f0 = (1 + CT_ROWS * g) * CT_WP;          // first output position (padded row 1+3g, padded col 0)
  const int nrows = min(CT_ROWS, G - CT_ROWS * g);   // valid ix rows in this group (1 for the last)

  // one-time setup: zero both stages (static zero k-chunks + rows that never hold data), weights, bias, barriers
  for (int e = tid; e < 2 * CT_STAGE / 16; e += 256) reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < CT_B_BYTES / 16; e += 256) reinterpret_cast<float4*>(smem + CT_OFF_B)[e] = __ldg(reinterpret_cast<const float4*>(wt) + e);
  if (tid < C) sbias[tid] = __ldg(bias + tid);
  if (warp == 7) tc::tmem_alloc(tmem_slot, CT_TMEM_COLS);
  if (tid == 0) {
    tc::mbar_init(&full[0], 64); tc::mbar_init(&full[1], 64);
    tc::mbar_init(&empty[0], 2); tc::mbar_init(&empty[1], 2);
    tc::mbar_init(&acc_full[0], 2); tc::mbar_init(&acc_full[1], 2);
    tc::mbar_init(&acc_empty[0], 128); tc::mbar_init(&acc_empty[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  unsigned long long* tlc = tl ? tl + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 : nullptr;
  long long w_a = 0, w_b = 0, w_c = 0;
  const long long t_begin = clock64();
  auto twait = [&](uint64_t* bar, uint32_t parity, long long& accum) {
    const long long t = clock64();
    tc::mbar_wait(bar, parity);
    accum += clock64() - t;
  };
  if (tlc && tid == 0) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tlc[28] = gt; tlc[30] = smid;
  }

  if (warp >= 4 && warp < 6) {
    // ====================== loader warps: A operand rows from a sliding iz window ======================
    const int lt = tid - 128;                                // 0..63
    const float* rowp[4];
    float vm[4], v0[4], vp[4], vn[4];                        // x[iz-1], x[iz], x[iz+1] of each owned row; vn = x[iz+2] prefetched
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = lt + 64 * q;
      rowp[q] = nullptr;
      if (r < CT_WIN) {
        const int fw = f0 - CT_HALO + r;
        if (fw >= 0) {
          const int ixp = fw / CT_WP, iyp = fw - ixp * CT_WP;
          if (ixp >= 1 && ixp <= G && iyp >= 1 && iyp <= G) rowp[q] = x + (((size_t)b * G + ixp - 1) * G + iyp - 1) * G;
        }
      }
      vm[q] = 0.f;
      v0[q] = rowp[q] ? __ldg(rowp[q]) : 0.f;
      vp[q] = rowp[q] ? __ldg(rowp[q] + 1) : 0.f;
      vn[q] = rowp[q] ? __ldg(rowp[q] + 2) : 0.f;
    }
#pragma unroll 1
    for (int iz = 0; iz < G; ++iz) {
      const int s = iz & 1;
      if (iz >= 2) twait(&empty[s], (uint32_t)(((iz - 2) >> 1) & 1), w_a);
      const long long t_w = clock64();
      uint8_t* st = smem + s * CT_STAGE;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (rowp[q]) {
          const int r = lt + 64 * q;
          float4 h, l;
          split4(make_float4(vm[q], v0[q], vp[q], 1.0f), h, l);   // 4th k = 1: the centre tap's k=3 weight is the bias
          *reinterpret_cast<float4*>(st + r * 16) = h;                    // hi k-chunk 0
          *reinterpret_cast<float4*>(st + 2 * CT_ACH + r * 16) = l;       // lo k-chunk 0
        }
      tc::fence_smem_to_async();
      tc::mbar_arrive(&full[s]);
      w_b += clock64() - t_w;
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // slide the window; the value needed two steps from now is fetched here
        vm[q] = v0[q];
        v0[q] = vp[q];
        vp[q] = vn[q];
        vn[q] = (rowp[q] && iz + 3 < G) ? __ldg(rowp[q] + iz + 3) : 0.f;
      }
    }
    if (tlc && lt == 0) { tlc[0] = (unsigned long long)w_a; tlc[1] = (unsigned long long)w_b; tlc[2] = (unsigned long long)(clock64() - t_begin); }
  } else if (warp >= 6) {
    // ====================== MMA-issue warps: 6 = Ahi.[Bhi;Blo]^T (N=64), 7 = Alo.Bhi^T (N=32) ======================
    const int kind = warp - 6;
    const uint32_t idesc = kind == 0 ? tc::make_idesc_tf32(128, 64) : tc::make_idesc_tf32(128, 32);
    const uint32_t b0 = tc::smem_u32(smem + CT_OFF_B);
#pragma unroll 1
    for (int iz = 0; iz < G; ++iz) {
      const int s = iz & 1, set = iz & 1;
      twait(&full[s], (uint32_t)((iz >> 1) & 1), w_a);
      if (iz >= 2) twait(&acc_empty[set], (uint32_t)(((iz - 2) >> 1) & 1), w_b);
      const long long t_i = clock64();
      tc::fence_after_sync();
      const uint32_t a0 = tc::smem_u32(smem + s * CT_STAGE) + (kind == 0 ? 0 : 2 * CT_ACH);
      const uint32_t d = tmem + set * 96 + (kind == 0 ? 0 : 64);
      if (tc::elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dx = tap / 3, dy = tap - dx * 3;
          tc::mma_tf32(d, tc::make_desc(a0 + (uint32_t)(dx * CT_WP + dy) * 16, CT_ACH, 128), tc::make_desc(b0 + tap * 2048, 1024, 128), idesc,
                       tap == 0 ? 0u : 1u);
        }
        tc::mma_commit(&empty[s]);
        tc::mma_commit(&acc_full[set]);
      }
      __syncwarp();
      w_c += clock64() - t_i;
    }
    if (tlc && (tid & 31) == 0) {
      tlc[4 + 4 * kind] = (unsigned long long)w_a; tlc[5 + 4 * kind] = (unsigned long long)w_b;
      tlc[6 + 4 * kind] = (unsigned long long)w_c; tlc[7 + 4 * kind] = (unsigned long long)(clock64() - t_begin);
    }
  } else {
    // ====================== drain + reductions (thread t <-> position f0 + t) ======================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = tid / CT_WP, iyp = tid - row * CT_WP, iy = iyp - 1;
    const bool valid = row < nrows && iyp >= 1 && iyp <= G;
    float xy[C];
    long long p_b1 = 0, p_sts = 0, p_b2 = 0, p_xz = 0, p_yz = 0;
#pragma unroll
    for (int j = 0; j < C; ++j) xy[j] = 0.f;
#pragma unroll 1
    for (int iz = 0; iz < G; ++iz) {
      const int set = iz & 1;
      twait(&acc_full[set], (uint32_t)((iz >> 1) & 1), w_a);
      const long long t_d = clock64();
      tc::fence_after_sync();
      float f[32], w[32];
      tc::tmem_ld32(tmem_row + set * 96, f);            // hi*hi
      tc::tmem_ld32(tmem_row + set * 96 + 32, w);       // hi*lo
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += w[j];
      tc::tmem_ld32(tmem_row + set * 96 + 64, w);       // lo*hi
      tc::fence_before_sync();
      tc::mbar_arrive(&acc_empty[set]);
      w_b += clock64() - t_d;
      const long long t_r = clock64();
      tc::named_bar_sync(1, 128);                        // the previous step's reductions have finished reading red
      const long long t_r1 = clock64();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float r = fmaxf(f[j] + w[j], 0.f);         // bias already added by the centre-tap MMA
        if (valid) {
          xy[j] += r;
          red[(j * CT_ROWS + row) * CT_RED_STRIDE + iy] = r;
        }
      }
      const long long t_r2 = clock64();
      tc::named_bar_sync(1, 128);
      const long long t_r3 = clock64();
      // xz[c][iz][ix] = sum over iy for the CTA's rows (4 interleaved partial sums, 128-bit reads)
      if (tid < C * CT_ROWS) {
        const int rr = tid % CT_ROWS;
        if (rr < nrows) {
          const float4* rp = reinterpret_cast<const float4*>(red + tid * CT_RED_STRIDE);
          float4 s4 = rp[0];
#pragma unroll
          for (int k = 1; k < G / 4; ++k) {
            const float4 v = rp[k];
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
          }
          xzacc[tid * G + iz] = (s4.x + s4.y) + (s4.z + s4.w);
        }
      }
      const long long t_r4 = clock64();
      // yz partial[iz][c][iy] = sum over this CTA's ix rows, four iy per thread
      float* part = yz_part + ((((size_t)b * CT_NG + g) * G + iz) * C) * G;
      for (int o = tid; o < C * G / 4; o += 128) {
        const int c = o / (G / 4), y4 = o - c * (G / 4);
        float4 s4 = *reinterpret_cast<const float4*>(red + (c * CT_ROWS) * CT_RED_STRIDE + 4 * y4);
        for (int rr = 1; rr < nrows; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(red + (c * CT_ROWS + rr) * CT_RED_STRIDE + 4 * y4);
          s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
        }
        st4(part + 4 * o, s4);
      }
      w_c += clock64() - t_r;
      p_b1 += t_r1 - t_r; p_sts += t_r2 - t_r1; p_b2 += t_r3 - t_r2; p_xz += t_r4 - t_r3; p_yz += clock64() - t_r4;
    }
    if (tlc && tid == 0) {
      tlc[20] = (unsigned long long)p_b1; tlc[21] = (unsigned long long)p_sts; tlc[22] = (unsigned long long)p_b2;
      tlc[23] = (unsigned long long)p_xz; tlc[24] = (unsigned long long)p_yz;
      tlc[16] = (unsigned long long)w_a; tlc[17] = (unsigned long long)w_b; tlc[18] = (unsigned long long)w_c;
      tlc[19] = (unsigned long long)(clock64() - t_begin);
    }
    // ---- outputs, TALL pre-split layout: image = plane*B + b, 8 k-chunks of 4 channels ----
    if (valid) {   // xy plane (1): row = iy, col = ix
      const long pos = TALL_MARGIN + tall_pos(G, 1 * B + b, iy, CT_ROWS * g + row);
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float4 h, l;
        split4(make_float4(xy[4 * kc] / 40.0f, xy[4 * kc + 1] / 40.0f, xy[4 * kc + 2] / 40.0f, xy[4 * kc + 3] / 40.0f), h, l);
        st4(pre_tall + ((size_t)kc * ps + pos) * 4, h);
        st4(pre_tall + ((size_t)(8 + kc) * ps + pos) * 4, l);
      }
    }
    tc::named_bar_sync(1, 128);   // xzacc complete
    for (int o = tid; o < nrows * G * 8; o += 128) {   // xz plane (0): row = iz, col = ix
      const int kc = o % 8, z = (o / 8) % G, rr = o / (8 * G);
      const long pos = TALL_MARGIN + tall_pos(G, 0 * B + b, z, CT_ROWS * g + rr);
      float4 h, l;
      split4(make_float4(xzacc[((4 * kc + 0) * CT_ROWS + rr) * G + z] / 40.0f, xzacc[((4 * kc + 1) * CT_ROWS + rr) * G + z] / 40.0f,
                         xzacc[((4 * kc + 2) * CT_ROWS + rr) * G + z] / 40.0f, xzacc[((4 * kc + 3) * CT_ROWS + rr) * G + z] / 40.0f), h, l);
      st4(pre_tall + ((size_t)kc * ps + pos) * 4, h);
      st4(pre_tall + ((size_t)(8 + kc) * ps + pos) * 4, l);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 7) tc::tmem_dealloc(tmem, CT_TMEM_COLS);
  if (tlc && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tlc[29] = gt;
  }
}

// yz[b][c][iz][iy] = (sum_g yz_part[b][g][iz][c][iy]) / 40  -> TALL plane 2.   grid (40 iz, B), block 320 = 40 iy x 8 kc
__global__ void __launch_bounds__(320)
yz_finish_tall_kernel(const float* __restrict__ yz_part, float* __restrict__ pre_tall, long ps, int B) {
  const int iz = blockIdx.x, b = blockIdx.y;
  const int iy = threadIdx.x % G, kc = threadIdx.x / G;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float s = 0.f;
#pragma unroll
    for (int gg = 0; gg < CT_NG; ++gg) s += yz_part[((((size_t)b * CT_NG + gg) * G + iz) * C + 4 * kc + j) * G + iy];
    v[j] = s / 40.0f;
  }
  float4 h, l;
  split4(make_float4(v[0], v[1], v[2], v[3]), h, l);
  const long pos = TALL_MARGIN + tall_pos(G, 2 * B + b, iz, iy);
  st4(pre_tall + ((size_t)kc * ps + pos) * 4, h);
  st4(pre_tall + ((size_t)(8 + kc) * ps + pos) * 4, l);
}

}  // namespace giga
