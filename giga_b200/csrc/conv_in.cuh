// Fused  Conv3d(1->32, k3, pad 1) + ReLU + the three tri-plane means.
//
// Replaces (reference, relative to src/vgn/ConvONets):
//   encoder/voxels.py:95-108   voxel coords + relu(conv_in(x)) -> an 8.19 MB/scene feature volume
//   encoder/voxels.py:57-66    3x  normalize_coordinate + coordinate2index + torch_scatter.scatter_mean
// With a 40^3 grid scattered onto 40^2 cells the index map is the identity per axis and
// every cell averages exactly the 40 voxels along the perpendicular axis (SURVEY.md 8a-a4,
// tests/test_oracle_golden.py::test_plane_mean_identity), so the feature volume is never
// materialised: each thread owns one (iy,iz) voxel column, marches along ix with a 3x3x3
// register window, and the three axis sums are formed on the fly:
//   yz[c][iz][iy] = sum_ix f   -> thread-private registers
//   xy[c][iy][ix] = sum_iz f   -> per-step shared-memory row reduction
//   xz[c][iz][ix] = sum_iy f   -> per-step reduction over the CTA's TY rows, then one partial
//                                 per CTA (deterministic order; finished by xz_finish_kernel)
// Weights (27x32) + bias live in the kernel parameter space (constant bank): every FFMA takes
// its weight as an immediate constant-bank operand, so the inner loop is pure FFMA.
// TSDF staging: a 4-D TMA tensor map over x[b][ix][iy][iz] (cuTensorMapEncodeTiled on the host, cp.async.bulk.tensor.4d -> UTMALDG in SASS):
// per march step ONE thread fetches the next ix slab of the CTA's rows plus a one-voxel halo -- box {48 iz, TY + 2 iy, 1 ix, 1 scene} at
// (iz -4, iy0 - 1, ix, b) -- into a 4-deep shared-memory ring, completion on an mbarrier; everything outside the volume is zero-filled by
// the TMA unit, which IS Conv3d's zero padding (no per-thread bounds masks, no slab -1 / 40 special cases).
// Outputs: the xy and yz plane means leave the kernel as TALL pre-split fp16 operands of the first U-Net layer (unet_tall.cuh); the xz
// partials are finished by xz_finish_tall_kernel straight into the same layout (no NCHW staging buffer, no separate layout pass).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder entry point is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include "unet_tall.cuh"

namespace giga {

struct ConvInParams {
  float2 w[27][16];  // [tap = dx*9 + dy*3 + dz][cout pair]   (cross-correlation, voxels.py:36)
  float2 b[16];
};

// Two tilings: TY = 5 iy rows per CTA (8 CTAs per scene, 200 threads) for batches that fill the machine, TY = 1 (40 CTAs
// per scene, 40 threads) for the small batches of the planner / single-scene latency path.  The xz sum over iy is always
// formed in the SAME order -- groups of 5 rows left to right, then the 8 group sums ascending (xz_finish_kernel) -- so a
// scene's result does not depend on the batch it is evaluated in.
constexpr int CI_GROUP = 5;         // canonical iy grouping of the xz sum
constexpr int CI_RED_STRIDE = 44;   // 16-byte aligned rows; float4 reads are bank-conflict free (44 = 12 mod 32)
constexpr int CI_SLAB_W = 48;       // staged iz extent: iz -4 .. 43 -- the box must START on a 16-byte boundary of the tensor row (a start at
                                    // iz = -1 is an illegal instruction: tools/tma_probe.cu), so the halo column iz = -1 sits at index 3
constexpr int CI_NSTAGE = 4;        // slab ring depth
// The 32 output channels are split over CS CTAs (blockIdx.z): one voxel column per thread gives only 1600 threads per
// scene, too few warps per SM to keep the FMA pipe fed; with CS = 2 every SM holds twice the warps, each with half the
// accumulators (and registers).
template <int TY, int CS>
struct ConvInCfg {
  static constexpr int NT = G / TY;                  // iy tiles per scene
  static constexpr int CPT = C / CS;                 // channels per CTA / thread
  static constexpr int THREADS = TY * G;
  static constexpr int RED = CPT * TY * CI_RED_STRIDE; // floats
  static constexpr int SLAB_ROWS = TY + 2;           // iy rows of a staged slab (one halo row either side)
  static constexpr int SLAB_BYTES = SLAB_ROWS * CI_SLAB_W * 4;
  static constexpr int SLAB_STRIDE = (SLAB_BYTES + 127) / 128 * 128;   // TMA destinations are 128-byte aligned
  static constexpr int OFF_SLAB = (3 * RED * 4 + 127) / 128 * 128;   // red (two buffers) + xyacc
  static constexpr int OFF_BAR = OFF_SLAB + CI_NSTAGE * SLAB_STRIDE;
  static constexpr int SMEM_BYTES = OFF_BAR + CI_NSTAGE * 8;     // red x2 + xyacc + slab ring + mbarriers
  static_assert(CI_GROUP % TY == 0 && G % TY == 0, "tiles must nest in the canonical groups");
  static_assert(CPT % 2 == 0 && C % CS == 0, "channel pairs");
};
__host__ inline int conv_in_ty(int B) { return B >= 8 ? 5 : 1; }

// CZ = which channel slice (compile time: the weights must stay constant-bank operands with immediate offsets -- a
// runtime channel offset turns every weight fetch into an indexed load and costs 3.4x)
__device__ __forceinline__ void tma_load_slab(void* dst_smem, const CUtensorMap* tmap, int iz, int iy, int ix, int b, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                   tc::smem_u32(dst_smem)),
               "l"(tmap), "r"(iz), "r"(iy), "r"(ix), "r"(b), "r"(tc::smem_u32(bar))
               : "memory");
}

template <int CI_TY, int CS, int CZ>
__device__ __forceinline__ void conv_in_body(const CUtensorMap* tmap, float* __restrict__ tall, long ps, float* __restrict__ xz_part, int B,
                                             const ConvInParams& P, float* smem) {
  using Cfg = ConvInCfg<CI_TY, CS>;
  constexpr int CI_NT = Cfg::NT, CI_THREADS = Cfg::THREADS, CI_RED = Cfg::RED, CPT = Cfg::CPT;
  constexpr int c0 = CZ * CPT;       // this CTA's first output channel
  float* red0 = smem;               // [2][32*TY][44]: step ix writes buffer ix & 1, so ONE CTA barrier per step suffices (the cross-thread
                                    // sums of step ix overlap the FMA phase of step ix + 1 of the other warps)
  float* xyacc = red0 + 2 * CI_RED; // [32*TY][44]  (col = ix)

  const int tile = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int iyl = tid / G, iz = tid % G;
  const int iy0 = tile * CI_TY;

  // ---- slab ring: slab j (= ix plane j of the CTA's rows + halo) lives in stage j % 4; slabs 0 .. 40 are fetched (slab 40 lies outside
  //      the volume: all zeros), slab j is consumed at the top of step j - 1 and its stage is refilled with slab j + 4 after that step's
  //      closing barrier ----
  uint8_t* slabs = reinterpret_cast<uint8_t*>(smem) + Cfg::OFF_SLAB;
  uint64_t* sbar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem) + Cfg::OFF_BAR);
  if (tid == 0) {
    for (int i = 0; i < CI_NSTAGE; ++i) tc::mbar_init(&sbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();   // the TSDF (and the buffers written below) belong to the stream order before this kernel
  auto fetch = [&](int j) {   // one thread
    uint64_t* bar = &sbar[j % CI_NSTAGE];
    tc::mbar_arrive_expect_tx(bar, (uint32_t)Cfg::SLAB_BYTES);
    tma_load_slab(slabs + (j % CI_NSTAGE) * Cfg::SLAB_STRIDE, tmap, -4, iy0 - 1, j, b, bar);
  };
  if (tid == 0)
    for (int j = 0; j < CI_NSTAGE; ++j) fetch(j);
  // this thread's 3 x 3 (dy, dz) taps inside a slab: rows iyl .. iyl + 2, columns iz + 3 .. iz + 5 (the slab starts at iy0 - 1, iz -4)
  const int soff = iyl * CI_SLAB_W + iz + 3;
  auto read_slab = [&](int j, float* dst) {
    tc::mbar_wait(&sbar[j % CI_NSTAGE], (uint32_t)((j / CI_NSTAGE) & 1));
    const float* sl = reinterpret_cast<const float*>(slabs + (j % CI_NSTAGE) * Cfg::SLAB_STRIDE) + soff;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) dst[dy * 3 + dz] = sl[dy * CI_SLAB_W + dz];
  };

  float2 yz[CPT / 2];   // channel pairs (FFMA2 / FADD2: half the issue slots of the scalar loop, same bits)
#pragma unroll
  for (int c = 0; c < CPT / 2; ++c) yz[c] = make_float2(0.f, 0.f);

  float win[3][9];  // [dx][dy*3+dz]; win[dx] holds slab ix+dx-1
#pragma unroll
  for (int t = 0; t < 9; ++t) win[1][t] = 0.f;               // slab -1 (padding)
  read_slab(0, win[2]);

#pragma unroll 1
  for (int ix = 0; ix < G; ++ix) {
    // Let the next kernel become resident only when this one is nearly done: its CTAs' shared memory comes out of the
    // L1 carve-out this kernel's TSDF reads live in (an early trigger costs tens of us per step).
    if (ix == G - 3) pdl_launch();
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      win[0][t] = win[1][t];
      win[1][t] = win[2][t];
    }
    read_slab(ix + 1, win[2]);
    float* red = red0 + (ix & 1) * CI_RED;
    float2 f[CPT / 2];
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) f[c] = P.b[c0 / 2 + c];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = win[dx][t];
#pragma unroll
        for (int c = 0; c < CPT / 2; ++c) fma2(f[c], P.w[dx * 9 + t][c0 / 2 + c], v);
      }
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c) {
      const float2 r = make_float2(fmaxf(f[c].x, 0.f), fmaxf(f[c].y, 0.f));
      add2(yz[c], r);
      red[((2 * c) * CI_TY + iyl) * CI_RED_STRIDE + iz] = r.x;
      red[((2 * c + 1) * CI_TY + iyl) * CI_RED_STRIDE + iz] = r.y;
    }
    __syncthreads();
    if (tid == 0 && ix + CI_NSTAGE <= G) fetch(ix + CI_NSTAGE);   // stage ix % 4 held slab ix, consumed by every thread at the top of step ix - 1
    // xy[c][iy][ix] = sum over iz (ascending, the reference's scatter order)
    if (tid < CPT * CI_TY) {
      const float* r = red + tid * CI_RED_STRIDE;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < G; k += 4) {
        const float4 q = ld4(r + k);
        s += q.x; s += q.y; s += q.z; s += q.w;
      }
      xyacc[tid * CI_RED_STRIDE + ix] = s;
    }
    // xz partial[ix][c][iz] = sum over this CTA's TY rows
    float* part = xz_part + (((size_t)b * CI_NT + tile) * G + ix) * (C * G) + c0 * G;
    for (int o = tid; o < CPT * (G / 4); o += CI_THREADS) {   // four iz per thread
      const int c = o / (G / 4), z = (o % (G / 4)) * 4;
      float4 s = ld4(red + (c * CI_TY) * CI_RED_STRIDE + z);
#pragma unroll
      for (int r = 1; r < CI_TY; ++r) {
        const float4 q = ld4(red + (c * CI_TY + r) * CI_RED_STRIDE + z);
        s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
      }
      st4(part + c * G + z, s);
    }
    // no second barrier: the next step writes the OTHER red buffer, and a thread reaches the next barrier only after these sums
  }
  __syncthreads();   // the last step's row sums (xyacc) are complete

  // ---- outputs as TALL pre-split operands of the first U-Net layer (unet_tall.cuh): image = plane * B + b, 16-byte elements of 8 channels ----
  // yz[c][iz][iy] = (sum over ix) / 40: this thread holds all CPT channels of pixel (row iz, col iy0 + iyl) of image 2B + b
  {
    const long pos = TALL_MARGIN + tall_pos(G, 2 * B + b, iz, iy0 + iyl);
#pragma unroll
    for (int k8 = 0; k8 < CPT / 8; ++k8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[2 * q] = yz[4 * k8 + q].x / 40.0f;
        v[2 * q + 1] = yz[4 * k8 + q].y / 40.0f;
      }
      uint4 h, l;
      split8(v, h, l);
      const int kc = c0 / 8 + k8;
      stu4(tall + ((size_t)kc * ps + pos) * 4, h);
      stu4(tall + ((size_t)(C / 8 + kc) * ps + pos) * 4, l);
    }
  }
  // xy[c][iy][ix] = (sum over iz) / 40 from the per-step row sums in shared memory: pixel (row iy0 + r, col ix) of image B + b
  for (int o = tid; o < (CPT / 8) * CI_TY * G; o += CI_THREADS) {
    const int ix = o % G, r = (o / G) % CI_TY, k8 = o / (G * CI_TY);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = xyacc[((8 * k8 + q) * CI_TY + r) * CI_RED_STRIDE + ix] / 40.0f;
    uint4 h, l;
    split8(v, h, l);
    const long pos = TALL_MARGIN + tall_pos(G, B + b, iy0 + r, ix);
    const int kc = c0 / 8 + k8;
    stu4(tall + ((size_t)kc * ps + pos) * 4, h);
    stu4(tall + ((size_t)(C / 8 + kc) * ps + pos) * 4, l);
  }
}

// grid (NT, B, CS), block THREADS
template <int CI_TY, int CS>
__global__ void __launch_bounds__(CI_TY * G, CI_TY == 5 ? (CS == 1 ? 2 : 4) : 8)
conv_in_planes_kernel(const __grid_constant__ CUtensorMap tmap,   // TSDF x[b][ix][iy][iz] as a 4-D tensor, box {48, TY + 2, 1, 1}
                      float* __restrict__ tall, long ps,      // TALL pre-split planes [hi|lo][4][ps][8 halfs], images (xz, xy, yz) x B
                      float* __restrict__ xz_part,            // [B][CI_NT][40 ix][32][40 iz]
                      int B, const __grid_constant__ ConvInParams P) {
  extern __shared__ __align__(128) float smem_ci[];
  if constexpr (CS == 1) {
    conv_in_body<CI_TY, 1, 0>(&tmap, tall, ps, xz_part, B, P, smem_ci);
  } else if constexpr (CS == 2) {
    if (blockIdx.z == 0) conv_in_body<CI_TY, 2, 0>(&tmap, tall, ps, xz_part, B, P, smem_ci);
    else conv_in_body<CI_TY, 2, 1>(&tmap, tall, ps, xz_part, B, P, smem_ci);
  } else {
    static_assert(CS <= 4, "channel split");
    switch (blockIdx.z) {
      case 0: conv_in_body<CI_TY, 4, 0>(&tmap, tall, ps, xz_part, B, P, smem_ci); break;
      case 1: conv_in_body<CI_TY, 4, 1>(&tmap, tall, ps, xz_part, B, P, smem_ci); break;
      case 2: conv_in_body<CI_TY, 4, 2>(&tmap, tall, ps, xz_part, B, P, smem_ci); break;
      default: conv_in_body<CI_TY, 4, 3>(&tmap, tall, ps, xz_part, B, P, smem_ci); break;
    }
  }
}

// xz[b][c][iz][ix] = (sum over iy, canonical order) / 40 from the per-CTA partials xz_part[b][t][ix][c][iz], written as TALL pre-split
// elements (image b, pixel (row iz, col ix)); also zeroes the U-Net's tile-dependency counters (every layer of the step follows this kernel).
// thread = (scene, k-chunk of 8 channels, ix, iz) with iz fastest (coalesced partial reads); grid ceil(B * 4 * 1600 / 256), block 256
template <int CI_NT>
__global__ void __launch_bounds__(256)
xz_finish_tall_kernel(const float* __restrict__ xz_part, float* __restrict__ tall, long ps, int B, unsigned* __restrict__ flags, int n_flags) {
  pdl_launch();
  pdl_wait();
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t < n_flags) flags[t] = 0u;
  if (t >= (long)B * 4 * G2) return;
  const int z = (int)(t % G), ix = (int)((t / G) % G), kc = (int)((t / G2) % 4), b = (int)(t / (4 * G2));
  constexpr int PER = CI_NT / (G / CI_GROUP);   // partials per canonical group: 1 (TY = 5) or 5 (TY = 1)
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = 8 * kc + q;
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G / CI_GROUP; ++g) {
      float sg = xz_part[((((size_t)b * CI_NT + g * PER) * G + ix) * C + c) * G + z];
#pragma unroll
      for (int r = 1; r < PER; ++r) sg += xz_part[((((size_t)b * CI_NT + g * PER + r) * G + ix) * C + c) * G + z];
      s += sg;
    }
    v[q] = s / 40.0f;
  }
  uint4 h, l;
  split8(v, h, l);
  const long pos = TALL_MARGIN + tall_pos(G, b, z, ix);
  stu4(tall + ((size_t)kc * ps + pos) * 4, h);
  stu4(tall + ((size_t)(C / 8 + kc) * ps + pos) * 4, l);
}

}  // namespace giga
