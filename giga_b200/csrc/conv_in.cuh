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
#pragma once
#include "common.cuh"

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
// The 32 output channels are split over CS CTAs (blockIdx.z): one voxel column per thread gives only 1600 threads per
// scene, too few warps per SM to keep the FMA pipe fed; with CS = 2 every SM holds twice the warps, each with half the
// accumulators (and registers).
template <int TY, int CS>
struct ConvInCfg {
  static constexpr int NT = G / TY;                  // iy tiles per scene
  static constexpr int CPT = C / CS;                 // channels per CTA / thread
  static constexpr int THREADS = TY * G;
  static constexpr int RED = CPT * TY * CI_RED_STRIDE; // floats
  static constexpr int SMEM_BYTES = 2 * RED * 4;     // red + xyacc
  static_assert(CI_GROUP % TY == 0 && G % TY == 0, "tiles must nest in the canonical groups");
  static_assert(CPT % 2 == 0 && C % CS == 0, "channel pairs");
};
__host__ inline int conv_in_ty(int B) { return B >= 8 ? 5 : 1; }

// CZ = which channel slice (compile time: the weights must stay constant-bank operands with immediate offsets -- a
// runtime channel offset turns every weight fetch into an indexed load and costs 3.4x)
template <int CI_TY, int CS, int CZ>
__device__ __forceinline__ void conv_in_body(const float* __restrict__ x, float* __restrict__ pre, float* __restrict__ xz_part, int B,
                                             const ConvInParams& P, float* smem) {
  pdl_wait();
  using Cfg = ConvInCfg<CI_TY, CS>;
  constexpr int CI_NT = Cfg::NT, CI_THREADS = Cfg::THREADS, CI_RED = Cfg::RED, CPT = Cfg::CPT;
  constexpr int c0 = CZ * CPT;       // this CTA's first output channel
  float* red = smem;                // [32*TY][44]
  float* xyacc = red + CI_RED;      // [32*TY][44]  (col = ix)

  const int tile = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int iyl = tid / G, iz = tid % G;
  const int iy0 = tile * CI_TY;

  // The thread's 3x3 (dy,dz) neighbourhood is read straight from global/L1 (the TSDF is 256 KB per scene and
  // L2 resident); slab ix+2 is prefetched into registers while slab ix+1's 864 FMAs run.  Out-of-volume taps
  // are zero (Conv3d padding=1): per-thread masks, fixed for the whole march.
  const float* xb = x + (size_t)b * G3;
  int off[9];
  bool ok[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
      const int gy = iy0 + iyl + dy - 1, gz = iz + dz - 1;
      ok[dy * 3 + dz] = gy >= 0 && gy < G && gz >= 0 && gz < G;
      off[dy * 3 + dz] = ok[dy * 3 + dz] ? gy * G + gz : 0;
    }

  float2 yz[CPT / 2];   // channel pairs (FFMA2 / FADD2: half the issue slots of the scalar loop, same bits)
#pragma unroll
  for (int c = 0; c < CPT / 2; ++c) yz[c] = make_float2(0.f, 0.f);

  float win[3][9];  // [dx][dy*3+dz]; win[dx] holds slab ix+dx-1
  float nxt[9];     // prefetched slab
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    win[1][t] = 0.f;                                         // slab -1 (padding)
    win[2][t] = ok[t] ? __ldg(xb + off[t]) : 0.f;            // slab 0
    nxt[t] = ok[t] ? __ldg(xb + G2 + off[t]) : 0.f;          // slab 1
  }

#pragma unroll 1
  for (int ix = 0; ix < G; ++ix) {
    // Let the next kernel become resident only when this one is nearly done: its CTAs' shared memory comes out of the
    // L1 carve-out this kernel's TSDF reads live in (an early trigger costs tens of us per step).
    if (ix == G - 3) pdl_launch();
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      win[0][t] = win[1][t];
      win[1][t] = win[2][t];
      win[2][t] = nxt[t];
    }
    if (ix + 2 < G) {
      const float* xs2 = xb + (size_t)(ix + 2) * G2;
#pragma unroll
      for (int t = 0; t < 9; ++t) nxt[t] = ok[t] ? __ldg(xs2 + off[t]) : 0.f;
    } else {
#pragma unroll
      for (int t = 0; t < 9; ++t) nxt[t] = 0.f;              // slab 40 (padding)
    }
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
    __syncthreads();
  }

  // yz[c][iz][iy]: transpose through smem so that rows of TY consecutive iy are written together
  float* stage = red;  // [CPT][40][TY] floats <= CI_RED
#pragma unroll
  for (int c = 0; c < CPT; ++c) stage[(c * G + iz) * CI_TY + iyl] = ((c & 1) ? yz[c / 2].y : yz[c / 2].x) / 40.0f;
  __syncthreads();
  float* pre_yz = pre + ((size_t)(2 * B + b) * C + c0) * G2;
  for (int o = tid; o < CPT * G * CI_TY; o += CI_THREADS) {
    const int cz = o / CI_TY, r = o % CI_TY;   // cz = c*40 + iz
    pre_yz[cz * G + iy0 + r] = stage[o];
  }
  float* pre_xy = pre + ((size_t)(1 * B + b) * C + c0) * G2;
  for (int o = tid; o < CPT * CI_TY * G; o += CI_THREADS) {
    const int j = o / G, ix = o % G;  // j = c*TY + iyl
    const int c = j / CI_TY, r = j % CI_TY;
    pre_xy[(c * G + iy0 + r) * G + ix] = xyacc[j * CI_RED_STRIDE + ix] / 40.0f;
  }
}

// grid (NT, B, CS), block THREADS
template <int CI_TY, int CS>
__global__ void __launch_bounds__(CI_TY * G, CI_TY == 5 ? (CS == 1 ? 2 : 4) : 8)
conv_in_planes_kernel(const float* __restrict__ x,   // [B][40][40][40]
                      float* __restrict__ pre,       // [3][B][32][40][40]  (xz, xy, yz), NCHW
                      float* __restrict__ xz_part,   // [B][CI_NT][40 ix][32][40 iz]
                      int B, const __grid_constant__ ConvInParams P) {
  extern __shared__ __align__(16) float smem_ci[];
  if constexpr (CS == 1) {
    conv_in_body<CI_TY, 1, 0>(x, pre, xz_part, B, P, smem_ci);
  } else if constexpr (CS == 2) {
    if (blockIdx.z == 0) conv_in_body<CI_TY, 2, 0>(x, pre, xz_part, B, P, smem_ci);
    else conv_in_body<CI_TY, 2, 1>(x, pre, xz_part, B, P, smem_ci);
  } else {
    static_assert(CS <= 4, "channel split");
    switch (blockIdx.z) {
      case 0: conv_in_body<CI_TY, 4, 0>(x, pre, xz_part, B, P, smem_ci); break;
      case 1: conv_in_body<CI_TY, 4, 1>(x, pre, xz_part, B, P, smem_ci); break;
      case 2: conv_in_body<CI_TY, 4, 2>(x, pre, xz_part, B, P, smem_ci); break;
      default: conv_in_body<CI_TY, 4, 3>(x, pre, xz_part, B, P, smem_ci); break;
    }
  }
}

// xz[b][c][iz][ix] = (sum over iy, canonical order) / 40 from the per-CTA partials xz_part[b][t][ix][c][iz]      grid (32, B), block 256
template <int CI_NT>
__global__ void __launch_bounds__(256)
xz_finish_kernel(const float* __restrict__ xz_part, float* __restrict__ pre, int B) {
  __shared__ float tile[G * 41];
  pdl_launch();
  pdl_wait();
  const int c = blockIdx.x, b = blockIdx.y;
  for (int e = threadIdx.x; e < G2; e += 256) {
    const int ix = e / G, z = e % G;
    constexpr int PER = CI_NT / (G / CI_GROUP);   // partials per canonical group: 1 (TY = 5) or 5 (TY = 1)
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G / CI_GROUP; ++g) {
      float sg = xz_part[((((size_t)b * CI_NT + g * PER) * G + ix) * C + c) * G + z];
#pragma unroll
      for (int r = 1; r < PER; ++r) sg += xz_part[((((size_t)b * CI_NT + g * PER + r) * G + ix) * C + c) * G + z];
      s += sg;
    }
    tile[z * 41 + ix] = s / 40.0f;
  }
  __syncthreads();
  float* o = pre + ((size_t)(0 * B + b) * C + c) * G2;
  for (int e = threadIdx.x; e < G2; e += 256) o[e] = tile[(e / G) * 41 + (e % G)];
}

}  // namespace giga
