// Native training step of the GIGA network: device-side parameter packing, the differentiable forward on the fp32 FMA-pipe kernels
// (activations kept), and hand-written BACKWARD kernels for every layer.
//
// Replaces (reference): scripts/train_giga.py:199-211 `_update` = net(x, pos, p_tsdf=pos_occ) -> loss_fn -> loss.backward() -> optimizer.step(),
// i.e. the autograd graph PyTorch builds over conv_onet/models/__init__.py:42-67 (ATen/cuDNN kernels: conv3d/conv2d/conv_transpose2d
// backward-data and backward-filter, max_pool2d_with_indices_backward, grid_sampler_2d_backward, addmm/mm for the Linear layers,
// threshold_backward for the ReLUs, scatter_mean's backward), about 900 launches per step in the reference.
//
// Gradient flow (reverse of giga_encode / decode_points_kernel; every buffer below is fp32):
//   head outputs --(sigmoid / normalize backward)--> decode_points_bwd_kernel: per 128-point tile recompute the head's forward,
//       back-propagate through fc_out, the 5 ResNet blocks and fc_c, reduce the weight gradients over the tile (outer products through
//       shared memory) and add them to the parameter gradients with vector reductions (REDG.ADD.F32x4); feature gradients are scattered
//       bilinearly into the plane gradients (grid_sampler backward) the same way
//   plane gradients [3][B][1600][32] -> conv_final_bwd_kernel (1x1 conv: data + filter gradient)
//   -> per U-Net conv: conv3x3_wgrad_kernel (filter + bias gradient) and conv3x3_kernel<EPI 1> run as the data gradient (the same
//      register-tiled kernel as the forward on flipped / transposed weights, epilogue = ReLU mask of the activation it flows into)
//   -> convT2x2_dgrad_kernel / convT2x2_wgrad_kernel, pool_bwd_kernel (first-maximum routing as ATen)
//   -> conv_in_bwd_kernel: recomputes Conv3d(1->32) per voxel for the ReLU mask, forms d f = mask * (g_xz + g_xy + g_yz) / 40 (the
//      plane means' backward) and reduces the 27x32 filter gradient; the 8 MB/scene feature volume is never materialised in the
//      backward either.
// Weight-gradient sums use floating-point atomics: their order (not their value beyond ~1e-6 relative) varies run to run, as with
// PyTorch's default cuDNN algorithms.
#pragma once
#include "common.cuh"
#include "conv_in.cuh"
#include "decoder.cuh"
#include "unet.cuh"
#include "unet_tall.cuh"

namespace giga {

// ---------------------------------------------------------------------------------------------------------------------------------
// helpers
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {   // 16-byte aligned vector reduction
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) { atomicAdd(p, a); }
// (d0, d1) += (a0, a1) * (b0, b1) as ONE packed FFMA2
__device__ __forceinline__ void fma2p(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n .reg .b64 d, a, b;\n mov.b64 d, {%0, %1};\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n fma.rn.f32x2 d, a, b, d;\n mov.b64 {%0, %1}, d;\n}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// (d0, d1) += (a0, a1) * s as ONE packed FFMA2 on two scalar registers (the allocator pairs them)
__device__ __forceinline__ void fma2s(float& d0, float& d1, float a0, float a1, float s) {
  asm("{\n .reg .b64 d, a, b;\n mov.b64 d, {%0, %1};\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %4};\n fma.rn.f32x2 d, a, b, d;\n mov.b64 {%0, %1}, d;\n}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(s));
}

// ---------------------------------------------------------------------------------------------------------------------------------
// conv_in forward for training: the inference kernel body with its 27x32 weights in __constant__ memory, refreshed device-to-device
// every step from the live parameters (the inference kernel takes them as host-packed kernel parameters).
__constant__ ConvInParams c_conv_in_train;

template <int CI_TY, int CS>
__global__ void __launch_bounds__(CI_TY * G, CI_TY == 5 ? 2 : 8)
conv_in_planes_train_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ tall, long ps, float* __restrict__ xz_part, int B) {
  extern __shared__ __align__(128) float smem_ci[];
  if constexpr (CS == 1) {
    conv_in_body<CI_TY, 1, 0>(&tmap, tall, ps, xz_part, B, c_conv_in_train, smem_ci);
  } else {
    static_assert(CS == 4, "channel split");
    switch (blockIdx.z) {
      case 0: conv_in_body<CI_TY, 4, 0>(&tmap, tall, ps, xz_part, B, c_conv_in_train, smem_ci); break;
      case 1: conv_in_body<CI_TY, 4, 1>(&tmap, tall, ps, xz_part, B, c_conv_in_train, smem_ci); break;
      case 2: conv_in_body<CI_TY, 4, 2>(&tmap, tall, ps, xz_part, B, c_conv_in_train, smem_ci); break;
      default: conv_in_body<CI_TY, 4, 3>(&tmap, tall, ps, xz_part, B, c_conv_in_train, smem_ci); break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Every TALL pre-split activation of the tensor-core forward -> the NCHW fp32 form the backward kernels read, in ONE launch
// (tall_to_nchw_kernel per layer took 15 launches): blockIdx.y = layer, blocks past a layer's size exit.
struct ExpandEntry { const float* src; float* dst; long ps; int hw, c8; };
struct ExpandArgs { ExpandEntry e[16]; int n_img; };

__global__ void __launch_bounds__(256) tall_expand_all_kernel(const __grid_constant__ ExpandArgs A) {
  const ExpandEntry& E = A.e[blockIdx.y];
  const int hw2 = E.hw * E.hw;
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long)A.n_img * E.c8 * hw2) return;
  const int pix = (int)(t % hw2);
  const int kc = (int)((t / hw2) % E.c8);
  const int img = (int)(t / ((long)hw2 * E.c8));
  const long pos = TALL_MARGIN + tall_pos(E.hw, img, pix / E.hw, pix % E.hw);
  float v[8];
  join8(ldu4(E.src + ((size_t)kc * E.ps + pos) * 4), ldu4(E.src + ((size_t)(E.c8 + kc) * E.ps + pos) * 4), v);
  float* p = E.dst + ((size_t)img * (8 * E.c8) + 8 * kc) * hw2 + pix;
#pragma unroll
  for (int j = 0; j < 8; ++j) p[(size_t)j * hw2] = v[j];
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Device-side parameter packing: reference-layout tensors (state_dict) -> the operand layouts of the FMA-pipe kernels.  src is
// [O][I][T] (out, in, taps);  mode 0: dst[(i*T + t)*ld + o]   (forward conv / Linear, input-major)
//                             mode 1: dst[(o*T + (T-1-t))*In + (i - i0)] for i in [i0, i0 + In)   (data gradient: flipped taps, roles swapped)
//                             mode 2: dst[e] = src[e]
struct PackEntry {
  const float* src;
  float* dst;
  int O, I, T, mode, ld, i0, In, pad;
};

__global__ void __launch_bounds__(256) train_pack_kernel(const PackEntry* __restrict__ tab) {
  const PackEntry e = tab[blockIdx.x];
  const int n = e.O * e.I * e.T;
  for (int x = blockIdx.y * 256 + threadIdx.x; x < n; x += gridDim.y * 256) {
    const float v = e.src[x];
    if (e.mode == 2) { e.dst[x] = v; continue; }
    const int t = x % e.T, i = (x / e.T) % e.I, o = x / (e.T * e.I);
    if (e.mode == 0) {
      e.dst[((size_t)i * e.T + t) * e.ld + o] = v;
    } else if (i >= e.i0 && i < e.i0 + e.In) {
      e.dst[((size_t)o * e.T + (e.T - 1 - t)) * e.In + (i - e.i0)] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Decoder backward.  grid (ceil(N/128), B), block 128 (thread = query point), one head per launch.
struct HeadGrads {   // parameter gradients in the reference's layouts (Linear weight = [out][in])
  float *fcp_w, *fcp_b;
  float *fcc_w[5], *fcc_b[5], *w0[5], *b0[5], *w1[5], *b1[5];
  float *out_w, *out_b;
};

constexpr int DB_PTS = 128;
constexpr int DB_ST = 140;   // staging row stride: points contiguous, rows 16-byte aligned, room for the skew below
// Row r of a staging tile starts at r * DB_ST + 4 * ((r >> 3) & 3): in the tile-wide outer products a warp reads rows 8 kg + i (kg = 0..3)
// of the input tile at once, and without the skew those four rows share their banks (8 * stride = 0 mod 32): a 4-way conflict on every load
// (ncu: 57 % of the shared-memory wavefronts were conflicts).
__device__ __forceinline__ int db_row(int r) { return r * DB_ST + 4 * ((r >> 3) & 3); }
constexpr int DB_SMEM_FLOATS = 96 * DB_ST + DW_BLK + 64 * DB_ST;
constexpr int DB_SMEM_BYTES = DB_SMEM_FLOATS * 4;   // 110,464 B -> two CTAs per SM
constexpr int DB_SAVE = 6 * 32 * DB_PTS;            // floats of scratch per tile: the hidden state before each block + the final one

// dW[j][k0 + 8 kg + i] += sum_pt sG[j][pt] * sA[8 kg + i][pt]   (thread = (j, kg)); bias: db[j] += sum_pt sG[j][pt]
__device__ __forceinline__ void wgrad_tile32(const float* sA, const float* sG, float* dW, int ldw, float* db, int tid) {
  const int j = tid >> 2, kg = tid & 3;
  float acc[8], acb[8];   // even / odd points: packed FFMA2 over point pairs, summed at the end
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = acb[i] = 0.f;
  float bs = 0.f;
  const float* gr = sG + db_row(j);
  const float* ar = sA + db_row(8 * kg);      // rows 8 kg .. 8 kg + 7 share one skew
#pragma unroll 2
  for (int p = 0; p < DB_PTS; p += 4) {
    const float4 g = ld4(gr + p);
    bs += (g.x + g.y) + (g.z + g.w);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = ld4(ar + i * DB_ST + p);
      fma2p(acc[i], acb[i], g.x, g.y, a.x, a.y);
      fma2p(acc[i], acb[i], g.z, g.w, a.z, a.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += acb[i];
  float* d = dW + (size_t)j * ldw + 8 * kg;
  red_add4(d, acc[0], acc[1], acc[2], acc[3]);
  red_add4(d + 4, acc[4], acc[5], acc[6], acc[7]);
  if (db && kg == 0) red_add(db + j, bs);
}

// y[k] = sum_j Wt[k][j] * g[j]  (Wt input-major in shared memory: the transposed product of the forward's h += Wt[k][:] * x[k])
__device__ __forceinline__ void matvec_t32(const float* Wt, const float* g, float* y) {
  // 32 independent dot-product chains (j outer, k inner): a lone warp per scheduler needs the instruction-level parallelism
#pragma unroll
  for (int k = 0; k < 32; ++k) y[k] = 0.f;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(Wt + k * 32 + 4 * j4);
      y[k] = fmaf(a.x, g[4 * j4 + 0], y[k]); y[k] = fmaf(a.y, g[4 * j4 + 1], y[k]);
      y[k] = fmaf(a.z, g[4 * j4 + 2], y[k]); y[k] = fmaf(a.w, g[4 * j4 + 3], y[k]);
    }
  }
}

// h[j] += sum_k Wt[k][j] * x[k]   (the forward's layer product, as decode_points_kernel)
__device__ __forceinline__ void matvec32(const float* Wt, const float* x, float* h) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float r = x[k];
    const float4* wr = reinterpret_cast<const float4*>(Wt + k * 32);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 w = wr[j4];
      fma2s(h[4 * j4 + 0], h[4 * j4 + 1], w.x, w.y, r);
      fma2s(h[4 * j4 + 2], h[4 * j4 + 3], w.z, w.w, r);
    }
  }
}

// warp_gather (decoder.cuh) writing the skewed row layout: feat[db_row(plane * 32 + channel) + point]
__device__ __forceinline__ void warp_gather_skew(const float* __restrict__ planes, int B, int b, const float* __restrict__ pts, int n0, int N,
                                                 float* feat, int pt0, float* tinfo_w) {
  const int lane = threadIdx.x & 31;
  {
    const int n = min(n0 + lane, N - 1);
    TexInfo t;
    point_taps(pts + ((size_t)b * N + n) * 3, t);
    int* ti = reinterpret_cast<int*>(tinfo_w) + lane * 24;
    float* tf = tinfo_w + lane * 24 + 12;
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
#pragma unroll 2
  for (int q = 0; q < 32; ++q) {
    const int4* oi = reinterpret_cast<const int4*>(tinfo_w + q * 24);
    const float4* wf = reinterpret_cast<const float4*>(tinfo_w + q * 24 + 12);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const int4 o = oi[pl];
      const float4 w = wf[pl];
      const float v0 = __ldg(pb[pl] + o.x), v1 = __ldg(pb[pl] + o.y), v2 = __ldg(pb[pl] + o.z), v3 = __ldg(pb[pl] + o.w);
      feat[db_row(pl * 32 + lane) + pt0 + q] = v0 * w.x + v1 * w.y + v2 * w.z + v3 * w.w;   // ATen order (as warp_gather)
    }
  }
  __syncwarp();
}

struct DecBwdJob {
  int head, N, tiles;
  const float* pts;    // [B][N][3]
  const float* gout;   // gradient of the loss w.r.t. this head's output: [B][N] or [B][N][4] (rot)
  float* gplanes;      // [3][B][40][40][32] (+=), or null (detached features)
  float* gpts;         // [B][N][3] (+=) gradient w.r.t. the query positions (grad_refine), or null
  long save_off;       // this job's scratch: [B * tiles][DB_SAVE]
  HeadGrads GR;
};
struct DecBwdArgs { DecBwdJob job[4]; };   // every head with a gradient, at its own point set: blockIdx.z picks the job (one launch: the
                                           // latency-bound single-point grasp tiles run beside the TSDF head's thousands of tiles)

// grid (max tiles, B, jobs), block 128
__global__ void __launch_bounds__(DB_PTS, 2)
decode_points_bwd_kernel(const float* __restrict__ planes,   // [3][B][40][40][32]
                         const float* __restrict__ hw,       // [4][DW_HEAD] packed head parameters (input-major)
                         int B, const __grid_constant__ DecBwdArgs A, float* __restrict__ save) {
  if ((int)blockIdx.x >= A.job[blockIdx.z].tiles) return;
  const int head = A.job[blockIdx.z].head, N = A.job[blockIdx.z].N;
  const float* __restrict__ pts = A.job[blockIdx.z].pts;
  const float* __restrict__ gout = A.job[blockIdx.z].gout;
  float* __restrict__ gplanes = A.job[blockIdx.z].gplanes;
  float* __restrict__ gpts = A.job[blockIdx.z].gpts;
  const HeadGrads& GR = A.job[blockIdx.z].GR;
  extern __shared__ __align__(16) float smem[];
  float* feat = smem;                     // [96][DB_ST]
  float* wbuf = feat + 96 * DB_ST;        // [DW_BLK]
  float* sA = wbuf + DW_BLK;              // [32][DB_ST]  layer inputs of the tile
  float* sG = sA + 32 * DB_ST;            // [32][DB_ST]  layer output gradients of the tile
  float* tinfo = sA;                      // gather scratch (dead before the first staging)
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * DB_PTS, n = n0 + tid;
  const bool valid = n < N;
  const int nc = valid ? n : N - 1;
  const int od = head == 1 ? 4 : 1;

  warp_gather_skew(planes, B, b, pts, n0 + warp * 32, N, feat, warp * 32, tinfo + warp * 32 * 24);
  const float* pp = pts + ((size_t)b * N + nc) * 3;
  const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  const float* W = hw + (size_t)head * DW_HEAD;
  float* sv = save + A.job[blockIdx.z].save_off + ((size_t)b * A.job[blockIdx.z].tiles + blockIdx.x) * DB_SAVE + tid;   // [6][32][128], this thread's column

  // ---- forward recompute (decode_points_kernel's arithmetic), keeping the hidden state that enters each ResNet block ----
  float h[32];
#pragma unroll
  for (int j = 0; j < 32; ++j)
    h[j] = __ldg(W + DW_FCP + 96 + j) + __ldg(W + DW_FCP + j) * px + __ldg(W + DW_FCP + 32 + j) * py + __ldg(W + DW_FCP + 64 + j) * pz;
#pragma unroll 1
  for (int blk = 0; blk < 5; ++blk) {
    __syncthreads();
    const float* Wb = W + DW_BLOCK0 + blk * DW_BLK;
    for (int e = tid; e < DW_BLK / 4; e += DB_PTS) st4(wbuf + e * 4, ld4(Wb + e * 4));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] += wbuf[DW_BLK_BC + j];
#pragma unroll 4
    for (int k = 0; k < 96; ++k) {
      const float f = feat[db_row(k) + tid];
      const float4* wr = reinterpret_cast<const float4*>(wbuf + DW_BLK_FCC + k * 32);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w = wr[j4];
        fma2s(h[4 * j4 + 0], h[4 * j4 + 1], w.x, w.y, f);
        fma2s(h[4 * j4 + 2], h[4 * j4 + 3], w.z, w.w, f);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) sv[(blk * 32 + j) * DB_PTS] = h[j];
    float t[32], r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { t[j] = wbuf[DW_BLK_B0 + j]; r[j] = fmaxf(h[j], 0.f); }
    matvec32(wbuf + DW_BLK_W0, r, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) { h[j] += wbuf[DW_BLK_B1 + j]; r[j] = fmaxf(t[j], 0.f); }
    matvec32(wbuf + DW_BLK_W1, r, h);
  }
  float o[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) o[m] = __ldg(W + DW_OUT + 128 + m);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float r = fmaxf(h[k], 0.f);
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + DW_OUT + k * 4));
    o[0] = fmaf(w.x, r, o[0]); o[1] = fmaf(w.y, r, o[1]); o[2] = fmaf(w.z, r, o[2]); o[3] = fmaf(w.w, r, o[3]);
  }

  // ---- head epilogue backward (models/__init__.py:119-123): sigmoid / F.normalize / identity ----
  float go[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const size_t idx = (size_t)b * N + n;
    if (head == 0) {
      const float q = 1.f / (1.f + expf(-o[0]));
      go[0] = gout[idx] * (1.f - q) * q;                               // sigmoid_backward: grad * (1 - y) * y
    } else if (head == 1) {
      const float4 gg = ld4(gout + idx * 4);
      const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
      if (nrm > 1e-12f) {                                              // v / max(|v|, eps): d/dv = (g - y (y . g)) / |v|
        const float inv = 1.f / nrm;
        const float y0 = o[0] * inv, y1 = o[1] * inv, y2 = o[2] * inv, y3 = o[3] * inv;
        const float dot = y0 * gg.x + y1 * gg.y + y2 * gg.z + y3 * gg.w;
        go[0] = (gg.x - y0 * dot) * inv; go[1] = (gg.y - y1 * dot) * inv;
        go[2] = (gg.z - y2 * dot) * inv; go[3] = (gg.w - y3 * dot) * inv;
      } else {                                                          // clamped denominator: a constant
        go[0] = gg.x * 1e12f; go[1] = gg.y * 1e12f; go[2] = gg.z * 1e12f; go[3] = gg.w * 1e12f;
      }
    } else {
      go[0] = gout[idx];
    }
  }

  // ---- fc_out backward: out = Wout relu(h) + bout ----
  float g[32];
  __syncthreads();   // every thread is done with the block-4 weights and (long ago) with the gather scratch
#pragma unroll
  for (int k = 0; k < 32; ++k) sA[db_row(k) + tid] = fmaxf(h[k], 0.f);
#pragma unroll
  for (int m = 0; m < 4; ++m) sG[db_row(m) + tid] = go[m];
  __syncthreads();
  {
    const int m = tid >> 5, k = tid & 31;
    float acc = 0.f, bs = 0.f;
    for (int p = 0; p < DB_PTS; p += 4) {
      const float4 gv = ld4(sG + db_row(m) + p), av = ld4(sA + db_row(k) + p);
      acc = fmaf(gv.x, av.x, acc); acc = fmaf(gv.y, av.y, acc); acc = fmaf(gv.z, av.z, acc); acc = fmaf(gv.w, av.w, acc);
      bs += (gv.x + gv.y) + (gv.z + gv.w);
    }
    if (m < od) {
      red_add(GR.out_w + m * 32 + k, acc);
      if (k == 0) red_add(GR.out_b + m, bs);
    }
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + DW_OUT + k * 4));
    const float s = w.x * go[0] + w.y * go[1] + w.z * go[2] + w.w * go[3];
    g[k] = h[k] > 0.f ? s : 0.f;
  }

  // ---- the five blocks, last to first.  g = dL/d(hidden state leaving block blk) ----
#pragma unroll 1
  for (int blk = 4; blk >= 0; --blk) {
    __syncthreads();   // staging buffers and wbuf free
    const float* Wb = W + DW_BLOCK0 + blk * DW_BLK;
    for (int e = tid; e < DW_BLK / 4; e += DB_PTS) st4(wbuf + e * 4, ld4(Wb + e * 4));
    float hb[32], t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) hb[j] = sv[(blk * 32 + j) * DB_PTS];
    __syncthreads();
    {   // recompute t = fc_0(relu(hb))
      float r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { t[j] = wbuf[DW_BLK_B0 + j]; r[j] = fmaxf(hb[j], 0.f); }
      matvec32(wbuf + DW_BLK_W0, r, t);
    }
    // fc_1: out += W1 relu(t) + b1
#pragma unroll
    for (int k = 0; k < 32; ++k) { sA[db_row(k) + tid] = fmaxf(t[k], 0.f); sG[db_row(k) + tid] = g[k]; }
    __syncthreads();
    wgrad_tile32(sA, sG, GR.w1[blk], 32, GR.b1[blk], tid);
    float gt[32];
    matvec_t32(wbuf + DW_BLK_W1, g, gt);
#pragma unroll
    for (int k = 0; k < 32; ++k) gt[k] = t[k] > 0.f ? gt[k] : 0.f;
    __syncthreads();
    // fc_0: t = W0 relu(hb) + b0
#pragma unroll
    for (int k = 0; k < 32; ++k) { sA[db_row(k) + tid] = fmaxf(hb[k], 0.f); sG[db_row(k) + tid] = gt[k]; }
    __syncthreads();
    wgrad_tile32(sA, sG, GR.w0[blk], 32, GR.b0[blk], tid);
    {
      float ga[32];
      matvec_t32(wbuf + DW_BLK_W0, gt, ga);
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] += hb[k] > 0.f ? ga[k] : 0.f;
    }
    __syncthreads();
    // fc_c[blk]: hb = (state leaving block blk - 1) + Wc feat + bc   ->  the same g flows on to block blk - 1
#pragma unroll
    for (int k = 0; k < 32; ++k) sG[db_row(k) + tid] = g[k];
    __syncthreads();
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
      wgrad_tile32(feat + pl * 32 * DB_ST, sG, GR.fcc_w[blk] + pl * 32, 96, pl == 0 ? GR.fcc_b[blk] : nullptr, tid);
#pragma unroll
    for (int j = 0; j < 32; ++j) sv[(blk * 32 + j) * DB_PTS] = g[j];   // kept for the feature gradient below
  }

  // ---- fc_p: h0 = Wp p + bp ----
  __syncthreads();
  sA[db_row(0) + tid] = px; sA[db_row(1) + tid] = py; sA[db_row(2) + tid] = pz;
#pragma unroll
  for (int k = 0; k < 32; ++k) sG[db_row(k) + tid] = g[k];
  __syncthreads();
  {
    const int j = tid >> 2, c = tid & 3;
    float acc = 0.f;
    for (int p = 0; p < DB_PTS; p += 4) {
      const float4 gv = ld4(sG + db_row(j) + p);
      if (c < 3) {
        const float4 av = ld4(sA + db_row(c) + p);
        acc = fmaf(gv.x, av.x, acc); acc = fmaf(gv.y, av.y, acc); acc = fmaf(gv.z, av.z, acc); acc = fmaf(gv.w, av.w, acc);
      } else {
        acc += (gv.x + gv.y) + (gv.z + gv.w);
      }
    }
    if (c < 3) red_add(GR.fcp_w + j * 3 + c, acc);
    else red_add(GR.fcp_b + j, acc);
  }
  float dp[3] = {0.f, 0.f, 0.f};
  if (gpts) {   // d h0 / d p = Wp (decoder.py:163, fc_p)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      dp[0] = fmaf(__ldg(W + DW_FCP + j), g[j], dp[0]);
      dp[1] = fmaf(__ldg(W + DW_FCP + 32 + j), g[j], dp[1]);
      dp[2] = fmaf(__ldg(W + DW_FCP + 64 + j), g[j], dp[2]);
    }
  }
  if (!gplanes && !gpts) return;

  // ---- feature gradient: d feat[k] = sum_blk sum_j Wc_blk[j][k] g_blk[j], scattered bilinearly into the plane gradients ----
  float gf[96];
#pragma unroll
  for (int k = 0; k < 96; ++k) gf[k] = 0.f;
#pragma unroll 1
  for (int blk = 0; blk < 5; ++blk) {
    __syncthreads();
    const float* Wb = W + DW_BLOCK0 + blk * DW_BLK + DW_BLK_FCC;
    for (int e = tid; e < 96 * 32 / 4; e += DB_PTS) st4(wbuf + e * 4, ld4(Wb + e * 4));
    float gb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) gb[j] = sv[(blk * 32 + j) * DB_PTS];
    __syncthreads();
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      float y[32];
      matvec_t32(wbuf + pl * 32 * 32, gb, y);
#pragma unroll
      for (int k = 0; k < 32; ++k) gf[pl * 32 + k] += y[k];
    }
  }
  if (!valid) return;
  if (gpts) {
    // grid_sampler_2d_backward w.r.t. the grid (bilinear, border, align_corners = True) chained through decoder.py:119-121 (g = 2 t - 1) and
    // normalize_coordinate (common.py:253-260: the one-sided clamps are constants).  Plane axes: xz (u = x, v = z), xy (x, y), yz (y, z).
    const float pv[3] = {px, py, pz};
    const int au[3] = {0, 0, 1}, av[3] = {2, 1, 2};
    float pix[3], dpix[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float t = __fdiv_rn(pv[a], 1.00001f) + 0.5f, dt = 1.f / 1.00001f;
      if (t >= 1.f) { t = 0.99999f; dt = 0.f; }
      if (t < 0.f) { t = 0.f; dt = 0.f; }
      float ix = ((2.0f * t - 1.0f) + 1.f) * 19.5f, d = 39.f * dt;
      if (ix <= 0.f) { ix = 0.f; d = 0.f; } else if (ix >= 39.f) { ix = 39.f; d = 0.f; }   // clip_coordinates_set_grad
      pix[a] = ix; dpix[a] = d;
    }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const float ix = pix[au[pl]], iy = pix[av[pl]];
      const float x0f = floorf(ix), y0f = floorf(iy);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const float tx = ix - x0f, ty = iy - y0f, sx = 1.f - tx, sy = 1.f - ty;
      const bool xin = x0 + 1 < G, yin = y0 + 1 < G;            // out-of-range corners read as zero (ATen safe_get)
      const float* pb = planes + ((size_t)pl * B + b) * (G2 * C);
      const float* nw = pb + (y0 * G + x0) * C;
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 tnw = __ldg(reinterpret_cast<const float4*>(nw) + c4);
        const float4 tne = xin ? __ldg(reinterpret_cast<const float4*>(nw + C) + c4) : z4;
        const float4 tsw = yin ? __ldg(reinterpret_cast<const float4*>(nw + G * C) + c4) : z4;
        const float4 tse = (xin && yin) ? __ldg(reinterpret_cast<const float4*>(nw + G * C + C) + c4) : z4;
        const float* gq = gf + pl * 32 + 4 * c4;
        gix += gq[0] * ((tne.x - tnw.x) * sy + (tse.x - tsw.x) * ty) + gq[1] * ((tne.y - tnw.y) * sy + (tse.y - tsw.y) * ty) +
               gq[2] * ((tne.z - tnw.z) * sy + (tse.z - tsw.z) * ty) + gq[3] * ((tne.w - tnw.w) * sy + (tse.w - tsw.w) * ty);
        giy += gq[0] * ((tsw.x - tnw.x) * sx + (tse.x - tne.x) * tx) + gq[1] * ((tsw.y - tnw.y) * sx + (tse.y - tne.y) * tx) +
               gq[2] * ((tsw.z - tnw.z) * sx + (tse.z - tne.z) * tx) + gq[3] * ((tsw.w - tnw.w) * sx + (tse.w - tne.w) * tx);
      }
      dp[au[pl]] = fmaf(gix, dpix[au[pl]], dp[au[pl]]);
      dp[av[pl]] = fmaf(giy, dpix[av[pl]], dp[av[pl]]);
    }
    float* gp = gpts + ((size_t)b * N + n) * 3;
    red_add(gp, dp[0]); red_add(gp + 1, dp[1]); red_add(gp + 2, dp[2]);
  }
  if (!gplanes) return;
  TexInfo ti;
  point_taps(pp, ti);
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    float* base = gplanes + ((size_t)pl * B + b) * (G2 * C);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float w = ti.w[pl][q];
      if (w == 0.f) continue;
      float* d = base + ti.off[pl][q];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        red_add4(d + 4 * c4, w * gf[pl * 32 + 4 * c4], w * gf[pl * 32 + 4 * c4 + 1], w * gf[pl * 32 + 4 * c4 + 2], w * gf[pl * 32 + 4 * c4 + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// conv_final backward (unet.py:238, 1x1 conv 32 -> 32, channels-last output):
//   g_in[ci][p] = (act[ci][p] > 0) * sum_co g[p][co] W[co][ci];   dW[co][ci] += sum_p g[p][co] act[ci][p];   db[co] += sum_p g[p][co]
// grid = min(tiles, a few waves), block 256; a CTA walks 64-pixel tiles and flushes its filter gradient once.
constexpr int FB_PIX = 64;
__global__ void __launch_bounds__(256)
conv_final_bwd_kernel(const float* __restrict__ gpl,    // [n_img][1600][32]  plane gradients
                      const float* __restrict__ act,    // [n_img][32][1600]  u1c2 (post-ReLU)
                      const float* __restrict__ wt,     // [ci][co] packed
                      float* __restrict__ gin,          // [n_img][32][1600]
                      float* __restrict__ dW,           // [co][ci]
                      float* __restrict__ db, int n_tiles) {
  __shared__ __align__(16) float Gs[FB_PIX * 33];   // [p][co]
  __shared__ __align__(16) float As[C * (FB_PIX + 1)];   // [ci][p]
  __shared__ __align__(16) float Ws[C * 33];        // [ci][co] (padded)
  const int tid = threadIdx.x;
  for (int e = tid; e < C * C; e += 256) Ws[(e / C) * 33 + (e % C)] = __ldg(wt + e);
  const int wco = tid >> 3, wcg = tid & 7;   // filter-gradient ownership: co, 4 ci
  float wacc[4] = {0.f, 0.f, 0.f, 0.f}, bacc = 0.f;
  constexpr int TPI = G2 / FB_PIX;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int img = tile / TPI, pix0 = (tile % TPI) * FB_PIX;
    __syncthreads();
    for (int e = tid; e < FB_PIX * C; e += 256) Gs[(e / C) * 33 + (e % C)] = __ldg(gpl + ((size_t)img * G2 + pix0) * C + e);
    for (int e = tid; e < C * FB_PIX; e += 256) As[(e / FB_PIX) * (FB_PIX + 1) + (e % FB_PIX)] = __ldg(act + ((size_t)img * C + e / FB_PIX) * G2 + pix0 + e % FB_PIX);
    __syncthreads();
    {   // data gradient: thread = (pixel p, 8 ci)
      const int p = tid & 63, cg = tid >> 6;
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 8
      for (int co = 0; co < C; ++co) {
        const float gv = Gs[p * 33 + co];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(gv, Ws[(cg * 8 + k) * 33 + co], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ci = cg * 8 + k;
        gin[((size_t)img * C + ci) * G2 + pix0 + p] = As[ci * (FB_PIX + 1) + p] > 0.f ? acc[k] : 0.f;
      }
    }
    {   // filter gradient
#pragma unroll 4
      for (int p = 0; p < FB_PIX; ++p) {
        const float gv = Gs[p * 33 + wco];
#pragma unroll
        for (int k = 0; k < 4; ++k) wacc[k] = fmaf(gv, As[(wcg * 4 + k) * (FB_PIX + 1) + p], wacc[k]);
        bacc += gv;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red_add(dW + wco * C + wcg * 4 + k, wacc[k]);
  if (wcg == 0) red_add(db + wco, bacc);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Filter gradient of a 3x3 convolution (stride 1, zero pad 1):
//   dW[co][ci][dy][dx] += sum_{img,y,x} gz[co][y][x] * in[ci][y+dy-1][x+dx-1];   db[co] += sum gz[co][y][x]
// A CTA owns a 32 co x 32 ci tile of the filter and walks (image, row band) tiles of the activations: warp = 4 output channels,
// lane = input channel, 36 accumulators per thread, flushed once with atomics.  grid (co tiles * ci tiles, P), block 256.
template <int HW_, int R_>
struct WgradCfg {
  static constexpr int HW = HW_, R = R_;
  static constexpr int GPR = (HW + 3) / 4;
  static constexpr int GW = 4 * GPR;                 // gradient row (padded to whole groups)
  static constexpr int CS = 4 * GPR + 8;             // input row: pixel x at column x + 4
  static constexpr int PSX = ((R + 2) * CS / 4 % 2 == 1) ? (R + 2) * CS : (R + 2) * CS + 4;   // odd number of 16-byte units: conflict-free lane stride
  static constexpr int GST = 36;                     // gradient tile: [row][pixel][32 co] with a 36-float pixel stride (16-byte aligned, staging
                                                     // stores at most 4-way conflicted)
  static constexpr int NB = HW / R;
  static constexpr int SMEM_BYTES = (32 * PSX + R * GW * GST) * 4;
  static_assert(HW % R == 0, "row band");
};

template <class K>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_kernel(const float* __restrict__ gz,    // [n_img][COUT][HW][HW]  gradient w.r.t. the conv's pre-activation
                     const float* __restrict__ in,    // [n_img][CIN_SRC][HW][HW]
                     int n_img, int COUT, int CIN_SRC,
                     float* __restrict__ dW,          // [COUT][CIN_TOTAL][3][3], already offset to this source's first input channel
                     int CIN_TOTAL, float* __restrict__ db,     // db: null for the second source of a concat
                     unsigned* __restrict__ next_tile) {        // [co tiles * ci tiles] zeroed work counters: tiles are handed out dynamically
  constexpr int HW = K::HW, R = K::R, GPR = K::GPR, GW = K::GW, CS = K::CS, PSX = K::PSX, GST = K::GST;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                 // [32 ci][R + 2][CS]
  float* gs = smem + 32 * PSX;      // [R][GW][GST]: the 32 output channels of a pixel are contiguous
  const int n_ci = CIN_SRC / 32;
  const int co0 = (blockIdx.x / n_ci) * 32, ci0 = (blockIdx.x % n_ci) * 32;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  // output-channel pairs (4w, 4w+1), (4w+2, 4w+3): the inner loop is FFMA2 (bit-identical to scalar fma, half the issue slots)
  float2 acc[2][9];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[k][t] = make_float2(0.f, 0.f);
  float2 bacc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  const int n_tiles = n_img * K::NB;
  __shared__ int s_tile;
#pragma unroll 1
  for (;;) {
    __syncthreads();   // the previous tile's reads of xs / gs (and of s_tile) are done
    if (tid == 0) s_tile = (int)atomicAdd(next_tile + blockIdx.x, 1u);   // (a static stride leaves the SMs 4 or 5 tiles each: 14 % tail)
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    const int img = tile / K::NB, y0 = (tile % K::NB) * R;
    for (int e = tid; e < 32 * (R + 2) * CS; e += 256) {
      const int c = e / ((R + 2) * CS), rem = e % ((R + 2) * CS), r = rem / CS, col = rem % CS;
      const int y = y0 + r - 1, xx = col - 4;
      float v = 0.f;
      if (y >= 0 && y < HW && xx >= 0 && xx < HW) v = __ldg(in + (((size_t)img * CIN_SRC + ci0 + c) * HW + y) * HW + xx);
      xs[c * PSX + r * CS + col] = v;
    }
    for (int e = tid; e < 32 * R * GW; e += 256) {
      const int c = e / (R * GW), rem = e % (R * GW), r = rem / GW, xx = rem % GW;
      gs[(r * GW + xx) * GST + c] = xx < HW ? __ldg(gz + (((size_t)img * COUT + co0 + c) * HW + y0 + r) * HW + xx) : 0.f;
    }
    __syncthreads();
    const float* xl = xs + lane * PSX;
    const float* gl = gs + 4 * w;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* xrow = xl + (r + dy) * CS;
        float4 prev = ld4(xrow), cur = ld4(xrow + 4);
#pragma unroll 2
        for (int xg = 0; xg < GPR; ++xg) {
          const float4 nxt = ld4(xrow + 4 * (xg + 2));
          const float xv[6] = {prev.w, cur.x, cur.y, cur.z, cur.w, nxt.x};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 gq = ld4(gl + (r * GW + 4 * xg + p) * GST);   // 4 output channels at pixel 4 xg + p (warp-uniform address)
            const float2 g01 = make_float2(gq.x, gq.y), g23 = make_float2(gq.z, gq.w);
            if (dy == 0) { add2(bacc[0], g01); add2(bacc[1], g23); }
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              fma2(acc[0][dy * 3 + dx], g01, xv[p + dx]);
              fma2(acc[1][dy * 3 + dx], g23, xv[p + dx]);
            }
          }
          prev = cur;
          cur = nxt;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* d = dW + ((size_t)(co0 + 4 * w + k) * CIN_TOTAL + ci0 + lane) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) red_add(d + t, (k & 1) ? acc[k >> 1][t].y : acc[k >> 1][t].x);
    if (db && ci0 == 0 && lane == 0) red_add(db + co0 + 4 * w + k, (k & 1) ? bacc[k >> 1].y : bacc[k >> 1].x);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// ConvTranspose2d(k2, s2) backward.  Forward: out[co][2i+a][2j+b] = bias[co] + sum_ci in[ci][i][j] W[ci][co][a][b].
//   data:   gin[ci][i][j] = (in[ci][i][j] > 0) * sum_co sum_ab gout[co][2i+a][2j+b] W[ci][co][a][b]     (in = a ReLU output)
//   filter: dW[ci][co][a][b] += sum_{img,i,j} in[ci][i][j] gout[co][2i+a][2j+b];   db[co] += sum gout[co]
// Both: warp = 4 channels of one side, lane = channel of the other side (conflict-free lane stride), one input row band per tile.
template <int HWI_, int CIN_, int COUT_>
struct ConvTBwdCfg {
  static constexpr int HWI = HWI_, CIN = CIN_, COUT = COUT_, HO = 2 * HWI_;
  static constexpr int GPR = (HWI + 3) / 4;
  static constexpr int XS = 4 * GPR;                 // input row (padded)
  static constexpr int GSW = 8 * GPR;                // output-gradient row (padded)
  static constexpr int R = HWI == 10 ? 10 : 4;       // input rows per tile
  static constexpr int NB = HWI / R;
  static constexpr int PSX = (R * XS / 4 % 2 == 1) ? R * XS : R * XS + 4;
  static constexpr int PSG = (2 * R * GSW / 4 % 2 == 1) ? 2 * R * GSW : 2 * R * GSW + 4;
};

// filter gradient: CTA = 32 ci x 32 co tile; warp = 4 co, lane = ci; acc[4 co][4 ab].  grid (CIN/32 * COUT/32, P), block 256
template <class K>
__global__ void __launch_bounds__(256)
convT2x2_wgrad_kernel(const float* __restrict__ in,     // [n_img][CIN][HWI][HWI]
                      const float* __restrict__ gout,   // [n_img][COUT][HO][HO]
                      int n_img, float* __restrict__ dW, float* __restrict__ db) {
  constexpr int HWI = K::HWI, HO = K::HO, R = K::R, GPR = K::GPR, XS = K::XS, GSW = K::GSW, PSX = K::PSX;
  constexpr int PSGB = 2 * R * GSW;   // broadcast reads: no stride constraint
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                 // [32 ci][R][XS]
  float* gs = smem + 32 * PSX;      // [32 co][2R][GSW]
  constexpr int n_co = K::COUT / 32;
  const int ci0 = (blockIdx.x / n_co) * 32, co0 = (blockIdx.x % n_co) * 32;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  float acc[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[k][q] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  const int n_tiles = n_img * K::NB;
#pragma unroll 1
  for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
    const int img = tile / K::NB, i0 = (tile % K::NB) * R;
    __syncthreads();
    for (int e = tid; e < 32 * R * XS; e += 256) {
      const int c = e / (R * XS), rem = e % (R * XS), r = rem / XS, col = rem % XS;
      xs[c * PSX + r * XS + col] = col < HWI ? __ldg(in + (((size_t)img * K::CIN + ci0 + c) * HWI + i0 + r) * HWI + col) : 0.f;
    }
    for (int e = tid; e < 32 * 2 * R * GSW; e += 256) {
      const int c = e / (2 * R * GSW), rem = e % (2 * R * GSW), r = rem / GSW, col = rem % GSW;
      gs[c * PSGB + r * GSW + col] = col < HO ? __ldg(gout + (((size_t)img * K::COUT + co0 + c) * HO + 2 * i0 + r) * HO + col) : 0.f;
    }
    __syncthreads();
    const float* xl = xs + lane * PSX;
#pragma unroll 1
    for (int r = 0; r < R; ++r)
#pragma unroll 1
      for (int xg = 0; xg < GPR; ++xg) {
        const float4 xq = ld4(xl + r * XS + 4 * xg);
        const float xv[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float* gr = gs + (4 * w + k) * PSGB + (2 * r + a) * GSW + 8 * xg;
            const float4 g0 = ld4(gr), g1 = ld4(gr + 4);
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};   // (pixel p, b) at 2p + b
            bacc[k] += ((g0.x + g0.y) + (g0.z + g0.w)) + ((g1.x + g1.y) + (g1.z + g1.w));
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              acc[k][2 * a + 0] = fmaf(xv[p], gv[2 * p + 0], acc[k][2 * a + 0]);
              acc[k][2 * a + 1] = fmaf(xv[p], gv[2 * p + 1], acc[k][2 * a + 1]);
            }
          }
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* d = dW + ((size_t)(ci0 + lane) * K::COUT + co0 + 4 * w + k) * 4;
    red_add4(d, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    if (ci0 == 0 && lane == 0) red_add(db + co0 + 4 * w + k, bacc[k]);
  }
}

// data gradient: CTA = (image, row band) x 32 ci; warp = 4 ci (weights broadcast), lane = (row, 4-pixel group) slot.
// grid (NB * CIN/32, n_img), block 256.  Weights W[ci][co][4] for the CTA's 32 ci are staged once: 32 * COUT * 4 floats.
template <class K>
__global__ void __launch_bounds__(256)
convT2x2_dgrad_kernel(const float* __restrict__ gout,   // [n_img][COUT][HO][HO]
                      const float* __restrict__ wgt,    // [CIN][COUT][2][2]
                      const float* __restrict__ in,     // [n_img][CIN][HWI][HWI]  (ReLU mask)
                      float* __restrict__ gin) {        // [n_img][CIN][HWI][HWI]
  constexpr int HWI = K::HWI, HO = K::HO, R = K::R, GPR = K::GPR, GSW = K::GSW, PSG = K::PSG, COUT = K::COUT;
  constexpr int CC = 32;                 // output-gradient channels staged per chunk
  constexpr int NSLOT = R * GPR;         // (row, group) slots of the band
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                      // [32 ci][COUT][4]
  float* gs = smem + 32 * COUT * 4;      // [CC co][2R][GSW]  (lane-strided by slot, broadcast over co)
  const int img = blockIdx.y;
  const int band = blockIdx.x % K::NB, ci0 = (blockIdx.x / K::NB) * 32;
  const int i0 = band * R;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 32 * COUT; e += 256) st4(ws + e * 4, ld4(wgt + ((size_t)ci0 * COUT + e) * 4));
  // a warp covers its 4 ci for slots lane, lane + 32, ...
  constexpr int SPL = (NSLOT + 31) / 32;
  float acc[SPL][4][4];   // [slot][ci][pixel]
#pragma unroll
  for (int s = 0; s < SPL; ++s)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[s][k][p] = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < COUT; c0 += CC) {
    __syncthreads();
    for (int e = tid; e < CC * 2 * R * GSW; e += 256) {
      const int c = e / (2 * R * GSW), rem = e % (2 * R * GSW), r = rem / GSW, col = rem % GSW;
      gs[c * PSG + r * GSW + col] = col < HO ? __ldg(gout + (((size_t)img * COUT + c0 + c) * HO + 2 * i0 + r) * HO + col) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SPL; ++s) {
      const int slot = lane + 32 * s;
      if (slot >= NSLOT) break;
      const int r = slot / GPR, xg = slot % GPR;
#pragma unroll 2
      for (int c = 0; c < CC; ++c) {
        const float* gr = gs + c * PSG + (2 * r) * GSW + 8 * xg;
        const float4 a0 = ld4(gr), a1 = ld4(gr + 4), b0 = ld4(gr + GSW), b1 = ld4(gr + GSW + 4);
        const float g0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};   // a = 0: (pixel p, b) at 2p + b
        const float g1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};   // a = 1
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 wv = ld4(ws + ((4 * w + k) * COUT + c0 + c) * 4);   // (a,b) = 00 01 10 11
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float v = acc[s][k][p];
            v = fmaf(g0[2 * p], wv.x, v); v = fmaf(g0[2 * p + 1], wv.y, v);
            v = fmaf(g1[2 * p], wv.z, v); v = fmaf(g1[2 * p + 1], wv.w, v);
            acc[s][k][p] = v;
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    const int slot = lane + 32 * s;
    if (slot >= NSLOT) break;
    const int r = slot / GPR, xg = slot % GPR;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t base = (((size_t)img * K::CIN + ci0 + 4 * w + k) * HWI + i0 + r) * HWI + 4 * xg;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (4 * xg + p < HWI) gin[base + p] = __ldg(in + base + p) > 0.f ? acc[s][k][p] : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// MaxPool2d(2,2) backward onto the (ReLU'd) activation it pooled: the window's FIRST maximum (row-major scan, as ATen's
// max_pool2d_with_indices) receives the gradient; a zero maximum sits on the flat part of the ReLU and passes nothing.
// g_full already holds the gradient that reached the activation through the skip connection (+=).  thread = pooled pixel.
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ full, const float* __restrict__ gpool, float* __restrict__ gfull, int HWP, long total) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int x = (int)(t % HWP), y = (int)((t / HWP) % HWP);
  const long plane = t / ((long)HWP * HWP);
  const int HWF = 2 * HWP;
  const size_t base = (size_t)plane * HWF * HWF + (size_t)(2 * y) * HWF + 2 * x;
  const float v00 = full[base], v01 = full[base + 1], v10 = full[base + HWF], v11 = full[base + HWF + 1];
  float m = v00;
  size_t at = base;
  if (v01 > m) { m = v01; at = base + 1; }
  if (v10 > m) { m = v10; at = base + HWF; }
  if (v11 > m) { m = v11; at = base + HWF + 1; }
  if (m > 0.f) gfull[at] += gpool[t];
}

// ---------------------------------------------------------------------------------------------------------------------------------
// conv_in backward (encoder/voxels.py:107 + the plane means of :57-66).  g_pre = gradients of the three pre-U-Net planes
// [3][B][32][40][40] (xz: [iz][ix], xy: [iy][ix], yz: [iz][iy]).  A CTA walks (scene, ix) slabs; per slab, in five sub-tiles of
// 8 iz x 40 iy voxels:  phase 1 (thread = voxel): recompute z_c = conv(x) + b_c for the 32 channels, d f_c = (z_c > 0) * (g_xz + g_xy
// + g_yz) / 40 into shared memory;  phase 2 (warp = (dx, dy) tap pair, lane = channel): dW[c][dx][dy][0..2] += sum_vox d f_c * x(shifted);
// warp 9 sums the bias gradient.  Accumulators live in registers across all slabs of the CTA, flushed once.
constexpr int CB_THREADS = 320;
constexpr int CB_XR = 43;                   // x-slab row stride ([iz + 1][iy + 1], iy fastest)
constexpr int CB_XS = 42 * CB_XR;           // one ix plane of the staged volume
constexpr int CB_GF = 324;                  // d f row stride: 16-byte aligned rows, an odd number of 16-byte units (conflict-free lane = channel float4 reads)
constexpr int CB_XTOT = round_up(3 * CB_XS, 4);   // the weight rows behind the slabs are read as float4
constexpr int CB_SMEM_FLOATS = CB_XTOT + 28 * 32 + 2 * 32 * 40 + 32 * CB_GF;
constexpr int CB_SMEM_BYTES = CB_SMEM_FLOATS * 4;

__global__ void __launch_bounds__(CB_THREADS)
conv_in_bwd_kernel(const float* __restrict__ x,       // [B][40 ix][40 iy][40 iz]
                   const float* __restrict__ wpk,     // [27][32] + [32] bias (packed, tap = dx*9 + dy*3 + dz)
                   const float* __restrict__ gpre,    // [3][B][32][1600]
                   int B, float* __restrict__ dW,     // [32][27]
                   float* __restrict__ db) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                    // [3 dx][42 (iz+1)][43 (iy+1)]
  float* ws = xs + CB_XTOT;             // [27][32], [32]
  float* gxz = ws + 28 * 32;           // [32][40 iz]   at this ix
  float* gxy = gxz + 32 * 40;          // [32][40 iy]
  float* gfs = gxy + 32 * 40;          // [32][CB_GF]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 28 * 32; e += CB_THREADS) ws[e] = __ldg(wpk + e);
  float acc[3] = {0.f, 0.f, 0.f};      // warp < 9: taps (dx, dy) = (warp / 3, warp % 3), dz = 0..2, channel = lane; warp 9: acc[0] = bias gradient
  const int tdx = warp / 3, tdy = warp % 3;
#pragma unroll 1
  for (int slab = blockIdx.x; slab < B * G; slab += gridDim.x) {
    const int b = slab / G, ix = slab % G;
    __syncthreads();
    for (int e = tid; e < 3 * 42 * 42; e += CB_THREADS) {   // iz fastest in the source: coalesced reads, transposing writes
      const int zz = e % 42, yy = (e / 42) % 42, d = e / (42 * 42);
      const int sx = ix + d - 1, sy = yy - 1, sz = zz - 1;
      float v = 0.f;
      if (sx >= 0 && sx < G && sy >= 0 && sy < G && sz >= 0 && sz < G) v = __ldg(x + (((size_t)b * G + sx) * G + sy) * G + sz);
      xs[d * CB_XS + zz * CB_XR + yy] = v;
    }
    for (int e = tid; e < 32 * 40; e += CB_THREADS) {
      const int c = e / 40, k = e % 40;
      gxz[e] = __ldg(gpre + (((size_t)0 * B + b) * C + c) * G2 + k * G + ix);   // [iz = k][ix]
      gxy[e] = __ldg(gpre + (((size_t)1 * B + b) * C + c) * G2 + k * G + ix);   // [iy = k][ix]
    }
    const float* gyz = gpre + ((size_t)2 * B + b) * C * G2;                      // [c][iz][iy]
#pragma unroll 1
    for (int zt = 0; zt < 5; ++zt) {
      __syncthreads();   // staging complete / previous sub-tile's phase 2 done
      {   // phase 1: voxel (iy = tid % 40, iz = 8 zt + tid / 40)
        const int iy = tid % G, izl = tid / G, iz = 8 * zt + izl;
        float2 z2[16];   // channel pairs: FFMA2 (each component the forward's own fma sequence: the ReLU mask is the forward's, bit for bit)
        float gy[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) gy[c] = __ldg(gyz + (size_t)c * G2 + iz * G + iy);   // in flight during the FMAs below
#pragma unroll
        for (int c = 0; c < 16; ++c) z2[c] = make_float2(ws[27 * 32 + 2 * c], ws[27 * 32 + 2 * c + 1]);
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
              const float xv = xs[d * CB_XS + (iz + dz) * CB_XR + iy + dy];
              const float4* wr = reinterpret_cast<const float4*>(ws + (d * 9 + dy * 3 + dz) * 32);
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 wv = wr[c4];
                fma2(z2[2 * c4], make_float2(wv.x, wv.y), xv);
                fma2(z2[2 * c4 + 1], make_float2(wv.z, wv.w), xv);
              }
            }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float gsum = (gxz[c * 40 + iz] + gxy[c * 40 + iy]) + gy[c];
          gfs[c * CB_GF + tid] = ((c & 1) ? z2[c >> 1].y : z2[c >> 1].x) > 0.f ? gsum / 40.0f : 0.f;
        }
      }
      __syncthreads();
      if (warp < 9) {   // phase 2
        const float* gl = gfs + lane * CB_GF;
        const float* xb = xs + tdx * CB_XS + tdy;
#pragma unroll 1
        for (int iy = 0; iy < G; iy += 4) {
          // voxels (iy .. iy + 3, iz = 8 zt + k): x(ix + dx - 1, iy + dy - 1, iz + dz - 1) = xs[dx][iz + dz][iy + dy]; four iy per 128-bit
          // load of d f (the loop is shared-memory-issue bound)
          float xw[10][4];
#pragma unroll
          for (int k = 0; k < 10; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) xw[k][j] = xb[(8 * zt + k) * CB_XR + iy + j];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 gv = ld4(gl + k * G + iy);
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
              acc[dz] = fmaf(gv.x, xw[k + dz][0], acc[dz]); acc[dz] = fmaf(gv.y, xw[k + dz][1], acc[dz]);
              acc[dz] = fmaf(gv.z, xw[k + dz][2], acc[dz]); acc[dz] = fmaf(gv.w, xw[k + dz][3], acc[dz]);
            }
          }
        }
      } else {
        const float* gl = gfs + lane * CB_GF;
        float s = 0.f;
        for (int v = 0; v < 320; ++v) s += gl[v];
        acc[0] += s;
      }
    }
  }
  if (warp < 9) {
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) red_add(dW + lane * 27 + tdx * 9 + tdy * 3 + dz, acc[dz]);
  } else {
    red_add(db + lane, acc[0]);
  }
}

}  // namespace giga
