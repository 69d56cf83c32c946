// fp32 direct-convolution kernels for the shared tri-plane U-Net (encoder/unet.py:48-114,225-239).
//
// All 3B planes of a batch are stacked into one image batch (the reference runs the same
// weights three times on B images, voxels.py:69-70).  Activations inside the U-Net are NCHW
// fp32 in HBM (rows contiguous -> the smem input tile is filled with coalesced row reads and
// needs no transpose); the last layer writes channels-last for the decoder's gather.
//
// conv3x3_kernel: register-tiled implicit GEMM on the FMA pipe.  A CTA owns R rows of one image
// and CT output channels and walks the input channels in chunks of CC.  Per chunk it stages
//   As[CC][R+2][CS]   zero-padded input rows (pixel x at column x+4, so 4-pixel groups are 16 B aligned)
//   Bs[CC][9][CT]     weights, output channel fastest
// A thread owns a 2x4 pixel patch x TCO output channels: per input channel it reads 4 rows x 6
// values once (shared by the 3 dy and 3 dx taps) and 9 x TCO weights as broadcast 128-bit loads,
// for 72*TCO FMAs.  Bias + ReLU (+ 2x2 max-pool of the patch, DownConv's `pool`) are fused in the
// epilogue; the concat of UpConv (unet.py:109) is two source pointers instead of a copy.
#pragma once
#include "common.cuh"

namespace giga {

// EPI 0: forward epilogue (bias + ReLU, optional pooled copy).  EPI 1: the kernel runs as the DATA GRADIENT of a 3x3 conv (train_bwd.cuh):
// src0 = gradient w.r.t. the conv's pre-activation, wp = flipped / role-swapped weights, `bias` = the activation the gradient flows
// into (its ReLU mask; null = no mask), no bias, no ReLU.
template <int HW_, int CIN0_, int CIN1_, int COUT_, int R_, int CT_, int TCO_, int CC_, bool POOL_, int EPI_ = 0>
struct Conv3x3Cfg {
  static constexpr int EPI = EPI_;
  static constexpr int HW = HW_, CIN0 = CIN0_, CIN1 = CIN1_, CIN = CIN0_ + CIN1_, COUT = COUT_;
  static constexpr int R = R_, CT = CT_, TCO = TCO_, CC = CC_;
  static constexpr bool POOL = POOL_;
  static constexpr int GPR = (HW + 3) / 4;       // 4-pixel groups per row
  static constexpr int CS = 4 * GPR + 8;         // smem columns
  static constexpr int RS = R + 2;               // smem rows
  static constexpr int PS = RS * CS;             // smem plane stride
  static constexpr int NPT = (R / 2) * GPR;      // pixel threads
  static constexpr int NTC = CT / TCO;           // cout threads
  static constexpr int NTHREADS = round_up(NPT * NTC, 32);
  static constexpr int NB = HW / R;              // row bands per image
  static constexpr int NCT = COUT / CT;          // cout tiles
  static constexpr int A_FLOATS = CC * PS;
  static constexpr int B_FLOATS = CC * 9 * CT;
  static constexpr int SMEM_BYTES = (A_FLOATS + B_FLOATS) * 4;
  static_assert(R % 2 == 0 && HW % R == 0, "row band");
  static_assert(COUT % CT == 0 && CIN % CC == 0 && CT % TCO == 0, "channel tiling");
  static_assert(TCO == 4 || TCO == 8, "TCO");
  static_assert(CIN0 % CC == 0, "a chunk must not straddle the two concat sources");
};

// grid (NB*NCT, n_img), block NTHREADS, dynamic smem SMEM_BYTES
template <class K>
__global__ void __launch_bounds__(K::NTHREADS)
conv3x3_kernel(const float* __restrict__ src0,  // [n_img][CIN0][HW][HW]
               const float* __restrict__ src1,  // [n_img][CIN1][HW][HW] or null
               const float* __restrict__ wp,    // [CIN][9][COUT]
               const float* __restrict__ bias,  // [COUT]
               float* __restrict__ out,         // [n_img][COUT][HW][HW]
               float* __restrict__ pooled) {    // [n_img][COUT][HW/2][HW/2] (POOL)
  constexpr int HW = K::HW, CS = K::CS, PS = K::PS, CT = K::CT, TCO = K::TCO, CC = K::CC;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + K::A_FLOATS;

  const int img = blockIdx.y;
  const int band = blockIdx.x % K::NB, ct = blockIdx.x / K::NB;
  const int y0 = band * K::R, co_base = ct * CT;
  const int tid = threadIdx.x;
  const int tc = tid % K::NTC, tp = tid / K::NTC;
  const bool active = tp < K::NPT;
  const int j = active ? tp / K::GPR : 0, g = active ? tp % K::GPR : 0;

  // output-channel PAIRS: the inner loop is FFMA2 (packed fp32 pairs: each component an IEEE fma, bit-identical to the scalar loop, half
  // the issue slots -- the loop competes with its own shared-memory loads for issue)
  float2 acc2[2][4][TCO / 2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int k = 0; k < TCO / 2; ++k) acc2[r][p][k] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int c0 = 0; c0 < K::CIN; c0 += CC) {
    __syncthreads();
    {  // stage inputs (zero padded); chunk lies wholly in src0 or src1
      const float* s;
      int cbase, cn;
      if (c0 < K::CIN0) { s = src0; cbase = c0; cn = K::CIN0; }
      else { s = src1; cbase = c0 - K::CIN0; cn = K::CIN1; }
      const float* sp = s + ((size_t)img * cn + cbase) * (HW * HW);
      for (int e = tid; e < K::A_FLOATS; e += K::NTHREADS) {
        const int c = e / PS, rem = e % PS, r = rem / CS, col = rem % CS;
        const int y = y0 + r - 1, xx = col - 4;
        float v = 0.f;
        if (y >= 0 && y < HW && xx >= 0 && xx < HW) v = __ldg(sp + (c * HW + y) * HW + xx);
        As[e] = v;
      }
      const float* wsrc = wp + (size_t)c0 * 9 * K::COUT + co_base;
      for (int e = tid; e < K::B_FLOATS / 4; e += K::NTHREADS) {
        const int idx = e * 4, cco = idx % CT, ctap = idx / CT;
        st4(Bs + idx, ld4(wsrc + (size_t)ctap * K::COUT + cco));
      }
    }
    __syncthreads();
    if (active) {
      const float* ap0 = As + (2 * j) * CS + 4 * g + 3;
      const float* bp0 = Bs + 4 * tc;
#pragma unroll 2
      for (int c = 0; c < CC; ++c) {
        const float* ap = ap0 + c * PS;
        float a[4][6];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          a[rr][0] = ap[rr * CS];
          const float4 m = ld4(ap + rr * CS + 1);
          a[rr][1] = m.x; a[rr][2] = m.y; a[rr][3] = m.z; a[rr][4] = m.w;
          a[rr][5] = ap[rr * CS + 5];
        }
        const float* bp = bp0 + c * 9 * CT;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            float2 w[TCO / 2];
            {
              const float4 w0 = ld4(bp + (dy * 3 + dx) * CT);
              w[0] = make_float2(w0.x, w0.y); w[1] = make_float2(w0.z, w0.w);
              if constexpr (TCO == 8) {
                const float4 w1 = ld4(bp + (dy * 3 + dx) * CT + CT / 2);
                w[2] = make_float2(w1.x, w1.y); w[3] = make_float2(w1.z, w1.w);
              }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float v0 = a[dy][p + dx], v1 = a[dy + 1][p + dx];
#pragma unroll
              for (int k = 0; k < TCO / 2; ++k) {
                fma2(acc2[0][p][k], w[k], v0);
                fma2(acc2[1][p][k], w[k], v1);
              }
            }
          }
      }
    }
  }

  if (!active) return;
  auto accv = [&](int r, int p, int k) { return (k & 1) ? acc2[r][p][k >> 1].y : acc2[r][p][k >> 1].x; };
  if constexpr (K::EPI == 1) {
#pragma unroll
    for (int k = 0; k < TCO; ++k) {
      const int co = co_base + (k < 4 ? 4 * tc + k : CT / 2 + 4 * tc + (k - 4));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const size_t base = (((size_t)img * K::COUT + co) * HW + y0 + 2 * j + r) * HW + 4 * g;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (4 * g + p < HW) out[base + p] = (!bias || __ldg(bias + base + p) > 0.f) ? accv(r, p, k) : 0.f;
      }
    }
    return;
  }
  // epilogue: bias + ReLU, full-resolution store, optional 2x2 max-pool of the thread's patch
#pragma unroll
  for (int k = 0; k < TCO; ++k) {
    const int co = co_base + (k < 4 ? 4 * tc + k : CT / 2 + 4 * tc + (k - 4));
    const float bv = __ldg(bias + co);
    float v[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int p = 0; p < 4; ++p) v[r][p] = fmaxf(accv(r, p, k) + bv, 0.f);
    float* o = out + ((size_t)img * K::COUT + co) * (HW * HW);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int y = y0 + 2 * j + r;
      if constexpr (HW % 4 == 0) {
        st4(o + y * HW + 4 * g, make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (4 * g + p < HW) o[y * HW + 4 * g + p] = v[r][p];
      }
    }
    if constexpr (K::POOL) {
      constexpr int HP = HW / 2;
      float* po = pooled + ((size_t)img * K::COUT + co) * (HP * HP) + (y0 / 2 + j) * HP;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float m = fmaxf(fmaxf(v[0][2 * q], v[0][2 * q + 1]), fmaxf(v[1][2 * q], v[1][2 * q + 1]));
        if (2 * g + q < HP) po[2 * g + q] = m;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d(k=2, s=2) (unet.py:25-31): out[co][2i+a][2j+b] = bias[co] + sum_ci in[ci][i][j] W[ci][co][a][b]
// -- no overlap, i.e. a per-pixel GEMM with 4*COUT outputs.  Thread = 1x4 input pixels x 4 couts x 4 (a,b).
// Weights are used in the reference's own [ci][co][a][b] layout (16 B per (ci,co)).
template <int HWI_, int CIN_, int COUT_, int R_, int CT_, int CC_>
struct ConvTCfg {
  static constexpr int HWI = HWI_, CIN = CIN_, COUT = COUT_, R = R_, CT = CT_, CC = CC_;
  static constexpr int GPR = (HWI + 3) / 4;
  static constexpr int CS = 4 * GPR;
  static constexpr int NPT = R * GPR;
  static constexpr int NTC = CT / 4;
  static constexpr int NTHREADS = round_up(NPT * NTC, 32);
  static constexpr int NB = HWI / R, NCT = COUT / CT;
  static constexpr int A_FLOATS = CC * R * CS;
  static constexpr int B_FLOATS = CC * CT * 4;
  static constexpr int SMEM_BYTES = (A_FLOATS + B_FLOATS) * 4;
  static_assert(HWI % R == 0 && COUT % CT == 0 && CIN % CC == 0 && CT % 4 == 0 && HWI % 2 == 0, "tiling");
};

// grid (NB*NCT, n_img), block NTHREADS
template <class K>
__global__ void __launch_bounds__(K::NTHREADS)
convT2x2_kernel(const float* __restrict__ src,   // [n_img][CIN][HWI][HWI]
                const float* __restrict__ w,     // [CIN][COUT][2][2]
                const float* __restrict__ bias,  // [COUT]
                float* __restrict__ out) {       // [n_img][COUT][2HWI][2HWI]
  constexpr int HWI = K::HWI, CS = K::CS, CT = K::CT, CC = K::CC, HO = 2 * K::HWI;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + K::A_FLOATS;
  const int img = blockIdx.y;
  const int band = blockIdx.x % K::NB, ct = blockIdx.x / K::NB;
  const int y0 = band * K::R, co_base = ct * CT;
  const int tid = threadIdx.x;
  const int tc = tid % K::NTC, tp = tid / K::NTC;
  const bool active = tp < K::NPT;
  const int r = active ? tp / K::GPR : 0, g = active ? tp % K::GPR : 0;

  float acc[4][4][4];  // [px][co][ab]
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[p][k][q] = 0.f;

#pragma unroll 1
  for (int c0 = 0; c0 < K::CIN; c0 += CC) {
    __syncthreads();
    const float* sp = src + ((size_t)img * K::CIN + c0) * (HWI * HWI);
    for (int e = tid; e < K::A_FLOATS; e += K::NTHREADS) {
      const int c = e / (K::R * CS), rem = e % (K::R * CS), rr = rem / CS, col = rem % CS;
      As[e] = col < HWI ? __ldg(sp + (c * HWI + y0 + rr) * HWI + col) : 0.f;
    }
    const float* wsrc = w + ((size_t)c0 * K::COUT + co_base) * 4;
    for (int e = tid; e < K::B_FLOATS / 4; e += K::NTHREADS) {
      const int c = e / CT, k = e % CT;
      st4(Bs + e * 4, ld4(wsrc + ((size_t)c * K::COUT + k) * 4));
    }
    __syncthreads();
    if (active) {
      const float* ap = As + r * CS + 4 * g;
      const float* bp = Bs + 16 * tc;
#pragma unroll 4
      for (int c = 0; c < CC; ++c) {
        const float4 av = ld4(ap + c * (K::R * CS));
        const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 wv = ld4(bp + c * (CT * 4) + 4 * k);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc[p][k][0] = fmaf(a[p], wv.x, acc[p][k][0]);
            acc[p][k][1] = fmaf(a[p], wv.y, acc[p][k][1]);
            acc[p][k][2] = fmaf(a[p], wv.z, acc[p][k][2]);
            acc[p][k][3] = fmaf(a[p], wv.w, acc[p][k][3]);
          }
        }
      }
    }
  }
  if (!active) return;
  const int i = y0 + r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co_base + 4 * tc + k;
    const float bv = __ldg(bias + co);
    float* o = out + ((size_t)img * K::COUT + co) * (HO * HO);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      float* orow = o + (2 * i + a) * HO + 8 * g;
      if (4 * g + 1 < HWI)
        st4(orow, make_float4(acc[0][k][2 * a] + bv, acc[0][k][2 * a + 1] + bv, acc[1][k][2 * a] + bv, acc[1][k][2 * a + 1] + bv));
      if (4 * g + 3 < HWI)
        st4(orow + 4, make_float4(acc[2][k][2 * a] + bv, acc[2][k][2 * a + 1] + bv, acc[3][k][2 * a] + bv, acc[3][k][2 * a + 1] + bv));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// conv_final: 1x1 conv 32->32 with bias, no activation (unet.py:238), NCHW in -> channels-last out
// (the layout the decoder gathers from: one 128 B line per texel).
constexpr int F_PIX = 64;  // pixels per CTA
// grid (1600/F_PIX, n_img), block 256
__global__ void __launch_bounds__(256)
conv1x1_nhwc_kernel(const float* __restrict__ src,  // [n_img][32][1600]
                    const float* __restrict__ wt,   // [ci][co]
                    const float* __restrict__ bias, float* __restrict__ out) {  // [n_img][1600][32]
  __shared__ __align__(16) float As[C * F_PIX];
  __shared__ __align__(16) float Ws[C * C];
  const int img = blockIdx.y, pix0 = blockIdx.x * F_PIX, tid = threadIdx.x;
  const float* sp = src + (size_t)img * C * G2 + pix0;
  for (int e = tid; e < C * F_PIX; e += 256) As[e] = __ldg(sp + (e / F_PIX) * G2 + (e % F_PIX));
  for (int e = tid; e < C * C; e += 256) Ws[e] = __ldg(wt + e);
  __syncthreads();
  const int cg = tid % 4, p = tid / 4;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = __ldg(bias + cg * 8 + k);
#pragma unroll 8
  for (int ci = 0; ci < C; ++ci) {
    const float a = As[ci * F_PIX + p];
    const float4 w0 = ld4(Ws + ci * C + cg * 8), w1 = ld4(Ws + ci * C + cg * 8 + 4);
    acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
    acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
    acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
    acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
  }
  float* o = out + ((size_t)img * G2 + pix0 + p) * C + cg * 8;
  st4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
  st4(o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

}  // namespace giga
