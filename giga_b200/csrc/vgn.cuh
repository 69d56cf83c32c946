// VGN baseline network (the volumetric grasping CNN GIGA is compared against) -- fp32 CUDA-core 3-D convolutions.
//
// Replaces (reference, relative to src/vgn): networks.py:48-63 `ConvNet.forward`, :172-188 `Encoder` (three stride-2 Conv3d + ReLU:
// 1->16 k5, 16->32 k3, 32->64 k3: 40^3 -> 20^3 -> 10^3 -> 5^3), :191-212 `Decoder` (Conv3d 64->64 k3 @5^3, F.interpolate(10) nearest,
// 64->32 k3 @10^3, interpolate(20), 32->16 k5 @20^3, interpolate(40)) and the three k5 heads at 40^3 with their epilogues
// (sigmoid(qual), F.normalize(rot, dim=1), raw width).  ATen semantics: cross-correlation, zero padding k/2, stride-2 output o reads
// inputs 2o + d - k/2, nearest interpolation to an exact multiple = source index dst >> 1.
// One generic kernel: a CTA computes 128 consecutive output voxels (z fastest) of one scene for ALL output channels; the weights of
// CI_CH input channels at a time are staged in shared memory as [ci][tap][co] and read as warp-uniform 128-bit broadcasts, the inputs
// come through L1 (neighbouring threads and taps re-read the same lines).  The nearest up-sampling in front of a conv is an index map of
// its input reads (the up-sampled tensors are never written).  This is the second model family on the conv path,
// not the headline kernel: fp32 FMA, ~50 % of the FMA pipe at best.
#pragma once
#include "common.cuh"

namespace giga {

// EPI 0: y = relu(conv + b) -> [B][COUT][DOUT^3];  EPI 1 (heads, COUT = 8: qual, rot x4, width, 2 unused): qual [B][DOUT^3] = sigmoid,
// rot [B][DOUT^3][4] = normalised, width [B][DOUT^3]
template <int CIN, int COUT, int K, int STRIDE, bool UPIN, int DIN, int DOUT, int CI_CH, int EPI>
__global__ void __launch_bounds__(128) conv3d_kernel(const float* __restrict__ x,     // [B][CIN][DIN^3]
                                                    const float* __restrict__ w,     // [CIN][K^3][COUT]
                                                    const float* __restrict__ bias,  // [COUT]
                                                    float* __restrict__ y, float* __restrict__ y_rot, float* __restrict__ y_width) {
  constexpr int K3 = K * K * K, PAD = K / 2;
  constexpr int DLOG = UPIN ? 2 * DIN : DIN;   // logical input resolution (after the nearest up-sampling)
  static_assert(DLOG / STRIDE == DOUT && COUT % 4 == 0 && CIN % CI_CH == 0, "shapes");
  __shared__ __align__(16) float sw[CI_CH * K3 * COUT];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int o = blockIdx.x * 128 + tid;
  const bool valid = o < DOUT * DOUT * DOUT;
  const int oc = valid ? o : DOUT * DOUT * DOUT - 1;
  const int oz = oc % DOUT, oy = (oc / DOUT) % DOUT, ox = oc / (DOUT * DOUT);
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = __ldg(bias + c);
  const float* xb = x + (size_t)b * CIN * DIN * DIN * DIN;
#pragma unroll 1
  for (int ci0 = 0; ci0 < CIN; ci0 += CI_CH) {
    __syncthreads();
    for (int e = tid; e < CI_CH * K3 * COUT / 4; e += 128)
      reinterpret_cast<float4*>(sw)[e] = __ldg(reinterpret_cast<const float4*>(w + (size_t)ci0 * K3 * COUT) + e);
    __syncthreads();
#pragma unroll 1
    for (int cc = 0; cc < CI_CH; ++cc) {
      const float* xc = xb + (size_t)(ci0 + cc) * DIN * DIN * DIN;
      const float* wc = sw + cc * K3 * COUT;
#pragma unroll 1
      for (int dx = 0; dx < K; ++dx) {
        const int lx = ox * STRIDE + dx - PAD;
        if (lx < 0 || lx >= DLOG) continue;
        const int ix = UPIN ? lx >> 1 : lx;
#pragma unroll 1
        for (int dy = 0; dy < K; ++dy) {
          const int ly = oy * STRIDE + dy - PAD;
          if (ly < 0 || ly >= DLOG) continue;
          const int iy = UPIN ? ly >> 1 : ly;
          const float* xr = xc + ((size_t)ix * DIN + iy) * DIN;
          const float* wr = wc + ((dx * K + dy) * K) * COUT;
#pragma unroll
          for (int dz = 0; dz < K; ++dz) {
            const int lz = oz * STRIDE + dz - PAD;
            const bool in = lz >= 0 && lz < DLOG;
            const float v = in ? __ldg(xr + (UPIN ? lz >> 1 : lz)) : 0.f;
#pragma unroll
            for (int c4 = 0; c4 < COUT / 4; ++c4) {
              const float4 q = *reinterpret_cast<const float4*>(wr + dz * COUT + 4 * c4);
              acc[4 * c4 + 0] = fmaf(q.x, v, acc[4 * c4 + 0]);
              acc[4 * c4 + 1] = fmaf(q.y, v, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(q.z, v, acc[4 * c4 + 2]);
              acc[4 * c4 + 3] = fmaf(q.w, v, acc[4 * c4 + 3]);
            }
          }
        }
      }
    }
  }
  if (!valid) return;
  constexpr int V = DOUT * DOUT * DOUT;
  if (EPI == 0) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) y[((size_t)b * COUT + c) * V + o] = fmaxf(acc[c], 0.f);
  } else {
    y[(size_t)b * V + o] = 1.f / (1.f + expf(-acc[0]));                                    // torch.sigmoid(conv_qual(x))
    const float nrm = sqrtf(acc[1] * acc[1] + acc[2] * acc[2] + acc[3] * acc[3] + acc[4] * acc[4]);
    const float d = fmaxf(nrm, 1e-12f);                                                    // F.normalize(dim=1, eps=1e-12)
    st4(y_rot + ((size_t)b * V + o) * 4, make_float4(acc[1] / d, acc[2] / d, acc[3] / d, acc[4] / d));
    y_width[(size_t)b * V + o] = acc[5];
  }
}

// packed VGN parameter blob (floats): per layer weights [CIN][K^3][COUT] then bias [COUT]
struct VgnLayout {
  long w[7], b[7], total;
};
constexpr int kVgnCin[7] = {1, 16, 32, 64, 64, 32, 16};
constexpr int kVgnCout[7] = {16, 32, 64, 64, 32, 16, 8};   // the three heads are one 8-channel layer (qual, rot x4, width, 2 zero)
constexpr int kVgnK[7] = {5, 3, 3, 3, 3, 5, 5};
inline VgnLayout make_vgn_layout() {
  VgnLayout L;
  long o = 0;
  for (int i = 0; i < 7; ++i) {
    L.w[i] = o; o += (long)kVgnCin[i] * kVgnK[i] * kVgnK[i] * kVgnK[i] * kVgnCout[i];
    L.b[i] = o; o += kVgnCout[i];
    o = (o + 3) / 4 * 4;   // 16-byte aligned weight blocks (float4 staging)
  }
  L.total = o;
  return L;
}
// activation workspace per scene (floats): e1 16x20^3, e2 32x10^3, e3 64x5^3, d1 64x5^3, d2 32x10^3, d3 16x20^3
constexpr long kVgnAct[6] = {16L * 8000, 32L * 1000, 64L * 125, 64L * 125, 32L * 1000, 16L * 8000};

}  // namespace giga
