// Fused implicit decoder: tri-plane bilinear gather + per-point ResNet-MLP heads + head epilogues.
//
// Replaces (reference, relative to src/vgn/ConvONets):
//   conv_onet/models/decoder.py:117-122  sample_plane_feature (normalize_coordinate common.py:238-261,
//                                        F.grid_sample bilinear / border / align_corners=True) x3 planes
//                                        -- recomputed per head in the reference, done ONCE per point here
//   conv_onet/models/decoder.py:160-176  fc_p, 5 x (fc_c[i] + ResnetBlockFC), fc_out
//   layers.py:39-47                      ResnetBlockFC: x + fc_1(relu(fc_0(relu(x))))
//   conv_onet/models/__init__.py:119-123 sigmoid(qual), F.normalize(rot), raw width / occupancy logit
// fp32 FMA-pipe version (the exact-parity path).  A CTA owns DEC_PTS points of one scene:
//   phase 1  each warp gathers its 32 points cooperatively: lane = channel, so every texel is one
//            coalesced 128 B read of the channels-last plane; 12 texels -> 3 x 32 features per point,
//            parked in shared memory as feat[96][DEC_PTS] (never written to HBM);
//   phase 2  thread = point.  The 32-wide hidden state lives in registers; each block's weights
//            (fc_c 96x32, fc_0, fc_1 32x32, biases = 20.9 KB) are staged in shared memory once per CTA
//            and read as warp-uniform 128-bit broadcasts.
#pragma once
#include "common.cuh"

namespace giga {

// packed per-head parameter blob (floats); all matrices stored input-major: Wt[k][j] = W[j][k]
constexpr int DW_FCP = 0;                      // Wt[3][32] + b[32]
constexpr int DW_BLOCK0 = 128;                 // 5 blocks
constexpr int DW_BLK_FCC = 0;                  //   Wt[96][32]
constexpr int DW_BLK_BC = 3072;                //   b[32]
constexpr int DW_BLK_W0 = 3104;                //   Wt[32][32]
constexpr int DW_BLK_B0 = 4128;                //   b[32]
constexpr int DW_BLK_W1 = 4160;                //   Wt[32][32]
constexpr int DW_BLK_B1 = 5184;                //   b[32]
constexpr int DW_BLK = 5216;
constexpr int DW_OUT = DW_BLOCK0 + 5 * DW_BLK; // 26208: Wt[32][4] (zero padded) + b[4]
constexpr int DW_HEAD = DW_OUT + 132;          // 26340 floats per head

constexpr int DEC_PTS = 128;
constexpr int DEC_FSTRIDE = DEC_PTS + 1;
constexpr int DEC_SMEM_FLOATS = 96 * DEC_FSTRIDE + DW_BLK + DEC_PTS * 24;
constexpr int DEC_SMEM_BYTES = DEC_SMEM_FLOATS * 4;  // 82,688 B

struct TexInfo {  // per point, per plane: 4 texel offsets (floats, within the scene's plane) + 4 weights
  int off[3][4];
  float w[3][4];
};

// common.py:253-260 with padding = 0: t = v / (1 + 0 + 10e-6) + 0.5, one-sided clamps
__device__ __forceinline__ float normalize_axis(float v) {
  float t = __fdiv_rn(v, 1.00001f) + 0.5f;
  if (t >= 1.f) t = 0.99999f;  // 1 - 10e-6
  if (t < 0.f) t = 0.f;
  return t;
}

// decoder.py:119-121 + ATen grid_sampler (bilinear, border, align_corners=True) on a 40x40 plane.
// u -> column (W), v -> row (H).
__device__ __forceinline__ void bilinear_taps(float u, float v, int off[4], float w[4]) {
  const float gx = 2.0f * u - 1.0f, gy = 2.0f * v - 1.0f;
  float ix = (gx + 1.f) * 19.5f, iy = (gy + 1.f) * 19.5f;  // ((g+1)/2)*(40-1)
  ix = fminf(fmaxf(ix, 0.f), 39.f);
  iy = fminf(fmaxf(iy, 0.f), 39.f);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int x1 = min(x0 + 1, G - 1), y1 = min(y0 + 1, G - 1);  // out-of-range corner has weight 0
  const float tx = ix - x0f, ty = iy - y0f;
  const float sx = 1.f - tx, sy = 1.f - ty;  // ATen CPU kernel: e = 1 - w, n = 1 - s
  off[0] = (y0 * G + x0) * C; w[0] = sx * sy;  // nw
  off[1] = (y0 * G + x1) * C; w[1] = tx * sy;  // ne
  off[2] = (y1 * G + x0) * C; w[2] = sx * ty;  // sw
  off[3] = (y1 * G + x1) * C; w[3] = tx * ty;  // se
}

__device__ __forceinline__ void point_taps(const float* __restrict__ p, TexInfo& t) {
  const float nx = normalize_axis(p[0]), ny = normalize_axis(p[1]), nz = normalize_axis(p[2]);
  bilinear_taps(nx, nz, t.off[0], t.w[0]);  // xz: u = x, v = z
  bilinear_taps(nx, ny, t.off[1], t.w[1]);  // xy: u = x, v = y
  bilinear_taps(ny, nz, t.off[2], t.w[2]);  // yz: u = y, v = z
}

// Gather the 3x32 features of the warp's 32 points into feat[k][pt] (k = plane*32 + channel).
// tinfo: this warp's scratch, [32 points][24] words.
__device__ __forceinline__ void warp_gather(const float* __restrict__ planes, int B, int b, const float* __restrict__ pts,
                                            int n0, int N, float* feat, int fstride, int pt0, float* tinfo_w) {
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
      // ATen order: nw*w_nw + ne*w_ne + sw*w_sw + se*w_se
      const float f = v0 * w.x + v1 * w.y + v2 * w.z + v3 * w.w;
      feat[(pl * 32 + lane) * fstride + pt0 + q] = f;
    }
  }
  __syncwarp();
}

// grid (ceil(N/DEC_PTS), B), block DEC_PTS, dynamic smem DEC_SMEM_BYTES
__global__ void __launch_bounds__(DEC_PTS)
decode_points_kernel(const float* __restrict__ planes,  // [3][B][40][40][32]
                     const float* __restrict__ pts,     // [B][N][3]
                     const float* __restrict__ hw,      // [4][DW_HEAD] packed head parameters
                     int B, int N, unsigned heads,
                     float* __restrict__ qual, float* __restrict__ rot, float* __restrict__ width,
                     float* __restrict__ occ) {
  extern __shared__ __align__(16) float smem[];
  float* feat = smem;                          // [96][DEC_FSTRIDE]
  float* wbuf = feat + 96 * DEC_FSTRIDE;       // [DW_BLK]
  float* tinfo = wbuf + DW_BLK;                // [DEC_PTS][24]
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * DEC_PTS;
  const int n = n0 + tid;
  const bool valid = n < N;
  const int nc = valid ? n : N - 1;

  warp_gather(planes, B, b, pts, n0 + warp * 32, N, feat, DEC_FSTRIDE, warp * 32, tinfo + warp * 32 * 24);

  const float* pp = pts + ((size_t)b * N + nc) * 3;
  const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);

#pragma unroll 1
  for (int head = 0; head < 4; ++head) {
    if (!(heads & (1u << head))) continue;
    const float* W = hw + (size_t)head * DW_HEAD;
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      h[j] = __ldg(W + DW_FCP + 96 + j) + __ldg(W + DW_FCP + j) * px + __ldg(W + DW_FCP + 32 + j) * py +
             __ldg(W + DW_FCP + 64 + j) * pz;

#pragma unroll 1
    for (int blk = 0; blk < 5; ++blk) {
      __syncthreads();  // wbuf free (also orders the gather before the first use of feat)
      const float* Wb = W + DW_BLOCK0 + blk * DW_BLK;
      for (int e = tid; e < DW_BLK / 4; e += DEC_PTS) st4(wbuf + e * 4, ld4(Wb + e * 4));
      __syncthreads();
      // net = net + fc_c[blk](c)
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] += wbuf[DW_BLK_BC + j];
#pragma unroll 4
      for (int k = 0; k < 96; ++k) {
        const float f = feat[k * DEC_FSTRIDE + tid];
        const float4* wr = reinterpret_cast<const float4*>(wbuf + DW_BLK_FCC + k * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = wr[j4];
          h[4 * j4 + 0] = fmaf(w.x, f, h[4 * j4 + 0]);
          h[4 * j4 + 1] = fmaf(w.y, f, h[4 * j4 + 1]);
          h[4 * j4 + 2] = fmaf(w.z, f, h[4 * j4 + 2]);
          h[4 * j4 + 3] = fmaf(w.w, f, h[4 * j4 + 3]);
        }
      }
      // ResnetBlockFC
      float t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = wbuf[DW_BLK_B0 + j];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float r = fmaxf(h[k], 0.f);
        const float4* wr = reinterpret_cast<const float4*>(wbuf + DW_BLK_W0 + k * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = wr[j4];
          t[4 * j4 + 0] = fmaf(w.x, r, t[4 * j4 + 0]);
          t[4 * j4 + 1] = fmaf(w.y, r, t[4 * j4 + 1]);
          t[4 * j4 + 2] = fmaf(w.z, r, t[4 * j4 + 2]);
          t[4 * j4 + 3] = fmaf(w.w, r, t[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] += wbuf[DW_BLK_B1 + j];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float r = fmaxf(t[k], 0.f);
        const float4* wr = reinterpret_cast<const float4*>(wbuf + DW_BLK_W1 + k * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = wr[j4];
          h[4 * j4 + 0] = fmaf(w.x, r, h[4 * j4 + 0]);
          h[4 * j4 + 1] = fmaf(w.y, r, h[4 * j4 + 1]);
          h[4 * j4 + 2] = fmaf(w.z, r, h[4 * j4 + 2]);
          h[4 * j4 + 3] = fmaf(w.w, r, h[4 * j4 + 3]);
        }
      }
    }
    // fc_out(relu(net))
    float o[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) o[m] = __ldg(W + DW_OUT + 128 + m);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float r = fmaxf(h[k], 0.f);
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + DW_OUT + k * 4));
      o[0] = fmaf(w.x, r, o[0]); o[1] = fmaf(w.y, r, o[1]);
      o[2] = fmaf(w.z, r, o[2]); o[3] = fmaf(w.w, r, o[3]);
    }
    if (valid) {
      const size_t idx = (size_t)b * N + n;
      const bool raw = (heads & 16u) != 0;                          // GIGA_HEAD_RAW
      if (head == 0) {
        qual[idx] = raw ? o[0] : 1.f / (1.f + expf(-o[0]));         // torch.sigmoid
      } else if (head == 1) {
        const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
        const float d = raw ? 1.f : fmaxf(nrm, 1e-12f);             // F.normalize(dim=2, eps=1e-12)
        st4(rot + idx * 4, make_float4(o[0] / d, o[1] / d, o[2] / d, o[3] / d));
      } else if (head == 2) {
        width[idx] = o[0];
      } else {
        occ[idx] = o[0];
      }
    }
  }
}

constexpr int SF_SMEM_BYTES = (96 * DEC_FSTRIDE + DEC_PTS * 24) * 4;  // 61,824 B
// Feature sampling only (query_feature / the concat feature): grid (ceil(N/128), B), block 128, smem SF_SMEM_BYTES.
// mode 0: out[b][n][96] (xz|xy|yz)   mode 1: out[b][n][32] = sum of the three planes
__global__ void __launch_bounds__(DEC_PTS)
sample_feature_kernel(const float* __restrict__ planes, const float* __restrict__ pts, int B, int N, int mode,
                      float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* feat = smem;                       // [96][DEC_FSTRIDE]
  float* tinfo = smem + 96 * DEC_FSTRIDE;   // [DEC_PTS][24]
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * DEC_PTS;
  warp_gather(planes, B, b, pts, n0 + warp * 32, N, feat, DEC_FSTRIDE, warp * 32, tinfo + warp * 32 * 24);
  __syncthreads();
  const int npts = min(DEC_PTS, N - n0);
  if (mode == 0) {
    for (int e = tid; e < npts * 96; e += DEC_PTS) {
      const int q = e / 96, k = e % 96;
      out[((size_t)b * N + n0 + q) * 96 + k] = feat[k * DEC_FSTRIDE + q];
    }
  } else {
    for (int e = tid; e < npts * 32; e += DEC_PTS) {
      const int q = e / 32, k = e % 32;
      out[((size_t)b * N + n0 + q) * 32 + k] =
          feat[k * DEC_FSTRIDE + q] + feat[(32 + k) * DEC_FSTRIDE + q] + feat[(64 + k) * DEC_FSTRIDE + q];
    }
  }
}

// Per-scene max / first arg-max of the grasp quality: grid B, block 256.
__global__ void __launch_bounds__(256)
scene_argmax_kernel(const float* __restrict__ qual, int N, float* __restrict__ best_val, int* __restrict__ best_idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  float v = -INFINITY;
  int i = 0x7fffffff;
  for (int n = tid; n < N; n += 256) {
    const float q = qual[(size_t)b * N + n];
    if (q > v || (q == v && n < i)) { v = q; i = n; }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, v, s);
    const int oi = __shfl_down_sync(0xffffffffu, i, s);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = v; si[tid >> 5] = i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > v || (sv[w] == v && si[w] < i)) { v = sv[w]; i = si[w]; }
    best_val[b] = v;
    best_idx[b] = i;
  }
}

}  // namespace giga
