// Planner post-processing on the device: what
// /root/reference/src/vgn/detection_implicit.py does on the host with scipy.ndimage between predict()
// and the Grasp list -- process() :115-143, bound() :87-97, select() :146-174 -- for B scenes at once,
// so that only the surviving grasps (a few dozen x 7 floats) cross PCIe instead of 64,000 x 6 floats.
//
// Volumes are [B][40][40][40] with flat voxel index v = (ix*40 + iy)*40 + iz = index of the lattice
// query point (VGNImplicit.__init__ :28-31 meshgrid order).  HBM-bound integer/float work: one thread per
// voxel, neighbours through L1/L2 (a scene's volume is 256 KB), no tensor cores.
//
// Bit-exactness with scipy.ndimage (its algorithms are restated by the CPU checker used in tests/):
//   * gaussian_filter: three separable passes (axis 0, 1, 2), each accumulated in fp64 in scipy's
//     symmetric correlate1d order -- centre*w[r], then for ii = -r..-1: += (x[l+ii] + x[l-ii])*w[ii+r] --
//     with explicit round-to-nearest mul/add (no FMA contraction) and rounded to fp32 between passes;
//     mode "nearest" = clamped indices; weights computed on the host in fp64 (gaussian_kernel1d);
//   * binary_dilation(outside, iterations=2, mask=~inside), 6-connected, border 0;
//   * maximum_filter(size=s), mode "reflect", window offsets -(s/2) .. s-1-s/2.
#pragma once
#include "common.cuh"

namespace giga {

constexpr int PL_MAXR = 8;   // gaussian radius limit (sigma <= 2 with truncate 4)

struct SelectParams {        // device copy of giga_select_params (+ the precomputed fp64 kernel)
  double w[2 * PL_MAXR + 1];
  int radius;
  float min_width, max_width, out_th;
  int lim_x, lim_y, lim_z;
  float low_th, threshold;
  int force_detection, max_filter_size;
};

// One separable gaussian pass along AXIS (0 = ix, 1 = iy, 2 = iz).  grid ceil(B*64000/256), block 256.
template <int AXIS>
__device__ __forceinline__ float gauss_tap(const float* __restrict__ vol, int ix, int iy, int iz, const SelectParams& P) {
  constexpr int stride = AXIS == 0 ? G2 : (AXIS == 1 ? G : 1);
  const int l = AXIS == 0 ? ix : (AXIS == 1 ? iy : iz);
  const float* line = vol + ((ix * G + iy) * G + iz) - l * stride;
  const int r = P.radius;
  double tmp = __dmul_rn((double)line[l * stride], P.w[r]);
  for (int ii = -r; ii < 0; ++ii) {
    const int a = max(l + ii, 0), b = min(l - ii, G - 1);
    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)line[a * stride], (double)line[b * stride]), P.w[ii + r]));
  }
  return (float)tmp;
}

template <int AXIS>
__global__ void __launch_bounds__(256)
gauss_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, const SelectParams P) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= B * G3) return;
  const int b = t / G3, v = t - b * G3;
  const int ix = v / G2, iy = (v / G) % G, iz = v % G;
  dst[t] = gauss_tap<AXIS>(src + (size_t)b * G3, ix, iy, iz, P);
}

__device__ __forceinline__ bool pl_outside(const float* __restrict__ tsdf, int ix, int iy, int iz, float out_th) {
  if ((unsigned)ix >= (unsigned)G || (unsigned)iy >= (unsigned)G || (unsigned)iz >= (unsigned)G) return false;   // border_value 0
  return tsdf[(ix * G + iy) * G + iz] > out_th;
}
__device__ __forceinline__ bool pl_mask(const float* __restrict__ tsdf, int ix, int iy, int iz, float out_th) {   // ~inside
  const float t = tsdf[(ix * G + iy) * G + iz];
  return !(1e-3f < t && t < out_th);
}
// first dilation iteration at u (u inside the volume)
__device__ __forceinline__ bool pl_d1(const float* __restrict__ tsdf, int ix, int iy, int iz, float th) {
  if (!pl_mask(tsdf, ix, iy, iz, th)) return false;   // inside voxels are never modified and start false
  return pl_outside(tsdf, ix, iy, iz, th) || pl_outside(tsdf, ix - 1, iy, iz, th) || pl_outside(tsdf, ix + 1, iy, iz, th) ||
         pl_outside(tsdf, ix, iy - 1, iz, th) || pl_outside(tsdf, ix, iy + 1, iz, th) || pl_outside(tsdf, ix, iy, iz - 1, th) ||
         pl_outside(tsdf, ix, iy, iz + 1, th);
}
__device__ __forceinline__ bool pl_d1b(const float* __restrict__ tsdf, int ix, int iy, int iz, float th) {
  if ((unsigned)ix >= (unsigned)G || (unsigned)iy >= (unsigned)G || (unsigned)iz >= (unsigned)G) return false;
  return pl_d1(tsdf, ix, iy, iz, th);
}

// Third gaussian pass (axis 2) fused with process()'s masks, bound() and select()'s LOW_TH:
//   qvol  (optional) the processed quality volume, i.e. what process()+bound() return
//   qlow  the same with values < LOW_TH zeroed (input of the NMS)
//   flag[b] |= any(qlow >= threshold)          (force_detection's test, detection_implicit.py:149)
__global__ void __launch_bounds__(256)
gauss_z_mask_kernel(const float* __restrict__ src, const float* __restrict__ tsdf, const float* __restrict__ width,
                    float* __restrict__ qvol, float* __restrict__ qlow, int* __restrict__ flag, int B, const SelectParams P) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  bool above = false;
  int b = 0;
  if (t < B * G3) {
    b = t / G3;
    const int v = t - b * G3;
    const int ix = v / G2, iy = (v / G) % G, iz = v % G;
    float q = gauss_tap<2>(src + (size_t)b * G3, ix, iy, iz, P);
    const float* ts = tsdf + (size_t)b * G3;
    const float th = P.out_th;
    bool valid = false;
    if (pl_mask(ts, ix, iy, iz, th))
      valid = pl_d1(ts, ix, iy, iz, th) || pl_d1b(ts, ix - 1, iy, iz, th) || pl_d1b(ts, ix + 1, iy, iz, th) ||
              pl_d1b(ts, ix, iy - 1, iz, th) || pl_d1b(ts, ix, iy + 1, iz, th) || pl_d1b(ts, ix, iy, iz - 1, th) ||
              pl_d1b(ts, ix, iy, iz + 1, th);
    if (!valid) q = 0.f;
    const float w = width[t];
    if (w < P.min_width || w > P.max_width) q = 0.f;
    // bound(): qual_vol[:lim] = 0, qual_vol[-lim:] = 0 (numpy: lim == 0 makes [-0:] the whole axis), z only from below
    if (ix < P.lim_x || ix >= G - P.lim_x || P.lim_x == 0) q = 0.f;
    if (iy < P.lim_y || iy >= G - P.lim_y || P.lim_y == 0) q = 0.f;
    if (iz < P.lim_z) q = 0.f;
    if (qvol) qvol[t] = q;
    if (q < P.low_th) q = 0.f;
    qlow[t] = q;
    above = q >= P.threshold;
  }
  // a block never straddles more than two scenes (64000 = 250 * 256: it straddles none)
  if (__syncthreads_or(above) && threadIdx.x == 0) atomicOr(&flag[b], 1);
}

// Non-maximum suppression + candidate compaction.  cand_* are [B][64000]; count[b] via atomics (the order of
// the candidate list is arbitrary; grasp_rank_kernel makes the output deterministic).
__global__ void __launch_bounds__(256)
grasp_nms_kernel(const float* __restrict__ qlow, const int* __restrict__ flag, int* __restrict__ count,
                 float* __restrict__ cand_val, int* __restrict__ cand_idx, int B, const SelectParams P) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= B * G3) return;
  const int b = t / G3, v = t - b * G3;
  const bool best_only = P.force_detection && !flag[b];
  const float thr = best_only ? 0.f : P.threshold;     // values are >= 0 (smoothed sigmoid outputs): "< 0" never fires
  float q = qlow[t];
  if (q < thr) q = 0.f;
  if (q == 0.f) return;                                // where(q == max, q, 0) is 0 either way
  const float* vol = qlow + (size_t)b * G3;
  const int ix = v / G2, iy = (v / G) % G, iz = v % G;
  const int s = P.max_filter_size, lo = s / 2;
  float m = q;
  for (int dx = -lo; dx < s - lo; ++dx) {
    int x = ix + dx;
    x = x < 0 ? -x - 1 : (x >= G ? 2 * G - 1 - x : x);   // reflect: (d c b a | a b c d | d c b a)
    for (int dy = -lo; dy < s - lo; ++dy) {
      int y = iy + dy;
      y = y < 0 ? -y - 1 : (y >= G ? 2 * G - 1 - y : y);
      const float* row = vol + (x * G + y) * G;
      for (int dz = -lo; dz < s - lo; ++dz) {
        int z = iz + dz;
        z = z < 0 ? -z - 1 : (z >= G ? 2 * G - 1 - z : z);
        float o = row[z];
        if (o < thr) o = 0.f;
        m = fmaxf(m, o);
      }
    }
  }
  if (q == m) {
    const int slot = atomicAdd(&count[b], 1);
    cand_val[(size_t)b * G3 + slot] = q;
    cand_idx[(size_t)b * G3 + slot] = v;
  }
}

// Descending sort by rank counting (ties: larger voxel index first = reversed stable argsort) and gather of
// the grasp parameters.  grid (ceil(64000/256), B): CTA y handles candidates [256 y, 256 y + 256) of scene b
// and exits at once when there are none.  Outputs [B][K]; n_out[b] = number of grasps found (may exceed K).
__global__ void __launch_bounds__(256)
grasp_rank_kernel(const int* __restrict__ count, const int* __restrict__ flag, const float* __restrict__ cand_val,
                  const int* __restrict__ cand_idx, const float* __restrict__ rot, const float* __restrict__ width, int K,
                  int* __restrict__ n_out, float* __restrict__ score, int* __restrict__ index, float* __restrict__ out_rot,
                  float* __restrict__ out_width, const SelectParams P) {
  __shared__ float sv[256];
  __shared__ int si[256];
  const int b = blockIdx.y, n = count[b];
  const bool best_only = P.force_detection && !flag[b];
  if (blockIdx.x == 0 && threadIdx.x == 0) n_out[b] = best_only ? min(n, 1) : n;
  if ((int)blockIdx.x * 256 >= n) return;
  const float* cv = cand_val + (size_t)b * G3;
  const int* ci = cand_idx + (size_t)b * G3;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float vi = i < n ? cv[i] : 0.f;
  const int xi = i < n ? ci[i] : 0;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    sv[threadIdx.x] = j < n ? cv[j] : -1.f;
    si[threadIdx.x] = j < n ? ci[j] : -1;
    __syncthreads();
    const int m = min(256, n - j0);
    for (int k = 0; k < m; ++k) rank += (sv[k] > vi) || (sv[k] == vi && si[k] > xi);
  }
  const int kmax = best_only ? 1 : K;
  if (i < n && rank < kmax) {
    const size_t o = (size_t)b * K + rank;
    score[o] = vi;
    index[o] = xi;
    const float4 r = ld4(rot + ((size_t)b * G3 + xi) * 4);
    st4(out_rot + o * 4, r);
    out_width[o] = width[(size_t)b * G3 + xi];
  }
}

// lattice [N][3] -> [B][N][3] (the decoder kernels take per-scene points)
__global__ void __launch_bounds__(256)
broadcast_points_kernel(const float* __restrict__ src, float* __restrict__ dst, int n3, int B) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n3) return;
  const float v = src[t];
  for (int b = 0; b < B; ++b) dst[(size_t)b * n3 + t] = v;
}

}  // namespace giga
