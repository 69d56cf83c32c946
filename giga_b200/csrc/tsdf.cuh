// TSDF integration of depth images into a uniform voxel volume, and the read-out of the network's input grid.  SURVEY.md 8f rank 2.
//
// Replaces (reference): vgn/perception.py:65-115 `TSDFVolume` -- `integrate` (:79-104) = open3d.pipelines.integration.UniformTSDFVolume
// .integrate on an RGBD image made with depth_scale = 1, depth_trunc = 2 (Open3D 0.12.0: UniformTSDFVolume.cpp
// IntegrateWithDepthToCameraDistanceMultiplier, ImageFactory.cpp ConvertDepthToFloatImage / CreateDepthToCameraDistanceMultiplierFloatImage)
// and `get_grid` (:106-115) = ExtractVoxelGrid + a Python loop over the voxels ("very slow (~35 ms / 50 ms of the whole pipeline)", :107).
// One launch integrates ALL views of a scene: a thread owns one (x, y) voxel column, transforms its first voxel centre once per view and
// marches along z by adding the scaled third column of the extrinsic -- the same float operations in the same order as Open3D's loop
// (explicit round-to-nearest mul / add / div: no fused multiply-add), views in submission order, so the running averages are reproduced
// exactly.  HBM-bound byte work: 2 x 4 B per voxel per view, the depth image is read through L2.
#pragma once
#include "common.cuh"

namespace giga {

constexpr int TSDF_MAX_VIEWS = 16;   // per launch
struct TsdfViews {
  int n;
  float E[TSDF_MAX_VIEWS][12];       // float extrinsic, rows 0..2
  float Es2[TSDF_MAX_VIEWS][3];      // (extrinsic * voxel_length) third column
};
struct TsdfCam {
  int width, height;
  float fx, fy, cx, cy, inv_fx, inv_fy;
  float voxel_length, half, trunc, trunc_inv, depth_scale, depth_trunc, safe_w, safe_h;
};

// grid ceil(R * R / 128), block 128
__global__ void __launch_bounds__(128)
tsdf_integrate_kernel(float* __restrict__ tsdf, float* __restrict__ weight,   // [R][R][R] (x, y, z)
                      int R, const float* __restrict__ depth,                  // [n][H][W]
                      const __grid_constant__ TsdfViews V, const __grid_constant__ TsdfCam K) {
  const int col = blockIdx.x * 128 + threadIdx.x;
  if (col >= R * R) return;
  const int x = col / R, y = col % R;
  const float px = __fadd_rn(K.half, __fmul_rn(K.voxel_length, (float)x));
  const float py = __fadd_rn(K.half, __fmul_rn(K.voxel_length, (float)y));
  const float pz = K.half;
  float* tv = tsdf + (size_t)col * R;
  float* wv = weight + (size_t)col * R;
  for (int v = 0; v < V.n; ++v) {
    const float* E = V.E[v];
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      c[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(E[4 * i], px), __fmul_rn(E[4 * i + 1], py)), __fmul_rn(E[4 * i + 2], pz)), E[4 * i + 3]);
    const float* dimg = depth + (size_t)v * K.width * K.height;
    for (int z = 0; z < R; ++z) {
      const float X = c[0], Y = c[1], Z = c[2];
      c[0] = __fadd_rn(c[0], V.Es2[v][0]); c[1] = __fadd_rn(c[1], V.Es2[v][1]); c[2] = __fadd_rn(c[2], V.Es2[v][2]);
      if (!(Z > 0.f)) continue;
      const float u_f = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(X, K.fx), Z), K.cx), 0.5f);
      const float v_f = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(Y, K.fy), Z), K.cy), 0.5f);
      if (!(u_f >= 0.0001f && u_f < K.safe_w && v_f >= 0.0001f && v_f < K.safe_h)) continue;
      const int u = (int)u_f, vv = (int)v_f;
      float d = __fdiv_rn(__ldg(dimg + (size_t)vv * K.width + u), K.depth_scale);
      if (d >= K.depth_trunc) d = 0.f;
      if (!(d > 0.f)) continue;
      const float xx = __fmul_rn((float)u - K.cx, K.inv_fx), yy = __fmul_rn((float)vv - K.cy, K.inv_fy);
      const float mult = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xx, xx), __fmul_rn(yy, yy)), 1.0f));
      const float sdf = __fmul_rn(__fsub_rn(d, Z), mult);
      if (sdf > -K.trunc) {
        const float t = fminf(1.0f, __fmul_rn(sdf, K.trunc_inv));
        const float w = wv[z];
        tv[z] = __fdiv_rn(__fadd_rn(__fmul_rn(tv[z], w), t), __fadd_rn(w, 1.0f));
        wv[z] = __fadd_rn(w, 1.0f);
      }
    }
  }
}

// get_grid: grid[x][y][z] = (w != 0 && -0.98 <= f < 0.98) ? float((f + 1.0) * 0.5) : 0
__global__ void __launch_bounds__(256)
tsdf_grid_kernel(const float* __restrict__ tsdf, const float* __restrict__ weight, float* __restrict__ grid, long n) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float f = tsdf[i], w = weight[i];
  grid[i] = (w != 0.0f && f < 0.98f && f >= -0.98f) ? (float)(((double)f + 1.0) * 0.5) : 0.0f;
}

}  // namespace giga
