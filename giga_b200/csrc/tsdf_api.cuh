// C-ABI entry points of the TSDF integration / read-out (kernels in tsdf.cuh); included at the end of giga_api.cu.
#pragma once

extern "C" {

int giga_tsdf_integrate(giga_ctx* ctx, float* tsdf, float* weight, int resolution, double size, double sdf_trunc, const float* depth, int n_views,
                        int width, int height, double fx, double fy, double cx, double cy, const double* extrinsics, double depth_scale,
                        double depth_trunc, void* stream) {
  if (!ctx || !tsdf || !weight || !depth || !extrinsics || resolution < 1 || resolution > 1024 || n_views < 1 || width < 1 || height < 1 ||
      !(size > 0) || !(sdf_trunc > 0) || !(depth_scale > 0))
    return fail(GIGA_EINVAL, "giga_tsdf_integrate: bad argument");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  TsdfCam K;
  K.width = width; K.height = height;
  K.fx = (float)fx; K.fy = (float)fy; K.cx = (float)cx; K.cy = (float)cy;
  K.inv_fx = 1.0f / K.fx; K.inv_fy = 1.0f / K.fy;
  K.voxel_length = (float)(size / resolution);
  K.half = K.voxel_length * 0.5f;
  K.trunc = (float)sdf_trunc;
  K.trunc_inv = 1.0f / K.trunc;
  K.depth_scale = (float)depth_scale; K.depth_trunc = (float)depth_trunc;
  K.safe_w = (float)width - 0.0001f; K.safe_h = (float)height - 0.0001f;
  for (int v0 = 0; v0 < n_views; v0 += TSDF_MAX_VIEWS) {
    TsdfViews V;
    V.n = std::min(TSDF_MAX_VIEWS, n_views - v0);
    for (int v = 0; v < V.n; ++v) {
      const double* E = extrinsics + (size_t)(v0 + v) * 16;
      for (int i = 0; i < 12; ++i) V.E[v][i] = (float)E[i];
      for (int i = 0; i < 3; ++i) {
        volatile float e = V.E[v][4 * i + 2];      // one rounded float multiply (Eigen: extrinsic_f * voxel_length_f)
        V.Es2[v][i] = e * K.voxel_length;
      }
    }
    LaunchScope ls(ctx, "tsdf:integrate", st);
    tsdf_integrate_kernel<<<ceil_div(resolution * resolution, 128), 128, 0, st>>>(tsdf, weight, resolution, depth + (size_t)v0 * width * height, V, K);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_tsdf_grid(giga_ctx* ctx, const float* tsdf, const float* weight, int resolution, float* grid, void* stream) {
  if (!ctx || !tsdf || !weight || !grid || resolution < 1 || resolution > 1024) return fail(GIGA_EINVAL, "giga_tsdf_grid: bad argument");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  const long n = (long)resolution * resolution * resolution;
  {
    LaunchScope ls(ctx, "tsdf:grid", st);
    tsdf_grid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tsdf, weight, grid, n);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

}  // extern "C"
