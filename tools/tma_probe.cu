// TMA tensor-map probe (not product): 4-D map over x[b][ix][iy][iz] fp32, box {44, 7, 1, 1}, negative / out-of-volume coordinates.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../giga_b200/csrc/tc.cuh"
#include "../giga_b200/csrc/unet_tall.cuh"
using namespace giga;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int iz, int iy, int ix, int b, float* out, int rows, int bw) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar, (uint32_t)(rows * bw * 4));
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     tc::smem_u32(smem)),
                 "l"(&tmap), "r"(iz), "r"(iy), "r"(ix), "r"(b), "r"(tc::smem_u32(&bar))
                 : "memory");
  }
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 16) && !ok; ++i)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(tc::smem_u32(&bar)), "r"(0u) : "memory");
  if (threadIdx.x == 0) out[rows * bw] = ok ? 1.f : -1.f;   // status word: -1 = the copy never completed
  for (int e = threadIdx.x; e < rows * bw; e += blockDim.x) out[e] = reinterpret_cast<float*>(smem)[e];
}

int main() {
  const int B = 3, G = 40;
  std::vector<float> h((size_t)B * G * G * G);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.001f + 1.0f;
  float *d, *dout;
  cudaMalloc(&d, h.size() * 4 + 1024);
  cudaMalloc(&dout, 7 * 44 * 4);
  float* base = d + 64;   // offset view (256 B)
  cudaMemcpy(base, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(ce), (int)qr, fn);
  EncodeFn encode = (EncodeFn)fn;
  cudaFree(dout);
  cudaMalloc(&dout, (7 * 64 + 4) * 4);
  struct V { int bw, rows, c[4]; };
  const V vs[] = {{40, 7, {0, 0, 0, 0}}, {40, 7, {0, -1, 0, 0}}, {40, 7, {0, 36, 3, 1}}, {44, 7, {0, 0, 0, 0}}, {44, 7, {-1, -1, 0, 0}}, {48, 7, {-4, -1, 0, 0}},
                  {44, 7, {-1, 4, 40, 1}}, {44, 3, {-1, 38, 17, 2}}, {64, 7, {-4, -1, 5, 1}}};
  for (const V& v : vs) {
    CUtensorMap tm;
    const cuuint64_t dims[4] = {40, 40, 40, (cuuint64_t)B};
    const cuuint64_t strides[3] = {160, 6400, 256000};
    const cuuint32_t box[4] = {(cuuint32_t)v.bw, (cuuint32_t)v.rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box {%d,%d}: encode -> %d\n", v.bw, v.rows, (int)r); continue; }
    cudaMemset(dout, 0, (7 * 64 + 4) * 4);
    probe<<<1, 128, 8192>>>(tm, v.c[0], v.c[1], v.c[2], v.c[3], dout, v.rows, v.bw);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(v.rows * v.bw + 1);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < v.rows; ++r2)
      for (int z = 0; z < v.bw; ++z) {
        const int gz = v.c[0] + z, gy = v.c[1] + r2, gx = v.c[2], gb = v.c[3];
        float want = 0.f;
        if (gz >= 0 && gz < G && gy >= 0 && gy < G && gx >= 0 && gx < G) want = h[((size_t)(gb * G + gx) * G + gy) * G + gz];
        if (o[r2 * v.bw + z] != want) ++bad;
      }
    printf("box {%d,%d,1,1} at (iz %d, iy %d, ix %d, b %d): %s, status %g, mismatches %d / %d\n", v.bw, v.rows, v.c[0], v.c[1], v.c[2], v.c[3],
           cudaGetErrorString(e), o[v.rows * v.bw], bad, v.rows * v.bw);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
