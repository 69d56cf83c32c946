// Device-side parameter commit: the operand layouts of the tensor-core kernels (fp16 hi / scaled-lo splits) written by kernels straight from
// device-resident reference-layout tensors -- what giga_ctx_commit_params packs on the host (giga_api.cu: pack_conv_tc, the decoder_ws
// blobs) without the device -> host -> pack -> device round trip (4.6 ms per commit, once per optimizer step in a training loop that
// evaluates with the inference path).  The results are bit-identical to the host packer's (tests/test_gpu_parity.py compares the blobs).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "decoder.cuh"
#include "decoder_ws.cuh"
#include "unet_tall.cuh"

namespace giga {

__device__ __forceinline__ void split_half_dev(float v, float lo_scale, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn((v - __half2float(h)) * lo_scale);
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}

// conv weights -> [ntile][chunk 16 ch][tap][kc 2][hi,lo][n NTILE][8 halfs]; mode 0: Conv2d [co][ci][3][3]; mode 1: ConvTranspose2d
// [ci][co][2][2] with ntile = a*2+b; mode 2: conv_final [n 32][k 32] -> [hi|lo][kc 4][n 32][8 halfs]
struct TcPackEntry {
  const float* src;
  uint16_t* dst;
  int NNT, NC, NTAPS, NTILE, CIN, COUT, mode, pad;
};

// grid (entries, blocks), block 256
__global__ void __launch_bounds__(256) pack_conv_tc_kernel(const TcPackEntry* __restrict__ tab) {
  const TcPackEntry e = tab[blockIdx.x];
  if (e.mode == 2) {
    for (int x = blockIdx.y * 256 + threadIdx.x; x < 1024; x += gridDim.y * 256) {
      const int k = x & 31, n = x >> 5;
      uint16_t hi, lo;
      split_half_dev(e.src[n * 32 + k], LO_SCALE, hi, lo);
      e.dst[((k / 8) * 32 + n) * 8 + (k % 8)] = hi;
      e.dst[1024 + ((k / 8) * 32 + n) * 8 + (k % 8)] = lo;
    }
    return;
  }
  const long total = (long)e.NNT * e.NC * e.NTAPS * 2 * e.NTILE * 8;
  for (long x = (long)blockIdx.y * 256 + threadIdx.x; x < total; x += (long)gridDim.y * 256) {
    const int j = (int)(x & 7);
    long r = x >> 3;
    const int n = (int)(r % e.NTILE); r /= e.NTILE;
    const int kc = (int)(r & 1); r >>= 1;
    const int tap = (int)(r % e.NTAPS); r /= e.NTAPS;
    const int c = (int)(r % e.NC);
    const int nt = (int)(r / e.NC);
    const int ci = c * 16 + kc * 8 + j;
    const float v = e.mode == 0 ? e.src[((long)(nt * e.NTILE + n) * e.CIN + ci) * 9 + tap] : e.src[((long)ci * e.COUT + n) * 4 + nt];
    uint16_t hi, lo;
    split_half_dev(v, LO_SCALE, hi, lo);
    const long base = ((((long)nt * e.NC + c) * e.NTAPS + tap) * 2) * 2 * e.NTILE * 8;
    e.dst[base + ((kc * 2 + 0) * e.NTILE + n) * 8 + j] = hi;
    e.dst[base + ((kc * 2 + 1) * e.NTILE + n) * 8 + j] = lo;
  }
}

// the decoder heads' tensors in the reference layouts (Linear weight = [out][in])
struct HeadParams {
  const float *fcc_w[5], *fcc_b[5], *w0[5], *b0[5], *w1[5], *b1[5];
};
struct DecPackArgs {
  HeadParams head[4];
  unsigned heads;          // bit h: head h present
  uint8_t* wblob[5];       // per job type (0: the three grasp heads, 1 + h: head h alone), null = absent
  float* hc;               // [4][HC_SIZE]
  const float* dheads;     // [4][DW_HEAD] the fp32 FMA-pipe blob (already packed): source of the head constants
  float* scale;            // [4][2]: 2^s, 2^-s
};

// per-head power-of-two pre-scale: max |w| * 2^s in [512, 1024) over fc_c / fc_0 / fc_1 of the five blocks.  grid 4, block 256
__global__ void __launch_bounds__(256) head_scale_kernel(const __grid_constant__ DecPackArgs A) {
  __shared__ float red[8];
  const int h = blockIdx.x, tid = threadIdx.x;
  if (!(A.heads & (1u << h))) return;
  float m = 0.f;
  for (int i = 0; i < 5; ++i) {
    for (int e = tid; e < 32 * 96; e += 256) m = fmaxf(m, fabsf(A.head[h].fcc_w[i][e]));
    for (int e = tid; e < 32 * 32; e += 256) m = fmaxf(m, fmaxf(fabsf(A.head[h].w0[i][e]), fabsf(A.head[h].w1[i][e])));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    int sexp = 0;
    if (m > 0.f && isfinite(m)) {
      int e;
      frexpf(m, &e);   // m = f * 2^e, f in [0.5, 1)
      sexp = min(max(10 - e, -14), 24);
    }
    A.scale[2 * h] = ldexpf(1.f, sexp);
    A.scale[2 * h + 1] = ldexpf(1.f, -sexp);
  }
}

// grid (5 job types, 5 blocks), block 256: one block of one job type's streamed weight blob; job type 0 / block 0 also writes the head constants
__global__ void __launch_bounds__(256) pack_decoder_ws_kernel(const __grid_constant__ DecPackArgs A) {
  const int type = blockIdx.x, blk = blockIdx.y, tid = threadIdx.x;
  if (type == 0 && blk == 0) {
    for (int e = tid; e < 4 * HC_SIZE; e += 256) {
      const int h = e / HC_SIZE, r = e % HC_SIZE;
      if (!(A.heads & (1u << h))) continue;
      const float* H = A.dheads + (size_t)h * DW_HEAD;
      float v = 0.f;
      if (r < HC_OUT) v = H[DW_FCP + r];
      else if (r < HC_B1) v = H[DW_OUT + (r - HC_OUT)];
      else if (r < HC_INV) v = H[DW_BLOCK0 + 4 * DW_BLK + DW_BLK_B1 + (r - HC_B1)];
      else if (r == HC_INV) v = A.scale[2 * h + 1];
      A.hc[e] = v;
    }
  }
  uint8_t* blob = A.wblob[type];
  if (!blob) return;
  const int nh = type == 0 ? 3 : 1, head0 = type == 0 ? 0 : type - 1;
  uint8_t* base = blob + (size_t)blk * wd_block_bytes(nh);
  uint16_t* fcc = reinterpret_cast<uint16_t*>(base);
  uint16_t* ch = reinterpret_cast<uint16_t*>(base + wd_fcc_bytes(nh));
  float* bias = reinterpret_cast<float*>(base + wd_fcc_bytes(nh) + nh * 8192);
  for (int c = 0; c < nh; ++c) {
    const HeadParams& P = A.head[head0 + c];
    const float ws = A.scale[2 * (head0 + c)];
    for (int x = tid; x < 96 * 32; x += 256) {      // fc_c: value W[j][k]
      const int k = x % 96, j = x / 96;
      uint16_t hi, lo;
      split_half_dev(P.fcc_w[blk][x] * ws, 1.f, hi, lo);
      const size_t e = ((size_t)(k / 8) * (nh * 32) + c * 32 + j) * 8 + (k % 8);
      fcc[e] = hi;
      fcc[(size_t)12 * nh * 32 * 8 + e] = lo;
    }
    for (int f = 0; f < 2; ++f) {
      const float* Wf = f == 0 ? P.w0[blk] : P.w1[blk];
      for (int x = tid; x < 32 * 32; x += 256) {
        const int k = x & 31, j = x >> 5;
        uint16_t hi, lo;
        split_half_dev(Wf[x] * ws, 1.f, hi, lo);
        const size_t e = (size_t)f * nh * 2048 + (size_t)c * 2048 + (k / 8) * 256 + j * 8 + (k % 8);
        ch[e] = hi;
        ch[e + 1024] = lo;
      }
    }
    if (tid < 32) {
      bias[c * 32 + tid] = P.b0[blk][tid];
      bias[nh * 32 + c * 32 + tid] = P.fcc_b[blk][tid] + (blk > 0 ? P.b1[blk - 1][tid] : 0.f);   // bc_b + b1_{b-1}
    }
  }
}

}  // namespace giga
