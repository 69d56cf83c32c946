// Training-step tail: the fused GIGA loss (value + gradient with respect to the network outputs) and the Adam update.
//
// Replaces (reference): scripts/train_giga.py:161-195 -- loss_fn = mean_b[ BCE(qual) + label * (min_i(1 - |<rot, target_i>|) + 0.01 * MSE(40 width)) +
// mean_m BCE(occ) ] with F.binary_cross_entropy's log clamp (-100) and ATen's backward (x - t) / max((1 - x) x, 1e-12) -- about fifteen ATen
// launches and their autograd nodes in the reference, ONE launch here; and :208-209 torch.optim.Adam(net.parameters(), lr) -- the
// single-tensor update rule of torch/optim/adam.py (exp_avg.lerp_(g, 1 - b1); exp_avg_sq = b2 v + (1 - b2) g^2; p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps))
// over ONE flat parameter buffer (164 tensors, 581,863 elements for GIGA) in one launch.
#pragma once
#include "common.cuh"

namespace giga {

__device__ __forceinline__ float bce_value(float x, float t) {   // aten/native/Loss.cpp binary_cross_entropy: logs clamped at -100
  return (t - 1.f) * fmaxf(logf(1.f - x), -100.f) - t * fmaxf(logf(x), -100.f);
}
__device__ __forceinline__ float bce_grad(float x, float t) {    // binary_cross_entropy_backward: (x - t) / max((1 - x) x, 1e-12)
  return (x - t) / fmaxf((1.f - x) * x, 1e-12f);
}

// grid B, block 256.  part [B][4] per-scene loss terms (qual, rot, width, occ); the last CTA to finish reduces them in scene order
// (deterministic) into loss_out[5] = means of (qual, rot, width, occ, all) and re-arms the counter.
__global__ void __launch_bounds__(256) giga_loss_kernel(const float* __restrict__ label_pred, const float* __restrict__ rot_pred,
                                                         const float* __restrict__ width_pred, const float* __restrict__ occ_pred,
                                                         const float* __restrict__ label, const float* __restrict__ rotations,
                                                         const float* __restrict__ width, const float* __restrict__ occ, int B, int M,
                                                         float* __restrict__ part, unsigned int* __restrict__ done, float* __restrict__ loss_out,
                                                         float* __restrict__ g_label, float* __restrict__ g_rot, float* __restrict__ g_width,
                                                         float* __restrict__ g_occ) {
  __shared__ float red[8];
  __shared__ bool last;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float invB = 1.f / (float)B;
  // occupancy term: mean over the M query points of the scene
  float s = 0.f;
  const float gscale = invB / (float)M;
  for (int m = tid; m < M; m += 256) {
    const float x = occ_pred[(size_t)b * M + m], t = occ[(size_t)b * M + m];
    s += bce_value(x, t);
    if (g_occ) g_occ[(size_t)b * M + m] = bce_grad(x, t) * gscale;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float so = 0.f;
    for (int i = 0; i < 8; ++i) so += red[i];
    const float l_occ = M > 0 ? so / (float)M : 0.f;
    const float q = label_pred[b], lab = label[b];
    const float l_qual = bce_value(q, lab);
    float d0 = 0.f, d1 = 0.f;
    for (int k = 0; k < 4; ++k) {
      d0 += rot_pred[4 * b + k] * rotations[8 * b + k];
      d1 += rot_pred[4 * b + k] * rotations[8 * b + 4 + k];
    }
    const float l0 = 1.f - fabsf(d0), l1 = 1.f - fabsf(d1);
    const float l_rot = fminf(l0, l1);
    const float dw = 40.f * width_pred[b] - 40.f * width[b];
    const float l_width = dw * dw;
    part[4 * b + 0] = l_qual; part[4 * b + 1] = l_rot; part[4 * b + 2] = l_width; part[4 * b + 3] = l_occ;
    if (g_label) g_label[b] = bce_grad(q, lab) * invB;
    if (g_width) g_width[b] = lab * 0.01f * (2.f * dw * 40.f) * invB;
    if (g_rot) {
      // torch.min(loss0, loss1) backward sends the gradient to the smaller operand (halves on a tie); |x|' = sign(x), sign(0) = 0
      const float w0 = l0 < l1 ? 1.f : (l0 == l1 ? 0.5f : 0.f), w1 = 1.f - w0;
      const float s0 = d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f), s1 = d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f);
      for (int k = 0; k < 4; ++k)
        g_rot[4 * b + k] = -lab * invB * (w0 * s0 * rotations[8 * b + k] + w1 * s1 * rotations[8 * b + 4 + k]);
    }
    __threadfence();
    last = atomicAdd(done, 1u) == (unsigned)B - 1;
  }
  __syncthreads();
  if (!last || tid != 0) return;
  __threadfence();
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < B; ++i) {
    const volatile float* p = part + 4 * i;
    const float lq = p[0], lr = p[1], lw = p[2], lo = p[3];
    a[0] += lq; a[1] += lr; a[2] += lw; a[3] += lo;
    a[4] += lq + label[i] * (lr + 0.01f * lw) + lo;
  }
  for (int k = 0; k < 5; ++k) loss_out[k] = a[k] * invB;
  *done = 0;
}

struct AdamScalars {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, step_size, bc2_sqrt, eps, weight_decay;
};

// grid = multiple of the SM count, block 256; 128-bit accesses over the flat buffers (5 streams: p, g, m, v read; p, m, v written = 28 B / element)
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                         long n, AdamScalars S) {
  const long n4 = n >> 2;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    if (S.weight_decay != 0.f) gg = fmaf(pp, S.weight_decay, gg);              // grad = grad.add(param, alpha=weight_decay)
    mm = mm + S.one_minus_beta1 * (gg - mm);                                    // exp_avg.lerp_(grad, 1 - beta1)   (weight < 0.5 branch)
    vv = vv * S.beta2 + S.one_minus_beta2 * gg * gg;                            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vv) / S.bc2_sqrt + S.eps;
    pp = pp - S.step_size * (mm / denom);                                       // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (n4 << 2) + threadIdx.x;
    upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace giga
