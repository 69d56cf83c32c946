// Device-side MISE (multi-resolution iso-surface extraction) bookkeeping for Generator3D's occupancy sweep.
//
// Replaces (reference, relative to src/vgn/ConvONets): utils/libmise/mise.pyx (the Cython octree: query :112-136, update :87-110,
// subdivide_voxels :184-232, subdivide_voxel :234-281, get_voxel_idx :284-345, to_dense :138-176) as driven by
// conv_onet/generation.py:127-143 (generate_from_latent).  The octree is restated DENSELY on the finest lattice -- resolution
// R = resolution0 * 2^depth, (R+1)^3 grid points, R^3 fine cells:
//   pstate[point]  bit 0 = the grid point exists (was added by the constructor or a subdivision), bit 1 = its value is known
//   val[point]     the evaluated occupancy logit (fp32: the reference stores the fp32 network output widened to double)
//   level[cell]    level of the LEAF voxel that contains the fine cell (0 = coarsest, size 2^depth; depth = finest, size 1)
// One MISE iteration = collect the unknown points (-> query positions), evaluate them with the TSDF head (the decoder kernel),
// scatter the values, mark the active leaves (a leaf is active when the KNOWN grid points of its closed box contain a value >= threshold
// and a value <= threshold: mise.pyx marks, for every known point, the leaves containing point + {-1,0}^3), subdivide them (level + 1,
// 27 new grid points at half spacing).  Leaves created in an iteration are first considered in the next one, exactly like the
// reference (marking reads the leaf structure of the start of the iteration: two kernels).  The final grid is order independent, so
// the set-wise restatement is bit-exact in values, known-sets and refinement pattern (the CPU test suite pins the same dense
// formulation against the compiled reference octree, the GPU tests pin these kernels against that).
#pragma once
#include "common.cuh"

namespace giga {

struct MiseDims {
  int R;       // finest resolution
  int L;       // R + 1 lattice points per axis
  int depth;   // upsampling steps
};

// grid ceil(L^3 / 256): constructor (mise.pyx:42-85): all cells leaves of level 0, grid points at multiples of 2^depth exist, nothing known
__global__ void __launch_bounds__(256) mise_init_kernel(MiseDims d, unsigned char* __restrict__ pstate, float* __restrict__ val,
                                                         unsigned char* __restrict__ level, unsigned char* __restrict__ active, int* __restrict__ count) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long np = (long)d.L * d.L * d.L, nc = (long)d.R * d.R * d.R;
  if (t == 0) *count = 0;
  if (t < nc) { level[t] = 0; active[t] = 0; }
  if (t >= np) return;
  const int z = (int)(t % d.L), y = (int)((t / d.L) % d.L), x = (int)(t / ((long)d.L * d.L));
  const int m = (1 << d.depth) - 1;
  pstate[t] = ((x & m) == 0 && (y & m) == 0 && (z & m) == 0) ? 1 : 0;
  val[t] = 0.f;
}

// query() (mise.pyx:112-136) + the position arithmetic of generation.py:131-135: pointsf = box_size * (points / R - 0.5) in float64,
// rounded to fp32 by torch.FloatTensor.  grid ceil(L^3 / 256)
__global__ void __launch_bounds__(256) mise_collect_kernel(MiseDims d, const unsigned char* __restrict__ pstate, double box_size, int* __restrict__ count,
                                                            int* __restrict__ qidx, float* __restrict__ qpts, int cap) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long)d.L * d.L * d.L) return;
  if ((pstate[t] & 3) != 1) return;   // exists and not known
  const int pos = atomicAdd(count, 1);
  if (pos >= cap) return;
  const int z = (int)(t % d.L), y = (int)((t / d.L) % d.L), x = (int)(t / ((long)d.L * d.L));
  qidx[pos] = (int)t;
  qpts[3 * pos + 0] = (float)(box_size * ((double)x / (double)d.R - 0.5));
  qpts[3 * pos + 1] = (float)(box_size * ((double)y / (double)d.R - 0.5));
  qpts[3 * pos + 2] = (float)(box_size * ((double)z / (double)d.R - 0.5));
}

// update() (mise.pyx:87-110), first half: store the values, mark the points known; resets the query counter.  grid ceil(n / 256)
__global__ void __launch_bounds__(256) mise_scatter_kernel(const int* __restrict__ qidx, const float* __restrict__ occ, int n, unsigned char* __restrict__ pstate,
                                                            float* __restrict__ val, int* __restrict__ count) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t == 0) *count = 0;
  if (t >= n) return;
  const int i = qidx[t];
  val[i] = occ[t];
  pstate[i] |= 2;
}

// subdivide_voxels(), marking (mise.pyx:196-226): the thread of a leaf's origin cell scans the known grid points of the leaf's closed box.
// grid ceil(R^3 / 256)
__global__ void __launch_bounds__(256) mise_mark_kernel(MiseDims d, const unsigned char* __restrict__ pstate, const float* __restrict__ val,
                                                         const unsigned char* __restrict__ level, unsigned char* __restrict__ active, double threshold) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long)d.R * d.R * d.R) return;
  const int z = (int)(t % d.R), y = (int)((t / d.R) % d.R), x = (int)(t / ((long)d.R * d.R));
  const int lv = level[t];
  unsigned char act = 0;
  if (lv < d.depth) {
    const int s = 1 << (d.depth - lv);
    if ((x & (s - 1)) == 0 && (y & (s - 1)) == 0 && (z & (s - 1)) == 0) {
      bool pos = false, neg = false;
      for (int i = 0; i <= s; ++i)
        for (int j = 0; j <= s; ++j)
          for (int k = 0; k <= s; ++k) {
            const long p = ((long)(x + i) * d.L + (y + j)) * d.L + (z + k);
            if (pstate[p] & 2) {
              const double v = (double)val[p];
              pos |= v >= threshold;
              neg |= v <= threshold;
            }
          }
      act = (pos && neg) ? 1 : 0;
    }
  }
  active[t] = act;
}

// subdivide_voxel() (mise.pyx:234-281) for every active leaf: its cells move one level down, the 27 grid points at half spacing exist.
// grid ceil(R^3 / 256)
__global__ void __launch_bounds__(256) mise_apply_kernel(MiseDims d, unsigned char* __restrict__ pstate, unsigned char* __restrict__ level,
                                                          const unsigned char* __restrict__ active) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long)d.R * d.R * d.R) return;
  const int z = (int)(t % d.R), y = (int)((t / d.R) % d.R), x = (int)(t / ((long)d.R * d.R));
  const int lv = level[t];
  if (lv >= d.depth) return;
  const int s = 1 << (d.depth - lv);
  const int ox = x & ~(s - 1), oy = y & ~(s - 1), oz = z & ~(s - 1);
  if (!active[((long)ox * d.R + oy) * d.R + oz]) return;
  level[t] = (unsigned char)(lv + 1);
  if (x == ox && y == oy && z == oz) {
    const int h = s >> 1;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          const long p = ((long)(ox + i * h) * d.L + (oy + j * h)) * d.L + (oz + k * h);
          pstate[p] |= 1;   // concurrent writers of one point all write the same value (the known bit does not change in this kernel)
        }
  }
}

// to_dense() (mise.pyx:138-176): values of the grid points, NaN elsewhere, then completed along x, y, z (each from its lower neighbour).
// AXIS 0 also initialises the output.  thread = one line of the axis; grid ceil(L^2 / 256)
template <int AXIS>
__global__ void __launch_bounds__(256) mise_dense_kernel(MiseDims d, const unsigned char* __restrict__ pstate, const float* __restrict__ val, float* __restrict__ out) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= d.L * d.L) return;
  const int a = t / d.L, b = t % d.L;
  const long L = d.L;
  // line index -> (stride along the axis, base offset): x lines over (j, k), y lines over (i, k), z lines over (i, j)
  const long stride = AXIS == 0 ? L * L : (AXIS == 1 ? L : 1);
  const long base = AXIS == 0 ? a * L + b : (AXIS == 1 ? a * L * L + b : (a * L + b) * L);
  float prev = 0.f;
  for (int i = 0; i < d.L; ++i) {
    const long p = base + i * stride;
    float v;
    if (AXIS == 0) v = (pstate[p] & 1) ? val[p] : __int_as_float(0x7fc00000);
    else v = out[p];
    if (i > 0 && v != v) v = prev;
    out[p] = v;
    prev = v;
  }
}

}  // namespace giga
