// ground.cu -- the two iterative solvers of the ground stage as persistent kernels (sm_100a).
//
// Replaces, in pcdet/models/registration/preprocessors/preprocessor_utils.py of the reference,
//   * iterative_reweighted_ransac (:32-80) + the 30-ratio loop that drives it (:147-170): per super-pillar
//     IRLS plane fits, up to 30 x 50 iterations, each ~20 torch launches + a batched eigh + a blocking
//     .max() in the reference;
//   * l1_minimization (:313-350): up to 10 000 AdamW iterations on the pillar height grid with a blocking
//     loss.item() per iteration.
// Both loops run to completion inside ONE launch with their stopping rules evaluated on the device.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pcs {

// ------------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition (cyclic Jacobi from the identity, like the batched syevj the reference
// reaches through torch.linalg.eigh on CUDA).  Returns the eigenvector of the smallest eigenvalue.
// ------------------------------------------------------------------------------------------------
__device__ void smallest_eigvec3(const double a_in[6], float n_out[3]) {
  // a = [xx, xy, xz, yy, yz, zz]
  double a[3][3] = {{a_in[0], a_in[1], a_in[2]}, {a_in[1], a_in[3], a_in[4]}, {a_in[2], a_in[4], a_in[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-18 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = a[p][q];
        if (apq == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A J
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T A
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {  // V <- V J
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  n_out[0] = (float)v[0][m];
  n_out[1] = (float)v[1][m];
  n_out[2] = (float)v[2][m];
}

constexpr int kAcc = 10;         // S0, S1[3], S2[6]
constexpr int kRansacThreads = 256;
constexpr int kGroup = 128;      // super-pillars whose planes a block keeps in shared memory at a time

struct RansacArgs {
  const float4 *vox;       // [Nv] (unused, x, y, z) sorted by super-pillar
  const int *cidx;         // [Nv] super-pillar id (ascending)
  const int *seg_start;    // [C+1]
  const float *origin;     // [C][3] local origin per super-pillar (any point near its voxels)
  const float *cmin_z;     // [C]
  const float *cmax_z;     // [C]
  const float *ratios;     // [n_ratios]
  float *w;                // [Nv] IRLS weights (scratch)
  double *acc;             // [3][C][kAcc] rotating moment accumulators (zeroed by the caller)
  int *nhit;               // [3][C] rotating hit counters (zeroed by the caller)
  unsigned int *gmax;      // [3] max |dw| as float bits (zeroed by the caller)
  float *center;           // [C][3] current plane centre (written by the pillar's owner block)
  float *normal;           // [C][3] current plane normal
  float *best_center;      // [C][3] out
  float *best_normal;      // [C][3] out (initialised to (0,0,1) by the caller)
  float *best_conf;        // [C]    out (initialised to 0 by the caller)
  int *iters_out;          // [n_ratios] IRLS iterations used per ratio (diagnostics)
  long long Nv;
  int C;
  int n_ratios;
  float sigma2;
  float stopping_delta;
  int max_iter;
};

struct Plane {
  float cx, cy, cz, nx, ny, nz, ox, oy, oz;
};

// plane of super-pillar p from the accumulated moments: centre = (sum w x)/(sum w + 1e-6), normal = eigenvector of
// the smallest eigenvalue of mean_i w_i d_i d_i^T (preprocessor_utils.py:47-71); moments are relative to origin[p]
__device__ __noinline__ void fit_plane(const RansacArgs &A, int p, const double *acc_buf, Plane &pl) {
  const double *s = acc_buf + (long long)p * kAcc;
  const double S0 = s[0];
  const double ox = A.origin[p * 3 + 0], oy = A.origin[p * 3 + 1], oz = A.origin[p * 3 + 2];
  const double inv = 1.0 / (S0 + 1e-6);
  const double cax = (s[1] + ox * S0) * inv, cay = (s[2] + oy * S0) * inv, caz = (s[3] + oz * S0) * inv;
  const double cx = cax - ox, cy = cay - oy, cz = caz - oz;
  const int n = A.seg_start[p + 1] - A.seg_start[p];
  const double invn = 1.0 / (double)(n > 0 ? n : 1);
  double cov[6];
  cov[0] = (s[4] - 2.0 * cx * s[1] + S0 * cx * cx) * invn;
  cov[1] = (s[5] - cx * s[2] - cy * s[1] + S0 * cx * cy) * invn;
  cov[2] = (s[6] - cx * s[3] - cz * s[1] + S0 * cx * cz) * invn;
  cov[3] = (s[7] - 2.0 * cy * s[2] + S0 * cy * cy) * invn;
  cov[4] = (s[8] - cy * s[3] - cz * s[2] + S0 * cy * cz) * invn;
  cov[5] = (s[9] - 2.0 * cz * s[3] + S0 * cz * cz) * invn;
  float nrm[3];
  smallest_eigvec3(cov, nrm);
  pl.cx = (float)cax;
  pl.cy = (float)cay;
  pl.cz = (float)caz;
  pl.nx = nrm[0];
  pl.ny = nrm[1];
  pl.nz = nrm[2];
  pl.ox = (float)ox;
  pl.oy = (float)oy;
  pl.oz = (float)oz;
}

__device__ __forceinline__ void warp_flush(float *acc10, int &hits, int pid, double *acc_buf, int *nhit_buf, int lane) {
#pragma unroll
  for (int k = 0; k < kAcc; k++) {
    double v = (double)acc10[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(acc_buf + (long long)pid * kAcc + k, v);
    acc10[k] = 0.f;
  }
  int h = hits;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
  if (lane == 0 && h != 0) atomicAdd(nhit_buf + pid, h);
  hits = 0;
}

// One IRLS half-step of a block over its contiguous voxel range [v0, v1):
//   fit the planes of the super-pillars in the range from acc_fit (skipped when `first`: weights come from
//   the height prior), evaluate them on every voxel (new weight, |dw|, hit) and accumulate the moments of the
//   NEXT fit into acc_next.  Pillars are handled in groups of kGroup whose planes live in shared memory.
//   The block that owns a pillar (holds its first voxel) publishes the plane and clears the spare buffers.
__device__ void ransac_step(const RansacArgs &A, long long v0, long long v1, bool first, float ratio,
                            const double *acc_fit, double *acc_next, int *nhit_next, unsigned int *gmax_next,
                            double *acc_spare, int *nhit_spare, Plane *s_plane) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = kRansacThreads / 32;
  const float sigma = sqrtf(A.sigma2);
  float dmax = 0.f;
  if (v1 > v0) {
    const int p_first = A.cidx[v0], p_last = A.cidx[v1 - 1];
    for (int g0 = p_first; g0 <= p_last; g0 += kGroup) {
      const int g1 = min(g0 + kGroup, p_last + 1);  // pillars [g0, g1)
      __syncthreads();
      for (int p = g0 + threadIdx.x; p < g1; p += kRansacThreads) {
        Plane pl;
        if (!first) {
          fit_plane(A, p, acc_fit, pl);
        } else {
          pl.ox = A.origin[p * 3 + 0];
          pl.oy = A.origin[p * 3 + 1];
          pl.oz = A.origin[p * 3 + 2];
          pl.cx = A.cmin_z[p] * ratio + A.cmax_z[p] * (1.0f - ratio);  // prior height (preprocessor_utils.py:148)
          pl.cy = pl.cz = pl.nx = pl.ny = pl.nz = 0.f;
        }
        s_plane[p - g0] = pl;
        const long long ps = A.seg_start[p];
        if (ps >= v0 && ps < v1 && A.seg_start[p + 1] > ps) {  // owner of a non-empty pillar
          if (!first) {
            A.center[p * 3 + 0] = pl.cx;
            A.center[p * 3 + 1] = pl.cy;
            A.center[p * 3 + 2] = pl.cz;
            A.normal[p * 3 + 0] = pl.nx;
            A.normal[p * 3 + 1] = pl.ny;
            A.normal[p * 3 + 2] = pl.nz;
          }
          for (int k = 0; k < kAcc; k++) acc_spare[(long long)p * kAcc + k] = 0.0;
          nhit_spare[p] = 0;
        }
      }
      __syncthreads();
      // voxels of this pillar group inside the block's range
      const long long a = max(v0, (long long)A.seg_start[g0]);
      const long long b = min(v1, (long long)A.seg_start[g1]);
      float acc[kAcc];
#pragma unroll
      for (int k = 0; k < kAcc; k++) acc[k] = 0.f;
      int hits = 0;
      int cur = -1;
      constexpr int U = 4;  // warp steps whose loads are issued together (memory-level parallelism)
      for (long long base0 = a + warp * 32; base0 < b; base0 += (long long)U * nwarp * 32) {
        int pidv[U];
        float4 pv[U];
        float wold[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const long long i = base0 + (long long)u * nwarp * 32 + lane;
          pidv[u] = -1;
          if (i < b) {
            pidv[u] = __ldg(A.cidx + i);
            pv[u] = __ldg(A.vox + i);
            wold[u] = first ? 0.f : A.w[i];
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const long long i = base0 + (long long)u * nwarp * 32 + lane;
          const int pid = pidv[u];
          const bool valid = pid >= 0;
          float wnew = 0.f, xr = 0.f, yr = 0.f, zr = 0.f;
          int hit = 0;
          if (valid) {
            const float4 p = pv[u];
            const Plane &pl = s_plane[pid - g0];
            if (first) {
              const float zd = pl.cx - p.w;
              wnew = A.sigma2 / (zd * zd + A.sigma2);
            } else {
              const float dx = p.y - pl.cx, dy = p.z - pl.cy, dz = p.w - pl.cz;
              const float err = fabsf(dx * pl.nx + dy * pl.ny + dz * pl.nz);
              hit = err < sigma;
              const float nw = A.sigma2 / (err * err + A.sigma2);
              const float dw = 0.25f / (dx * dx + dy * dy + dz * dz + 0.25f);
              wnew = nw * dw;
              dmax = fmaxf(dmax, fabsf(wnew - wold[u]));
            }
            A.w[i] = wnew;
            xr = p.y - pl.ox;
            yr = p.z - pl.oy;
            zr = p.w - pl.oz;
          }
          const int pl0 = __shfl_sync(0xffffffffu, pid, 0);
          const bool uniform = __all_sync(0xffffffffu, (!valid) || pid == pl0) && pl0 >= 0;
          if (uniform) {
            if (pl0 != cur) {
              if (cur >= 0) warp_flush(acc, hits, cur, acc_next, nhit_next, lane);
              cur = pl0;
            }
            if (valid) {
              const float wx = wnew * xr, wy = wnew * yr, wz = wnew * zr;
              acc[0] += wnew;
              acc[1] += wx;
              acc[2] += wy;
              acc[3] += wz;
              acc[4] += wx * xr;
              acc[5] += wx * yr;
              acc[6] += wx * zr;
              acc[7] += wy * yr;
              acc[8] += wy * zr;
              acc[9] += wz * zr;
              hits += hit;
            }
          } else if (__any_sync(0xffffffffu, valid)) {
            if (cur >= 0) warp_flush(acc, hits, cur, acc_next, nhit_next, lane);
            cur = -1;
            if (valid) {  // a pillar boundary inside this warp step: per-lane atomics
              const double w = wnew, x = xr, y = yr, z = zr;
              double *d = acc_next + (long long)pid * kAcc;
              atomicAdd(d + 0, w);
              atomicAdd(d + 1, w * x);
              atomicAdd(d + 2, w * y);
              atomicAdd(d + 3, w * z);
              atomicAdd(d + 4, w * x * x);
              atomicAdd(d + 5, w * x * y);
              atomicAdd(d + 6, w * x * z);
              atomicAdd(d + 7, w * y * y);
              atomicAdd(d + 8, w * y * z);
              atomicAdd(d + 9, w * z * z);
              if (hit) atomicAdd(nhit_next + pid, 1);
            }
          }
        }
      }
      if (cur >= 0) warp_flush(acc, hits, cur, acc_next, nhit_next, lane);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0 && dmax > 0.f) atomicMax(gmax_next, __float_as_uint(dmax));
}

__global__ void __launch_bounds__(kRansacThreads, 3) ground_ransac_kernel(RansacArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ Plane s_plane[kGroup];
  // contiguous, 32-aligned voxel range of this block
  const long long per = (((A.Nv + gridDim.x - 1) / gridDim.x) + 31) / 32 * 32;
  const long long v0 = min((long long)blockIdx.x * per, A.Nv), v1 = min(v0 + per, A.Nv);
  const long long accN = (long long)A.C * kAcc;
  for (int r = 0; r < A.n_ratios; r++) {
    const float ratio = A.ratios[r];
    // weights from the height prior; moments of fit 0 -> buffer 0 (buffers 1 and 2 are clear)
    ransac_step(A, v0, v1, true, ratio, nullptr, A.acc, A.nhit, A.gmax + 0, A.acc + 2 * accN, A.nhit + 2 * A.C, s_plane);
    grid.sync();
    int it = 0, last_eval = 0;
    for (; it < A.max_iter; it++) {
      const int bf = it % 3, bn = (it + 1) % 3, bs = (it + 2) % 3;
      if (blockIdx.x == 0 && threadIdx.x == 0) A.gmax[bs] = 0u;
      ransac_step(A, v0, v1, false, ratio, A.acc + bf * accN, A.acc + bn * accN, A.nhit + bn * A.C, A.gmax + bn,
                  A.acc + bs * accN, A.nhit + bs * A.C, s_plane);
      last_eval = bn;
      grid.sync();
      if (__uint_as_float(A.gmax[bn]) < A.stopping_delta) {
        ++it;
        break;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.iters_out) A.iters_out[r] = it;
    // owners keep the plane that explains the most voxels (preprocessor_utils.py:160-170) and clear all buffers
    if (v1 > v0) {
      const int p_first = A.cidx[v0], p_last = A.cidx[v1 - 1];
      for (int p = p_first + threadIdx.x; p <= p_last; p += kRansacThreads) {
        const long long ps = A.seg_start[p];
        if (!(ps >= v0 && ps < v1 && A.seg_start[p + 1] > ps)) continue;
        const float nh = (float)A.nhit[last_eval * A.C + p];
        if (A.best_conf[p] < nh) {
          A.best_conf[p] = nh;
          for (int k = 0; k < 3; k++) {
            A.best_normal[p * 3 + k] = A.normal[p * 3 + k];
            A.best_center[p * 3 + k] = A.center[p * 3 + k];
          }
        }
        for (int bsel = 0; bsel < 3; bsel++) {
          for (int k = 0; k < kAcc; k++) A.acc[bsel * accN + (long long)p * kAcc + k] = 0.0;
          A.nhit[bsel * A.C + p] = 0;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.gmax[0] = A.gmax[1] = A.gmax[2] = 0u;
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// L1 height-field smoothing with AdamW, one persistent CTA
// ------------------------------------------------------------------------------------------------
struct L1Args {
  const float *min_z;   // [X*Y]
  const float *weight;  // [X*Y]
  float *h;             // [X*Y] out (initial value = start point, zeros in the reference)
  float *m;             // [X*Y] scratch (zeroed)
  float *v;             // [X*Y] scratch (zeroed)
  int *info;            // [2] out: iterations executed, early-stop flag
  float *loss_out;      // [1] last loss
  int X, Y;
  float lr, lr_gamma;
  int decay_step;       // MultiStepLR milestone (<=0: none)
  float rigid_weight;
  int max_iters;
  float beta1, beta2, eps, weight_decay;
  float stop_tol;       // 1e-4
};

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(1024) l1_heightfield_kernel(L1Args A) {
  extern __shared__ float sh[];  // h[X*Y] (when it fits), then reduction scratch
  const int P = A.X * A.Y;
  const int X = A.X, Y = A.Y;
  float *h = A.h;  // global fallback
  __shared__ double red[32];
  __shared__ int s_stop;
  __shared__ double s_loss;
  const bool use_smem = (size_t)P * sizeof(float) <= 200 * 1024;
  if (use_smem) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) sh[i] = A.h[i];
    h = sh;
  }
  if (threadIdx.x == 0) s_stop = 0;
  __syncthreads();
  const float n0 = (float)P, n1 = (float)((X - 2) * Y), n2 = (float)(X * (Y - 2)), n3 = (float)((X - 2) * (Y - 2));
  const float k1 = A.rigid_weight / n1, k2 = A.rigid_weight / n2, k3 = A.rigid_weight / n3;
  double last_loss = 1e10;
  int countdown = 3;
  float lr = A.lr;
  double b1t = 1.0, b2t = 1.0;
  int it = 0;
#define HH(i, j) h[(i) * Y + (j)]
#define WW(i, j) (A.weight[(i) * Y + (j)] + 1e-2f)
  for (; it < A.max_iters; it++) {
    // ---- loss and gradient (gather form) ------------------------------------------------------------
    double lsum = 0.0;
    // each thread keeps the gradients of its cells in registers across the sync below
    float g[64];
    int nc = 0;
    for (int c = threadIdx.x; c < P; c += blockDim.x, nc++) {
      const int i = c / Y, j = c - i * Y;
      const float hc = HH(i, j);
      const float w = A.weight[c];
      const float t0 = (hc - A.min_z[c]) * w;
      float grad = sgnf(t0) * w / n0;
      double l = fabsf(t0) / n0;
      // second differences: along i ("left"), along j ("up"), both diagonals; centre terms contribute -2,
      // the centred terms of the two neighbours along the same line contribute +1 each
      const bool ci = (i >= 1 && i <= X - 2), cj = (j >= 1 && j <= Y - 2);
      if (ci) {
        const float t = (HH(i - 1, j) - 2.f * hc + HH(i + 1, j)) * WW(i, j);
        grad += -2.f * sgnf(t) * WW(i, j) * k1;
        l += (double)fabsf(t) * k1;
      }
      if (i + 1 <= X - 2) {
        const float t = (hc - 2.f * HH(i + 1, j) + HH(i + 2, j)) * WW(i + 1, j);
        grad += sgnf(t) * WW(i + 1, j) * k1;
      }
      if (i - 1 >= 1) {
        const float t = (HH(i - 2, j) - 2.f * HH(i - 1, j) + hc) * WW(i - 1, j);
        grad += sgnf(t) * WW(i - 1, j) * k1;
      }
      if (cj) {
        const float t = (HH(i, j - 1) - 2.f * hc + HH(i, j + 1)) * WW(i, j);
        grad += -2.f * sgnf(t) * WW(i, j) * k2;
        l += (double)fabsf(t) * k2;
      }
      if (j + 1 <= Y - 2) {
        const float t = (hc - 2.f * HH(i, j + 1) + HH(i, j + 2)) * WW(i, j + 1);
        grad += sgnf(t) * WW(i, j + 1) * k2;
      }
      if (j - 1 >= 1) {
        const float t = (HH(i, j - 2) - 2.f * HH(i, j - 1) + hc) * WW(i, j - 1);
        grad += sgnf(t) * WW(i, j - 1) * k2;
      }
      if (ci && cj) {
        const float ta = (HH(i - 1, j - 1) - 2.f * hc + HH(i + 1, j + 1)) * WW(i, j);
        const float tb = (HH(i + 1, j - 1) - 2.f * hc + HH(i - 1, j + 1)) * WW(i, j);
        grad += -2.f * (sgnf(ta) + sgnf(tb)) * WW(i, j) * k3;
        l += ((double)fabsf(ta) + (double)fabsf(tb)) * k3;
      }
      // as the (i-1, j-1) corner of the centre (i+1, j+1) [t1] ...
      if (i + 1 <= X - 2 && j + 1 <= Y - 2 && i + 1 >= 1 && j + 1 >= 1) {
        const float t = (hc - 2.f * HH(i + 1, j + 1) + HH(i + 2, j + 2)) * WW(i + 1, j + 1);
        grad += sgnf(t) * WW(i + 1, j + 1) * k3;
      }
      // ... as the (i+1, j+1) corner of the centre (i-1, j-1) [t1]
      if (i - 1 >= 1 && j - 1 >= 1 && i - 1 <= X - 2 && j - 1 <= Y - 2) {
        const float t = (HH(i - 2, j - 2) - 2.f * HH(i - 1, j - 1) + hc) * WW(i - 1, j - 1);
        grad += sgnf(t) * WW(i - 1, j - 1) * k3;
      }
      // t2 uses h[i+1, j-1] and h[i-1, j+1] around the centre (i, j):
      // this cell is the (i+1, j-1) corner of the centre (i-1, j+1) ...
      if (i - 1 >= 1 && i - 1 <= X - 2 && j + 1 >= 1 && j + 1 <= Y - 2) {
        const float t = (hc - 2.f * HH(i - 1, j + 1) + HH(i - 2, j + 2)) * WW(i - 1, j + 1);
        grad += sgnf(t) * WW(i - 1, j + 1) * k3;
      }
      // ... and the (i-1, j+1) corner of the centre (i+1, j-1)
      if (i + 1 >= 1 && i + 1 <= X - 2 && j - 1 >= 1 && j - 1 <= Y - 2) {
        const float t = (HH(i + 2, j - 2) - 2.f * HH(i + 1, j - 1) + hc) * WW(i + 1, j - 1);
        grad += sgnf(t) * WW(i + 1, j - 1) * k3;
      }
      g[nc] = grad;  // nc < 64 because P <= 65536 (checked by the host wrapper)
      lsum += l;
    }
    // block reduction of the loss
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    __syncthreads();  // all gradient reads of h are done
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int wdx = 0; wdx < (int)(blockDim.x >> 5); wdx++) t += red[wdx];
      s_loss = t;
    }
    // ---- AdamW step (torch.optim.AdamW: decoupled decay, bias-corrected moments) ----------------------
    b1t *= A.beta1;
    b2t *= A.beta2;
    const float bc1 = (float)(1.0 - b1t), bc2s = (float)sqrt(1.0 - b2t);
    nc = 0;
    for (int c = threadIdx.x; c < P; c += blockDim.x, nc++) {
      const float grad = g[nc];
      float p = h[c] * (1.f - lr * A.weight_decay);
      const float m = A.beta1 * A.m[c] + (1.f - A.beta1) * grad;
      const float v = A.beta2 * A.v[c] + (1.f - A.beta2) * grad * grad;
      A.m[c] = m;
      A.v[c] = v;
      const float denom = sqrtf(v) / bc2s + A.eps;
      p -= (lr / bc1) * (m / denom);
      h[c] = p;
    }
    if (A.decay_step > 0 && it + 1 == A.decay_step) lr *= A.lr_gamma;  // MultiStepLR after scheduler.step()
    __syncthreads();
    // ---- stopping rule (preprocessor_utils.py:339-346), evaluated by every thread on the same loss --------
    const double loss = (double)(float)s_loss;
    if ((last_loss - loss) < (double)A.stop_tol) countdown -= 1;
    else countdown = 3;
    if (countdown == 0) {
      ++it;
      break;
    }
    last_loss = loss;
  }
#undef HH
#undef WW
  __syncthreads();
  if (use_smem)
    for (int i = threadIdx.x; i < P; i += blockDim.x) A.h[i] = sh[i];
  if (threadIdx.x == 0) {
    A.info[0] = it;
    A.info[1] = countdown == 0;
    A.loss_out[0] = (float)s_loss;
  }
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_ground_ransac(pcs_stream_t s, const float *vox, const int32_t *cidx, const int32_t *seg_start,
                      const float *origin, const float *cmin_z, const float *cmax_z, const float *ratios, int64_t Nv,
                      int C, int n_ratios, float sigma2, float stopping_delta, int max_iter, float *w, double *acc,
                      int32_t *nhit, uint32_t *gmax, float *center, float *normal, float *best_center,
                      float *best_normal, float *best_conf, int32_t *iters_out) {
  if (Nv < 0 || C < 1 || n_ratios < 1 || ((uintptr_t)vox & 15) || !cidx || !seg_start || !acc || !nhit || !gmax)
    return set_error(PCS_ERR_BAD_ARG, "pcs_ground_ransac: bad args");
  RansacArgs A;
  A.vox = (const float4 *)vox;
  A.cidx = cidx;
  A.seg_start = seg_start;
  A.origin = origin;
  A.cmin_z = cmin_z;
  A.cmax_z = cmax_z;
  A.ratios = ratios;
  A.w = w;
  A.acc = acc;
  A.nhit = nhit;
  A.gmax = gmax;
  A.center = center;
  A.normal = normal;
  A.best_center = best_center;
  A.best_normal = best_normal;
  A.best_conf = best_conf;
  A.iters_out = iters_out;
  A.Nv = Nv;
  A.C = C;
  A.n_ratios = n_ratios;
  A.sigma2 = sigma2;
  A.stopping_delta = stopping_delta;
  A.max_iter = max_iter;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ground_ransac_kernel, kRansacThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long blocks = (Nv + 2047) / 2048;  // at least ~2k voxels per block
  long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void *args[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)ground_ransac_kernel, dim3((unsigned)blocks),
                                              dim3(kRansacThreads), args, 0, as_stream(s));
  g_launches++;
  if (e != cudaSuccess) return set_error((int)e, "ground_ransac_kernel (cooperative launch)");
  return check_launch("ground_ransac_kernel");
}

int pcs_l1_heightfield(pcs_stream_t s, const float *min_z, const float *weight, float *h, float *m, float *v, int X,
                       int Y, float lr, float lr_gamma, int decay_step, float rigid_weight, int max_iters,
                       int32_t *info, float *loss_out) {
  if (X < 3 || Y < 3 || (long long)X * Y > 65536 || !min_z || !weight || !h || !m || !v || !info || !loss_out)
    return set_error(PCS_ERR_BAD_ARG, "pcs_l1_heightfield: bad args (3 <= X,Y and X*Y <= 65536)");
  L1Args A;
  A.min_z = min_z;
  A.weight = weight;
  A.h = h;
  A.m = m;
  A.v = v;
  A.info = info;
  A.loss_out = loss_out;
  A.X = X;
  A.Y = Y;
  A.lr = lr;
  A.lr_gamma = lr_gamma;
  A.decay_step = decay_step;
  A.rigid_weight = rigid_weight;
  A.max_iters = max_iters;
  A.beta1 = 0.9f;
  A.beta2 = 0.999f;
  A.eps = 1e-8f;
  A.weight_decay = 1e-2f;  // torch.optim.AdamW defaults (preprocessor_utils.py:316)
  A.stop_tol = 1e-4f;
  size_t smem = (size_t)X * Y * sizeof(float);
  if (smem > 200 * 1024) smem = 0;
  cudaFuncSetAttribute(l1_heightfield_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  PCS_LAUNCH(l1_heightfield_kernel, 1, 1024, smem, as_stream(s), A);
  return 0;
}

}  // extern "C"
