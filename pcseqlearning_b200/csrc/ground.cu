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

#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pcs {

// ------------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition (cyclic Jacobi from the identity, like the batched syevj the reference
// reaches through torch.linalg.eigh on CUDA).  Returns the eigenvector of the smallest eigenvalue.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void smallest_eigvec3(const double a_in[6], float n_out[3]) {
  // a = [xx, xy, xz, yy, yz, zz]
  const Eig3 e = jacobi_eig3(a_in[0], a_in[1], a_in[2], a_in[3], a_in[4], a_in[5]);
  int m = 0;
  double dm = e.d0;
  if (e.d1 < dm) {
    m = 1;
    dm = e.d1;
  }
  if (e.d2 < dm) m = 2;
  n_out[0] = (float)(m == 0 ? e.v00 : (m == 1 ? e.v01 : e.v02));
  n_out[1] = (float)(m == 0 ? e.v10 : (m == 1 ? e.v11 : e.v12));
  n_out[2] = (float)(m == 0 ? e.v20 : (m == 1 ? e.v21 : e.v22));
}

constexpr int kAcc = 10;         // S0, S1[3], S2[6]
constexpr int kRansacThreads = 256;
constexpr int kGroup = 128;      // super-pillars whose planes a block keeps in shared memory at a time
constexpr int kG = 3;            // height ratios processed together by one team of blocks
constexpr int kMaxRatios = 32;

// All height ratios of the reference's loop (preprocessor_utils.py:147-170) are independent IRLS problems that
// differ only in their initial weights, so they are iterated CONCURRENTLY: one sweep over the voxels advances
// every ratio by one IRLS iteration.  Nothing per (ratio, voxel) is stored: the previous weight needed by the
// max|dw| stopping rule is recomputed from the previous plane, bit-identically.
struct RansacArgs {
  const float4 *vox;       // [Nv] (unused, x, y, z) sorted by super-pillar
  const int *cidx;         // [Nv] super-pillar id (ascending)
  const int *seg_start;    // [C+1]
  const float *origin;     // [C][3] local origin per super-pillar (any point near its voxels)
  const float *cmin_z;     // [C]
  const float *cmax_z;     // [C]
  const float *ratios;     // [R]
  double *acc;             // [3][R][C][kAcc] rotating moment accumulators (zeroed by the caller)
  int *nhit;               // [3][R][C] rotating hit counters (zeroed by the caller)
  unsigned int *gmax;      // [3][R] max |dw| as float bits (zeroed by the caller)
  float *planes;           // [2][R][C][6] published planes (centre, normal) of the last two iterations
  int *fin;                // [R][2] buffers holding the final planes / hit counts of every ratio
  float *best_center;      // [C][3] out
  float *best_normal;      // [C][3] out (initialised to (0,0,1) by the caller)
  float *best_conf;        // [C]    out (initialised to 0 by the caller)
  int *iters_out;          // [R] IRLS iterations used per ratio
  long long Nv;
  int C;
  int R;
  float sigma2;
  float stopping_delta;
  int max_iter;
  int prefetch;  // 1: prefetch the voxels of the warp's next step into L1 while the current step is evaluated
};

struct PlaneN {  // plane being evaluated: centre, normal
  float cx, cy, cz, nx, ny, nz;
};

// plane of super-pillar p from accumulated moments: centre = (sum w x)/(sum w + 1e-6), normal = eigenvector of the
// smallest eigenvalue of mean_i w_i d_i d_i^T (preprocessor_utils.py:47-71); moments are relative to origin[p]
__device__ __noinline__ void fit_plane(const RansacArgs &A, int p, const double *s, PlaneN &pl) {
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
}

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// IRLS weight of a voxel for a plane (preprocessor_utils.py:72-75); err returned for the hit test.
// One fast reciprocal for both factors: sigma2 * 0.25 / ((err^2 + sigma2) (|d|^2 + 0.25)); the weights feed a
// tolerance-compared fit, and the previous weight is recomputed with the same expression (bit-identical).
__device__ __forceinline__ float plane_weight(const PlaneN &pl, float x, float y, float z, float sigma2, float &err) {
  const float dx = x - pl.cx, dy = y - pl.cy, dz = z - pl.cz;
  err = fabsf(dx * pl.nx + dy * pl.ny + dz * pl.nz);
  const float den = (err * err + sigma2) * (dx * dx + dy * dy + dz * dz + 0.25f);
  return sigma2 * 0.25f * fast_rcp(den);  // den >= sigma2 / 4 > 0: no range fix-up needed
}

__device__ __forceinline__ float prior_weight(float prior_z, float z, float sigma2) {
  const float zd = prior_z - z;
  return sigma2 * fast_rcp(zd * zd + sigma2);
}

__device__ __forceinline__ void warp_flush(float *acc10, int &hits, double *dst, int *hit_dst, int lane) {
#pragma unroll
  for (int k = 0; k < kAcc; k++) {
    double v = (double)acc10[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(dst + k, v);
    acc10[k] = 0.f;
  }
  int h = hits;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
  if (lane == 0 && h != 0) atomicAdd(hit_dst, h);
  hits = 0;
}

// Phase A of an IRLS iteration (all blocks): fit plane `it` of every (active ratio, non-empty super-pillar) once,
// publish it, and clear the spare accumulators of the pair.  s_act: the active ratio ids.
__device__ void ransac_fit(const RansacArgs &A, int it, const int *s_act, int n_act) {
  const long long RC = (long long)A.R * A.C;
  const int bf = it % 3, bs = (it + 2) % 3;
  const double *acc_fit = A.acc + (long long)bf * RC * kAcc;
  double *acc_spare = A.acc + (long long)bs * RC * kAcc;
  int *nhit_spare = A.nhit + (long long)bs * RC;
  float *pub_new = A.planes + (long long)(it & 1) * RC * 6;
  const long long tasks = (long long)n_act * A.C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tasks;
       t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % A.C), r = s_act[t / A.C];
    if (A.seg_start[p + 1] <= A.seg_start[p]) continue;  // empty super-pillar
    const long long rp = (long long)r * A.C + p;
    PlaneN pl;
    fit_plane(A, p, acc_fit + rp * kAcc, pl);
    float *d = pub_new + rp * 6;
    d[0] = pl.cx; d[1] = pl.cy; d[2] = pl.cz; d[3] = pl.nx; d[4] = pl.ny; d[5] = pl.nz;
    for (int k = 0; k < kAcc; k++) acc_spare[rp * kAcc + k] = 0.0;
    nhit_spare[rp] = 0;
  }
}

// Phase B of an IRLS iteration of a team (its nr still-active ratios, ids packed 8 bits each in `rids`) over the
// block's contiguous voxel range [v0, v1).
//   it < 0  : weights from the height prior, moments of fit 0
//   it >= 0 : new plane = planes[it & 1] (phase A); previous plane = prior (it == 0) or planes[(it-1) & 1]; evaluate
//             the new plane (weight, |dw|, hit) and accumulate the moments of fit it+1 into acc[(it+1) % 3].
__device__ void ransac_step(const RansacArgs &A, long long v0, long long v1, int it, unsigned int rids, int nr,
                            float (*s_new)[kGroup][8], float (*s_old)[kGroup][8], float (*s_org)[3]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = kRansacThreads / 32;
  const float sigma = sqrtf(A.sigma2);
  const long long RC = (long long)A.R * A.C;
  const int bn = (it + 1) % 3;
  const bool first = it < 0;
  double *acc_next = A.acc + (long long)(first ? 0 : bn) * RC * kAcc;
  int *nhit_next = A.nhit + (long long)(first ? 0 : bn) * RC;
  unsigned int *gmax_next = A.gmax + (first ? 0 : bn) * A.R;
  const float *pub_new = A.planes + (long long)(it & 1) * RC * 6;
  const float *pub_old = A.planes + (long long)((it - 1) & 1) * RC * 6;
  float dmax[kG];
#pragma unroll
  for (int g = 0; g < kG; g++) dmax[g] = 0.f;
  unsigned int viol = 0;  // warp-uniform: ratios for which this warp already found |dw| >= delta
  if (v1 > v0) {
    const int p_first = A.cidx[v0], p_last = A.cidx[v1 - 1];
    for (int g0 = p_first; g0 <= p_last; g0 += kGroup) {
      const int g1 = min(g0 + kGroup, p_last + 1);  // pillars [g0, g1)
      __syncthreads();
      for (int t = threadIdx.x; t < (g1 - g0) * 3; t += kRansacThreads)
        s_org[t / 3][t % 3] = A.origin[(long long)g0 * 3 + t];
      // planes of the group from global memory (written in phase A, before the grid barrier): plain loads
      for (int t = threadIdx.x; t < (g1 - g0) * nr * 6; t += kRansacThreads) {
        const int k = t % 6, lp = (t / 6) % (g1 - g0), g = t / (6 * (g1 - g0));
        const int p = g0 + lp, r = (rids >> (8 * g)) & 0xffu;
        const long long rp = (long long)r * A.C + p;
        if (first) {
          // "plane" of the prior: only its height is used (:148)
          if (k == 0) s_new[g][lp][0] = A.cmin_z[p] * A.ratios[r] + A.cmax_z[p] * (1.0f - A.ratios[r]);
        } else {
          s_new[g][lp][k] = __ldcg(pub_new + rp * 6 + k);
          if (it == 0) {
            if (k == 0) s_old[g][lp][0] = A.cmin_z[p] * A.ratios[r] + A.cmax_z[p] * (1.0f - A.ratios[r]);
          } else {
            s_old[g][lp][k] = __ldcg(pub_old + rp * 6 + k);
          }
        }
      }
      __syncthreads();
      const long long a = max(v0, (long long)A.seg_start[g0]);
      const long long b = min(v1, (long long)A.seg_start[g1]);
      float acc[kG][kAcc];
      int hits[kG];
#pragma unroll
      for (int g = 0; g < kG; g++) {
        hits[g] = 0;
#pragma unroll
        for (int k = 0; k < kAcc; k++) acc[g][k] = 0.f;
      }
      int cur = -1;
      for (long long base = a + warp * 32; base < b; base += nwarp * 32) {
        const long long i = base + lane;
        const bool valid = i < b;
        int pid = -1;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          pid = __ldg(A.cidx + i);
          p = __ldg(A.vox + i);
        }
        if (A.prefetch && i + nwarp * 32 < b) {
          prefetch_l1(A.vox + i + nwarp * 32);
          if ((lane & 7) == 0) prefetch_l1(A.cidx + i + nwarp * 32);
        }
        const int pl0 = __shfl_sync(0xffffffffu, pid, 0);
        const bool uniform = __all_sync(0xffffffffu, (!valid) || pid == pl0) && pl0 >= 0;
        if (uniform && pl0 != cur) {
          if (cur >= 0) {
#pragma unroll
            for (int g = 0; g < kG; g++)
              if (g < nr)
                warp_flush(acc[g], hits[g], acc_next + ((long long)((rids >> (8 * g)) & 0xffu) * A.C + cur) * kAcc,
                           nhit_next + (long long)((rids >> (8 * g)) & 0xffu) * A.C + cur, lane);
          }
          cur = pl0;
        }
        if (!uniform && cur >= 0) {
#pragma unroll
          for (int g = 0; g < kG; g++)
            if (g < nr)
              warp_flush(acc[g], hits[g], acc_next + ((long long)((rids >> (8 * g)) & 0xffu) * A.C + cur) * kAcc,
                         nhit_next + (long long)((rids >> (8 * g)) & 0xffu) * A.C + cur, lane);
          cur = -1;
        }
        // segments of equal super-pillar id inside this warp step (ids ascend, so segments are lane ranges):
        // only needed when the step is not uniform
        int seg_end = 32;
        bool seg_head = false;
        if (!uniform) {
          const int pid_prev = __shfl_up_sync(0xffffffffu, pid, 1);
          seg_head = lane == 0 || pid != pid_prev;
          const unsigned int heads = __ballot_sync(0xffffffffu, seg_head);
          const unsigned int above = heads & ~((2u << lane) - 1u);
          seg_end = above ? __ffs(above) - 1 : 32;
        }
        const int lp = valid ? pid - g0 : 0;
        float xr = 0.f, yr = 0.f, zr = 0.f;
        if (valid) {
          xr = p.y - s_org[lp][0];
          yr = p.z - s_org[lp][1];
          zr = p.w - s_org[lp][2];
        }
#pragma unroll
        for (int g = 0; g < kG; g++) {
          if (g >= nr) continue;
          float wnew = 0.f;
          int hit = 0;
          if (valid) {
            if (first) {
              wnew = prior_weight(s_new[g][lp][0], p.w, A.sigma2);
            } else {
              const float4 na = *reinterpret_cast<const float4 *>(s_new[g][lp]);
              const float2 nb = *reinterpret_cast<const float2 *>(s_new[g][lp] + 4);
              const PlaneN pn = {na.x, na.y, na.z, na.w, nb.x, nb.y};
              float err;
              wnew = plane_weight(pn, p.y, p.z, p.w, A.sigma2, err);
              hit = err < sigma;
              // The stopping rule only asks whether max|dw| < delta.  Once this warp has seen one voxel with
              // |dw| >= delta for ratio g the answer is known, and the previous weight is no longer recomputed.
              if (!((viol >> g) & 1u)) {
                float wold;
                const float *o = s_old[g][lp];
                if (it == 0) {
                  wold = prior_weight(o[0], p.w, A.sigma2);
                } else {
                  const float4 oa = *reinterpret_cast<const float4 *>(o);
                  const float2 ob = *reinterpret_cast<const float2 *>(o + 4);
                  const PlaneN po = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y};
                  float e2;
                  wold = plane_weight(po, p.y, p.z, p.w, A.sigma2, e2);
                }
                dmax[g] = fmaxf(dmax[g], fabsf(wnew - wold));
              }
            }
          }
          const float wx = wnew * xr, wy = wnew * yr, wz = wnew * zr;
          if (uniform) {  // invalid lanes add zeros
            acc[g][0] += wnew;
            acc[g][1] += wx;
            acc[g][2] += wy;
            acc[g][3] += wz;
            acc[g][4] += wx * xr;
            acc[g][5] += wx * yr;
            acc[g][6] += wx * zr;
            acc[g][7] += wy * yr;
            acc[g][8] += wy * zr;
            acc[g][9] += wz * zr;
            hits[g] += hit;
          } else {
            // a super-pillar boundary inside this warp step: segmented sums, one set of atomics per segment
            float m[kAcc] = {wnew, wx, wy, wz, wx * xr, wx * yr, wx * zr, wy * yr, wy * zr, wz * zr};
            int h = hit;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const bool take = lane + o < seg_end;
#pragma unroll
              for (int k = 0; k < kAcc; k++) {
                const float t = __shfl_down_sync(0xffffffffu, m[k], o);
                if (take) m[k] += t;
              }
              const int th = __shfl_down_sync(0xffffffffu, h, o);
              if (take) h += th;
            }
            if (seg_head && valid) {
              const long long rg = (rids >> (8 * g)) & 0xffu;
              double *d = acc_next + (rg * A.C + pid) * kAcc;
#pragma unroll
              for (int k = 0; k < kAcc; k++) atomicAdd(d + k, (double)m[k]);
              if (h) atomicAdd(nhit_next + rg * A.C + pid, h);
            }
          }
        }
        if (!first) {
#pragma unroll
          for (int g = 0; g < kG; g++)
            if (!((viol >> g) & 1u) && __any_sync(0xffffffffu, dmax[g] >= A.stopping_delta)) viol |= 1u << g;
        }
      }
      if (cur >= 0) {
#pragma unroll
        for (int g = 0; g < kG; g++)
          if (g < nr)
            warp_flush(acc[g], hits[g], acc_next + ((long long)((rids >> (8 * g)) & 0xffu) * A.C + cur) * kAcc,
                       nhit_next + (long long)((rids >> (8 * g)) & 0xffu) * A.C + cur, lane);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kG; g++) {
    float d = dmax[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
    if (lane == 0 && d > 0.f && g < nr) atomicMax(gmax_next + ((rids >> (8 * g)) & 0xffu), __float_as_uint(d));
  }
}

#ifdef PCS_RANSAC_TRACE
// debug build only (tools/trace_ransac.py): per-block timestamps of one IRLS iteration
__device__ unsigned long long g_ransac_trace[2048 * 4];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int smid() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#endif

// Adaptive load balance of the voxel sweep.  The cost of a voxel range depends on how many super-pillar boundaries
// it holds (sparse far-field ranges are several times more expensive per voxel than dense ones), so every block
// times its sweep and the team re-cuts its ranges so that, assuming a uniform cost density inside each old range,
// all blocks would have taken the same time.  Scratch lives in static device arrays (one cooperative launch owns
// the GPU at a time); only the order of the atomic additions depends on the cut.
constexpr int kMaxRansacGrid = 2048;
__device__ unsigned int g_bal_dur[kMaxRansacGrid];
__device__ long long g_bal_v0[kMaxRansacGrid];

template <int kMinBlocks>
__global__ void __launch_bounds__(kRansacThreads, kMinBlocks) ground_ransac_kernel(RansacArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) float s_new[kG][kGroup][8];
  __shared__ __align__(16) float s_old[kG][kGroup][8];
  __shared__ float s_org[kGroup][3];
  __shared__ int s_act[kMaxRatios];  // ids of the still-active ratios
  __shared__ int s_nact;
  __shared__ long long s_cut[2];
  // Teams of blocks: every iteration the still-active ratios are dealt out to ceil(n_active / kG) teams (sizes
  // differ by at most one) and every team gets a share of the grid proportional to its ratio count, so the SMs of
  // ratios that have converged go to the ones still iterating.  Inside a team every block owns a contiguous,
  // 32-aligned voxel range.  The assignment only depends on `done`, which every thread derives identically.
  const unsigned int all_done = (A.R >= 32) ? 0xffffffffu : ((1u << A.R) - 1u);
  unsigned int done = 0;
  unsigned int rids = 0;
  int nr = 0;
  long long v0 = 0, v1 = 0;
  int tb0 = 0, tb1 = 0;  // blocks of my team
  auto assign = [&]() {
    const unsigned int act = all_done & ~done;
    const int n_act = __popc(act);
    __syncthreads();
    if (threadIdx.x == 0) {
      int k = 0;
      for (unsigned int m = act; m; m &= m - 1) s_act[k++] = __ffs(m) - 1;
      s_nact = k;
    }
    __syncthreads();
    nr = 0;
    rids = 0;
    v0 = v1 = 0;
    if (n_act == 0) return;
    const int n_teams = (n_act + kG - 1) / kG;
    const int base = n_act / n_teams, rem = n_act % n_teams;
    int first_r = 0;  // index (among active ratios) of the team's first ratio
    for (int t = 0; t < n_teams; t++) {
      const int sz = base + (t < rem ? 1 : 0);
      const int b0 = (int)((long long)gridDim.x * first_r / n_act);
      const int b1 = (int)((long long)gridDim.x * (first_r + sz) / n_act);
      if ((int)blockIdx.x >= b0 && (int)blockIdx.x < b1) {
        nr = sz;
        unsigned int m = act;
        for (int k = 0; k < first_r; k++) m &= m - 1;  // skip the ratios of earlier teams
        for (int g = 0; g < sz; g++) {
          rids |= (unsigned int)(__ffs(m) - 1) << (8 * g);
          m &= m - 1;
        }
        const int bpt = b1 - b0, brank = (int)blockIdx.x - b0;
        tb0 = b0;
        tb1 = b1;
        const long long per = (((A.Nv + bpt - 1) / bpt) + 31) / 32 * 32;
        v0 = min((long long)brank * per, A.Nv);
        v1 = min(v0 + per, A.Nv);
        return;
      }
      first_r += sz;
    }
  };

  // new cut of the team's ranges from the sweep times of the last iteration (every block derives its two cut
  // points with the same arithmetic as its neighbours)
  auto rebalance = [&]() {
    if (threadIdx.x == 0) {
      const int bpt = tb1 - tb0, brank = (int)blockIdx.x - tb0;
      double total = 0.0;
      for (int j = tb0; j < tb1; j++) total += (double)__ldcg(&g_bal_dur[j]);
      for (int e = 0; e < 2; e++) {
        const int k = brank + e;
        long long v = (k >= bpt) ? A.Nv : 0;
        if (k > 0 && k < bpt) {
          const double target = total * (double)k / (double)bpt;
          double cum = 0.0;
          int j = tb0;
          for (; j < tb1; j++) {
            const double d = (double)__ldcg(&g_bal_dur[j]);
            if (cum + d >= target) break;
            cum += d;
          }
          if (j >= tb1) {
            v = A.Nv;
          } else {
            const long long a = __ldcg(&g_bal_v0[j]);
            const long long b = (j + 1 < tb1) ? __ldcg(&g_bal_v0[j + 1]) : A.Nv;
            const double d = (double)__ldcg(&g_bal_dur[j]);
            const double frac = d > 0.0 ? (target - cum) / d : 0.0;
            v = a + (long long)(frac * (double)(b - a));
            v = (v + 16) / 32 * 32;
            v = max(a, min(v, b));
          }
        }
        s_cut[e] = min(v, A.Nv);
      }
    }
    __syncthreads();
    v0 = s_cut[0];
    v1 = s_cut[1];
    __syncthreads();
  };

  assign();
  ransac_step(A, v0, v1, -1, rids, nr, s_new, s_old, s_org);
  grid.sync();
  int it = 0;
  for (; it < A.max_iter && done != all_done; it++) {
    const int bn = (it + 1) % 3, bs = (it + 2) % 3;
    if (blockIdx.x == 0)
      for (int r = threadIdx.x; r < A.R; r += kRansacThreads) A.gmax[bs * A.R + r] = 0u;
    ransac_fit(A, it, s_act, s_nact);
    grid.sync();
#ifdef PCS_RANSAC_TRACE
    if (it == 20 && threadIdx.x == 0 && blockIdx.x < 2048) g_ransac_trace[blockIdx.x * 4 + 0] = gtimer();
#endif
    const long long t_sweep = clock64();
    if (nr > 0) ransac_step(A, v0, v1, it, rids, nr, s_new, s_old, s_org);
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long dt = clock64() - t_sweep;
      g_bal_dur[blockIdx.x] = (unsigned int)min(dt, 0xffffffffLL);
      g_bal_v0[blockIdx.x] = v0;
    }
#ifdef PCS_RANSAC_TRACE
    __syncthreads();
    if (it == 20 && threadIdx.x == 0 && blockIdx.x < 2048) {
      g_ransac_trace[blockIdx.x * 4 + 1] = gtimer();
      g_ransac_trace[blockIdx.x * 4 + 2] = ((unsigned long long)smid() << 32) | (unsigned int)nr;
    }
#endif
    grid.sync();
#ifdef PCS_RANSAC_TRACE
    if (it == 20 && threadIdx.x == 0 && blockIdx.x < 2048) g_ransac_trace[blockIdx.x * 4 + 3] = gtimer();
#endif
    // per-ratio stopping rule (preprocessor_utils.py:76-78), evaluated identically by every thread
    const unsigned int before = done;
    for (int r = 0; r < A.R; r++) {
      if ((done >> r) & 1u) continue;
      const bool conv = __uint_as_float(A.gmax[bn * A.R + r]) < A.stopping_delta;
      if (conv || it + 1 == A.max_iter) {
        done |= 1u << r;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          A.fin[r * 2 + 0] = it & 1;  // planes buffer of the last fit
          A.fin[r * 2 + 1] = bn;      // hit counters of the last evaluation
          A.iters_out[r] = it + 1;
        }
      }
    }
    if (done != before)
      assign();  // new teams: back to the even cut
    else if (nr > 0 && (it < 4 || (it & (it - 1)) == 0))
      rebalance();
  }
  grid.sync();
  // owners (team 0) keep, ratio after ratio, the plane that explains the most voxels (:160-170)
  {
    const long long per = (((A.Nv + gridDim.x - 1) / gridDim.x) + 31) / 32 * 32;
    v0 = min((long long)blockIdx.x * per, A.Nv);
    v1 = min(v0 + per, A.Nv);
  }
  if (v1 > v0) {
    const long long RC = (long long)A.R * A.C;
    const int p_first = A.cidx[v0], p_last = A.cidx[v1 - 1];
    for (int p = p_first + threadIdx.x; p <= p_last; p += kRansacThreads) {
      const long long ps = A.seg_start[p];
      if (!(ps >= v0 && ps < v1 && A.seg_start[p + 1] > ps)) continue;
      float bc = A.best_conf[p];
      for (int r = 0; r < A.R; r++) {
        const long long rp = (long long)r * A.C + p;
        const float nh = (float)A.nhit[(long long)A.fin[r * 2 + 1] * RC + rp];
        if (bc < nh) {
          bc = nh;
          const float *pl = A.planes + ((long long)A.fin[r * 2 + 0] * RC + rp) * 6;
          for (int k = 0; k < 3; k++) {
            A.best_center[p * 3 + k] = pl[k];
            A.best_normal[p * 3 + k] = pl[3 + k];
          }
        }
      }
      A.best_conf[p] = bc;
    }
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
                      int C, int n_ratios, float sigma2, float stopping_delta, int max_iter, double *acc,
                      int32_t *nhit, uint32_t *gmax, float *planes, int32_t *fin, float *best_center,
                      float *best_normal, float *best_conf, int32_t *iters_out) {
  if (Nv < 0 || C < 1 || n_ratios < 1 || n_ratios > kMaxRatios || ((uintptr_t)vox & 15) || !cidx || !seg_start ||
      !acc || !nhit || !gmax || !planes || !fin || !iters_out)
    return set_error(PCS_ERR_BAD_ARG, "pcs_ground_ransac: bad args (1 <= n_ratios <= 32)");
  RansacArgs A;
  A.vox = (const float4 *)vox;
  A.cidx = cidx;
  A.seg_start = seg_start;
  A.origin = origin;
  A.cmin_z = cmin_z;
  A.cmax_z = cmax_z;
  A.ratios = ratios;
  A.acc = acc;
  A.nhit = nhit;
  A.gmax = gmax;
  A.planes = planes;
  A.fin = fin;
  A.best_center = best_center;
  A.best_normal = best_normal;
  A.best_conf = best_conf;
  A.iters_out = iters_out;
  A.Nv = Nv;
  A.C = C;
  A.R = n_ratios;
  A.sigma2 = sigma2;
  A.stopping_delta = stopping_delta;
  A.max_iter = max_iter;
  {
    const char *e = getenv("PCS_RANSAC_PREFETCH");
    A.prefetch = e ? atoi(e) : 0;
  }
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // 3 CTAs / SM at 80 registers (616 B of spills, all in the per-iteration plane-fit phase) or 2 CTAs / SM with 128
  // registers and no spills (PCS_RANSAC_OCC=2).  Measured on B200, 198 frames: 37.4 ms vs 41.1 ms -- the voxel sweep
  // is latency bound and the third resident CTA is worth more than the spills cost, so 3 stays the default.
  static int occ = -1;
  if (occ < 0) {
    const char *o = getenv("PCS_RANSAC_OCC");
    occ = (o && atoi(o) == 2) ? 2 : 3;
  }
  const void *kern = occ == 3 ? (const void *)ground_ransac_kernel<3> : (const void *)ground_ransac_kernel<2>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRansacThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > occ) per_sm = occ;
  const int n_teams = (n_ratios + kG - 1) / kG;
  long long bpt = (Nv + 4095) / 4096;  // at least ~4k voxels per block
  long long cap = ((long long)sms * per_sm) / n_teams;
  if (bpt > cap) bpt = cap;
  if (bpt < 1) bpt = 1;
  long long blocks = bpt * n_teams;
  if (blocks < n_ratios) blocks = n_ratios;  // the in-kernel team assignment needs one block per active ratio
  if (blocks > kMaxRansacGrid) blocks = kMaxRansacGrid;
  void *args[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks),
                                              dim3(kRansacThreads), args, 0, as_stream(s));
  g_launches++;
  if (e != cudaSuccess) return set_error((int)e, "ground_ransac_kernel (cooperative launch)");
  return check_launch("ground_ransac_kernel");
}

#ifdef PCS_RANSAC_TRACE
int pcs_debug_ransac_trace(unsigned long long *host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_ransac_trace, sizeof(unsigned long long) * 2048 * 4);
}
#endif

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

// ------------------------------------------------------------------------------------------------
// Curvature pruning of the super-pillar planes ("Truncated Least Squares", preprocessor_utils.py:175-193):
// for 100 decreasing thresholds, k-nearest-neighbour curvature of every surviving plane centre, drop the planes
// whose mean curvature is not below the threshold.  One persistent CTA; the reference runs ~25 launches and a
// blocking .max() per threshold.
// ------------------------------------------------------------------------------------------------
namespace pcs {

constexpr int kPruneThreads = 64;
constexpr int kPruneMaxK = 16;
constexpr int kPruneMaxN = 8192;

// one selection pass of the warp over the centres: smallest (d2 bits, index) pair greater than (last_d, last_j),
// optionally restricted to active centres.  Returns index (0x7fffffff if none) and its d2 bits in wd.
__device__ __forceinline__ int prune_next_nearest(const float *sx, const float *sy, const float *sz, const int *active,
                                                  int n, float xi, float yi, float zi, unsigned int last_d, int last_j,
                                                  int lane, unsigned int &wd) {
  unsigned int bd = 0xffffffffu;
  int bj = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    if (active && !active[j]) continue;
    const float dx = sx[j] - xi, dy = sy[j] - yi, dz = sz[j] - zi;
    const unsigned int db = __float_as_uint(dx * dx + dy * dy + dz * dz);
    const bool after = db > last_d || (db == last_d && j > last_j);
    if (after && (db < bd || (db == bd && j < bj))) {
      bd = db;
      bj = j;
    }
  }
  wd = __reduce_min_sync(0xffffffffu, bd);
  return __reduce_min_sync(0xffffffffu, bd == wd ? bj : 0x7fffffff);
}

__device__ __forceinline__ float prune_term(const float *sx, const float *sy, const float *sz, int j, float xi, float yi,
                                            float zi, float nx, float ny, float nz) {
  const float dx = sx[j] - xi, dy = sy[j] - yi, dz = sz[j] - zi;
  return fabsf(dx * nx + dy * ny + dz * nz) / (sqrtf(dx * dx + dy * dy + dz * dz) + 1e-4f);
}

// state (global, zero-initialised by the caller): [0] / [2] max-curvature bits of even / odd evaluation rounds (the
// slot of the next round is cleared while nobody uses it), [1] removed count.
// Every warp owns `pw` planes.  For each of them it first builds the list of its 32 nearest centres (sorted by
// (d2, index), the order the reference's knn returns them in): as long as K of those are still active, the K nearest
// ACTIVE centres are the first K active entries of the list, so a round costs one ballot per plane instead of a scan
// over all centres.  A plane that has lost too many of its 32 falls back to the full selection scan.
__global__ void __launch_bounds__(kPruneThreads) plane_prune_kernel(const float *__restrict__ xyz,
                                                                    const float *__restrict__ normal, int n, int K,
                                                                    const float *__restrict__ thresholds, int n_thr,
                                                                    int *__restrict__ keep, float *__restrict__ curv,
                                                                    unsigned int *__restrict__ state, int pw) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float sm[];  // x[n], y[n], z[n], active[n], cand[2][pw][32], c[2][pw]
  float *sx = sm, *sy = sm + n, *sz = sm + 2 * n;
  int *active = reinterpret_cast<int *>(sm + 3 * n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *cand = reinterpret_cast<int *>(sm + 4 * n) + warp * pw * 32;
  float *cq = sm + 4 * n + (kPruneThreads / 32) * pw * 32 + warp * pw;
  const int first = (blockIdx.x * (kPruneThreads / 32) + warp) * pw;  // my planes: [first, first + pw)
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    sx[j] = xyz[j * 3 + 0];
    sy[j] = xyz[j * 3 + 1];
    sz[j] = xyz[j * 3 + 2];
    active[j] = 1;
  }
  for (int q = lane; q < pw; q += 32)
    if (first + q < n) keep[first + q] = 1;
  __syncthreads();
  for (int q = 0; q < pw; q++) {
    const int i = first + q;
    if (i >= n) break;
    const float xi = sx[i], yi = sy[i], zi = sz[i];
    unsigned int last_d = 0u;
    int last_j = -1, mine = -1;
    for (int k = 0; k < 32; k++) {
      unsigned int wd;
      const int wj = prune_next_nearest(sx, sy, sz, nullptr, n, xi, yi, zi, last_d, last_j, lane, wd);
      if (wj == 0x7fffffff) break;  // fewer than 32 centres
      if (lane == k) mine = wj;
      last_d = wd;
      last_j = wj;
    }
    cand[q * 32 + lane] = mine;
  }
  __syncwarp();
  int count = n, round = 0;
  bool changed = true;
  float maxc = 0.f;
  for (int t = 0; t < n_thr; t++) {
    if (count < K) break;  // fewer planes than neighbours (grid-uniform)
    if (changed) {
      float m = 0.f;
      for (int q = 0; q < pw; q++) {
        const int i = first + q;
        if (i >= n) break;
        if (!active[i]) continue;
        const float xi = sx[i], yi = sy[i], zi = sz[i];
        const float nx = normal[i * 3 + 0], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
        const int j = cand[q * 32 + lane];
        const bool a = j >= 0 && active[j];
        unsigned int ball = __ballot_sync(0xffffffffu, a);
        float acc = 0.f;
        if (__popc(ball) >= K) {
          const float term = a ? prune_term(sx, sy, sz, j, xi, yi, zi, nx, ny, nz) : 0.f;
          for (int k = 0; k < K; k++) {  // ascending (d2, index), the summation order of the scan below
            acc += __shfl_sync(0xffffffffu, term, __ffs(ball) - 1);
            ball &= ball - 1;
          }
        } else {
          unsigned int last_d = 0u;
          int last_j = -1;
          for (int k = 0; k < K; k++) {
            unsigned int wd;
            const int wj = prune_next_nearest(sx, sy, sz, active, n, xi, yi, zi, last_d, last_j, lane, wd);
            last_d = wd;
            last_j = wj;
            acc += prune_term(sx, sy, sz, wj, xi, yi, zi, nx, ny, nz);
          }
        }
        const float c = acc / (float)K;
        if (lane == 0) {
          cq[q] = c;
          curv[i] = c;
        }
        m = fmaxf(m, c);
      }
      __syncwarp();
      unsigned int *slot = state + ((round & 1) ? 2 : 0);
      if (lane == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
      grid.sync();
      maxc = __uint_as_float(__ldcg(slot));
      changed = false;
      ++round;
    }
    if (thresholds[t] > maxc) continue;  // :186-187; nothing changes, no barrier needed
    // remove every plane whose curvature is not below the threshold (at least the arg-max goes)
    const float thr = thresholds[t];
    bool drop = false;
    for (int q = lane; q < pw; q += 32) {
      const int i = first + q;
      if (i < n && active[i] && !(cq[q] < thr)) {
        keep[i] = 0;
        drop = true;
      }
    }
    // (a lane drops at most one plane per 32 owned; pw <= 32)
    const unsigned int dropped = __ballot_sync(0xffffffffu, drop);
    if (lane == 0 && dropped) atomicAdd(state + 1, (unsigned int)__popc(dropped));
    if (blockIdx.x == 0 && threadIdx.x == 0) state[(round & 1) ? 2 : 0] = 0u;  // slot of the NEXT evaluation round
    grid.sync();
    count = n - (int)__ldcg(state + 1);
    for (int j = threadIdx.x; j < n; j += blockDim.x) active[j] = __ldcg(keep + j);
    __syncthreads();
    changed = true;
  }
}

}  // namespace pcs

extern "C" int pcs_plane_prune(pcs_stream_t s, const float *xyz, const float *normal, int n, int K,
                               const float *thresholds, int n_thr, int32_t *keep, float *curv, uint32_t *state) {
  using namespace pcs;
  if (n < 1 || n > kPruneMaxN || K < 1 || K > kPruneMaxK || !xyz || !normal || !thresholds || !keep || !curv || !state)
    return set_error(PCS_ERR_BAD_ARG, "pcs_plane_prune: bad args (n <= 8192, K <= 16)");  // state: uint32[4]
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int wpb = kPruneThreads / 32;
  int pw = (n + sms * wpb - 1) / (sms * wpb);  // planes per warp, one CTA per SM at most
  if (pw < 1) pw = 1;
  if (pw > 32) return set_error(PCS_ERR_BAD_ARG, "pcs_plane_prune: too many planes for this device");
  int blocks = (n + pw * wpb - 1) / (pw * wpb);
  size_t smem = ((size_t)n * 4 + (size_t)wpb * pw * 33) * sizeof(float);
  cudaFuncSetAttribute(plane_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  void *args[] = {(void *)&xyz, (void *)&normal, (void *)&n, (void *)&K, (void *)&thresholds, (void *)&n_thr,
                  (void *)&keep, (void *)&curv, (void *)&state, (void *)&pw};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)plane_prune_kernel, dim3(blocks), dim3(kPruneThreads), args, smem,
                                              as_stream(s));
  g_launches++;
  if (e != cudaSuccess) return set_error((int)e, "plane_prune_kernel (cooperative launch)");
  return check_launch("plane_prune_kernel");
}

// ------------------------------------------------------------------------------------------------
// Velocity smoothing of the tracker (cluster_tracking.py:162-199): AdamW on the per-component xy velocities over
// the frames [a, b]: loss = w0 * mean (v - d)^2 + w * mean |v[f] - v[f+1]|, lr 1e-2 with MultiStepLR([100,200,300]),
// 3-strike stopping rule with a blocking loss.item() per iteration in the reference; here one persistent CTA.
// torch.optim.AdamW decays EVERY element of the parameter tensor (also the ones without gradient); those only see
// the accumulated factor prod(1 - lr_t * wd), applied once at the end.
// ------------------------------------------------------------------------------------------------
namespace pcs {

struct VeloArgs {
  float *velos;        // [C][F][3] in/out
  const float *diffs;  // [C][F][3]
  float *m;            // [C][nf][2] scratch (zeroed)
  float *v;            // [C][nf][2] scratch (zeroed)
  int *info;           // [2]: iterations run, stopped early
  int C, F, a, b;
  float w0, w;
  int num_itr;
  float stopping;
};

__global__ void __launch_bounds__(1024) smooth_velo_kernel(VeloArgs A) {
  __shared__ double red[32];
  __shared__ double s_loss;
  const int nf = A.b - A.a + 1;
  const int S = A.C * nf * 2;  // optimised elements: component, frame in [a, b], x/y
  const float n1 = (float)S, n2 = (float)(A.C * (nf - 1) * 2);
  float lr = 1e-2f;
  const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, wd = 1e-2f;
  double b1t = 1.0, b2t = 1.0, decay = 1.0;
  double last_loss = 1e10;
  int countdown = 3, it = 0;
#define VEL(c, f, k) A.velos[((long long)(c) * A.F + (f)) * 3 + (k)]
  for (; it < A.num_itr; it++) {
    double lsum = 0.0;
    float g[24];
    int nc = 0;
    for (int e = threadIdx.x; e < S; e += blockDim.x, nc++) {
      const int k = e & 1, fi = (e >> 1) % nf, c = (e >> 1) / nf, f = A.a + fi;
      const float vv = VEL(c, f, k);
      const float r = vv - A.diffs[((long long)c * A.F + f) * 3 + k];
      float grad = A.w0 * 2.f * r / n1;
      double l = (double)A.w0 * r * r / n1;
      if (f < A.b) {
        const float d = vv - VEL(c, f + 1, k);
        grad += A.w * ((d > 0.f) - (d < 0.f)) / n2;
        l += (double)A.w * fabsf(d) / n2;
      }
      if (f > A.a) {
        const float d = VEL(c, f - 1, k) - vv;
        grad -= A.w * ((d > 0.f) - (d < 0.f)) / n2;
      }
      g[nc] = grad;
      lsum += l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    __syncthreads();  // all reads of the velocities are done
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
      s_loss = t;
    }
    b1t *= beta1;
    b2t *= beta2;
    const float bc1 = (float)(1.0 - b1t), bc2s = (float)sqrt(1.0 - b2t);
    nc = 0;
    for (int e = threadIdx.x; e < S; e += blockDim.x, nc++) {
      const int k = e & 1, fi = (e >> 1) % nf, c = (e >> 1) / nf, f = A.a + fi;
      const float grad = g[nc];
      float p = VEL(c, f, k) * (1.f - lr * wd);
      const float m = beta1 * A.m[e] + (1.f - beta1) * grad;
      const float v = beta2 * A.v[e] + (1.f - beta2) * grad * grad;
      A.m[e] = m;
      A.v[e] = v;
      p -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
      VEL(c, f, k) = p;
    }
    decay *= (double)(1.f - lr * wd);
    if (it + 1 == 100 || it + 1 == 200 || it + 1 == 300) lr *= 0.1f;  // MultiStepLR after scheduler.step()
    __syncthreads();
    const double loss = (double)(float)s_loss;
    if (last_loss - loss < (double)A.stopping) countdown -= 1;
    else countdown = 3;
    if (countdown <= 0) {
      ++it;
      break;
    }
    last_loss = loss;
  }
  // elements without gradient: weight decay only
  const long long total = (long long)A.C * A.F * 3;
  const float dec = (float)decay;
  for (long long e = threadIdx.x; e < total; e += blockDim.x) {
    const int k = (int)(e % 3), f = (int)((e / 3) % A.F);
    if (k < 2 && f >= A.a && f <= A.b) continue;
    A.velos[e] *= dec;
  }
#undef VEL
  if (threadIdx.x == 0) {
    A.info[0] = it;
    A.info[1] = countdown <= 0;
  }
}

}  // namespace pcs

extern "C" int pcs_smooth_velo(pcs_stream_t s, float *velos, const float *diffs, float *m, float *v, int C, int F,
                               int a, int b, float w0, float w, int num_itr, float stopping, int32_t *info) {
  if (!velos || !diffs || !m || !v || !info || C < 1 || a < 0 || b >= F || a >= b ||
      (long long)C * (b - a + 1) * 2 > 24LL * 1024)
    return pcs::set_error(PCS_ERR_BAD_ARG, "pcs_smooth_velo: bad args (a < b < F, C * (b-a+1) * 2 <= 24576)");
  pcs::VeloArgs A;
  A.velos = velos;
  A.diffs = diffs;
  A.m = m;
  A.v = v;
  A.info = info;
  A.C = C;
  A.F = F;
  A.a = a;
  A.b = b;
  A.w0 = w0;
  A.w = w;
  A.num_itr = num_itr;
  A.stopping = stopping;
  PCS_LAUNCH(pcs::smooth_velo_kernel, 1, 1024, 0, pcs::as_stream(s), A);
  return 0;
}
