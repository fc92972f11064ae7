// icp.cu -- per-cluster trimmed two-way ICP (TLS registration) as ONE persistent cooperative kernel (sm_100a).
//
// Replaces register_to_next_frame (pcdet/models/registration/preprocessors/registration_utils.py:83-206):
// up to max_iter iterations of
//   two radius graphs with K = 1 (moving -> ref and ref -> moving, :131-138; each a voxel-hash build + search),
//   per-component centroids / centred covariance (:150-164), rotation regulariser (:165), batched 3x3 SVD in
//   fp64 (:167-174), transform accumulation and point update (:175-179), loss-based 3-strike stopping rule
//   (:180-186),
// followed by the truncated-mean residual of the last iteration (:156) and the matched-fraction graph (:189-199).
// The reference spends ~150 launches and >= 8 host syncs per iteration; here the whole loop runs on the device
// with grid-wide barriers between phases and the stopping rule evaluated in-kernel.
//
// The reference grid (over the static ref points) is built once by pcs_hash_build; the moving points move every
// iteration, so their grid (same cell geometry) is rebuilt in-kernel: insert/count -> range assignment ->
// scatter.  Per-component sums are fp64 raw moments accumulated with one 17-lane atomic instruction per edge;
// centroids / covariance / loss follow from them exactly.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pcs {

constexpr int kMom = 17;  // n, sum m (3), sum r (3), sum m r^T (9), sum |m - r|^2
constexpr int kIcpThreads = 256;

struct IcpGridGeom {
  float lo[4];
  float vs[4];
  long long dims[4];
};

struct IcpArgs {
  // geometry shared by all three grids (device arrays written by pcs_grid_params, voxel size by value)
  const float *lo_d;        // [4]
  const long long *dims_d;  // [4]
  float vs[4];
  // static grid over the non-stationary ref points (targets of the forward search)
  const pcs_slot_t *ref_table;
  long long ref_mask;
  const float4 *ref_sorted;
  const int *ref_sidx;
  // static grid over ALL ref points (matched-fraction graph after the loop)
  const pcs_slot_t *all_table;
  long long all_mask;
  const float4 *all_sorted;
  const int *all_sidx;
  // moving grid, rebuilt every iteration
  pcs_slot_t *mov_table;
  long long mov_mask;  // H - 1
  float4 *mov_sorted;
  int *mov_sidx;
  int *counters;  // [4]: [1] scatter cursor, [2] error flag
  // points
  float4 *mov;          // [nm] non-stationary moving points (frame, x, y, z), updated in place
  const int *mov_comp;  // [nm]
  const float4 *ref;    // [nr] non-stationary ref points
  int nm, nr, C;
  int df;  // frame offset ref - moving
  float radius;
  double angle_reg;
  int max_iter;
  double stopping_delta;
  // per-iteration state
  int *nn_fwd;     // [nm] ref index of the nearest ref point (or -1)
  int *nn_bwd;     // [nr] moving index of the nearest moving point (or -1)
  double *mom;     // [C][kMom]
  double *Ti;      // [C][12] transform of the current iteration (R row-major, t)
  double *T;       // [C][12] accumulated transform (initialised to identity by the caller)
  double *mu;      // [C][6] centroids (moving, ref) of the current iteration
  double *l1_sum;  // [C][2] sum of residual norms / sum of clamped residual norms
  double *state;   // [4]: last_error, loss, -, -
  int *istate;     // [4]: countdown, iterations run, is_last flag, stop flag
  // outputs
  double *l1_err;          // [C]
  int *match_cnt;          // [C] matched moving voxels per component (for comp_edge_ratio)
};

__device__ __forceinline__ long long icp_coord(float p, float lo, float vs) {
  return (long long)rintf(__fdiv_rn(__fsub_rn(p, lo), vs)) + 1;
}

__device__ __forceinline__ long long icp_key(const IcpGridGeom &g, long long c0, long long c1, long long c2,
                                              long long c3) {
  return map2key4(c0, c1, c2, c3, g.dims);
}

__device__ __forceinline__ bool slot_lookup(const pcs_slot_t *table, long long mask, long long key, int &start,
                                            int &count) {
  long long slot = hash_key(key) & mask;
  for (long long probes = 0; probes <= mask; ++probes) {
    const int4 vv = *reinterpret_cast<const int4 *>(table + slot);
    const long long k = ((long long)vv.y << 32) | (unsigned int)vv.x;
    if (k == key) {
      start = vv.z;
      count = vv.w;
      return true;
    }
    if (k == PCS_EMPTY_KEY) return false;
    slot = (slot + 1) & mask;
  }
  return false;
}

// Nearest stored point to `qp` within radius (fp32 4-D distance, reference order), searched by one warp over the
// 27 cells [c + (dfo, -1..1, -1..1, -1..1)]; cells are pruned with the conservative distance bound of search.cu.
// `cursor_mode`: slot.start is the END of the cell range (scatter cursor not rewound) -> rows [start-count, start).
// Returns the original index (or -1); ties broken by ascending index.
__device__ int nn_search_warp(const IcpGridGeom &g, const pcs_slot_t *table, long long mask, const float4 *sorted,
                              const int *sidx, bool cursor_mode, const float4 qp, int dfo, float r2, int lane,
                              float *d2_out) {
  const float u1 = __fdiv_rn(__fsub_rn(qp.y, g.lo[1]), g.vs[1]);
  const float u2 = __fdiv_rn(__fsub_rn(qp.z, g.lo[2]), g.vs[2]);
  const float u3 = __fdiv_rn(__fsub_rn(qp.w, g.lo[3]), g.vs[3]);
  const float n1 = rintf(u1), n2 = rintf(u2), n3 = rintf(u3);
  const long long qc0 = icp_coord(qp.x, g.lo[0], g.vs[0]) + dfo;
  const long long qc1 = (long long)n1 + 1, qc2 = (long long)n2 + 1, qc3 = (long long)n3 + 1;
  int start = 0, count = 0;
  unsigned int sel = 0xffffffffu;
  if (lane < 27) {
    const int o1 = lane % 3 - 1, o2 = (lane / 3) % 3 - 1, o3 = lane / 9 - 1;
    const long long c1 = qc1 + o1, c2 = qc2 + o2, c3 = qc3 + o3;
    const bool clamped = qc0 < 0 || qc0 > g.dims[0] || c1 < 0 || c1 > g.dims[1] || c2 < 0 || c2 > g.dims[2] ||
                         c3 < 0 || c3 > g.dims[3];
    float dmin2 = 0.f;
    if (!clamped) {
      float gp[3];
      const float f[3] = {u1 - n1, u2 - n2, u3 - n3};
      const float uu[3] = {u1, u2, u3};
      const int oo[3] = {o1, o2, o3};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float gg = 0.f;
        if (oo[k] != 0) {
          gg = (oo[k] > 0) ? (0.5f - f[k]) : (0.5f + f[k]);
          gg -= 4e-6f * (fabsf(uu[k]) + 1.0f);
          gg = gg > 0.f ? gg * g.vs[k + 1] : 0.f;
        }
        gp[k] = gg;
      }
      dmin2 = (gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]) * 0.99999f;
    }
    if (dmin2 <= r2) {
      int s = 0, c = 0;
      if (slot_lookup(table, mask, icp_key(g, qc0, c1, c2, c3), s, c) && c > 0) {
        start = cursor_mode ? s - c : s;
        count = c;
        sel = (__float_as_uint(dmin2) & ~31u) | (unsigned int)lane;
      }
    }
  }
  unsigned long long best = ~0ull;
  float bound = r2;
  while (true) {
    const unsigned int pick = __reduce_min_sync(0xffffffffu, sel);
    if (pick == 0xffffffffu) break;
    if (__uint_as_float(pick & ~31u) > bound) break;
    const int src = pick & 31;
    if (lane == src) sel = 0xffffffffu;
    const int cs = __shfl_sync(0xffffffffu, start, src), cc = __shfl_sync(0xffffffffu, count, src);
    for (int j = lane; j < cc; j += 32) {
      const float d2 = dist2_ref(sorted[cs + j], qp);
      if (d2 <= bound) {
        const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)sidx[cs + j];
        best = k < best ? k : best;
      }
    }
    // tighten the bound with the warp's current best
    unsigned long long wb = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, wb, o);
      wb = t < wb ? t : wb;
    }
    best = wb;
    if (wb != ~0ull) bound = fminf(bound, __uint_as_float((unsigned int)(wb >> 32)));
  }
  if (best == ~0ull) return -1;
  if (d2_out) *d2_out = __uint_as_float((unsigned int)(best >> 32));
  return (int)(unsigned int)(best & 0xffffffffu);
}

// 3x3 SVD-based rotation: R = V diag(1,1,det(V U^T)) U^T for A = U S V^T (registration_utils.py:167-174).
// A is well conditioned here (the regulariser adds angle_reg * R_prev), so U = A V S^-1 from the eigen-decomposition
// of A^T A is accurate; R is the unique polar factor whatever the ordering / signs of the decomposition.
__device__ void kabsch_rotation(const double A[9], double R[9]) {
  // B = A^T A (symmetric), eigen-decomposition B = V diag(s^2) V^T
  double b00 = 0, b01 = 0, b02 = 0, b11 = 0, b12 = 0, b22 = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double x = A[k * 3 + 0], y = A[k * 3 + 1], z = A[k * 3 + 2];
    b00 += x * x;
    b01 += x * y;
    b02 += x * z;
    b11 += y * y;
    b12 += y * z;
    b22 += z * z;
  }
  const Eig3 e = jacobi_eig3(b00, b01, b02, b11, b12, b22);
  double ev[3] = {e.d0, e.d1, e.d2};
  double V[3][3] = {{e.v00, e.v01, e.v02}, {e.v10, e.v11, e.v12}, {e.v20, e.v21, e.v22}};
  // order singular values descending so that the reflection fix lands on the smallest one
  int i0 = 0, i1 = 1, i2 = 2;
  if (ev[i1] > ev[i0]) { int t = i0; i0 = i1; i1 = t; }
  if (ev[i2] > ev[i0]) { int t = i0; i0 = i2; i2 = t; }
  if (ev[i2] > ev[i1]) { int t = i1; i1 = i2; i2 = t; }
  const int idx[3] = {i0, i1, i2};
  double Vs[3][3], U[3][3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double sv = sqrt(fmax(ev[idx[j]], 0.0));
#pragma unroll
    for (int i = 0; i < 3; i++) Vs[i][j] = V[i][idx[j]];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double av = A[i * 3 + 0] * Vs[0][j] + A[i * 3 + 1] * Vs[1][j] + A[i * 3 + 2] * Vs[2][j];
      U[i][j] = sv > 1e-300 ? av / sv : 0.0;
    }
  }
  if (!(ev[i2] > 1e-24 * fmax(ev[i0], 1e-300))) {
    // rank-deficient A: complete U with the cross product of its first two columns
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  // M = V U^T ; d = det(M) is the third sign entry, as in the reference (:172)
  double M[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) M[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
  const double d = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                   M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R[i * 3 + j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + d * Vs[i][2] * U[j][2];
}

// ---- phases -----------------------------------------------------------------------------------------------
__device__ void phase_clear_table(const IcpArgs &A, long long tid, long long nth) {
  const int4 e = make_int4(-1, -1, 0, 0);
  int4 *t = reinterpret_cast<int4 *>(A.mov_table);
  for (long long i = tid; i <= A.mov_mask; i += nth) t[i] = e;
  if (tid == 0) A.counters[1] = 0;
}

// apply the previous iteration's transform (fp64 product stored back to fp32, registration_utils.py:179) and
// count the point into its cell
__device__ void phase_insert(const IcpArgs &A, const IcpGridGeom &geo, bool apply, long long tid, long long nth) {
  for (long long i = tid; i < A.nm; i += nth) {
    float4 p = A.mov[i];
    if (apply) {
      const double *t = A.Ti + (long long)A.mov_comp[i] * 12;
      const double x = p.y, y = p.z, z = p.w;
      p.y = (float)(t[0] * x + t[1] * y + t[2] * z + t[9]);
      p.z = (float)(t[3] * x + t[4] * y + t[5] * z + t[10]);
      p.w = (float)(t[6] * x + t[7] * y + t[8] * z + t[11]);
      A.mov[i] = p;
    }
    const long long key = icp_key(geo, icp_coord(p.x, geo.lo[0], geo.vs[0]), icp_coord(p.y, geo.lo[1], geo.vs[1]),
                                  icp_coord(p.z, geo.lo[2], geo.vs[2]), icp_coord(p.w, geo.lo[3], geo.vs[3]));
    long long slot = hash_key(key) & A.mov_mask;
    bool ok = false;
    for (long long probes = 0; probes <= A.mov_mask; ++probes) {
      const long long cur = *((volatile long long *)&A.mov_table[slot].key);
      if (cur == key) {
        ok = true;
        break;
      }
      if (cur == PCS_EMPTY_KEY) {
        const unsigned long long prev = atomicCAS((unsigned long long *)&A.mov_table[slot].key,
                                                  (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
        if (prev == (unsigned long long)PCS_EMPTY_KEY || (long long)prev == key) {
          ok = true;
          break;
        }
      }
      slot = (slot + 1) & A.mov_mask;
    }
    if (ok) atomicAdd(&A.mov_table[slot].count, 1);
    else atomicExch(&A.counters[2], PCS_ERR_TABLE_FULL);
  }
}

__device__ void phase_ranges(const IcpArgs &A, long long tid, long long nth) {
  for (long long i = tid; i <= A.mov_mask; i += nth) {
    const int c = A.mov_table[i].count;
    if (c > 0) A.mov_table[i].start = atomicAdd(&A.counters[1], c);
  }
}

__device__ void phase_scatter(const IcpArgs &A, const IcpGridGeom &geo, long long tid, long long nth) {
  for (long long i = tid; i < A.nm; i += nth) {
    const float4 p = A.mov[i];
    const long long key = icp_key(geo, icp_coord(p.x, geo.lo[0], geo.vs[0]), icp_coord(p.y, geo.lo[1], geo.vs[1]),
                                  icp_coord(p.z, geo.lo[2], geo.vs[2]), icp_coord(p.w, geo.lo[3], geo.vs[3]));
    long long slot = hash_key(key) & A.mov_mask;
    long long probes = 0;
    while (*((volatile long long *)&A.mov_table[slot].key) != key && probes <= A.mov_mask) {
      slot = (slot + 1) & A.mov_mask;
      ++probes;
    }
    if (probes > A.mov_mask) continue;  // insertion failed (flagged in counters[2])
    const int pos = atomicAdd(&A.mov_table[slot].start, 1);  // start becomes the END of the range (cursor mode)
    A.mov_sorted[pos] = p;
    A.mov_sidx[pos] = (int)i;
  }
}

// both nearest-neighbour searches of the iteration + raw moments of the edge set (one warp per query)
__device__ void phase_search(const IcpArgs &A, const IcpGridGeom &geo, float r2, long long warp_id, long long nwarps, int lane) {
  const long long total = (long long)A.nm + A.nr;
  for (long long w = warp_id; w < total; w += nwarps) {
    int mi, ri;
    float4 mp, rp;
    if (w < A.nm) {  // forward: moving point -> nearest ref (frame digit + df)
      mi = (int)w;
      mp = A.mov[mi];
      ri = nn_search_warp(geo, A.ref_table, A.ref_mask, A.ref_sorted, A.ref_sidx, false, mp, A.df, r2, lane, nullptr);
      if (lane == 0) A.nn_fwd[mi] = ri;
      if (ri < 0) continue;
      rp = A.ref[ri];
    } else {  // backward: ref point -> nearest moving (frame digit - df)
      ri = (int)(w - A.nm);
      rp = A.ref[ri];
      mi = nn_search_warp(geo, A.mov_table, A.mov_mask, A.mov_sorted, A.mov_sidx, true, rp, -A.df, r2, lane, nullptr);
      if (lane == 0) A.nn_bwd[ri] = mi;
      if (mi < 0) continue;
      mp = A.mov[mi];
    }
    const int c = A.mov_comp[mi];
    const double m[3] = {mp.y, mp.z, mp.w}, r[3] = {rp.y, rp.z, rp.w};
    double v = 0.0;
    if (lane == 0) v = 1.0;
    else if (lane < 4) v = m[lane - 1];
    else if (lane < 7) v = r[lane - 4];
    else if (lane < 16) v = m[(lane - 7) / 3] * r[(lane - 7) % 3];
    else if (lane == 16) {
      const double dx = m[0] - r[0], dy = m[1] - r[1], dz = m[2] - r[2];
      v = dx * dx + dy * dy + dz * dz;
    }
    if (lane < kMom) atomicAdd(A.mom + (long long)c * kMom + lane, v);
  }
}

// per-component solve (thread per component) -- centroids, covariance, regulariser, rotation, transform update
__device__ double phase_solve(const IcpArgs &A, long long tid, long long nth) {
  double loss_part = 0.0;
  for (long long c = tid; c < A.C; c += nth) {
    double *s = A.mom + c * kMom;
    const double n = s[0];
    double mu_m[3] = {0, 0, 0}, mu_r[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (n > 0.5) {
      // the reference takes fp32 segment means and casts them to fp64 (:150-151)
      for (int k = 0; k < 3; k++) {
        mu_m[k] = (double)(float)(s[1 + k] / n);
        mu_r[k] = (double)(float)(s[4 + k] / n);
      }
      // sum (m - mu_m)(r - mu_r)^T = S_mr - mu_m S_r^T - S_m mu_r^T + n mu_m mu_r^T
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          cov[i * 3 + j] = (s[7 + i * 3 + j] - mu_m[i] * s[4 + j] - s[1 + i] * mu_r[j] + n * mu_m[i] * mu_r[j]) / n;
      // sum |(m - mu_m) - (r - mu_r)|^2 = S_|m-r|^2 - 2 d.(S_m - S_r) + n |d|^2,  d = mu_m - mu_r
      double dd = 0.0, cross = 0.0;
      for (int k = 0; k < 3; k++) {
        const double d = mu_m[k] - mu_r[k];
        dd += d * d;
        cross += d * (s[1 + k] - s[4 + k]);
      }
      loss_part += s[16] - 2.0 * cross + n * dd;
    }
    double *T = A.T + c * 12;
    double Amat[9];
    for (int k = 0; k < 9; k++) Amat[k] = cov[k] + A.angle_reg * T[k];  // :165 regulariser on the accumulated R
    double R[9];
    kabsch_rotation(Amat, R);
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = mu_r[i] - (R[i * 3 + 0] * mu_m[0] + R[i * 3 + 1] * mu_m[1] + R[i * 3 + 2] * mu_m[2]);
    double *Ti = A.Ti + c * 12;
    for (int k = 0; k < 9; k++) Ti[k] = R[k];
    for (int k = 0; k < 3; k++) Ti[9 + k] = t[k];
    // T <- Ti @ T  (:178)
    double Rn[9], tn[3];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Rn[i * 3 + j] = R[i * 3 + 0] * T[0 * 3 + j] + R[i * 3 + 1] * T[1 * 3 + j] + R[i * 3 + 2] * T[2 * 3 + j];
      tn[i] = R[i * 3 + 0] * T[9] + R[i * 3 + 1] * T[10] + R[i * 3 + 2] * T[11] + t[i];
    }
    for (int k = 0; k < 9; k++) T[k] = Rn[k];
    for (int k = 0; k < 3; k++) T[9 + k] = tn[k];
    for (int k = 0; k < 3; k++) {
      A.mu[c * 6 + k] = mu_m[k];
      A.mu[c * 6 + 3 + k] = mu_r[k];
    }
    A.l1_sum[c * 2 + 0] = 0.0;
    A.l1_sum[c * 2 + 1] = 0.0;
    A.l1_err[c] = n;  // edge count, turned into the truncated mean by the last-iteration passes
    for (int k = 0; k < kMom; k++) s[k] = 0.0;
  }
  return loss_part;
}

// residual norm |P - Q| of one edge with the centroids of this iteration (positions before the update)
__device__ __forceinline__ double edge_residual(const IcpArgs &A, int mi, int ri, int &c) {
  const float4 mp = A.mov[mi], rp = A.ref[ri];
  c = A.mov_comp[mi];
  const double *mu = A.mu + (long long)c * 6;
  const double dx = ((double)mp.y - mu[0]) - ((double)rp.y - mu[3]);
  const double dy = ((double)mp.z - mu[1]) - ((double)rp.z - mu[4]);
  const double dz = ((double)mp.w - mu[2]) - ((double)rp.w - mu[5]);
  return sqrt(dx * dx + dy * dy + dz * dz);
}

// pass 0: sum of residuals per component; pass 1: sum of residuals clamped to mean +- 0.3 (truncated_robust_mean)
__device__ void phase_l1(const IcpArgs &A, int pass, long long tid, long long nth) {
  const long long total = (long long)A.nm + A.nr;
  for (long long e = tid; e < total; e += nth) {
    int mi, ri;
    if (e < A.nm) {
      mi = (int)e;
      ri = A.nn_fwd[mi];
    } else {
      ri = (int)(e - A.nm);
      mi = A.nn_bwd[ri];
    }
    if (mi < 0 || ri < 0) continue;
    int c;
    double d = edge_residual(A, mi, ri, c);
    if (pass == 1) {
      const double n = A.l1_err[c];
      const double mean = A.l1_sum[c * 2 + 0] / (n > 0.5 ? n : 1.0);
      d = fmin(fmax(d, mean - 0.3), mean + 0.3);
    }
    atomicAdd(A.l1_sum + (long long)c * 2 + pass, d);
  }
}

__global__ void __launch_bounds__(kIcpThreads) icp_register_kernel(IcpArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_red[kIcpThreads / 32];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const long long warp_id = tid >> 5, nwarps = nth >> 5;
  const float r2 = __fmul_rn(A.radius, A.radius);
  IcpGridGeom geo;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    geo.lo[k] = A.lo_d[k];
    geo.vs[k] = A.vs[k];
    geo.dims[k] = A.dims_d[k];
  }

  int it = 0;
  bool apply = false;
  for (; it < A.max_iter; it++) {
    phase_clear_table(A, tid, nth);
    grid.sync();
    phase_insert(A, geo, apply, tid, nth);
    grid.sync();
    phase_ranges(A, tid, nth);
    grid.sync();
    phase_scatter(A, geo, tid, nth);
    grid.sync();
    phase_search(A, geo, r2, warp_id, nwarps, lane);
    grid.sync();
    // solve + loss (block-reduced, one atomic per block)
    double part = phase_solve(A, tid, nth);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kIcpThreads / 32; w++) t += s_red[w];
      if (t != 0.0) atomicAdd(A.state + 1, t);
    }
    grid.sync();
    // stopping rule (registration_utils.py:180-186), evaluated identically by every thread
    const double loss = A.state[1];
    const double last = A.state[0];
    int countdown = A.istate[0];
    if (last - loss < A.stopping_delta) countdown -= 1;
    else countdown = 3;
    const bool stop = countdown <= 0;
    const bool is_last = stop || (it + 1 == A.max_iter);
    grid.sync();  // everyone has read state before it is overwritten
    if (tid == 0) {
      A.istate[0] = countdown;
      A.istate[1] = it + 1;
      A.state[0] = loss;
      A.state[1] = 0.0;
    }
    apply = true;
    if (is_last) {
      // truncated mean residual of this (last) iteration, computed on the positions before the update (:156)
      phase_l1(A, 0, tid, nth);
      grid.sync();
      phase_l1(A, 1, tid, nth);
      grid.sync();
      for (long long c = tid; c < A.C; c += nth) {
        const double n = A.l1_err[c];
        A.l1_err[c] = n > 0.5 ? A.l1_sum[c * 2 + 1] / n : 0.0;
      }
      ++it;
      break;
    }
  }
  // apply the last transform to the points (the reference moves the points before it breaks, :179)
  if (apply) {
    for (long long i = tid; i < A.nm; i += nth) {
      float4 p = A.mov[i];
      const double *t = A.Ti + (long long)A.mov_comp[i] * 12;
      const double x = p.y, y = p.z, z = p.w;
      p.y = (float)(t[0] * x + t[1] * y + t[2] * z + t[9]);
      p.z = (float)(t[3] * x + t[4] * y + t[5] * z + t[10]);
      p.w = (float)(t[6] * x + t[7] * y + t[8] * z + t[11]);
      A.mov[i] = p;
    }
  }
  grid.sync();
  // matched fraction: moving voxels with ANY ref voxel (stationary included) within the radius (:189-199)
  for (long long w = warp_id; w < A.nm; w += nwarps) {
    const float4 mp = A.mov[w];
    const int ri = nn_search_warp(geo, A.all_table, A.all_mask, A.all_sorted, A.all_sidx, false, mp, A.df,
                                  r2, lane, nullptr);
    if (lane == 0 && ri >= 0) atomicAdd(A.match_cnt + A.mov_comp[w], 1);
  }
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_register_icp(pcs_stream_t s, const float *geo_lo, const float *geo_vs, const int64_t *geo_dims,
                     const pcs_slot_t *ref_table, int64_t ref_H, const float *ref_sorted, const int32_t *ref_sidx,
                     const pcs_slot_t *all_table, int64_t all_H, const float *all_sorted, const int32_t *all_sidx,
                     pcs_slot_t *mov_table,
                     int64_t mov_H, float *mov_sorted, int32_t *mov_sidx, int32_t *counters, float *mov,
                     const int32_t *mov_comp, const float *ref, int nm, int nr, int C, int df, float radius,
                     double angle_reg, int max_iter, double stopping_delta, int32_t *nn_fwd, int32_t *nn_bwd,
                     double *mom, double *Ti, double *T, double *mu, double *l1_sum, double *state, int32_t *istate,
                     double *l1_err, int32_t *match_cnt) {
  if (nm < 0 || nr < 0 || C < 1 || (ref_H & (ref_H - 1)) || (all_H & (all_H - 1)) || (mov_H & (mov_H - 1)) || mov_H < 2 ||
      ((uintptr_t)mov & 15) || ((uintptr_t)ref & 15) || ((uintptr_t)mov_sorted & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_register_icp: bad args");
  IcpArgs A;
  A.lo_d = geo_lo;
  A.dims_d = (const long long *)geo_dims;
  for (int i = 0; i < 4; i++) A.vs[i] = geo_vs[i];
  A.ref_table = ref_table;
  A.ref_mask = ref_H - 1;
  A.ref_sorted = (const float4 *)ref_sorted;
  A.ref_sidx = ref_sidx;
  A.all_table = all_table;
  A.all_mask = all_H - 1;
  A.all_sorted = (const float4 *)all_sorted;
  A.all_sidx = all_sidx;
  A.mov_table = mov_table;
  A.mov_mask = mov_H - 1;
  A.mov_sorted = (float4 *)mov_sorted;
  A.mov_sidx = mov_sidx;
  A.counters = counters;
  A.mov = (float4 *)mov;
  A.mov_comp = mov_comp;
  A.ref = (const float4 *)ref;
  A.nm = nm;
  A.nr = nr;
  A.C = C;
  A.df = df;
  A.radius = radius;
  A.angle_reg = angle_reg;
  A.max_iter = max_iter;
  A.stopping_delta = stopping_delta;
  A.nn_fwd = nn_fwd;
  A.nn_bwd = nn_bwd;
  A.mom = mom;
  A.Ti = Ti;
  A.T = T;
  A.mu = mu;
  A.l1_sum = l1_sum;
  A.state = state;
  A.istate = istate;
  A.l1_err = l1_err;
  A.match_cnt = match_cnt;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_register_kernel, kIcpThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  long long work = (long long)nm + nr;
  long long blocks = (work + (kIcpThreads / 32) - 1) / (kIcpThreads / 32);  // one warp per query
  long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void *args[] = {&A};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)icp_register_kernel, dim3((unsigned)blocks), dim3(kIcpThreads), args,
                                              0, as_stream(s));
  g_launches++;
  if (e != cudaSuccess) return set_error((int)e, "icp_register_kernel (cooperative launch)");
  return check_launch("icp_register_kernel");
}

}  // extern "C"
