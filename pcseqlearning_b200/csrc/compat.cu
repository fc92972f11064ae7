// compat.cu -- the four entry points of the reference's torch_hash op on caller-supplied voxel coordinates (sm_100a).
//
// Replaces hash_insert_gpu / radius_graph_gpu / correspondence / points_in_radius_gpu
// (pcdet/ops/torch_hash/src/torch_hash.h:16-32, torch_hash_kernel.cu:54-222, 224-605) for callers that bring their
// OWN integer voxel coordinates (HashTable.find_corres*, the module-style RadiusGraph / ChamferDistance of
// torch_hash_modules.py) instead of letting the grid derive them from the points.  Same semantics as the reference:
//   key = map2key(coord, dims) with the upper clamp == dims_i (:31-47); the cells [coord + qmin, coord + qmax] are
//   visited with dimension 0 fastest (:254-259); fp32 distance over all D columns with one FMA per dimension (:364-368);
//   radius graph accepts d2 <= r*r (:370), points_in_radius d2 < r*r (:206), correspondence takes the nearest row with no
//   radius test (:137-140).
// Unlike the reference's multimap (one slot per point, probe chains as long as a voxel's occupancy) the table holds one
// slot per occupied cell and the point indices are counting-sorted by cell; ties at equal distance are resolved by
// ascending point index (the reference's order is a race).  max_num_neighbors = -1 returns ALL neighbours within the
// radius (the reference counts them but its fill kernel never writes a row for -1, :372-393).
#include "common.cuh"

namespace pcs {

constexpr int kCompatMaxD = 4;

struct CompatDims {
  long long d[kCompatMaxD];
  int nd;
};

struct CompatRange {
  int qmin[kCompatMaxD], range[kCompatMaxD];
  int nc;
};

__device__ __forceinline__ long long compat_key(const long long *c, const CompatDims &D) {
  long long k = 0;
  for (int i = 0; i < D.nd; i++) {
    long long v = c[i];
    v = v < 0 ? 0 : (v > D.d[i] ? D.d[i] : v);
    k = k * D.d[i] + v;
  }
  return k;
}

__device__ __forceinline__ unsigned int compat_find_or_claim(pcs_slot_t *table, unsigned int mask, long long key, bool claim) {
  unsigned int slot = hash_key(key) & mask;
  for (unsigned int probes = 0; probes <= mask; ++probes) {
    const long long cur = *((volatile long long *)&table[slot].key);
    if (cur == key) return slot;
    if (cur == PCS_EMPTY_KEY) {
      if (!claim) return 0xffffffffu;
      const unsigned long long prev =
          atomicCAS((unsigned long long *)&table[slot].key, (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
      if (prev == (unsigned long long)PCS_EMPTY_KEY || (long long)prev == key) return slot;
    }
    slot = (slot + 1) & mask;
  }
  return 0xffffffffu;
}

__global__ void __launch_bounds__(256) compat_clear_kernel(int4 *table, long long H, int *ctr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int4 e = make_int4(-1, -1, 0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) table[i] = e;
  if (blockIdx.x == 0 && threadIdx.x < 4) ctr[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256) compat_count_kernel(const long long *__restrict__ coords, int n, CompatDims D,
                                                           pcs_slot_t *table, unsigned int mask, int *ctr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int slot = compat_find_or_claim(table, mask, compat_key(coords + (long long)i * D.nd, D), true);
  if (slot == 0xffffffffu) atomicExch(&ctr[2], PCS_ERR_TABLE_FULL);
  else atomicAdd(&table[slot].count, 1);
}

__global__ void __launch_bounds__(256) compat_ranges_kernel(pcs_slot_t *table, long long H, int *ctr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) {
    const int c = table[i].count;
    if (c > 0) table[i].start = atomicAdd(&ctr[1], c);
  }
}

// slot.start becomes the END of the cell's rows (rows = [start - count, start))
__global__ void __launch_bounds__(256) compat_scatter_kernel(const long long *__restrict__ coords, int n, CompatDims D,
                                                             pcs_slot_t *table, unsigned int mask, int *__restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int slot = compat_find_or_claim(table, mask, compat_key(coords + (long long)i * D.nd, D), false);
  if (slot == 0xffffffffu) return;
  rows[atomicAdd(&table[slot].start, 1)] = i;
}

__device__ __forceinline__ float compat_d2(const float *a, const float *b, int nd) {
  float acc = 0.f;
  for (int i = 0; i < nd; i++) {
    const float d = __fsub_rn(a[i], b[i]);
    acc = __fmaf_rn(d, d, acc);
  }
  return acc;
}

// visits every stored row of the cells around one query; F(row index, d2)
template <typename F>
__device__ __forceinline__ void compat_visit(const pcs_slot_t *table, unsigned int mask, const int *rows,
                                             const float *values, const CompatDims &D, const CompatRange &R,
                                             const long long *qc, const float *qv, F f) {
  long long c[kCompatMaxD];
  for (int cell = 0; cell < R.nc; cell++) {
    int t = cell;
    for (int i = 0; i < D.nd; i++) {
      c[i] = qc[i] + t % R.range[i] + R.qmin[i];
      t /= R.range[i];
    }
    const unsigned int slot = compat_find_or_claim(const_cast<pcs_slot_t *>(table), mask, compat_key(c, D), false);
    if (slot == 0xffffffffu) continue;
    const int end = table[slot].start, cnt = table[slot].count;
    for (int j = end - cnt; j < end; j++) {
      const int r = rows[j];
      f(r, compat_d2(values + (long long)r * D.nd, qv, D.nd));
    }
  }
}

__global__ void __launch_bounds__(128) compat_degree_kernel(const pcs_slot_t *table, unsigned int mask, const int *rows,
                                                            const float *values, CompatDims D, CompatRange R,
                                                            const long long *qcoords, const float *qvalues, int m,
                                                            const float *radius, int K, int *degree) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const float r2 = __fmul_rn(radius[q], radius[q]);
  int cnt = 0;
  compat_visit(table, mask, rows, values, D, R, qcoords + (long long)q * D.nd, qvalues + (long long)q * D.nd,
               [&](int, float d2) { cnt += d2 <= r2; });
  degree[q] = (K >= 0 && cnt > K) ? K : cnt;
}

// keeps the `deg` smallest (d2, row) pairs of every query, ascending, by insertion into its output segment
__global__ void __launch_bounds__(128) compat_fill_kernel(const pcs_slot_t *table, unsigned int mask, const int *rows,
                                                          const float *values, CompatDims D, CompatRange R,
                                                          const long long *qcoords, const float *qvalues, int m,
                                                          const float *radius, const int *degree,
                                                          const long long *offsets, long long *edges, float *dists) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int cap = degree[q];
  if (cap == 0) return;
  const float r2 = __fmul_rn(radius[q], radius[q]);
  long long *e = edges + offsets[q] * 2;
  float *dd = dists + offsets[q];
  int n = 0;
  compat_visit(table, mask, rows, values, D, R, qcoords + (long long)q * D.nd, qvalues + (long long)q * D.nd,
               [&](int r, float d2) {
                 if (d2 > r2) return;
                 int pos = n < cap ? n : cap;
                 // (d2, row) strictly before the entry at pos - 1 -> shift
                 while (pos > 0 && (d2 < dd[pos - 1] || (d2 == dd[pos - 1] && (long long)r < e[(pos - 1) * 2]))) {
                   if (pos < cap) {
                     dd[pos] = dd[pos - 1];
                     e[pos * 2] = e[(pos - 1) * 2];
                   }
                   pos--;
                 }
                 if (pos < cap) {
                   dd[pos] = d2;
                   e[pos * 2] = r;
                   e[pos * 2 + 1] = q;
                   if (n < cap) n++;
                 }
               });
  for (int i = 0; i < n; i++) e[i * 2 + 1] = q;
}

__global__ void __launch_bounds__(128) compat_corres_kernel(const pcs_slot_t *table, unsigned int mask, const int *rows,
                                                            const float *values, CompatDims D, CompatRange R,
                                                            const long long *qcoords, const float *qvalues, int m,
                                                            long long *corres) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  float best = 1e10f;
  long long arg = -1;
  compat_visit(table, mask, rows, values, D, R, qcoords + (long long)q * D.nd, qvalues + (long long)q * D.nd,
               [&](int r, float d2) {
                 if (d2 < best || (d2 == best && arg >= 0 && r < arg)) {
                   best = d2;
                   arg = r;
                 }
               });
  corres[q] = arg;
}

__global__ void __launch_bounds__(128) compat_in_radius_kernel(const pcs_slot_t *table, unsigned int mask, const int *rows,
                                                               const float *values, CompatDims D, CompatRange R,
                                                               const long long *qcoords, const float *qvalues, int m,
                                                               float r2, long long *visited) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  compat_visit(table, mask, rows, values, D, R, qcoords + (long long)q * D.nd, qvalues + (long long)q * D.nd,
               [&](int r, float d2) {
                 if (d2 < r2) visited[r] = 1;
               });
}

}  // namespace pcs

using namespace pcs;

namespace {

bool make_dims(const int64_t *dims_host, int nd, CompatDims &D) {
  if (nd < 1 || nd > kCompatMaxD || !dims_host) return false;
  D.nd = nd;
  for (int i = 0; i < kCompatMaxD; i++) D.d[i] = i < nd ? dims_host[i] : 1;
  return true;
}

bool make_range(const int *qmin, const int *qmax, int nd, CompatRange &R) {
  if (!qmin || !qmax) return false;
  R.nc = 1;
  for (int i = 0; i < kCompatMaxD; i++) {
    R.qmin[i] = i < nd ? qmin[i] : 0;
    R.range[i] = i < nd ? qmax[i] - qmin[i] + 1 : 1;
    if (R.range[i] < 1) return false;
    R.nc *= R.range[i];
  }
  return true;
}

}  // namespace

extern "C" {

/* hash_insert_gpu (torch_hash.h:16-18): table <- the occupied cells of coords int64[n][nd] (dims (host) int64[nd]);
 * rows int32[n] = point indices grouped by cell.  table: 16-byte slots [H], H a power of two >= 2 * (number of
 * occupied cells) -- H >= 2 n is always enough; ctr int32[4] ([2] = error flag). */
int pcs_compat_hash_insert(pcs_stream_t s, const int64_t *coords, int64_t n, int nd, const int64_t *dims_host,
                           pcs_slot_t *table, int64_t H, int32_t *rows, int32_t *ctr) {
  CompatDims D;
  if (!make_dims(dims_host, nd, D) || n < 0 || n >= (1LL << 31) || !table || H < 2 || (H & (H - 1)) || !ctr ||
      (n > 0 && (!coords || !rows)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_compat_hash_insert: bad args");
  cudaStream_t st = as_stream(s);
  PCS_LAUNCH(compat_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, ctr);
  if (n == 0) return 0;
  PCS_LAUNCH(compat_count_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (const long long *)coords, (int)n, D, table,
             (unsigned int)(H - 1), ctr);
  PCS_LAUNCH(compat_ranges_kernel, grid_for(H, 256, 8), 256, 0, st, table, (long long)H, ctr);
  PCS_LAUNCH(compat_scatter_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (const long long *)coords, (int)n, D, table,
             (unsigned int)(H - 1), rows);
  return 0;
}

/* first pass of radius_graph_gpu (count_radius_graph_degree_kernel, :224-288): degree int32[m] = min(#accepted, K)
 * (K = -1: all).  values float[n][nd] are the inserted points, qcoords / qvalues the queries, radius float[m]. */
int pcs_compat_radius_degree(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows,
                             const float *values, int nd, const int64_t *dims_host, const int64_t *qcoords,
                             const float *qvalues, int64_t m, const int *qmin, const int *qmax, const float *radius,
                             int K, int32_t *degree) {
  CompatDims D;
  CompatRange R;
  if (!make_dims(dims_host, nd, D) || !make_range(qmin, qmax, nd, R) || !table || (H & (H - 1)) || m < 0 ||
      (m > 0 && (!qcoords || !qvalues || !radius || !degree)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_compat_radius_degree: bad args");
  if (m == 0) return 0;
  PCS_LAUNCH(compat_degree_kernel, (unsigned)((m + 127) / 128), 128, 0, as_stream(s), table, (unsigned int)(H - 1), rows,
             values, D, R, (const long long *)qcoords, qvalues, (int)m, radius, K, degree);
  return 0;
}

/* second pass (radius_graph_kernel, :290-409): edges int64[E][2] rows (ref index, query index) grouped by ascending
 * query and ascending (distance, index) inside a query; offsets int64[m + 1] = exclusive scan of degree;
 * dists float[E] scratch / output. */
int pcs_compat_radius_fill(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                           int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                           const int *qmin, const int *qmax, const float *radius, const int32_t *degree,
                           const int64_t *offsets, int64_t *edges, float *dists) {
  CompatDims D;
  CompatRange R;
  if (!make_dims(dims_host, nd, D) || !make_range(qmin, qmax, nd, R) || !table || (H & (H - 1)) || m < 0 ||
      (m > 0 && (!qcoords || !qvalues || !radius || !degree || !offsets)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_compat_radius_fill: bad args");
  if (m == 0) return 0;
  PCS_LAUNCH(compat_fill_kernel, (unsigned)((m + 127) / 128), 128, 0, as_stream(s), table, (unsigned int)(H - 1), rows,
             values, D, R, (const long long *)qcoords, qvalues, (int)m, radius, degree, (const long long *)offsets,
             (long long *)edges, dists);
  return 0;
}

/* correspondence (torch_hash.h:20-22, kernel :96-155): corres int64[m] = nearest inserted row over the visited cells,
 * NO radius test, -1 when all of them are empty. */
int pcs_nn_correspondence(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                          int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                          const int *qmin, const int *qmax, int64_t *corres) {
  CompatDims D;
  CompatRange R;
  if (!make_dims(dims_host, nd, D) || !make_range(qmin, qmax, nd, R) || !table || (H & (H - 1)) || m < 0 ||
      (m > 0 && (!qcoords || !qvalues || !corres)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_nn_correspondence: bad args");
  if (m == 0) return 0;
  PCS_LAUNCH(compat_corres_kernel, (unsigned)((m + 127) / 128), 128, 0, as_stream(s), table, (unsigned int)(H - 1), rows,
             values, D, R, (const long long *)qcoords, qvalues, (int)m, (long long *)corres);
  return 0;
}

/* points_in_radius_gpu (torch_hash.h:29-32, kernel :160-222): visited[row] = 1 for every inserted row with
 * d2 < radius^2 (strict) of any query. */
int pcs_points_in_radius(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                         int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                         const int *qmin, const int *qmax, float radius, int64_t *visited) {
  CompatDims D;
  CompatRange R;
  if (!make_dims(dims_host, nd, D) || !make_range(qmin, qmax, nd, R) || !table || (H & (H - 1)) || m < 0 ||
      (m > 0 && (!qcoords || !qvalues || !visited)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_points_in_radius: bad args");
  if (m == 0) return 0;
  const float r2 = radius * radius;
  PCS_LAUNCH(compat_in_radius_kernel, (unsigned)((m + 127) / 128), 128, 0, as_stream(s), table, (unsigned int)(H - 1),
             rows, values, D, R, (const long long *)qcoords, qvalues, (int)m, r2, (long long *)visited);
  return 0;
}

}  // extern "C"
