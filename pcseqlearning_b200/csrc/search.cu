// search.cu -- single-pass K-nearest radius search over the cell hash (sm_100a).
//
// Replaces count_radius_graph_degree_kernel + radius_graph_kernel
// (pcdet/ops/torch_hash/src/torch_hash_kernel.cu:224-409).  One warp owns one query:
//   1. the lanes look up the (up to 27) neighbour cells in parallel -- one 128-bit slot load each;
//   2. the cells' row ranges are concatenated into one virtual candidate range, which the warp
//      sweeps 32 candidates at a time with coalesced float4 loads of the cell-sorted points;
//   3. candidates that pass the reference's fp32 test d2 <= r*r are ballot-compacted and inserted
//      into a sorted K-list held one entry per lane (64-bit key = d2 bits : row index, so ties are
//      broken by ascending reference index -- the canonical order of SURVEY.md A.4);
//   4. the list is written as one coalesced row, or -- fused connected components -- every
//      (query, neighbour) pair is united in the union-find forest and nothing else is written.
#include "common.cuh"

namespace pcs {

struct QueryRange {
  int qmin[4];
  int range[4];
  int nc;
};

constexpr int kWarpsPerBlock = 8;

template <bool kFusedUF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
radius_search_kernel(const pcs_slot_t *__restrict__ table, long long mask, const float4 *__restrict__ sorted_pts,
                     const int *__restrict__ sorted_idx, SegGeom g, const float4 *__restrict__ queries,
                     long long m, const int *__restrict__ order, QueryRange qr, const float *__restrict__ radius,
                     float radius_scalar, int K, int *__restrict__ nbr_idx, float *__restrict__ nbr_d2,
                     int *__restrict__ nbr_cnt, int *__restrict__ uf_parent) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  __shared__ int s_pref[kWarpsPerBlock][33];
  __shared__ int s_start[kWarpsPerBlock][32];
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const unsigned long long kInf = ~0ull;

  for (long long w = (long long)blockIdx.x * kWarpsPerBlock + warp; w < m; w += nwarps) {
    const long long q = order ? (long long)order[w] : w;
    const float4 qp = queries[q];
    const float r = radius ? radius[q] : radius_scalar;
    const float r2 = __fmul_rn(r, r);
    const int seg = point_segment(qp.x, g.seg_div, g.n_seg);
    const float4 lo = s_lo[seg];
    const long long *dims = s_dims + seg * 4;
    const long long qc0 = voxel_coord(qp.x, lo.x, g.vs[0]);
    const long long qc1 = voxel_coord(qp.y, lo.y, g.vs[1]);
    const long long qc2 = voxel_coord(qp.z, lo.z, g.vs[2]);
    const long long qc3 = voxel_coord(qp.w, lo.w, g.vs[3]);

    unsigned long long best = kInf;  // lane j holds the j-th smallest (d2, index) key
    int accepted = 0;

    for (int cb = 0; cb < qr.nc; cb += 32) {
      // ---- 1. parallel cell lookups ----------------------------------------------------------
      int start = 0, count = 0;
      const int cell = cb + lane;
      if (cell < qr.nc) {
        int t = cell;
        const int o0 = t % qr.range[0] + qr.qmin[0];
        t /= qr.range[0];
        const int o1 = t % qr.range[1] + qr.qmin[1];
        t /= qr.range[1];
        const int o2 = t % qr.range[2] + qr.qmin[2];
        t /= qr.range[2];
        const int o3 = t % qr.range[3] + qr.qmin[3];
        const long long key =
            map2key4(qc0 + o0, qc1 + o1, qc2 + o2, qc3 + o3, dims) | ((long long)seg << PCS_SEG_SHIFT);
        long long slot = hash_key(key) & mask;
        for (long long probes = 0; probes <= mask; ++probes) {
          const int4 v = __ldg(reinterpret_cast<const int4 *>(table + slot));
          const long long k = ((long long)v.y << 32) | (unsigned int)v.x;
          if (k == key) {
            start = v.z;
            count = v.w;
            break;
          }
          if (k == PCS_EMPTY_KEY) break;
          slot = (slot + 1) & mask;
        }
      }
      const int incl = warp_incl_scan(count, lane);
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) continue;
      __syncwarp();
      s_pref[warp][lane] = incl - count;
      s_start[warp][lane] = start;
      if (lane == 31) s_pref[warp][32] = total;
      __syncwarp();

      // ---- 2./3. sweep the concatenated candidate range --------------------------------------
      int c = 0;
      const int total_round = (total + 31) & ~31;
      for (int j = lane; j < total_round; j += 32) {
        unsigned long long key64 = kInf;
        bool within = false;
        if (j < total) {
          while (j >= s_pref[warp][c + 1]) ++c;
          const int src = s_start[warp][c] + (j - s_pref[warp][c]);
          const float4 p = __ldg(sorted_pts + src);
          const float d2 = dist2_ref(p, qp);
          within = d2 <= r2;
          if (within)
            key64 = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__ldg(sorted_idx + src);
        }
        const unsigned int wmask = __ballot_sync(0xffffffffu, within);
        if (wmask == 0) continue;
        accepted += __popc(wmask);
        unsigned long long worst = __shfl_sync(0xffffffffu, best, K - 1);
        unsigned int cand = __ballot_sync(0xffffffffu, within && key64 < worst);
        while (cand) {
          const int srcl = __ffs(cand) - 1;
          cand &= cand - 1;
          const unsigned long long ck = __shfl_sync(0xffffffffu, key64, srcl);
          if (ck < worst) {  // warp-uniform
            const unsigned long long up = __shfl_up_sync(0xffffffffu, best, 1);
            const bool gt = best > ck;
            const bool gt_prev = (lane > 0) && (up > ck);
            if (gt) best = gt_prev ? up : ck;
            worst = __shfl_sync(0xffffffffu, best, K - 1);
          }
        }
      }
    }

    // ---- 4. emit ---------------------------------------------------------------------------------
    const int cnt = accepted < K ? accepted : K;
    const int idx = (int)(unsigned int)(best & 0xffffffffu);
    if (nbr_idx && lane < K) nbr_idx[q * K + lane] = lane < cnt ? idx : -1;
    if (nbr_d2 && lane < K) nbr_d2[q * K + lane] = lane < cnt ? __uint_as_float((unsigned int)(best >> 32)) : 0.f;
    if (nbr_cnt && lane == 0) nbr_cnt[q] = cnt;
    if (kFusedUF) {
      if (lane < cnt && idx != (int)q) uf_unite(uf_parent, (int)q, idx);
    }
  }
}

// padded lists -> int64[E][2] rows (ref, query) grouped by ascending query
__global__ void __launch_bounds__(256) lists_to_edges_kernel(const int *__restrict__ nbr_idx,
                                                             const float *__restrict__ nbr_d2,
                                                             const int *__restrict__ nbr_cnt,
                                                             const long long *__restrict__ offsets, long long m,
                                                             int K, long long *__restrict__ edges,
                                                             float *__restrict__ dists) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long q = t / K;
  int j = (int)(t - q * K);
  if (q >= m) return;
  if (j < nbr_cnt[q]) {
    long long e = offsets[q] + j;
    longlong2 row;
    row.x = nbr_idx[q * K + j];
    row.y = q;
    reinterpret_cast<longlong2 *>(edges)[e] = row;
    if (dists && nbr_d2) dists[e] = nbr_d2[q * K + j];
  }
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_radius_search(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                      const int32_t *sorted_idx, int seg_div, int n_seg, const float *seg_lo,
                      const int64_t *seg_dims, const float *vs, const float *queries, int64_t m,
                      const int32_t *order, const int *qmin, const int *qmax, const float *radius,
                      float radius_scalar, int K, int32_t *nbr_idx, float *nbr_d2, int32_t *nbr_cnt,
                      int32_t *uf_parent) {
  if (!table || H < 2 || (H & (H - 1)) || K < 1 || K > PCS_MAX_K || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS ||
      !qmin || !qmax || ((uintptr_t)queries & 15) || ((uintptr_t)sorted_pts & 15) || m < 0 || m >= (1LL << 31))
    return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: bad args (1 <= K <= 32, 16-byte aligned points)");
  if (!uf_parent && !nbr_idx && !nbr_cnt) return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: no output requested");
  if (m == 0) return 0;
  QueryRange qr;
  qr.nc = 1;
  for (int i = 0; i < 4; i++) {
    qr.qmin[i] = qmin[i];
    qr.range[i] = qmax[i] - qmin[i] + 1;
    if (qr.range[i] < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: qmax < qmin");
    qr.nc *= qr.range[i];
  }
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  long long blocks = (m + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = 148LL * 8 * 4;  // persistent-ish: a few waves of 8 resident CTAs per SM
  int grid = (int)(blocks < cap ? blocks : cap);
  if (uf_parent) {
    PCS_LAUNCH(radius_search_kernel<true>, grid, kWarpsPerBlock * 32, 0, as_stream(s), table, (long long)(H - 1),
               (const float4 *)sorted_pts, sorted_idx, g, (const float4 *)queries, (long long)m, order, qr, radius,
               radius_scalar, K, nbr_idx, nbr_d2, nbr_cnt, uf_parent);
  } else {
    PCS_LAUNCH(radius_search_kernel<false>, grid, kWarpsPerBlock * 32, 0, as_stream(s), table, (long long)(H - 1),
               (const float4 *)sorted_pts, sorted_idx, g, (const float4 *)queries, (long long)m, order, qr, radius,
               radius_scalar, K, nbr_idx, nbr_d2, nbr_cnt, uf_parent);
  }
  return 0;
}

int pcs_lists_to_edges(pcs_stream_t s, const int32_t *nbr_idx, const float *nbr_d2, const int32_t *nbr_cnt,
                       const int64_t *offsets, int64_t m, int K, int64_t *edges, float *dists) {
  if (!nbr_idx || !nbr_cnt || !offsets || K < 1 || ((uintptr_t)edges & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_lists_to_edges: bad args");
  if (m == 0) return 0;
  long long threads = (long long)m * K;
  PCS_LAUNCH(lists_to_edges_kernel, (unsigned)((threads + 255) / 256), 256, 0, as_stream(s), nbr_idx, nbr_d2,
             nbr_cnt, (const long long *)offsets, (long long)m, K, (long long *)edges, dists);
  return 0;
}

}  // extern "C"
