// search.cu -- single-pass K-nearest radius search over the cell hash (sm_100a).
//
// Replaces count_radius_graph_degree_kernel + radius_graph_kernel
// (pcdet/ops/torch_hash/src/torch_hash_kernel.cu:224-409).  One warp owns one query:
//   1. every lane derives one neighbour cell (up to 27), a conservative lower bound `dmin2` of the
//      squared distance from the query to anything stored in that cell, and -- unless the cell lies
//      beyond the radius -- looks the cell up with one 128-bit slot load;
//   2. cells are visited in ascending dmin2 (warp redux-min selection); the sweep stops as soon as
//      dmin2 exceeds min(r^2, current K-th best distance): nothing in the remaining cells can enter
//      the list, so the result is exactly the K nearest of the reference's 27-cell candidate set;
//   3. a cell's rows are swept 32 at a time with coalesced float4 loads of the cell-sorted points;
//      the reference's fp32 test d2 <= r*r (same FMA order) and a float compare against the K-th best
//      gate a ballot-compacted insertion into a sorted K-list held one entry per lane (64-bit key =
//      d2 bits : row index, so ties are broken by ascending reference index -- SURVEY.md A.4);
//   4. the list is written as one coalesced row, or -- fused connected components -- the roots of
//      the query and all its neighbours are hooked under their minimum in the union-find forest and
//      no edge is ever written.
#include <stdlib.h>

#include "common.cuh"

namespace pcs {

struct QueryRange {
  int qmin[4];
  int range[4];
  int nc;
  int append;  // 1: append-then-sort list building, 0: sorted insertion from the first candidate
  int prefetch;  // 1: every lane prefetches the first rows of the cell it found (the sweeps come later, one by one)
};

// Up to three union-find forests fed by one search (multi-radius cluster proposals): forest k receives the list
// entries with d2 <= r2[k]; a forest with need_full[k] set is only fed by queries whose list is full (cnt == K) --
// for those the K nearest within the search radius are also the K nearest within any larger radius.
struct UfTargets {
  int *parent[3];
  float r2[3];
  int need_full[3];
  int n;
};

constexpr int kWarpsPerBlock = 8;

constexpr unsigned int kFull = 0xffffffffu;

// lower bound (in metres) of |p_i - q_i| for points stored `o` cells away along one axis;
// u = (q - lo)/vs as computed for the key, f = u - rint(u).  The margin covers the fp32 rounding of u
// for both the query and the stored point (2 * |u| * 2^-23 cells) with a 16x safety factor.
__device__ __forceinline__ float axis_gap(int o, float f, float u, float vs) {
  if (o == 0) return 0.f;
  float g = (o > 0) ? ((float)o - 0.5f - f) : ((float)(-o) - 0.5f + f);
  g -= 4e-6f * (fabsf(u) + 1.0f);
  return g > 0.f ? g * vs : 0.f;
}

// ascending bitonic sort of one 64-bit key per lane
__device__ __forceinline__ unsigned long long warp_sort_asc(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(kFull, v, j);
      const bool take_min = (((lane & k) == 0) == ((lane & j) == 0));
      v = take_min ? (o < v ? o : v) : (o > v ? o : v);
    }
  }
  return v;
}

template <bool kFusedUF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
radius_search_kernel(const pcs_slot_t *__restrict__ table, long long mask, const float4 *__restrict__ sorted_pts,
                     const int *__restrict__ sorted_idx, SegGeom g, const float4 *__restrict__ queries,
                     long long m, const int *__restrict__ order, QueryRange qr, const float *__restrict__ radius,
                     float radius_scalar, int K, int *__restrict__ nbr_idx, float *__restrict__ nbr_d2,
                     int *nbr_cnt, UfTargets uf, const int *skip_full_cnt,  // may alias (cascade passes): no __restrict__
                    
                     const unsigned int *__restrict__ occ, int occ_shift) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  __shared__ unsigned long long s_list[kWarpsPerBlock][32];  // per-warp unsorted list while it is filling
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const unsigned long long kInf = ~0ull;

  for (long long w = (long long)blockIdx.x * kWarpsPerBlock + warp; w < m; w += nwarps) {
    // self-query mode (queries == NULL): the w-th row of the cell-sorted array queries its own grid
    const long long q = queries ? (order ? (long long)order[w] : w) : (long long)sorted_idx[w];
    // second pass of the multi-radius search: queries whose finer-radius list was already full are done
    if (skip_full_cnt && skip_full_cnt[q] >= K) continue;
    const float4 qp = queries ? queries[q] : sorted_pts[w];
    const float r = radius ? radius[q] : radius_scalar;
    const float r2 = __fmul_rn(r, r);
    const int seg = point_segment(qp.x, g.seg_div, g.n_seg);
    const float4 lo = s_lo[seg];
    const long long *dims = s_dims + seg * 4;
    // voxel coordinates exactly as voxel_coord() computes them, keeping u for the pruning bound
    const float u0 = __fdiv_rn(__fsub_rn(qp.x, lo.x), g.vs[0]);
    const float u1 = __fdiv_rn(__fsub_rn(qp.y, lo.y), g.vs[1]);
    const float u2 = __fdiv_rn(__fsub_rn(qp.z, lo.z), g.vs[2]);
    const float u3 = __fdiv_rn(__fsub_rn(qp.w, lo.w), g.vs[3]);
    const float n0 = rintf(u0), n1 = rintf(u1), n2 = rintf(u2), n3 = rintf(u3);
    const long long qc0 = (long long)n0 + 1, qc1 = (long long)n1 + 1, qc2 = (long long)n2 + 1,
                    qc3 = (long long)n3 + 1;

    unsigned long long best = kInf;  // lane j holds the j-th smallest (d2, index) key
    float worst_d2 = __int_as_float(0x7f800000);  // d2 of list entry K-1 (+inf while the list is not full)
    int fill = 0;                                 // number of valid list entries (<= K)
    // While the list is not full, accepted candidates are simply appended (lane = fill + rank); the list is
    // sorted once, when it fills up or at the end.  Only then does the insertion path below (with its K-th best
    // threshold) take over.  Most sparse-region queries never fill their list and never pay for an insertion.
    bool sorted = !qr.append;

    for (int cb = 0; cb < qr.nc; cb += 32) {
      // ---- 1. per-lane cell: offset, pruning bound, lookup -----------------------------------
      int start = 0, count = 0;
      unsigned int sel = 0xffffffffu;  // selection key: dmin2 bits (low 5 cleared) | lane ; ~0 = nothing to visit
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
        const long long c0 = qc0 + o0, c1 = qc1 + o1, c2 = qc2 + o2, c3 = qc3 + o3;
        float dmin2 = 0.f;
        // if map2key clamps a digit the cell aliases another one (reference quirk): no geometric bound then
        const bool clamped = c0 < 0 || c0 > dims[0] || c1 < 0 || c1 > dims[1] || c2 < 0 || c2 > dims[2] ||
                             c3 < 0 || c3 > dims[3];
        if (!clamped) {
          const float g1 = axis_gap(o1, u1 - n1, u1, g.vs[1]);
          const float g2 = axis_gap(o2, u2 - n2, u2, g.vs[2]);
          const float g3 = axis_gap(o3, u3 - n3, u3, g.vs[3]);
          dmin2 = g1 * g1 + g2 * g2 + g3 * g3;  // dimension 0 (frame) is left out: bound stays conservative
          dmin2 *= 0.99999f;
        }
        if (dmin2 <= r2) {
          const long long key = map2key4(c0, c1, c2, c3, dims) | ((long long)seg << PCS_SEG_SHIFT);
          const unsigned int h = hash_key(key);
          bool maybe = true;
          if (occ) {  // occupancy bitmap: most empty neighbour cells are rejected here, without a probe sequence
            const unsigned int b = h >> occ_shift;
            maybe = (__ldg(occ + (b >> 5)) >> (b & 31)) & 1u;
          }
          if (maybe) {
            const int klo = (int)(unsigned int)key, khi = (int)(key >> 32);
            unsigned int slot = h & (unsigned int)mask;
            for (unsigned int probes = 0; probes <= (unsigned int)mask; ++probes) {
              const int4 v = __ldg(reinterpret_cast<const int4 *>(table + slot));
              if (v.x == klo && v.y == khi) {
                start = v.z;
                count = v.w;
                break;
              }
              if ((v.x & v.y) == -1) break;  // PCS_EMPTY_KEY
              slot = (slot + 1) & (unsigned int)mask;
            }
          }
          if (count > 0) {
            sel = (__float_as_uint(dmin2) & ~31u) | (unsigned int)lane;
            if (qr.prefetch) {  // the up-to-27 cell rows are fetched concurrently instead of one miss per visit
              prefetch_l1(sorted_pts + start);
              prefetch_l1(sorted_idx + start);
            }
          }
        }
      }

      // ---- 2. visit cells in ascending dmin2 ---------------------------------------------------
      while (true) {
        const unsigned int pick = __reduce_min_sync(kFull, sel);
        if (pick == 0xffffffffu) break;
        const float pick_d = __uint_as_float(pick & ~31u);
        if (pick_d > fminf(r2, worst_d2)) break;  // every remaining cell is at least this far
        const int src_lane = pick & 31;
        if (lane == src_lane) sel = 0xffffffffu;
        const int cstart = __shfl_sync(kFull, start, src_lane);
        const int ccount = __shfl_sync(kFull, count, src_lane);

        // ---- 3. sweep the cell, 32 rows per step ----------------------------------------------
        for (int j0 = 0; j0 < ccount; j0 += 32) {
          const int j = j0 + lane;
          float d2 = __int_as_float(0x7f800000);
          if (j < ccount) d2 = dist2_ref(__ldg(sorted_pts + cstart + j), qp);
          // reference acceptance d2 <= r*r, and it must not be worse than the current K-th best
          const bool pass = (d2 <= r2) && (d2 <= worst_d2);
          unsigned int cand = __ballot_sync(kFull, pass);
          if (cand == 0) continue;
          unsigned long long key64 = kInf;
          if (pass)
            key64 = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__ldg(sorted_idx + cstart + j);
          if (!sorted) {
            const int npass = __popc(cand);
            if (fill + npass <= K) {
              // append: the passing lanes store their keys at list positions fill + rank (shared memory)
              if (pass) s_list[warp][fill + __popc(cand & ((1u << lane) - 1u))] = key64;
              fill += npass;
              continue;
            }
            // overflow: bring the list into registers, order it, then insert with a threshold
            __syncwarp();
            best = lane < fill ? s_list[warp][lane] : kInf;
            best = warp_sort_asc(best, lane);
            sorted = true;
          }
          unsigned long long worst = __shfl_sync(kFull, best, K - 1);
          while (cand) {
            const int srcl = __ffs(cand) - 1;
            cand &= cand - 1;
            const unsigned long long ck = __shfl_sync(kFull, key64, srcl);
            if (ck < worst) {  // warp-uniform
              const unsigned long long up = __shfl_up_sync(kFull, best, 1);
              const bool gt = best > ck;
              const bool gt_prev = (lane > 0) && (up > ck);
              if (gt) best = gt_prev ? up : ck;
              worst = __shfl_sync(kFull, best, K - 1);
              if (fill < K) ++fill;
            }
          }
          if (fill == K) worst_d2 = __uint_as_float((unsigned int)(worst >> 32));
        }
      }
    }

    // ---- 4. emit ---------------------------------------------------------------------------------
    if (!sorted) {
      __syncwarp();
      best = lane < fill ? s_list[warp][lane] : kInf;
      if (fill > 1 && (nbr_idx || nbr_d2)) best = warp_sort_asc(best, lane);
      __syncwarp();
    }
    const int cnt = fill;
    const int idx = (int)(unsigned int)(best & 0xffffffffu);
    if (nbr_idx && lane < K) nbr_idx[q * K + lane] = lane < cnt ? idx : -1;
    if (nbr_d2 && lane < K) nbr_d2[q * K + lane] = lane < cnt ? __uint_as_float((unsigned int)(best >> 32)) : 0.f;
    if (nbr_cnt && lane == 0) nbr_cnt[q] = cnt;
    if (kFusedUF) {
      const float my_d2 = __uint_as_float((unsigned int)(best >> 32));
      // fully unrolled over the (at most 3) forests: a runtime index into the by-value UfTargets arrays would put
      // the struct into local memory (LDL / STL in the hot loop)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k >= uf.n) break;
        if (uf.need_full[k] && cnt < K) continue;
        // hook the roots of the query and of its neighbours within r2[k] under the smallest of them
        int *parent = uf.parent[k];
        const bool use = lane < cnt && my_d2 <= uf.r2[k];
        const int node = use ? idx : (int)q;
        const int root = uf_find(parent, node);
        const int rmin = (int)__reduce_min_sync(kFull, (unsigned int)root);
        const unsigned int peers = __match_any_sync(kFull, root);
        if (root != rmin && lane == __ffs(peers) - 1) uf_unite(parent, root, rmin);
        // lanes without a neighbour carried the query itself; only a full warp of neighbours leaves it out
        if (__all_sync(kFull, use) && lane == 0) uf_unite(parent, (int)q, rmin);
      }
    }
  }
}

// ---- thread-per-query self search with fused union-find ("cell-coherent" variant) ------------------------------
// The cluster-proposal passes need no neighbour lists, only (a) the union of every query with its K nearest points
// within r and (b) min(count, K).  With the queries taken in cell order, the 32 lanes of a warp sit in a handful of
// neighbouring cells, so ONE thread per query walks its 27 cells on its own: slot and row loads of neighbouring
// lanes hit the same lines, nothing is exchanged between lanes, and the instruction count per query drops from
// ~1400 warp instructions (one warp per query) to that many THREAD instructions.  The K-list lives in shared memory,
// one column per thread ([K][threads]: conflict-free); while it is not full, candidates are appended, afterwards a
// candidate replaces the current worst entry (largest (d2, index) key) and the worst is searched again -- the set
// that remains is exactly the K smallest keys, the same set the warp kernel keeps.
constexpr int kTqThreads = 128;

__global__ void __launch_bounds__(kTqThreads, 6)
self_search_uf_kernel(const pcs_slot_t *__restrict__ table, long long mask, const float4 *__restrict__ sorted_pts,
                      const int *__restrict__ sorted_idx, SegGeom g, long long m, QueryRange qr, float r2, int K,
                      int *nbr_cnt, UfTargets uf, const int *skip_full_cnt,  // may alias: no __restrict__
                      const unsigned int *__restrict__ occ, int occ_shift) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  extern __shared__ unsigned long long s_klist[];  // [K][kTqThreads]
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  unsigned long long *list = s_klist + threadIdx.x;
  const long long w = (long long)blockIdx.x * kTqThreads + threadIdx.x;
  if (w >= m) return;
  const long long q = (long long)sorted_idx[w];
  if (skip_full_cnt && skip_full_cnt[q] >= K) return;
  const float4 qp = sorted_pts[w];
  const int seg = point_segment(qp.x, g.seg_div, g.n_seg);
  const float4 lo = s_lo[seg];
  const long long *dims = s_dims + seg * 4;
  const float u0 = __fdiv_rn(__fsub_rn(qp.x, lo.x), g.vs[0]);
  const float u1 = __fdiv_rn(__fsub_rn(qp.y, lo.y), g.vs[1]);
  const float u2 = __fdiv_rn(__fsub_rn(qp.z, lo.z), g.vs[2]);
  const float u3 = __fdiv_rn(__fsub_rn(qp.w, lo.w), g.vs[3]);
  const float n0 = rintf(u0), n1 = rintf(u1), n2 = rintf(u2), n3 = rintf(u3);
  const long long qc0 = (long long)n0 + 1, qc1 = (long long)n1 + 1, qc2 = (long long)n2 + 1, qc3 = (long long)n3 + 1;

  int fill = 0, worst_pos = 0;
  unsigned long long worst = 0ull;
  float worst_d2 = __int_as_float(0x7f800000);  // +inf while the list is not full
  for (int cell = 0; cell < qr.nc; ++cell) {
    int t = cell;
    const int o0 = t % qr.range[0] + qr.qmin[0];
    t /= qr.range[0];
    const int o1 = t % qr.range[1] + qr.qmin[1];
    t /= qr.range[1];
    const int o2 = t % qr.range[2] + qr.qmin[2];
    t /= qr.range[2];
    const int o3 = t % qr.range[3] + qr.qmin[3];
    const long long c0 = qc0 + o0, c1 = qc1 + o1, c2 = qc2 + o2, c3 = qc3 + o3;
    float dmin2 = 0.f;
    const bool clamped = c0 < 0 || c0 > dims[0] || c1 < 0 || c1 > dims[1] || c2 < 0 || c2 > dims[2] || c3 < 0 ||
                         c3 > dims[3];
    if (!clamped) {
      const float g1 = axis_gap(o1, u1 - n1, u1, g.vs[1]);
      const float g2 = axis_gap(o2, u2 - n2, u2, g.vs[2]);
      const float g3 = axis_gap(o3, u3 - n3, u3, g.vs[3]);
      dmin2 = (g1 * g1 + g2 * g2 + g3 * g3) * 0.99999f;
    }
    if (dmin2 > fminf(r2, worst_d2)) continue;
    const long long key = map2key4(c0, c1, c2, c3, dims) | ((long long)seg << PCS_SEG_SHIFT);
    const unsigned int h = hash_key(key);
    if (occ) {
      const unsigned int b = h >> occ_shift;
      if (!((__ldg(occ + (b >> 5)) >> (b & 31)) & 1u)) continue;
    }
    int start = 0, count = 0;
    {
      const int klo = (int)(unsigned int)key, khi = (int)(key >> 32);
      unsigned int slot = h & (unsigned int)mask;
      for (unsigned int probes = 0; probes <= (unsigned int)mask; ++probes) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(table + slot));
        if (v.x == klo && v.y == khi) {
          start = v.z;
          count = v.w;
          break;
        }
        if ((v.x & v.y) == -1) break;
        slot = (slot + 1) & (unsigned int)mask;
      }
    }
    for (int j = 0; j < count; ++j) {
      const float d2 = dist2_ref(__ldg(sorted_pts + start + j), qp);
      if (!(d2 <= r2) || !(d2 <= worst_d2)) continue;
      const unsigned long long key64 =
          ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__ldg(sorted_idx + start + j);
      if (fill < K) {
        list[fill * kTqThreads] = key64;
        ++fill;
        if (fill < K) continue;
      } else {
        if (!(key64 < worst)) continue;
        list[worst_pos * kTqThreads] = key64;
      }
      // the list is full: (re)locate its largest key
      worst = 0ull;
      for (int i = 0; i < K; ++i) {
        const unsigned long long v = list[i * kTqThreads];
        if (v >= worst) {
          worst = v;
          worst_pos = i;
        }
      }
      worst_d2 = __uint_as_float((unsigned int)(worst >> 32));
    }
  }
  if (nbr_cnt) nbr_cnt[q] = fill;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k >= uf.n) break;
    if (uf.need_full[k] && fill < K) continue;
    int *parent = uf.parent[k];
    const float rk2 = uf.r2[k];
    int rq = uf_find(parent, (int)q);
    for (int i = 0; i < fill; ++i) {
      const unsigned long long v = list[i * kTqThreads];
      if (!(__uint_as_float((unsigned int)(v >> 32)) <= rk2)) continue;
      const int idx = (int)(unsigned int)(v & 0xffffffffu);
      if (idx == (int)q) continue;
      const int ri = uf_find(parent, idx);
      if (ri != rq) {
        uf_unite(parent, rq, ri);
        rq = uf_find(parent, (int)q);
      }
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
                      int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full, int n_uf,
                      const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits) {
  if (occ && (occ_bits < 32 || occ_bits > (1LL << 32) || (occ_bits & (occ_bits - 1))))
    return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: occ_bits must be a power of two in [32, 2^32]");
  int occ_shift = 0;
  if (occ) {
    int lg = 0;
    while ((1LL << lg) < occ_bits) ++lg;
    occ_shift = 32 - lg;
  }
  if (!table || H < 2 || H > (1LL << 31) || (H & (H - 1)) || K < 1 || K > PCS_MAX_K || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS ||
      !qmin || !qmax || ((uintptr_t)queries & 15) || ((uintptr_t)sorted_pts & 15) || m < 0 || m >= (1LL << 31))
    return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: bad args (1 <= K <= 32, 16-byte aligned points)");
  if (n_uf < 0 || n_uf > 3 || (n_uf > 0 && (!uf_parents || !uf_r2 || !uf_need_full)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: at most 3 union-find targets");
  if (n_uf == 0 && !nbr_idx && !nbr_cnt) return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: no output requested");
  UfTargets uf;
  uf.n = n_uf;
  for (int k = 0; k < 3; k++) {
    uf.parent[k] = k < n_uf ? uf_parents[k] : nullptr;
    uf.r2[k] = k < n_uf ? uf_r2[k] : 0.f;
    uf.need_full[k] = k < n_uf ? uf_need_full[k] : 0;
    if (k < n_uf && !uf.parent[k]) return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: null union-find forest");
  }
  if (m == 0) return 0;
  QueryRange qr;
  qr.nc = 1;
  {
    const char *e = getenv("PCS_SEARCH_APPEND");
    qr.append = e ? atoi(e) : 1;
    const char *p = getenv("PCS_SEARCH_PREFETCH");
    qr.prefetch = p ? atoi(p) : 1;
  }
  for (int i = 0; i < 4; i++) {
    qr.qmin[i] = qmin[i];
    qr.range[i] = qmax[i] - qmin[i] + 1;
    if (qr.range[i] < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_radius_search: qmax < qmin");
    qr.nc *= qr.range[i];
  }
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  long long blocks = (m + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = 148LL * 8 * 4;  // a few waves of 8 resident CTAs per SM; warps stride over the queries
  int grid = (int)(blocks < cap ? blocks : cap);
  if (n_uf > 0) {
    PCS_LAUNCH(radius_search_kernel<true>, grid, kWarpsPerBlock * 32, 0, as_stream(s), table, (long long)(H - 1),
               (const float4 *)sorted_pts, sorted_idx, g, (const float4 *)queries, (long long)m, order, qr, radius,
               radius_scalar, K, nbr_idx, nbr_d2, nbr_cnt, uf, skip_full_cnt, occ, occ_shift);
  } else {
    PCS_LAUNCH(radius_search_kernel<false>, grid, kWarpsPerBlock * 32, 0, as_stream(s), table, (long long)(H - 1),
               (const float4 *)sorted_pts, sorted_idx, g, (const float4 *)queries, (long long)m, order, qr, radius,
               radius_scalar, K, nbr_idx, nbr_d2, nbr_cnt, uf, skip_full_cnt, occ, occ_shift);
  }
  return 0;
}

int pcs_self_search_uf(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                       const int32_t *sorted_idx, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                       const int64_t *seg_dims, const float *vs, const int *qmin, const int *qmax, float radius, int K,
                       int32_t *nbr_cnt, int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full,
                       int n_uf, const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits) {
  if (occ && (occ_bits < 32 || occ_bits > (1LL << 32) || (occ_bits & (occ_bits - 1))))
    return set_error(PCS_ERR_BAD_ARG, "pcs_self_search_uf: occ_bits must be a power of two in [32, 2^32]");
  int occ_shift = 0;
  if (occ) {
    int lg = 0;
    while ((1LL << lg) < occ_bits) ++lg;
    occ_shift = 32 - lg;
  }
  if (!table || H < 2 || H > (1LL << 31) || (H & (H - 1)) || K < 1 || K > PCS_MAX_K || n_seg < 1 ||
      n_seg > PCS_MAX_SEGMENTS || !qmin || !qmax || ((uintptr_t)sorted_pts & 15) || n < 0 || n >= (1LL << 31) ||
      (n > 0 && (!sorted_pts || !sorted_idx)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_self_search_uf: bad args (1 <= K <= 32, 16-byte aligned points)");
  if (n_uf < 1 || n_uf > 3 || !uf_parents || !uf_r2 || !uf_need_full)
    return set_error(PCS_ERR_BAD_ARG, "pcs_self_search_uf: 1 to 3 union-find targets");
  UfTargets uf;
  uf.n = n_uf;
  for (int k = 0; k < 3; k++) {
    uf.parent[k] = k < n_uf ? uf_parents[k] : nullptr;
    uf.r2[k] = k < n_uf ? uf_r2[k] : 0.f;
    uf.need_full[k] = k < n_uf ? uf_need_full[k] : 0;
    if (k < n_uf && !uf.parent[k]) return set_error(PCS_ERR_BAD_ARG, "pcs_self_search_uf: null union-find forest");
  }
  if (n == 0) return 0;
  QueryRange qr;
  qr.nc = 1;
  qr.append = qr.prefetch = 0;
  for (int i = 0; i < 4; i++) {
    qr.qmin[i] = qmin[i];
    qr.range[i] = qmax[i] - qmin[i] + 1;
    if (qr.range[i] < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_self_search_uf: qmax < qmin");
    qr.nc *= qr.range[i];
  }
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  const size_t smem = (size_t)K * kTqThreads * sizeof(unsigned long long);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(self_search_uf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kTqThreads * 8);
    attr_set = true;
  }
  const float r = radius;
  PCS_LAUNCH(self_search_uf_kernel, (unsigned)((n + kTqThreads - 1) / kTqThreads), kTqThreads, smem, as_stream(s),
             table, (long long)(H - 1), (const float4 *)sorted_pts, sorted_idx, g, (long long)n, qr, r * r, K,
             nbr_cnt, uf, skip_full_cnt, occ, occ_shift);
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
