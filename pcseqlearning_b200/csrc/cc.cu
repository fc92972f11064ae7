// cc.cu -- exclusive scan and lock-free union-find connected components (sm_100a).
//
// Replaces graph_utils.connected_components (pcdet/models/model_utils/graph_utils.py:40-53), i.e. the
// device->host copy of the edge list and single-threaded scipy.sparse.csgraph.connected_components,
// and the `cumsum(degree) - degree` + two blocking .item() of radius_graph_gpu
// (pcdet/ops/torch_hash/src/torch_hash_kernel.cu:534-538).
#include "common.cuh"

namespace pcs {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // items per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ long long block_excl_scan_ll(long long v, long long *total, long long *smem /*[8]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < kScanBlock / 32; w++) {
      long long x = smem[w];
      smem[w] = t;
      t += x;
    }
    smem[kScanBlock / 32] = t;
  }
  __syncthreads();
  long long base = smem[warp];
  *total = smem[kScanBlock / 32];
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const int *__restrict__ in, long long n,
                                                                 long long *__restrict__ block_sums) {
  __shared__ long long smem[kScanBlock / 32 + 1];
  long long base = (long long)blockIdx.x * kScanTile;
  long long v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long i = base + threadIdx.x * kScanItems + k;
    if (i < n) v += in[i];
  }
  long long total;
  block_excl_scan_ll(v, &total, smem);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of the block sums, grand total written to *grand
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(long long *__restrict__ block_sums, long long nb,
                                                               long long *__restrict__ grand) {
  __shared__ long long smem[kScanBlock / 32 + 1];
  long long carry = 0;
  for (long long b0 = 0; b0 < nb; b0 += kScanBlock) {
    long long i = b0 + threadIdx.x;
    long long v = i < nb ? block_sums[i] : 0;
    long long total;
    long long ex = block_excl_scan_ll(v, &total, smem);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const int *__restrict__ in, long long n,
                                                                const long long *__restrict__ block_sums,
                                                                long long *__restrict__ out) {
  __shared__ long long smem[kScanBlock / 32 + 1];
  long long base = (long long)blockIdx.x * kScanTile;
  int x[kScanItems];
  long long v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long i = base + threadIdx.x * kScanItems + k;
    x[k] = i < n ? in[i] : 0;
    v += x[k];
  }
  long long total;
  long long ex = block_excl_scan_ll(v, &total, smem) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    long long i = base + threadIdx.x * kScanItems + k;
    if (i < n) out[i] = ex;
    ex += x[k];
  }
}

// ---- union-find ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) uf_init_kernel(int *__restrict__ parent, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = (int)i;
}

__global__ void __launch_bounds__(256) uf_union_edges_kernel(int *__restrict__ parent,
                                                             const long long *__restrict__ e0,
                                                             const long long *__restrict__ e1, long long E) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) {
    int a = (int)e0[i], b = (int)e1[i];
    if (a != b) uf_unite(parent, a, b);
  }
}

__global__ void __launch_bounds__(256) seg_of_kernel(const float4 *__restrict__ pts, long long n, int seg_div,
                                                     int n_seg, int *__restrict__ seg_of) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg_of[i] = point_segment(pts[i].x, seg_div, n_seg);
}

constexpr int kLabBlock = 1024;

// pass 1: flatten (parent[i] = root) and count the roots of every segment inside each 1024-node tile
__global__ void __launch_bounds__(kLabBlock) uf_flatten_count_kernel(int *__restrict__ parent, long long n,
                                                                     const int *__restrict__ seg_of, int n_seg,
                                                                     int *__restrict__ tile_cnt /*[tiles][n_seg]*/) {
  __shared__ int hist[PCS_MAX_SEGMENTS];
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * kLabBlock + threadIdx.x;
  if (i < n) {
    // read-only walk: the only writer of parent[i] in this pass is thread i, so a stale path-halving
    // store can never overwrite the root written below
    volatile int *p = parent;
    int r = (int)i;
    while (true) {
      int pr = p[r];
      if (pr == r) break;
      r = pr;
    }
    parent[i] = r;
    if (r == (int)i) atomicAdd(&hist[seg_of ? seg_of[i] : 0], 1);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < n_seg; s += blockDim.x) tile_cnt[(long long)blockIdx.x * n_seg + s] = hist[s];
}

// pass 2: one block per segment -- exclusive scan of that segment's tile counts (in place), total -> n_comp
__global__ void __launch_bounds__(kScanBlock) uf_tile_scan_kernel(int *__restrict__ tile_cnt, long long tiles,
                                                                  int n_seg, long long *__restrict__ n_comp) {
  __shared__ long long smem[kScanBlock / 32 + 1];
  const int s = blockIdx.x;
  long long carry = 0;
  for (long long b0 = 0; b0 < tiles; b0 += kScanBlock) {
    long long t = b0 + threadIdx.x;
    long long v = t < tiles ? tile_cnt[t * n_seg + s] : 0;
    long long total;
    long long ex = block_excl_scan_ll(v, &total, smem);
    if (t < tiles) tile_cnt[t * n_seg + s] = (int)(carry + ex);
    carry += total;
  }
  if (threadIdx.x == 0) n_comp[s] = carry;
}

// pass 3: roots get label = (components of earlier segments) + (roots of the same segment in earlier
// tiles) + (roots of the same segment earlier in this tile)
__global__ void __launch_bounds__(kLabBlock) uf_root_labels_kernel(const int *__restrict__ parent, long long n,
                                                                   const int *__restrict__ seg_of, int n_seg,
                                                                   const int *__restrict__ tile_cnt,
                                                                   const long long *__restrict__ n_comp,
                                                                   long long *__restrict__ labels) {
  __shared__ int warp_cnt[kLabBlock / 32][PCS_MAX_SEGMENTS];
  __shared__ long long seg_off[PCS_MAX_SEGMENTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < (kLabBlock / 32) * PCS_MAX_SEGMENTS; k += blockDim.x)
    (&warp_cnt[0][0])[k] = 0;
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int s = 0; s < n_seg; s++) {
      seg_off[s] = t;
      t += n_comp[s];
    }
  }
  __syncthreads();
  long long i = (long long)blockIdx.x * kLabBlock + threadIdx.x;
  bool is_root = (i < n) && (parent[i] == (int)i);
  int seg = is_root ? (seg_of ? seg_of[i] : 0) : -1 - lane;  // distinct dummies never match
  unsigned int peers = __match_any_sync(0xffffffffu, seg);
  int rank_in_warp = __popc(peers & ((1u << lane) - 1));
  if (is_root && rank_in_warp == 0) warp_cnt[warp][seg] = __popc(peers);
  __syncthreads();
  if (is_root) {
    int before = 0;
    for (int w = 0; w < warp; w++) before += warp_cnt[w][seg];
    labels[i] = seg_off[seg] + tile_cnt[(long long)blockIdx.x * n_seg + seg] + before + rank_in_warp;
  }
}

// pass 4: every other node copies its root's label
__global__ void __launch_bounds__(256) uf_copy_labels_kernel(const int *__restrict__ parent, long long n,
                                                             long long *__restrict__ labels) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int r = parent[i];
    if (r != (int)i) labels[i] = labels[r];
  }
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int64_t pcs_exclusive_scan_tmp_bytes(int64_t n) { return ((n + kScanTile - 1) / kScanTile + 1) * 8; }

int pcs_exclusive_scan(pcs_stream_t s, const int32_t *in, int64_t n, int64_t *out, void *tmp, int64_t tmp_bytes) {
  if (n < 0 || !out || (n > 0 && !in) || tmp_bytes < pcs_exclusive_scan_tmp_bytes(n) || !tmp)
    return set_error(PCS_ERR_BAD_ARG, "pcs_exclusive_scan: bad args / tmp too small");
  cudaStream_t st = as_stream(s);
  long long nb = (n + kScanTile - 1) / kScanTile;
  long long *sums = (long long *)tmp;
  if (nb > 0) PCS_LAUNCH(scan_reduce_kernel, (unsigned)nb, kScanBlock, 0, st, in, (long long)n, sums);
  PCS_LAUNCH(scan_sums_kernel, 1, kScanBlock, 0, st, sums, nb, (long long *)out + n);
  if (nb > 0) PCS_LAUNCH(scan_apply_kernel, (unsigned)nb, kScanBlock, 0, st, in, (long long)n, sums, (long long *)out);
  return 0;
}

int pcs_uf_init(pcs_stream_t s, int32_t *parent, int64_t n) {
  if (n < 0 || n >= (1LL << 31) || (n > 0 && !parent)) return set_error(PCS_ERR_BAD_ARG, "pcs_uf_init: bad args");
  if (n == 0) return 0;
  PCS_LAUNCH(uf_init_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(s), parent, (long long)n);
  return 0;
}

int pcs_uf_union_edges(pcs_stream_t s, int32_t *parent, const int64_t *e0, const int64_t *e1, int64_t E) {
  if (E < 0 || (E > 0 && (!parent || !e0 || !e1))) return set_error(PCS_ERR_BAD_ARG, "pcs_uf_union_edges: bad args");
  if (E == 0) return 0;
  PCS_LAUNCH(uf_union_edges_kernel, (unsigned)((E + 255) / 256), 256, 0, as_stream(s), parent,
             (const long long *)e0, (const long long *)e1, (long long)E);
  return 0;
}

int64_t pcs_uf_labels_tmp_bytes(int64_t n, int n_seg) {
  return ((n + kLabBlock - 1) / kLabBlock + 1) * (int64_t)n_seg * 4;
}

int pcs_uf_labels(pcs_stream_t s, int32_t *parent, int64_t n, const int32_t *seg_of, int n_seg, int64_t *labels,
                  int64_t *n_comp, void *tmp, int64_t tmp_bytes) {
  if (n < 0 || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS || !n_comp || (n > 0 && (!parent || !labels)) || !tmp ||
      tmp_bytes < pcs_uf_labels_tmp_bytes(n, n_seg))
    return set_error(PCS_ERR_BAD_ARG, "pcs_uf_labels: bad args / tmp too small");
  cudaStream_t st = as_stream(s);
  long long tiles = (n + kLabBlock - 1) / kLabBlock;
  int *tile_cnt = (int *)tmp;
  if (tiles > 0)
    PCS_LAUNCH(uf_flatten_count_kernel, (unsigned)tiles, kLabBlock, 0, st, parent, (long long)n, seg_of, n_seg,
               tile_cnt);
  PCS_LAUNCH(uf_tile_scan_kernel, n_seg, kScanBlock, 0, st, tile_cnt, tiles, n_seg, (long long *)n_comp);
  if (tiles > 0) {
    PCS_LAUNCH(uf_root_labels_kernel, (unsigned)tiles, kLabBlock, 0, st, parent, (long long)n, seg_of, n_seg,
               tile_cnt, (const long long *)n_comp, (long long *)labels);
    PCS_LAUNCH(uf_copy_labels_kernel, (unsigned)((n + 255) / 256), 256, 0, st, parent, (long long)n,
               (long long *)labels);
  }
  return 0;
}

int pcs_point_segments(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, int32_t *seg_of) {
  if (n < 0 || n_seg < 1 || ((uintptr_t)pts & 15) || (n > 0 && !seg_of))
    return set_error(PCS_ERR_BAD_ARG, "pcs_point_segments: bad args");
  if (n == 0) return 0;
  PCS_LAUNCH(seg_of_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(s), (const float4 *)pts, (long long)n,
             seg_div < 1 ? 1 : seg_div, n_seg, seg_of);
  return 0;
}

}  // extern "C"

// ---- row gather -------------------------------------------------------------------------------------------------
// dst[i] = src[idx[i]] for rows of 1 / 4 / 8 / 12 / 16 bytes (int64 indices).  Replaces the fancy-indexing passes of
// the host code (simple_reg.py:126-130 re-ordering after the subsample, ground_plane_remover.py:238-247 mask filter,
// preprocessor_utils.py:416-419 voxel -> point broadcast): one coalesced index load, one (random) vector load and one
// coalesced vector store per row.
namespace pcs {

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T *__restrict__ src, const long long *__restrict__ idx,
                                                          long long n, T *__restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[idx[i]];
}

struct Row12 {
  float a, b, c;
};

}  // namespace pcs

extern "C" int pcs_gather_rows(pcs_stream_t s, const void *src, const int64_t *idx, int64_t n, int row_bytes, void *dst) {
  using namespace pcs;
  if (n < 0 || (n > 0 && (!src || !idx || !dst))) return set_error(PCS_ERR_BAD_ARG, "pcs_gather_rows: bad args");
  if (n == 0) return 0;
  const int grid = grid_for(n, 256, 16);
  cudaStream_t st = as_stream(s);
  const long long *ix = (const long long *)idx;
  switch (row_bytes) {
    case 1:
      PCS_LAUNCH(gather_rows_kernel<unsigned char>, grid, 256, 0, st, (const unsigned char *)src, ix, (long long)n,
                 (unsigned char *)dst);
      break;
    case 4:
      PCS_LAUNCH(gather_rows_kernel<int>, grid, 256, 0, st, (const int *)src, ix, (long long)n, (int *)dst);
      break;
    case 8:
      PCS_LAUNCH(gather_rows_kernel<long long>, grid, 256, 0, st, (const long long *)src, ix, (long long)n,
                 (long long *)dst);
      break;
    case 12:
      PCS_LAUNCH(gather_rows_kernel<Row12>, grid, 256, 0, st, (const Row12 *)src, ix, (long long)n, (Row12 *)dst);
      break;
    case 16:
      if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return set_error(PCS_ERR_BAD_ARG, "pcs_gather_rows: alignment");
      PCS_LAUNCH(gather_rows_kernel<int4>, grid, 256, 0, st, (const int4 *)src, ix, (long long)n, (int4 *)dst);
      break;
    default:
      return set_error(PCS_ERR_BAD_ARG, "pcs_gather_rows: row size must be 1, 4, 8, 12 or 16 bytes");
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Per-group min / max of a float column (torch_scatter.scatter(min|max) of preprocessor_utils.py:113-114 and the
// pillar statistics of format_pillars): order-preserving uint encoding + atomicMin / atomicMax.  A warp whose 32
// rows belong to one group (the common case: rows sorted by group) reduces with redux first and issues ONE pair of
// atomics; torch's scatter_reduce issues one contended atomic per row.
// ------------------------------------------------------------------------------------------------
namespace pcs {

__global__ void __launch_bounds__(256) minmax_init_kernel(unsigned int *__restrict__ mn, unsigned int *__restrict__ mx,
                                                          long long C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    mn[i] = 0xffffffffu;
    mx[i] = 0u;
  }
}

__global__ void __launch_bounds__(256) minmax_scatter_kernel(const float *__restrict__ val, long long stride,
                                                             const long long *__restrict__ ids, long long n,
                                                             long long C, unsigned int *__restrict__ mn,
                                                             unsigned int *__restrict__ mx) {
  const int lane = threadIdx.x & 31;
  const long long nround = (n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround;
       i += (long long)gridDim.x * blockDim.x) {
    const bool valid = i < n;
    long long g = -1;
    unsigned int o = 0u;
    if (valid) {
      g = ids[i];
      o = f2ord(val[i * stride]);
      if (g < 0 || g >= C) g = -1;  // out-of-range ids are ignored
    }
    const long long g0 = __shfl_sync(0xffffffffu, g, 0);
    if (__all_sync(0xffffffffu, g == g0 || !valid)) {
      const unsigned int lo = __reduce_min_sync(0xffffffffu, valid ? o : 0xffffffffu);
      const unsigned int hi = __reduce_max_sync(0xffffffffu, valid ? o : 0u);
      if (lane == 0 && g0 >= 0) {
        atomicMin(mn + g0, lo);
        atomicMax(mx + g0, hi);
      }
    } else if (g >= 0) {
      atomicMin(mn + g, o);
      atomicMax(mx + g, o);
    }
  }
}

__global__ void __launch_bounds__(256) minmax_finish_kernel(const unsigned int *__restrict__ mn,
                                                            const unsigned int *__restrict__ mx, long long C,
                                                            float *__restrict__ out_min, float *__restrict__ out_max) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const bool empty = mn[i] == 0xffffffffu && mx[i] == 0u;  // empty groups -> 0 (torch_scatter convention)
    out_min[i] = empty ? 0.f : ord2f(mn[i]);
    out_max[i] = empty ? 0.f : ord2f(mx[i]);
  }
}

}  // namespace pcs

extern "C" int pcs_group_minmax(pcs_stream_t s, const float *values, int64_t stride, const int64_t *ids, int64_t n,
                                int64_t C, uint32_t *tmp, float *out_min, float *out_max) {
  using namespace pcs;
  if (n < 0 || C < 1 || stride < 1 || !tmp || !out_min || !out_max || (n > 0 && (!values || !ids)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_group_minmax: bad args");
  cudaStream_t st = as_stream(s);
  PCS_LAUNCH(minmax_init_kernel, (unsigned)((C + 255) / 256), 256, 0, st, tmp, tmp + C, (long long)C);
  if (n > 0)
    PCS_LAUNCH(minmax_scatter_kernel, grid_for(n, 256, 8), 256, 0, st, values, (long long)stride,
               (const long long *)ids, (long long)n, (long long)C, tmp, tmp + C);
  PCS_LAUNCH(minmax_finish_kernel, (unsigned)((C + 255) / 256), 256, 0, st, tmp, tmp + C, (long long)C, out_min,
             out_max);
  return 0;
}
