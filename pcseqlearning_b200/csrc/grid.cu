// grid.cu -- voxel grid geometry + open-addressing cell hash with counting sort (sm_100a).
//
// Replaces, for the cluster-tracking path, the torch ops of RadiusGraph.build_graph
// (pcdet/models/model_utils/graph_utils.py:169-183) and hash_insert_gpu
// (pcdet/ops/torch_hash/src/torch_hash_kernel.cu:54-91).  See include/pcseq_b200.h.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pcs {

thread_local char g_err[512] = "";
long long g_launches = 0;

int set_error(int code, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s (code %d%s%s)", what, code, code > 0 ? ": " : "",
           code > 0 ? cudaGetErrorString((cudaError_t)code) : "");
  return code;
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, what);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bounds: per-segment min/max of (frame, x, y, z)
// ------------------------------------------------------------------------------------------------
__global__ void bounds_init_kernel(unsigned int *bounds, int n_seg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg * 8) bounds[i] = ((i & 7) < 4) ? 0xffffffffu : 0u;  // min slots = +max, max slots = 0
}

// One float4 per thread per step, grid-stride.  Lanes of a warp are grouped by segment with
// match_any so each distinct segment costs 8 redux + 8 shared atomics per warp; blocks flush their
// shared copy with global atomics once at the end.
__global__ void __launch_bounds__(256) bounds_update_kernel(const float4 *__restrict__ pts, long long n,
                                                            int seg_div, int n_seg,
                                                            unsigned int *__restrict__ bounds) {
  __shared__ unsigned int sb[PCS_MAX_SEGMENTS * 8];
  for (int i = threadIdx.x; i < n_seg * 8; i += blockDim.x) sb[i] = ((i & 7) < 4) ? 0xffffffffu : 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long nround = ((n + 31) / 32) * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    bool valid = i < n;
    float4 p = valid ? ldg_stream_f4(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    int seg = valid ? point_segment(p.x, seg_div, n_seg) : -1;
    unsigned int ox = f2ord(p.x), oy = f2ord(p.y), oz = f2ord(p.z), ow = f2ord(p.w);
    unsigned int todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
      int leader = __ffs(todo) - 1;
      int s = __shfl_sync(0xffffffffu, seg, leader);
      bool mine = valid && (seg == s);
      unsigned int grp = __ballot_sync(0xffffffffu, mine);
      unsigned int mn0 = __reduce_min_sync(0xffffffffu, mine ? ox : 0xffffffffu);
      unsigned int mn1 = __reduce_min_sync(0xffffffffu, mine ? oy : 0xffffffffu);
      unsigned int mn2 = __reduce_min_sync(0xffffffffu, mine ? oz : 0xffffffffu);
      unsigned int mn3 = __reduce_min_sync(0xffffffffu, mine ? ow : 0xffffffffu);
      unsigned int mx0 = __reduce_max_sync(0xffffffffu, mine ? ox : 0u);
      unsigned int mx1 = __reduce_max_sync(0xffffffffu, mine ? oy : 0u);
      unsigned int mx2 = __reduce_max_sync(0xffffffffu, mine ? oz : 0u);
      unsigned int mx3 = __reduce_max_sync(0xffffffffu, mine ? ow : 0u);
      if (lane == leader) {
        unsigned int *b = sb + s * 8;
        atomicMin(b + 0, mn0);
        atomicMin(b + 1, mn1);
        atomicMin(b + 2, mn2);
        atomicMin(b + 3, mn3);
        atomicMax(b + 4, mx0);
        atomicMax(b + 5, mx1);
        atomicMax(b + 6, mx2);
        atomicMax(b + 7, mx3);
      }
      todo &= ~grp;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_seg * 8; i += blockDim.x) {
    if ((i & 7) < 4) {
      if (sb[i] != 0xffffffffu) atomicMin(bounds + i, sb[i]);
    } else {
      if (sb[i] != 0u) atomicMax(bounds + i, sb[i]);
    }
  }
}

// lo = min - 2*vs ; hi = max + 2*vs ; dims = rint((hi - lo)/vs) + 3   (graph_utils.py:172-176, fp32)
__global__ void grid_params_kernel(const unsigned int *__restrict__ bounds, int n_seg, float vs0, float vs1,
                                   float vs2, float vs3, int pad, float *__restrict__ seg_lo,
                                   long long *__restrict__ seg_dims) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg * 4) return;
  int s = i >> 2, d = i & 3;
  float vs = d == 0 ? vs0 : (d == 1 ? vs1 : (d == 2 ? vs2 : vs3));
  unsigned int omin = bounds[s * 8 + d], omax = bounds[s * 8 + 4 + d];
  if (omin == 0xffffffffu && omax == 0u) {  // empty segment
    seg_lo[i] = 0.f;
    seg_dims[i] = 1;
    return;
  }
  float two_vs = __fmul_rn(vs, 2.0f);
  float lo = __fsub_rn(ord2f(omin), two_vs);
  float hi = __fadd_rn(ord2f(omax), two_vs);
  // pad > 0 (persistent ICP): `pad` extra cells of margin on every side, for point sets that move after the
  // grid has been laid out; pad == 0 is the reference geometry bit for bit
  seg_lo[i] = pad ? __fsub_rn(lo, __fmul_rn(vs, (float)pad)) : lo;
  seg_dims[i] = (long long)rintf(__fdiv_rn(__fsub_rn(hi, lo), vs)) + 3 + 2 * pad;
}

__global__ void __launch_bounds__(256) voxel_keys_kernel(const float4 *__restrict__ pts, long long n, SegGeom g,
                                                         long long *__restrict__ coords,
                                                         long long *__restrict__ keys) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long c[4];
  bool ovf;
  long long k = point_key(pts[i], g, s_lo, s_dims, c, &ovf);
  if (coords) {
    coords[i * 4 + 0] = c[0];
    coords[i * 4 + 1] = c[1];
    coords[i * 4 + 2] = c[2];
    coords[i * 4 + 3] = c[3];
  }
  if (keys) keys[i] = k & ((1LL << PCS_SEG_SHIFT) - 1);
}

// ------------------------------------------------------------------------------------------------
// hash build: clear -> count (find-or-insert unique cell keys) -> assign ranges -> scatter
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_clear_kernel(int4 *__restrict__ table, long long H,
                                                          int *__restrict__ counters, unsigned int *__restrict__ occ,
                                                          long long occ_words) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const int4 e = make_int4(-1, -1, 0, 0);  // key = -1, start = 0, count = 0
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) table[i] = e;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < occ_words; i += stride) occ[i] = 0u;
  if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
}

// find-or-insert; returns the slot.  Plain load first (most points land in an existing cell).
__device__ __forceinline__ long long find_or_insert(pcs_slot_t *table, long long mask, long long key,
                                                    int *counters, unsigned int *occ, int occ_shift) {
  const unsigned int h = hash_key(key);
  long long slot = h & mask;
  // A probe sequence this long means the table is (nearly) full: give up and flag it instead of walking the whole
  // table for every remaining point (an under-sized table cost 600 ms that way before the caller could rebuild it).
  const long long max_probes = mask < 8192 ? mask : 8192;
  for (long long probes = 0; probes <= max_probes; ++probes) {
    if ((probes & 255) == 255 && *((volatile int *)&counters[2]) != 0) return -1;  // someone already flagged it
    long long cur = *((volatile long long *)&table[slot].key);
    if (cur == key) return slot;
    if (cur == PCS_EMPTY_KEY) {
      unsigned long long prev = atomicCAS((unsigned long long *)&table[slot].key,
                                          (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
      if (prev == (unsigned long long)PCS_EMPTY_KEY) {
        atomicAdd(&counters[0], 1);
        if (occ) {  // the claimer publishes the cell in the occupancy bitmap
          const unsigned int b = h >> occ_shift;
          atomicOr(&occ[b >> 5], 1u << (b & 31));
        }
        return slot;
      }
      if ((long long)prev == key) return slot;
    }
    slot = (slot + 1) & mask;
  }
  atomicExch(&counters[2], PCS_ERR_TABLE_FULL);
  return -1;
}

// Pass 1: count points per cell.  Lanes holding the same key are aggregated (match_any) so a warp of
// spatially coherent points issues one probe sequence and one atomicAdd per distinct cell.
__global__ void __launch_bounds__(256) hash_count_kernel(const float4 *__restrict__ pts, long long n, SegGeom g,
                                                         pcs_slot_t *__restrict__ table, long long mask,
                                                         int *__restrict__ counters, unsigned int *__restrict__ occ,
                                                         int occ_shift) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long nround = ((n + 31) / 32) * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    bool valid = i < n;
    long long key = -2 - lane;  // distinct dummies for inactive lanes
    if (valid) {
      long long c[4];
      bool ovf;
      key = point_key(ldg_stream_f4(pts + i), g, s_lo, s_dims, c, &ovf);
      if (ovf) atomicExch(&counters[2], PCS_ERR_KEY_RANGE);
    }
    unsigned int peers = __match_any_sync(0xffffffffu, key);
    int leader = __ffs(peers) - 1;
    if (valid && lane == leader) {
      long long slot = find_or_insert(table, mask, key, counters, occ, occ_shift);
      if (slot >= 0) atomicAdd(&table[slot].count, __popc(peers));
    }
  }
}

// Pass 2: give every occupied slot a contiguous row range.  Block-level scan of the counts, one
// global atomic per block for the base (cell order in memory is arbitrary; nothing depends on it).
__global__ void __launch_bounds__(256) assign_ranges_kernel(pcs_slot_t *__restrict__ table, long long H,
                                                            int *__restrict__ counters) {
  __shared__ int warp_tot[8];
  __shared__ int block_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long base_i = (long long)blockIdx.x * (blockDim.x * 4) + threadIdx.x * 4;
  int c[4];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    long long i = base_i + k;
    c[k] = (i < H) ? table[i].count : 0;
    sum += c[k];
  }
  int incl = warp_incl_scan(sum, lane);
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) {
      int v = warp_tot[w];
      warp_tot[w] = t;
      t += v;
    }
    block_base = t > 0 ? atomicAdd(&counters[1], t) : 0;
  }
  __syncthreads();
  int off = block_base + warp_tot[warp] + incl - sum;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    long long i = base_i + k;
    if (i < H && c[k] > 0) table[i].start = off;  // start doubles as the scatter cursor in pass 3
    off += c[k];
  }
}

// Pass 3: scatter points into their cell ranges.  table[slot].start is advanced as a cursor; after the
// pass start == first row + count, which pass 4 rewinds.
__global__ void __launch_bounds__(256) hash_scatter_kernel(const float4 *__restrict__ pts, long long n, SegGeom g,
                                                           pcs_slot_t *__restrict__ table, long long mask,
                                                           float4 *__restrict__ sorted_pts,
                                                           int *__restrict__ sorted_idx) {
  __shared__ float4 s_lo[PCS_MAX_SEGMENTS];
  __shared__ long long s_dims[PCS_MAX_SEGMENTS * 4];
  load_geom(g, s_lo, s_dims);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long nround = ((n + 31) / 32) * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    bool valid = i < n;
    long long key = -2 - lane;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      long long c[4];
      bool ovf;
      p = ldg_stream_f4(pts + i);
      key = point_key(p, g, s_lo, s_dims, c, &ovf);
    }
    unsigned int peers = __match_any_sync(0xffffffffu, key);
    int leader = __ffs(peers) - 1;
    int base = 0;
    if (valid && lane == leader) {
      long long slot = hash_key(key) & mask;
      long long probes = 0;
      // bounded: a key that could not be inserted (table overflow, flagged in counters[2]) is simply not found
      while (*((volatile long long *)&table[slot].key) != key && probes <= mask) {
        slot = (slot + 1) & mask;
        ++probes;
      }
      base = (probes <= mask) ? atomicAdd(&table[slot].start, __popc(peers)) : -1;
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid && base >= 0) {
      int pos = base + __popc(peers & ((1u << lane) - 1));
      sorted_pts[pos] = p;
      sorted_idx[pos] = (int)i;
    }
  }
}

__global__ void __launch_bounds__(256) rewind_ranges_kernel(pcs_slot_t *__restrict__ table, long long H) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H) {
    int c = table[i].count;
    if (c > 0) table[i].start -= c;
  }
}

// ---- key-ordered cell ranges (pcs_hash_build_sorted) ---------------------------------------------
// assign_ranges_kernel hands out row ranges in SLOT order, i.e. in hash order: the 27 neighbour cells of a query end
// up scattered over sorted_pts.  The variant below numbers the occupied cells by ascending key instead (frame-major,
// z fastest), so a frame's cells are contiguous and self-queries walk the grid coherently.
__global__ void __launch_bounds__(256) fill_keys_kernel(long long *__restrict__ keys, long long cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) keys[i] = 0x7fffffffffffffffLL;
}

__global__ void __launch_bounds__(256) collect_cells_kernel(const pcs_slot_t *__restrict__ table, long long H,
                                                            long long *__restrict__ keys, int *__restrict__ slots,
                                                            long long cap, int *__restrict__ cursor) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  const long long k = table[i].key;
  if (k == PCS_EMPTY_KEY) return;
  const int r = atomicAdd(cursor, 1);
  if (r < cap) {
    keys[r] = k;
    slots[r] = (int)i;
  }
}

__global__ void __launch_bounds__(256) gather_counts_kernel(const pcs_slot_t *__restrict__ table,
                                                            const long long *__restrict__ keys_sorted,
                                                            const int *__restrict__ slots_sorted, long long cap,
                                                            int *__restrict__ cnt) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < cap) cnt[r] = keys_sorted[r] == 0x7fffffffffffffffLL ? 0 : table[slots_sorted[r]].count;
}

__global__ void __launch_bounds__(256) scatter_starts_kernel(pcs_slot_t *__restrict__ table,
                                                             const long long *__restrict__ keys_sorted,
                                                             const int *__restrict__ slots_sorted,
                                                             const long long *__restrict__ offs, long long cap) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < cap && keys_sorted[r] != 0x7fffffffffffffffLL) table[slots_sorted[r]].start = (int)offs[r];
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_version(void) { return 100; }
const char *pcs_last_error(void) { return g_err; }
int64_t pcs_launch_count(void) { return g_launches; }
void pcs_reset_launch_count(void) { g_launches = 0; }

int pcs_bounds_init(pcs_stream_t s, uint32_t *bounds, int n_seg) {
  if (!bounds || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS) return set_error(PCS_ERR_BAD_ARG, "pcs_bounds_init: bad args");
  PCS_LAUNCH(bounds_init_kernel, 1, 512, 0, as_stream(s), bounds, n_seg);
  return 0;
}

int pcs_bounds_update(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, uint32_t *bounds) {
  if (!bounds || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS || (n > 0 && !pts) || ((uintptr_t)pts & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_bounds_update: bad args (points must be 16-byte aligned)");
  if (n == 0) return 0;
  PCS_LAUNCH(bounds_update_kernel, grid_for(n, 256, 8), 256, 0, as_stream(s), (const float4 *)pts, (long long)n,
             seg_div < 1 ? 1 : seg_div, n_seg, bounds);
  return 0;
}

int pcs_grid_params(pcs_stream_t s, const uint32_t *bounds, int n_seg, const float *vs, int pad, float *seg_lo,
                    int64_t *seg_dims) {
  if (!bounds || !vs || !seg_lo || !seg_dims || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS)
    return set_error(PCS_ERR_BAD_ARG, "pcs_grid_params: bad args");
  PCS_LAUNCH(grid_params_kernel, 1, 256, 0, as_stream(s), bounds, n_seg, vs[0], vs[1], vs[2], vs[3], pad, seg_lo,
             (long long *)seg_dims);
  return 0;
}

int pcs_voxel_keys(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                   const int64_t *seg_dims, const float *vs, int64_t *coords, int64_t *keys) {
  if (n_seg < 1 || n_seg > PCS_MAX_SEGMENTS || ((uintptr_t)pts & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_voxel_keys: bad args");
  if (n == 0) return 0;
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  PCS_LAUNCH(voxel_keys_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(s), (const float4 *)pts,
             (long long)n, g, (long long *)coords, (long long *)keys);
  return 0;
}

int pcs_hash_build(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                   const int64_t *seg_dims, const float *vs, pcs_slot_t *table, int64_t H, float *sorted_pts,
                   int32_t *sorted_idx, int32_t *counters, uint32_t *occ, int64_t occ_bits) {
  if (occ && (occ_bits < 32 || occ_bits > (1LL << 32) || (occ_bits & (occ_bits - 1))))
    return set_error(PCS_ERR_BAD_ARG, "pcs_hash_build: occ_bits must be a power of two in [32, 2^32]");
  int occ_shift = 0;
  if (occ) {
    int lg = 0;
    while ((1LL << lg) < occ_bits) ++lg;
    occ_shift = 32 - lg;
  }
  if (!table || !counters || H < 2 || (H & (H - 1)) || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS || n < 0 ||
      n >= (1LL << 31) || ((uintptr_t)pts & 15) || ((uintptr_t)sorted_pts & 15) || ((uintptr_t)table & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_hash_build: bad args (H must be a power of two, buffers 16-byte aligned)");
  cudaStream_t st = as_stream(s);
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  PCS_LAUNCH(table_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, counters, occ,
             occ ? (long long)(occ_bits >> 5) : 0LL);
  if (n == 0) return 0;
  PCS_LAUNCH(hash_count_kernel, grid_for(n, 256, 8), 256, 0, st, (const float4 *)pts, (long long)n, g, table,
             (long long)(H - 1), counters, occ, occ_shift);
  PCS_LAUNCH(assign_ranges_kernel, (unsigned)((H + 1023) / 1024), 256, 0, st, table, (long long)H, counters);
  PCS_LAUNCH(hash_scatter_kernel, grid_for(n, 256, 8), 256, 0, st, (const float4 *)pts, (long long)n, g, table,
             (long long)(H - 1), (float4 *)sorted_pts, sorted_idx);
  PCS_LAUNCH(rewind_ranges_kernel, (unsigned)((H + 255) / 256), 256, 0, st, table, (long long)H);
  return 0;
}

/* ---- staged for the next round: cell ranges in key order (not used by the default path yet) ---- */
static inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

static int64_t sorted_ws_layout(int64_t cap, int64_t off[8]) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr,
                                  (unsigned long long *)nullptr, (const int *)nullptr, (int *)nullptr, (int)cap);
  int64_t o = 0;
  off[0] = o; o += align256(cap * 8);                              // keys_in
  off[1] = o; o += align256(cap * 8);                              // keys_out
  off[2] = o; o += align256(cap * 4);                              // slots_in
  off[3] = o; o += align256(cap * 4);                              // slots_out
  off[4] = o; o += align256(cap * 4);                              // cnt
  off[5] = o; o += align256((cap + 1) * 8);                        // offs
  off[6] = o; o += align256(pcs_exclusive_scan_tmp_bytes(cap));    // scan scratch
  off[7] = o; o += align256((int64_t)cub_bytes + 256);             // CUB scratch
  return o + 256;
}

int64_t pcs_hash_build_sorted_ws_bytes(int64_t n, int64_t H) {
  int64_t off[8];
  const int64_t cap = n < H ? n : H;
  return sorted_ws_layout(cap < 1 ? 1 : cap, off);
}

int pcs_hash_build_sorted(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                          const int64_t *seg_dims, const float *vs, pcs_slot_t *table, int64_t H, float *sorted_pts,
                          int32_t *sorted_idx, int32_t *counters, uint32_t *occ, int64_t occ_bits, void *ws,
                          int64_t ws_bytes) {
  if (occ && (occ_bits < 32 || occ_bits > (1LL << 32) || (occ_bits & (occ_bits - 1))))
    return set_error(PCS_ERR_BAD_ARG, "pcs_hash_build_sorted: occ_bits must be a power of two in [32, 2^32]");
  int occ_shift = 0;
  if (occ) {
    int lg = 0;
    while ((1LL << lg) < occ_bits) ++lg;
    occ_shift = 32 - lg;
  }
  if (!table || !counters || H < 2 || (H & (H - 1)) || n_seg < 1 || n_seg > PCS_MAX_SEGMENTS || n < 0 ||
      n >= (1LL << 31) || ((uintptr_t)pts & 15) || ((uintptr_t)sorted_pts & 15) || ((uintptr_t)table & 15) || !ws ||
      ((uintptr_t)ws & 15) || ws_bytes < pcs_hash_build_sorted_ws_bytes(n, H))
    return set_error(PCS_ERR_BAD_ARG, "pcs_hash_build_sorted: bad args / workspace too small");
  cudaStream_t st = as_stream(s);
  SegGeom g = make_geom(seg_lo, seg_dims, vs, seg_div, n_seg);
  PCS_LAUNCH(table_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, counters, occ,
             occ ? (long long)(occ_bits >> 5) : 0LL);
  if (n == 0) return 0;
  PCS_LAUNCH(hash_count_kernel, grid_for(n, 256, 8), 256, 0, st, (const float4 *)pts, (long long)n, g, table,
             (long long)(H - 1), counters, occ, occ_shift);
  const int64_t cap = n < H ? n : H;  // an upper bound of the number of occupied cells, known without a host sync
  int64_t off[8];
  sorted_ws_layout(cap, off);
  char *w = (char *)ws;
  long long *keys_in = (long long *)(w + off[0]), *keys_out = (long long *)(w + off[1]);
  int *slots_in = (int *)(w + off[2]), *slots_out = (int *)(w + off[3]), *cnt = (int *)(w + off[4]);
  long long *offs = (long long *)(w + off[5]);
  const unsigned blocks_cap = (unsigned)((cap + 255) / 256);
  PCS_LAUNCH(fill_keys_kernel, blocks_cap, 256, 0, st, keys_in, (long long)cap);
  // counters[3] is the append cursor of the collection pass (cleared by table_clear_kernel)
  PCS_LAUNCH(collect_cells_kernel, (unsigned)((H + 255) / 256), 256, 0, st, table, (long long)H, keys_in, slots_in,
             (long long)cap, counters + 3);
  size_t cub_bytes = (size_t)(ws_bytes - off[7]);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w + off[7], cub_bytes, (const unsigned long long *)keys_in,
                                                  (unsigned long long *)keys_out, slots_in, slots_out, (int)cap, 0, 64,
                                                  st);
  g_launches += 8;
  if (e != cudaSuccess) return set_error((int)e, "pcs_hash_build_sorted: cub::DeviceRadixSort::SortPairs");
  PCS_LAUNCH(gather_counts_kernel, blocks_cap, 256, 0, st, table, keys_out, slots_out, (long long)cap, cnt);
  int rc = pcs_exclusive_scan(s, cnt, cap, (int64_t *)offs, w + off[6], pcs_exclusive_scan_tmp_bytes(cap));
  if (rc != 0) return rc;
  PCS_LAUNCH(scatter_starts_kernel, blocks_cap, 256, 0, st, table, keys_out, slots_out, offs, (long long)cap);
  PCS_LAUNCH(hash_scatter_kernel, grid_for(n, 256, 8), 256, 0, st, (const float4 *)pts, (long long)n, g, table,
             (long long)(H - 1), (float4 *)sorted_pts, sorted_idx);
  PCS_LAUNCH(rewind_ranges_kernel, (unsigned)((H + 255) / 256), 256, 0, st, table, (long long)H);
  return 0;
}


}  // extern "C"
