// track.cu -- the cluster tracker batched over ALL (anchor frame, component key) instances of a sequence (sm_100a).
//
// Replaces the Python loops of ClusterTracking.forward / track_frame
// (pcdet/models/registration/preprocessors/cluster_tracking.py:430-787, 853-884) and register_to_next_frame
// (registration_utils.py:83-206).  The reference walks 3 component keys x 25 anchors one after the other and, per
// anchor, 16 target frames x 3 levels of ICP with ~150 launches and >= 8 host syncs per ICP iteration.  The anchors
// and keys are independent, so here every tracking step (anchor a -> frame a + dir * s, the same dir and s for all
// instances) is ONE batch:
//   * trk_sample_*      voxel down-sampling of all moving clouds at once (sample_frame, cluster_tracking.py:39-51):
//                       per-instance grid origin, fp64 means, majority `stationary`, upper-median component
//   * trk_icp_kernel    ONE persistent cooperative launch per level for all instances (one 1024-thread CTA per SM):
//                       per-iteration moving-grid rebuild, two-way K=1 nearest-neighbour search over work lists
//                       handed out by a work counter, exact neighbour caching, fp64 raw-moment reductions (segmented
//                       warp sums per component), Jacobi 3x3 SVD rotation, regulariser, and the reference's
//                       loss-based 3-strike stopping rule evaluated per instance on the device
//   * trk_smooth_kernel AdamW velocity smoothing (smooth_velo, :162-199), one thread-block cluster per instance
//   * trk_update / trk_extract  stopping tests (:675-691) and nearest-neighbour point extraction (:710-721)
// No host synchronisation happens between the first and the last step of a sequence.
//
// Grids here are NOT the reference's RadiusGraph grids: nearest neighbours within a radius do not depend on the
// cell layout, so cells are keyed (group, cx, cy, cz) with a sequence-global origin; only the distance arithmetic
// (fp32, FMA chain starting with the frame difference, torch_hash_kernel.cu:364-370) is the reference's.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <utility>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pcs {

constexpr unsigned int kAll = 0xffffffffu;
constexpr int kTrkThreads = 256;
constexpr int kMomN = 17;  // n, sum m (3), sum r (3), sum m r^T (9), sum |m - r|^2
constexpr int kRelFrames = 17;
constexpr int kAnchorRel = 8;

// ---------------------------------------------------------------------------------------------------------------
// packed-key cell grid
// ---------------------------------------------------------------------------------------------------------------
struct PGrid {
  pcs_slot_t *table;
  unsigned int mask;  // H - 1
  float4 *sorted;     // cell-sorted rows (x = payload bits, yzw = xyz)
  int *sidx;          // original row of every sorted row (nullptr = identity)
  int *cells;         // slots claimed by the current build
  int *ctr;           // [0] claimed cells, [1] scatter cursor, [2] error flag
  const int *tie;     // optional per-row tie-break key (component id): equal distances are resolved by (tie, row)
                      // instead of the row index alone -- row order of the moving voxels is not reproducible
  float lo0, lo1, lo2, inv_cs, cs;
};

__device__ __forceinline__ int clamp_cell(int c) { return c < 0 ? 0 : (c > 65535 ? 65535 : c); }

__device__ __forceinline__ float cell_u(float p, float lo, float inv) { return __fmul_rn(__fsub_rn(p, lo), inv); }

__device__ __forceinline__ long long pkey(int group, int cx, int cy, int cz) {
  return ((long long)group << 48) | ((long long)cx << 32) | ((long long)cy << 16) | (long long)cz;
}

__device__ __forceinline__ long long pg_point_key(const PGrid &g, int group, float x, float y, float z) {
  return pkey(group, clamp_cell((int)floorf(cell_u(x, g.lo0, g.inv_cs))), clamp_cell((int)floorf(cell_u(y, g.lo1, g.inv_cs))),
              clamp_cell((int)floorf(cell_u(z, g.lo2, g.inv_cs))));
}

__device__ __forceinline__ bool pg_lookup(const PGrid &g, long long key, int &start, int &count) {
  unsigned int slot = hash_key(key) & g.mask;
  const int klo = (int)(unsigned int)key, khi = (int)(key >> 32);
  for (unsigned int probes = 0; probes <= g.mask; ++probes) {
    const int4 v = *reinterpret_cast<const int4 *>(g.table + slot);
    if (v.x == klo && v.y == khi) {
      start = v.z;
      count = v.w;
      return true;
    }
    if ((v.x & v.y) == -1) return false;
    slot = (slot + 1) & g.mask;
  }
  return false;
}

// find-or-claim the slot of `key`; a claimer appends the slot to g.cells.  Returns the slot or -1 (table full).
__device__ __forceinline__ int pg_claim(const PGrid &g, long long key) {
  unsigned int slot = hash_key(key) & g.mask;
  for (unsigned int probes = 0; probes <= g.mask; ++probes) {
    const long long cur = *((volatile long long *)&g.table[slot].key);
    if (cur == key) return (int)slot;
    if (cur == PCS_EMPTY_KEY) {
      const unsigned long long prev =
          atomicCAS((unsigned long long *)&g.table[slot].key, (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
      if (prev == (unsigned long long)PCS_EMPTY_KEY) {
        g.cells[atomicAdd(&g.ctr[0], 1)] = (int)slot;
        return (int)slot;
      }
      if ((long long)prev == key) return (int)slot;
    }
    slot = (slot + 1) & g.mask;
  }
  atomicExch(&g.ctr[2], PCS_ERR_TABLE_FULL);
  return -1;
}

__device__ __forceinline__ void pg_count(const PGrid &g, int group, float x, float y, float z) {
  const int slot = pg_claim(g, pg_point_key(g, group, x, y, z));
  if (slot >= 0) atomicAdd(&g.table[slot].count, 1);
}

// ranges of the claimed cells (any order) ; afterwards slot.start is a scatter cursor
__device__ __forceinline__ void pg_ranges(const PGrid &g, long long tid, long long nth) {
  const int nc = g.ctr[0];
  for (long long i = tid; i < nc; i += nth) {
    const int slot = g.cells[i];
    const int c = g.table[slot].count;
    g.table[slot].start = atomicAdd(&g.ctr[1], c);
  }
}

// slot.start becomes the END of the cell's range ("cursor mode")
__device__ __forceinline__ void pg_scatter(const PGrid &g, int group, float x, float y, float z, unsigned int payload,
                                           int row) {
  const long long key = pg_point_key(g, group, x, y, z);
  unsigned int slot = hash_key(key) & g.mask;
  unsigned int probes = 0;
  while (g.table[slot].key != key && probes <= g.mask) {
    slot = (slot + 1) & g.mask;
    ++probes;
  }
  if (probes > g.mask) return;
  const int pos = atomicAdd(&g.table[slot].start, 1);
  g.sorted[pos] = make_float4(__uint_as_float(payload), x, y, z);
  if (g.sidx) g.sidx[pos] = row;
}

__device__ __forceinline__ void pg_clear_used(const PGrid &g, long long tid, long long nth) {
  const int nc = g.ctr[0];
  const int4 e = make_int4(-1, -1, 0, 0);
  for (long long i = tid; i < nc; i += nth) *reinterpret_cast<int4 *>(g.table + g.cells[i]) = e;
}

// reference distance: acc0 = (frame difference)^2, then one FMA per spatial dimension (ref - query)
__device__ __forceinline__ float dist2_acc(float acc, float rx, float ry, float rz, float qx, float qy, float qz) {
  float d = __fsub_rn(rx, qx);
  acc = __fmaf_rn(d, d, acc);
  d = __fsub_rn(ry, qy);
  acc = __fmaf_rn(d, d, acc);
  d = __fsub_rn(rz, qz);
  acc = __fmaf_rn(d, d, acc);
  return acc;
}

// Nearest stored row of `group` within the radius (d2 <= r2), searched by one warp over the 27 cells around the
// query; cells are visited in ascending lower bound and pruned against the best distance found so far.  Rows whose
// payload has a bit of `skipmask` set are ignored.  Returns the row's original index (warp-uniform) or -1; ties by
// ascending index.
// one cell swept by the warp, 32 rows per step; (best, bidx, bound) are warp-uniform on entry and on return
// per-lane record of the two nearest rows a search has scanned (any distance): gives, after the search, a lower bound
// of the distance to every row OTHER than the winner
struct Near2 {
  float d1, d2;      // smallest and second smallest squared distance seen by this lane
  unsigned int i1;   // row of d1
};

// warp-wide minimum of (best, bidx) in the order (d2, tie key, row)
__device__ __forceinline__ void nn_reduce_best(unsigned long long &best, unsigned int &bidx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(kAll, best, o);
    const unsigned int ti = __shfl_xor_sync(kAll, bidx, o);
    if (t < best || (t == best && ti < bidx)) {
      best = t;
      bidx = ti;
    }
  }
}

__device__ __forceinline__ void nn_sweep(const PGrid &g, int cs, int cc, float qx, float qy, float qz, float acc0,
                                         unsigned int skipmask, int lane, unsigned long long &best, unsigned int &bidx,
                                         float &bound, Near2 *near = nullptr, bool lazy = false) {
  for (int j = lane; j < cc; j += 32) {
    const float4 p = g.sorted[cs + j];
    if (__float_as_uint(p.x) & skipmask) continue;
    const float d2 = dist2_acc(acc0, p.y, p.z, p.w, qx, qy, qz);
    const unsigned int idx = g.sidx ? (unsigned int)g.sidx[cs + j] : (unsigned int)(cs + j);
    if (near) {
      if (d2 < near->d1) {
        near->d2 = near->d1;
        near->d1 = d2;
        near->i1 = idx;
      } else if (d2 < near->d2) {
        near->d2 = d2;
      }
    }
    if (d2 <= bound) {
      const unsigned long long k =
          ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)(g.tie ? g.tie[idx] : (int)idx);
      if (k < best || (k == best && idx < bidx)) {
        best = k;
        bidx = idx;
      }
    }
  }
  if (lazy) {
    // (best, bidx) stay per lane (nn_reduce_best merges them once, after the last cell); only the pruning bound is
    // shared: the smallest distance any lane holds (non-negative floats order like their bit patterns)
    const unsigned int m = __reduce_min_sync(kAll, (unsigned int)(best >> 32));
    if (m != 0xffffffffu) bound = fminf(bound, __uint_as_float(m));
    return;
  }
  nn_reduce_best(best, bidx);
  if (best != ~0ull) bound = fminf(bound, __uint_as_float((unsigned int)(best >> 32)));
}

__device__ int nn_search(const PGrid &g, bool cursor_mode, int group, float qx, float qy, float qz, float acc0,
                         float r2, unsigned int skipmask, int lane, float *d2_out = nullptr,
                         unsigned long long best0 = ~0ull, unsigned long long *key_out = nullptr,
                         float *others_lb2 = nullptr, float margin = 0.f) {
  const float ux = cell_u(qx, g.lo0, g.inv_cs), uy = cell_u(qy, g.lo1, g.inv_cs), uz = cell_u(qz, g.lo2, g.inv_cs);
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const int cx = clamp_cell((int)fx), cy = clamp_cell((int)fy), cz = clamp_cell((int)fz);
  int start = 0, count = 0;
  unsigned int sel = 0xffffffffu;
  const float kInf = __int_as_float(0x7f800000);
  float lbcell = kInf;  // lower bound of the rows in this lane's cell if the cell is never swept
  Near2 near = {kInf, kInf, 0xffffffffu};
  // best0: a known candidate (d2 bits << 32 | index) -- its distance prunes the cell lookups from the start
  const float r2_full = r2;
  if (best0 != ~0ull) r2 = fminf(r2, __uint_as_float((unsigned int)(best0 >> 32)));
  // margin > 0 (neighbour caching): cells up to `margin` metres beyond the current best distance are swept as well.
  // They cannot hold a better row, but their rows then enter the lower bound of "every other row" exactly instead of
  // through the distance to the cell wall, which for a query next to a wall is no bound at all.
  float r2_look = r2;
  if (margin > 0.f) {
    const float e = sqrtf(fmaxf(r2 - acc0, 0.f)) + margin;
    r2_look = fminf(r2_full, e * e + acc0);
  }
  if (lane < 27) {
    const int ox = lane % 3 - 1, oy = (lane / 3) % 3 - 1, oz = lane / 9 - 1;
    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
    if (nx >= 0 && nx <= 65535 && ny >= 0 && ny <= 65535 && nz >= 0 && nz <= 65535) {
      float gx = ox == 0 ? 0.f : (ox > 0 ? (fx + 1.f - ux) : (ux - fx));
      float gy = oy == 0 ? 0.f : (oy > 0 ? (fy + 1.f - uy) : (uy - fy));
      float gz = oz == 0 ? 0.f : (oz > 0 ? (fz + 1.f - uz) : (uz - fz));
      gx -= 4e-6f * (fabsf(ux) + 1.f);
      gy -= 4e-6f * (fabsf(uy) + 1.f);
      gz -= 4e-6f * (fabsf(uz) + 1.f);
      gx = (ox != 0 && gx > 0.f) ? gx * g.cs : 0.f;
      gy = (oy != 0 && gy > 0.f) ? gy * g.cs : 0.f;
      gz = (oz != 0 && gz > 0.f) ? gz * g.cs : 0.f;
      const float dmin2 = (gx * gx + gy * gy + gz * gz + acc0) * 0.99999f;
      if (dmin2 <= r2_look) {
        int s = 0, c = 0;
        if (pg_lookup(g, pkey(group, nx, ny, nz), s, c) && c > 0) {
          start = cursor_mode ? s - c : s;
          count = c;
          sel = (__float_as_uint(dmin2) & ~31u) | (unsigned int)lane;
        }
      } else {
        lbcell = dmin2;  // pruned unseen: whatever it holds is at least this far
      }
    }
  }
  // candidates are ordered by (d2, tie key, row); with g.tie == nullptr the tie key is the row itself
  unsigned long long best = ~0ull;
  unsigned int bidx = 0xffffffffu;
  if (best0 != ~0ull) {
    bidx = (unsigned int)(best0 & 0xffffffffu);
    best = (best0 & 0xffffffff00000000ull) | (unsigned int)(g.tie ? g.tie[bidx] : (int)bidx);
  }
  float bound = r2;
  while (true) {
    const unsigned int pick = __reduce_min_sync(kAll, sel);
    if (pick == 0xffffffffu) break;
    float reach = bound;
    if (margin > 0.f) {
      const float e = sqrtf(fmaxf(bound - acc0, 0.f)) + margin;
      reach = fminf(r2_full, e * e + acc0);
    }
    if (__uint_as_float(pick & ~31u) > reach) break;
    const int src = pick & 31;
    if (lane == src) sel = 0xffffffffu;
    const int cs = __shfl_sync(kAll, start, src), cc = __shfl_sync(kAll, count, src);
    nn_sweep(g, cs, cc, qx, qy, qz, acc0, skipmask, lane, best, bidx, bound, others_lb2 ? &near : nullptr, true);
  }
  nn_reduce_best(best, bidx);
  if (others_lb2) {
    // lower bound of the squared distance to every stored row except the winner: the nearest other scanned row, the
    // cells pruned or left unvisited, and the border of the 3x3x3 block (everything outside is a full cell away)
    const float border = 0.98f * g.cs;
    float lb = (best != ~0ull && near.i1 == bidx) ? near.d2 : near.d1;
    lb = fminf(fminf(lb, lbcell), border * border + acc0);
    if (sel != 0xffffffffu) lb = fminf(lb, __uint_as_float(sel & ~31u));  // found, but the sweep stopped before it
    *others_lb2 = __uint_as_float(__reduce_min_sync(kAll, __float_as_uint(lb)));  // lb >= 0: bit patterns order
  }
  if (best == ~0ull) return -1;
  if (d2_out) *d2_out = __uint_as_float((unsigned int)(best >> 32));
  if (key_out) *key_out = best;
  return (int)bidx;
}

// Outer shell of the 5x5x5 block (cells with a coordinate offset of +-2): continues a search started by nn_search on
// the inner 27 cells when the ball of the current bound sticks out of them.  `best` / `bound` are warp-uniform.
__device__ void nn_search_shell(const PGrid &g, bool cursor_mode, int group, float qx, float qy, float qz, float acc0,
                                unsigned int skipmask, int lane, unsigned long long &best, unsigned int &bidx,
                                float &bound) {
  const float ux = cell_u(qx, g.lo0, g.inv_cs), uy = cell_u(qy, g.lo1, g.inv_cs), uz = cell_u(qz, g.lo2, g.inv_cs);
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const int cx = clamp_cell((int)fx), cy = clamp_cell((int)fy), cz = clamp_cell((int)fz);
  for (int base = 0; base < 125; base += 32) {
    const int idx = base + lane;
    int start = 0, count = 0;
    if (idx < 125) {
      const int ox = idx % 5 - 2, oy = (idx / 5) % 5 - 2, oz = idx / 25 - 2;
      const bool outer = ox == 2 || ox == -2 || oy == 2 || oy == -2 || oz == 2 || oz == -2;
      const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
      if (outer && nx >= 0 && nx <= 65535 && ny >= 0 && ny <= 65535 && nz >= 0 && nz <= 65535) {
        float gx = ox == 0 ? 0.f : (ox > 0 ? (fx + (float)ox - ux) : (ux - fx + (float)(-ox - 1)));
        float gy = oy == 0 ? 0.f : (oy > 0 ? (fy + (float)oy - uy) : (uy - fy + (float)(-oy - 1)));
        float gz = oz == 0 ? 0.f : (oz > 0 ? (fz + (float)oz - uz) : (uz - fz + (float)(-oz - 1)));
        gx -= 4e-6f * (fabsf(ux) + 1.f);
        gy -= 4e-6f * (fabsf(uy) + 1.f);
        gz -= 4e-6f * (fabsf(uz) + 1.f);
        gx = (ox != 0 && gx > 0.f) ? gx * g.cs : 0.f;
        gy = (oy != 0 && gy > 0.f) ? gy * g.cs : 0.f;
        gz = (oz != 0 && gz > 0.f) ? gz * g.cs : 0.f;
        const float dmin2 = (gx * gx + gy * gy + gz * gz + acc0) * 0.99999f;
        if (dmin2 <= bound) {
          int s = 0, c = 0;
          if (pg_lookup(g, pkey(group, nx, ny, nz), s, c) && c > 0) {
            start = cursor_mode ? s - c : s;
            count = c;
          }
        }
      }
    }
    unsigned int todo = __ballot_sync(kAll, count > 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int cs = __shfl_sync(kAll, start, src), cc = __shfl_sync(kAll, count, src);
      nn_sweep(g, cs, cc, qx, qy, qz, acc0, skipmask, lane, best, bidx, bound);
    }
  }
}

// nearest row within the radius over (2 * rings + 1)^3 cells (rings = 1 or 2: cell size >= radius / rings)
__device__ __forceinline__ int nn_search_rings(const PGrid &g, bool cursor_mode, int group, float qx, float qy, float qz,
                                               float acc0, float r2, unsigned int skipmask, int lane, int rings,
                                               unsigned long long best0 = ~0ull, float *others_lb2 = nullptr,
                                               float margin = 0.f) {
  float d2 = 0.f;
  unsigned long long best = ~0ull;
  int r = nn_search(g, cursor_mode, group, qx, qy, qz, acc0, r2, skipmask, lane, &d2, best0, &best,
                    rings < 2 ? others_lb2 : nullptr, rings < 2 ? margin : 0.f);
  if (rings < 2) return r;
  if (others_lb2) *others_lb2 = 0.f;  // the two-ring search keeps no bound
  float bound = r >= 0 ? d2 : r2;
  // everything outside the inner block is farther than one cell (minus the fp32 slack of the cell coordinates)
  const float cover = 0.98f * g.cs;
  if (bound - acc0 <= cover * cover) return r;
  unsigned int bidx = r >= 0 ? (unsigned int)r : 0xffffffffu;
  if (r < 0) best = ~0ull;
  nn_search_shell(g, cursor_mode, group, qx, qy, qz, acc0, skipmask, lane, best, bidx, bound);
  if (best == ~0ull) return -1;
  return (int)bidx;
}

// One THREAD searches the inner 27 cells for the row nearest to the query with d2 <= bound.  Exact whenever the ball
// of `bound` lies inside the 27-cell block, i.e. (bound - acc0) <= (0.98 * cs)^2 -- the caller guarantees it (the
// bound is the distance to the neighbour of the previous ICP iteration).  Returns (d2 bits << 32 | index) or ~0.
__device__ unsigned long long nn_search_thread(const PGrid &g, bool cursor_mode, int group, float qx, float qy, float qz,
                                               float acc0, float bound, unsigned int skipmask) {
  const float ux = cell_u(qx, g.lo0, g.inv_cs), uy = cell_u(qy, g.lo1, g.inv_cs), uz = cell_u(qz, g.lo2, g.inv_cs);
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const int cx = clamp_cell((int)fx), cy = clamp_cell((int)fy), cz = clamp_cell((int)fz);
  // per-axis gaps to the lower / upper neighbour cell (metres, conservative)
  float glo[3], ghi[3];
  {
    const float u[3] = {ux, uy, uz}, f[3] = {fx, fy, fz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float m = 4e-6f * (fabsf(u[k]) + 1.f);
      const float a = u[k] - f[k] - m, b = f[k] + 1.f - u[k] - m;
      glo[k] = a > 0.f ? a * g.cs : 0.f;
      ghi[k] = b > 0.f ? b * g.cs : 0.f;
    }
  }
  unsigned long long best = ~0ull;
  unsigned int bidx = 0xffffffffu;
  for (int i = 0; i < 27; i++) {
    // own cell first (i = 0), then the others
    const int t = i == 0 ? 13 : (i <= 13 ? i - 1 : i);
    const int ox = t % 3 - 1, oy = (t / 3) % 3 - 1, oz = t / 9 - 1;
    const float gx = ox == 0 ? 0.f : (ox > 0 ? ghi[0] : glo[0]);
    const float gy = oy == 0 ? 0.f : (oy > 0 ? ghi[1] : glo[1]);
    const float gz = oz == 0 ? 0.f : (oz > 0 ? ghi[2] : glo[2]);
    if ((gx * gx + gy * gy + gz * gz + acc0) * 0.99999f > bound) continue;
    const int nx = cx + ox, ny = cy + oy, nz = cz + oz;
    if (nx < 0 || nx > 65535 || ny < 0 || ny > 65535 || nz < 0 || nz > 65535) continue;
    int s = 0, c = 0;
    if (!pg_lookup(g, pkey(group, nx, ny, nz), s, c) || c <= 0) continue;
    if (cursor_mode) s -= c;
    for (int j = 0; j < c; j++) {
      const float4 p = g.sorted[s + j];
      if (__float_as_uint(p.x) & skipmask) continue;
      const float d2 = dist2_acc(acc0, p.y, p.z, p.w, qx, qy, qz);
      if (d2 <= bound) {
        const unsigned int idx = g.sidx ? (unsigned int)g.sidx[s + j] : (unsigned int)(s + j);
        const unsigned long long k =
            ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)(g.tie ? g.tie[idx] : (int)idx);
        if (k < best || (k == best && idx < bidx)) {
          best = k;
          bidx = idx;
          bound = d2;
        }
      }
    }
  }
  return best == ~0ull ? ~0ull : ((best & 0xffffffff00000000ull) | bidx);
}

// upper_bound(off, n + 1 entries, x) - 1 : the segment that holds item x
__device__ __forceinline__ int seg_of_item(const int *off, int n, int x) {
  int lo = 0, hi = n;  // invariant: off[lo] <= x < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid;
    else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------------------------
// setup helpers for the static grids
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trk_cell_keys_kernel(const float4 *__restrict__ pts, const int *__restrict__ group,
                                                            int n, float lo0, float lo1, float lo2, float inv_cs,
                                                            long long *__restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  PGrid g;
  g.lo0 = lo0, g.lo1 = lo1, g.lo2 = lo2, g.inv_cs = inv_cs;
  keys[i] = pg_point_key(g, group[i], p.y, p.z, p.w);
}

__global__ void __launch_bounds__(256) trk_table_clear_kernel(int4 *__restrict__ table, long long H, int *ctr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int4 e = make_int4(-1, -1, 0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) table[i] = e;
  if (ctr && blockIdx.x == 0 && threadIdx.x < 4) ctr[threadIdx.x] = 0;
}

// table <- (key, start, count) of n unique cells (rows of the key-sorted point array)
__global__ void __launch_bounds__(256) trk_grid_fill_kernel(pcs_slot_t *table, unsigned int mask,
                                                            const long long *__restrict__ keys,
                                                            const int *__restrict__ starts,
                                                            const int *__restrict__ counts, int n, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long key = keys[i];
  unsigned int slot = hash_key(key) & mask;
  for (unsigned int probes = 0; probes <= mask; ++probes) {
    const unsigned long long prev =
        atomicCAS((unsigned long long *)&table[slot].key, (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
    if (prev == (unsigned long long)PCS_EMPTY_KEY) {
      table[slot].start = starts[i];
      table[slot].count = counts[i];
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicExch(err, PCS_ERR_TABLE_FULL);
}

// ---------------------------------------------------------------------------------------------------------------
// batched voxel sampler (sample_frame, cluster_tracking.py:39-51 on GridSampling3D, grid_sampling.py:22-46)
// ---------------------------------------------------------------------------------------------------------------
struct SampSlot {  // 16 bytes
  long long key;
  int head;  // linked list of the voxel's points (-1 = end)
  int cnt;
};

struct SampArgs {
  const float4 *pts;          // [n] (., x, y, z)
  const int *group;           // [n] instance (or frame) of every point
  const int *skey;            // [n] sort key of the point (component id); nullptr = the group is the sort key
  const unsigned char *bits;  // [n] up to three flag bits per point (stationary per component key)
  const int *act;             // [n_groups] or nullptr
  int n, n_groups, n_keys, ns_only;
  float s1, s2, s3;
  unsigned int *sb;  // [n_groups][6] ordered-uint bounds (min xyz, max xyz) of the groups' points
  SampSlot *table;
  unsigned int mask;
  double *vsum;  // [H][3]
  int *vbits;    // [H][3]
  int *vk;       // [H][2] min / max sort key
  int *pnext;    // [n]
  int *vlist;    // [n] claimed slots
  int *ctr;      // [0] voxels, [1] kept voxels, [2] error
  int *vres;     // [H][2]: sort key, flag bits
  int *kcount;   // [n_keys + 1] kept voxels per sort key
  int *koff;     // [n_keys + 1]
  int *kcur;     // [n_keys]
  int *vdeg;     // [n_keys] all voxels per sort key (optional)
  float4 *out_pts;  // [n] (flag bits, mean xyz) grouped by sort key
  int *out_key;
  int *out_group;
};

__device__ __forceinline__ void bounds_accumulate(unsigned int *sb, int group, bool valid, float x, float y, float z,
                                                  int lane) {
  // warp-aggregated when every lane belongs to the same group (the usual case: points are grouped)
  const int g0 = __shfl_sync(kAll, group, 0);
  const bool uniform = __all_sync(kAll, valid && group == g0);
  if (uniform) {
    const unsigned int ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
    const unsigned int mnx = __reduce_min_sync(kAll, ox), mny = __reduce_min_sync(kAll, oy), mnz = __reduce_min_sync(kAll, oz);
    const unsigned int mxx = __reduce_max_sync(kAll, ox), mxy = __reduce_max_sync(kAll, oy), mxz = __reduce_max_sync(kAll, oz);
    if (lane == 0) {
      unsigned int *b = sb + (long long)g0 * 6;
      atomicMin(b + 0, mnx);
      atomicMin(b + 1, mny);
      atomicMin(b + 2, mnz);
      atomicMax(b + 3, mxx);
      atomicMax(b + 4, mxy);
      atomicMax(b + 5, mxz);
    }
  } else if (valid) {
    unsigned int *b = sb + (long long)group * 6;
    atomicMin(b + 0, f2ord(x));
    atomicMin(b + 1, f2ord(y));
    atomicMin(b + 2, f2ord(z));
    atomicMax(b + 3, f2ord(x));
    atomicMax(b + 4, f2ord(y));
    atomicMax(b + 5, f2ord(z));
  }
}

__global__ void __launch_bounds__(256) trk_bounds_reset_kernel(unsigned int *sb, int n_groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_groups * 6) sb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) trk_group_bounds_kernel(const float4 *__restrict__ pts,
                                                               const int *__restrict__ group, int n, unsigned int *sb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  int g = 0;
  if (valid) {
    p = pts[i];
    g = group[i];
  }
  bounds_accumulate(sb, g, valid, p.y, p.z, p.w, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256) trk_samp_reset_kernel(SampArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= A.n_keys) {
    A.kcount[i] = 0;
    if (i < A.n_keys) A.kcur[i] = 0;
  }
  if (i < 2) A.ctr[i] = 0;  // [2] is a sticky error flag
}

__global__ void __launch_bounds__(256) trk_samp_insert_kernel(SampArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int g = A.group[i];
  if (A.act && !A.act[g]) return;
  const float4 p = A.pts[i];
  const unsigned int *b = A.sb + (long long)g * 6;
  // GridSampling3D: start = min of the group's points, cell = trunc((p - start) / size) in fp32 (IEEE division)
  const float sx = ord2f(b[0]), sy = ord2f(b[1]), sz = ord2f(b[2]);
  int c1 = (int)__fdiv_rn(__fsub_rn(p.y, sx), A.s1);
  int c2 = (int)__fdiv_rn(__fsub_rn(p.z, sy), A.s2);
  int c3 = (int)__fdiv_rn(__fsub_rn(p.w, sz), A.s3);
  if (c1 < 0 || c1 > 65535 || c2 < 0 || c2 > 65535 || c3 < 0 || c3 > 65535) {
    atomicExch(&A.ctr[2], PCS_ERR_KEY_RANGE);
    c1 = clamp_cell(c1);
    c2 = clamp_cell(c2);
    c3 = clamp_cell(c3);
  }
  const long long key = pkey(g, c1, c2, c3);
  unsigned int slot = hash_key(key) & A.mask;
  bool ok = false;
  for (unsigned int probes = 0; probes <= A.mask; ++probes) {
    const long long cur = *((volatile long long *)&A.table[slot].key);
    if (cur == key) {
      ok = true;
      break;
    }
    if (cur == PCS_EMPTY_KEY) {
      const unsigned long long prev =
          atomicCAS((unsigned long long *)&A.table[slot].key, (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
      if (prev == (unsigned long long)PCS_EMPTY_KEY) {
        A.vlist[atomicAdd(&A.ctr[0], 1)] = (int)slot;
        ok = true;
        break;
      }
      if ((long long)prev == key) {
        ok = true;
        break;
      }
    }
    slot = (slot + 1) & A.mask;
  }
  if (!ok) {
    atomicExch(&A.ctr[2], PCS_ERR_TABLE_FULL);
    return;
  }
  atomicAdd(&A.table[slot].cnt, 1);
  double *s = A.vsum + (long long)slot * 3;
  atomicAdd(s + 0, (double)p.y);
  atomicAdd(s + 1, (double)p.z);
  atomicAdd(s + 2, (double)p.w);
  if (A.bits) {
    const unsigned int f = A.bits[i];
    if (f & 1u) atomicAdd(A.vbits + (long long)slot * 3 + 0, 1);
    if (f & 2u) atomicAdd(A.vbits + (long long)slot * 3 + 1, 1);
    if (f & 4u) atomicAdd(A.vbits + (long long)slot * 3 + 2, 1);
  }
  if (A.skey) {
    const int k = A.skey[i];
    atomicMin(A.vk + (long long)slot * 2 + 0, k);
    atomicMax(A.vk + (long long)slot * 2 + 1, k);
    A.pnext[i] = atomicExch(&A.table[slot].head, i);
  }
}

// per voxel: flag majority, sort key (upper median of the members' keys, registration_utils.py:60-81), counts
__global__ void __launch_bounds__(256) trk_samp_finalize1_kernel(SampArgs A) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= A.ctr[0]) return;
  const int slot = A.vlist[v];
  const int cnt = A.table[slot].cnt;
  unsigned int fb = 0;
  if (A.bits) {
    const int *vb = A.vbits + (long long)slot * 3;
    // mean(flag) > 0.5  <=>  2 * (#set) > count
    fb = (2 * vb[0] > cnt ? 1u : 0u) | (2 * vb[1] > cnt ? 2u : 0u) | (2 * vb[2] > cnt ? 4u : 0u);
  }
  int key;
  if (A.skey) {
    const int kmin = A.vk[(long long)slot * 2], kmax = A.vk[(long long)slot * 2 + 1];
    key = kmin;
    if (kmin != kmax) {
      // element of rank cnt / 2 of the sorted member keys, by counting over the voxel's point list
      const int target = cnt / 2;
      for (int a = A.table[slot].head; a >= 0; a = A.pnext[a]) {
        const int va = A.skey[a];
        int less = 0, eq = 0;
        for (int c = A.table[slot].head; c >= 0; c = A.pnext[c]) {
          const int vc = A.skey[c];
          less += vc < va;
          eq += vc == va;
        }
        if (less <= target && target < less + eq) {
          key = va;
          break;
        }
      }
    }
  } else {
    key = (int)(A.table[slot].key >> 48);
  }
  A.vres[(long long)slot * 2 + 0] = key;
  A.vres[(long long)slot * 2 + 1] = (int)fb;
  if (A.vdeg) atomicAdd(A.vdeg + key, 1);
  if (!(A.ns_only && (fb & 1u))) atomicAdd(A.kcount + key, 1);
}

// single CTA: koff = exclusive scan of kcount (n_keys + 1 entries, the last one is the total) ; ctr[1] = total
__global__ void __launch_bounds__(1024) trk_scan_kernel(const int *__restrict__ in, int *__restrict__ out, int n,
                                                        int *total_out) {
  __shared__ int s_warp[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      const int wi = warp_incl_scan(w, lane);
      s_warp[lane] = wi - w;
      if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    const int carry = s_carry;
    if (i < n) out[i] = carry + s_warp[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_warp[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = s_carry;
    if (total_out) *total_out = s_carry;
  }
}

__global__ void __launch_bounds__(256) trk_samp_finalize2_kernel(SampArgs A) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  // the bounds were consumed by the insert pass: reset them for the next producer
  if (v < A.n_groups * 6) A.sb[v] = (v % 6) < 3 ? 0xffffffffu : 0u;
  if (v >= A.ctr[0]) return;
  const int slot = A.vlist[v];
  const int key = A.vres[(long long)slot * 2], fb = A.vres[(long long)slot * 2 + 1];
  const int cnt = A.table[slot].cnt;
  const int group = (int)(A.table[slot].key >> 48);
  double *s = A.vsum + (long long)slot * 3;
  if (!(A.ns_only && (fb & 1))) {
    const int pos = A.koff[key] + atomicAdd(A.kcur + key, 1);
    const double inv = 1.0 / (double)(cnt > 0 ? cnt : 1);
    A.out_pts[pos] = make_float4(__int_as_float(fb), (float)(s[0] * inv), (float)(s[1] * inv), (float)(s[2] * inv));
    A.out_key[pos] = key;
    A.out_group[pos] = group;
  }
  // self-cleaning: the slot and its rows are ready for the next sampling pass
  A.table[slot].key = PCS_EMPTY_KEY;
  A.table[slot].head = -1;
  A.table[slot].cnt = 0;
  s[0] = s[1] = s[2] = 0.0;
  A.vbits[(long long)slot * 3 + 0] = A.vbits[(long long)slot * 3 + 1] = A.vbits[(long long)slot * 3 + 2] = 0;
  A.vk[(long long)slot * 2 + 0] = 0x7fffffff;
  A.vk[(long long)slot * 2 + 1] = -1;
}

__global__ void __launch_bounds__(256) trk_samp_init_kernel(SampArgs A, long long H) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) {
    A.table[i].key = PCS_EMPTY_KEY;
    A.table[i].head = -1;
    A.table[i].cnt = 0;
    A.vsum[i * 3 + 0] = A.vsum[i * 3 + 1] = A.vsum[i * 3 + 2] = 0.0;
    A.vbits[i * 3 + 0] = A.vbits[i * 3 + 1] = A.vbits[i * 3 + 2] = 0;
    A.vk[i * 2 + 0] = 0x7fffffff;
    A.vk[i * 2 + 1] = -1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// batched ICP (register_to_next_frame, registration_utils.py:83-206) -- one persistent cooperative launch
// ---------------------------------------------------------------------------------------------------------------
enum { PH_RUN = 0, PH_FINISHING = 1, PH_FROZEN = 2 };

struct IcpB {
  int J, G;
  const int *act;        // [J]
  const int *ref_group;  // [J] group of the instance's reference voxels (non-stationary voxels of its key and frame)
  const int *ref_group_all;  // [J] group holding ALL voxels of the target frame (matched-fraction search)
  const int *skipmask;   // [J] stationary bit of the instance's component key
  const int *ref_off;    // [n_groups + 1] voxel range of every group in the (group-major) reference array
  const int *g_inst;     // [G]
  PGrid ref;             // static reference grid (sorted == the reference voxel array, sidx == nullptr)
  PGrid mov;             // moving grid, rebuilt every iteration
  float4 *mv;            // [n_mv] moving voxels grouped by component (., x, y, z), updated in place
  const int *mv_gid;
  const int *mv_inst;
  const int *n_mv;  // device count
  const int *vdeg;  // [G] voxels per component (stationary ones included)
  float r2, acc0, cover2;
  float margin;  // neighbour caching: extra sweep reach beyond the best distance (metres)
  int batch;
  int rings, mode;  // mode 0: warp search seeded with the previous neighbour, 1: + thread-level fast path, 2: unseeded
  double angle_reg, stopping_delta;
  int max_iter, want_l1, want_ratio;
  int *nn_fwd, *nn_bwd, *boff;
  int *mvbeg, *mvend;  // [J] rows of every instance in the moving voxel array (voxels are grouped by instance)
  // neighbour caching: per query a lower bound (3-D metres) of the distance to every target other than its cached
  // neighbour; per moving voxel its displacement in this iteration, per instance the largest of them (float bits)
  float *sec_fwd, *sec_bwd, *disp;
  int *dmax;
  float r3skip;
  double *mom, *Ti, *T, *mu, *l1_sum, *l1_n;
  int *phase, *cd, *iters, *itcnt;
  double *last, *loss;
  int *match_cnt;
  double *l1_err;
  float *ratio;
  long long *prof;  // optional [16]: nanoseconds per phase (B, C, D, E, F, G, H, tail), [8] iterations, [9] launches
};

// Jacobi rotation with one reciprocal square root instead of a square root and a division: the solve of an ICP
// iteration is a pure latency chain (one thread per component, every other warp waits at the barrier), and the fp64
// divisions and square roots are most of it.
#define TRK_JACOBI_ROT(app, aqq, apq, arp, arq, v0p, v0q, v1p, v1q, v2p, v2q)          \
  if ((apq) != 0.0) {                                                                  \
    const double theta = ((aqq) - (app)) / (2.0 * (apq));                              \
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
    const double c = rsqrt(t * t + 1.0), sn = t * c;                                   \
    (app) -= t * (apq);                                                                \
    (aqq) += t * (apq);                                                                \
    (apq) = 0.0;                                                                       \
    { const double x = (arp), y = (arq); (arp) = c * x - sn * y; (arq) = sn * x + c * y; } \
    { const double x = (v0p), y = (v0q); (v0p) = c * x - sn * y; (v0q) = sn * x + c * y; } \
    { const double x = (v1p), y = (v1q); (v1p) = c * x - sn * y; (v1q) = sn * x + c * y; } \
    { const double x = (v2p), y = (v2q); (v2p) = c * x - sn * y; (v2q) = sn * x + c * y; } \
  }

__device__ __forceinline__ Eig3 trk_jacobi_eig3(double a00, double a01, double a02, double a11, double a12, double a22) {
  Eig3 e;
  e.v00 = e.v11 = e.v22 = 1.0;
  e.v01 = e.v02 = e.v10 = e.v12 = e.v20 = e.v21 = 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double diag = fabs(a00) + fabs(a11) + fabs(a22);
    // off-diagonal mass below 1e-16 of the diagonal: the next sweep would square it (quadratic convergence)
    if (off == 0.0 || off <= 1e-16 * diag) break;
    TRK_JACOBI_ROT(a00, a11, a01, a02, a12, e.v00, e.v01, e.v10, e.v11, e.v20, e.v21)
    TRK_JACOBI_ROT(a00, a22, a02, a01, a12, e.v00, e.v02, e.v10, e.v12, e.v20, e.v22)
    TRK_JACOBI_ROT(a11, a22, a12, a01, a02, e.v01, e.v02, e.v11, e.v12, e.v21, e.v22)
  }
  e.d0 = a00;
  e.d1 = a11;
  e.d2 = a22;
  return e;
}

__device__ void kabsch_rotation_b(const double A[9], double R[9]) {
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
  const Eig3 e = trk_jacobi_eig3(b00, b01, b02, b11, b12, b22);
  double ev[3] = {e.d0, e.d1, e.d2};
  double V[3][3] = {{e.v00, e.v01, e.v02}, {e.v10, e.v11, e.v12}, {e.v20, e.v21, e.v22}};
  int i0 = 0, i1 = 1, i2 = 2;
  if (ev[i1] > ev[i0]) { int t = i0; i0 = i1; i1 = t; }
  if (ev[i2] > ev[i0]) { int t = i0; i0 = i2; i2 = t; }
  if (ev[i2] > ev[i1]) { int t = i1; i1 = i2; i2 = t; }
  const int idx[3] = {i0, i1, i2};
  double Vs[3][3], U[3][3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double sv2 = fmax(ev[idx[j]], 0.0);
    const double inv_sv = sv2 > 1e-300 ? rsqrt(sv2) : 0.0;  // 1 / singular value (sv > 1e-150 ... the guard of before)
#pragma unroll
    for (int i = 0; i < 3; i++) Vs[i][j] = V[i][idx[j]];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double av = A[i * 3 + 0] * Vs[0][j] + A[i * 3 + 1] * Vs[1][j] + A[i * 3 + 2] * Vs[2][j];
      U[i][j] = av * inv_sv;
    }
  }
  if (!(ev[i2] > 1e-24 * fmax(ev[i0], 1e-300))) {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
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

__device__ __forceinline__ void apply_T(float4 &p, const double *t) {
  const double x = p.y, y = p.z, z = p.w;
  p.y = (float)(t[0] * x + t[1] * y + t[2] * z + t[9]);
  p.z = (float)(t[3] * x + t[4] * y + t[5] * z + t[10]);
  p.w = (float)(t[6] * x + t[7] * y + t[8] * z + t[11]);
}

__global__ void __launch_bounds__(256) trk_icp_setup_kernel(IcpB A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.G) {
    double *T = A.T + (long long)i * 12;
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
    for (int k = 0; k < kMomN; k++) A.mom[(long long)i * kMomN + k] = 0.0;
    A.l1_sum[i * 2] = A.l1_sum[i * 2 + 1] = 0.0;
    A.l1_n[i] = 0.0;
    A.match_cnt[i] = 0;
    if (A.want_l1) A.l1_err[i] = 0.0;
  }
  if (i < A.J) {
    A.mvbeg[i] = A.mvend[i] = 0;
    if (A.sec_fwd) A.dmax[i] = 0;
    A.phase[i] = A.act[i] ? PH_RUN : PH_FROZEN;
    A.cd[i] = 3;
    A.iters[i] = 0;
    A.last[i] = 1e10;
    A.loss[i] = 0.0;
  }
  if (i < (A.max_iter + 2) * 3) A.itcnt[i] = 0;  // [..][2] finishing / running counts, then one work counter per iteration
  if (i == 0) {
    int o = 0;
    for (int j = 0; j < A.J; j++) {
      A.boff[j] = o;
      if (A.act[j]) o += A.ref_off[A.ref_group[j] + 1] - A.ref_off[A.ref_group[j]];
    }
    A.boff[A.J] = o;
    A.mov.ctr[0] = A.mov.ctr[1] = 0;
  }
}

__global__ void __launch_bounds__(256) trk_icp_ranges_kernel(IcpB A) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *A.n_mv;
  if (v >= n) return;
  const int inst = A.mv_inst[v];
  if (v == 0 || A.mv_inst[v - 1] != inst) A.mvbeg[inst] = v;
  if (v == n - 1 || A.mv_inst[v + 1] != inst) A.mvend[inst] = v + 1;
}

// edge of work item w: forward items are the moving voxels, backward items the instances' reference voxels
__device__ __forceinline__ bool icp_edge_of(const IcpB &A, int w, int nmv, int &mi, int &ri, int want_phase) {
  if (w < nmv) {
    mi = w;
    if (A.phase[A.mv_inst[mi]] != want_phase) return false;
    ri = A.nn_fwd[mi];
  } else {
    const int b = w - nmv;
    const int j = seg_of_item(A.boff, A.J, b);
    if (A.phase[j] != want_phase) return false;
    ri = A.ref_off[A.ref_group[j]] + (b - A.boff[j]);
    mi = A.nn_bwd[b];
  }
  return mi >= 0 && ri >= 0;
}

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

#define ICP_PROF(slot)                                   \
  if (A.prof && tid == 0) {                              \
    const long long now_ = gtime();                      \
    A.prof[slot] += now_ - t_prev;                       \
    t_prev = now_;                                       \
  }

constexpr int kMaxInst = 1024;

template <int kMinBlocks, int kThreads = kTrkThreads>
__global__ void __launch_bounds__(kThreads, kMinBlocks) trk_icp_kernel(IcpB A) {
  __shared__ int s_aoff[kMaxInst + 1];
  cg::grid_group grid = cg::this_grid();
  long long t_prev = gtime();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const long long warp_id = tid >> 5, nwarps = nth >> 5;
  const int nmv = *A.n_mv;
  const int nbwd = A.boff[A.J];
  const int total = nmv + nbwd;
  constexpr int CH = 8;  // consecutive work items per warp (moment sums stay in registers across them)

  for (int it = 0; it < A.max_iter; it++) {
    // ---- B: apply the previous transform (fp64 product stored back to fp32, :179) and count into cells --------
    for (long long v0 = tid - lane; v0 < nmv; v0 += nth) {  // warp-uniform trip count (warp collectives below)
      const long long v = v0 + lane;
      int j = -1;
      float d = 0.f;
      if (v < nmv) {
        j = A.mv_inst[v];
        if (A.phase[j] != PH_RUN) j = -1;
      }
      if (j >= 0) {
        float4 p = A.mv[v];
        if (A.iters[j] > 0) {
          const float ox = p.y, oy = p.z, oz = p.w;
          apply_T(p, A.Ti + (long long)A.mv_gid[v] * 12);
          A.mv[v] = p;
          const float dx = p.y - ox, dy = p.z - oy, dz = p.w - oz;
          d = sqrtf(dx * dx + dy * dy + dz * dz) * 1.00001f + 1e-7f;  // never below the true displacement
        }
        if (A.sec_fwd) A.disp[v] = d;
        pg_count(A.mov, j, p.y, p.z, p.w);
      }
      if (A.sec_fwd) {
        // largest displacement per instance: one atomic per warp when its voxels belong to one instance
        const int jlo = __reduce_min_sync(kAll, j < 0 ? 0x7fffffff : j), jhi = __reduce_max_sync(kAll, j);
        const int dbits = __float_as_int(d);
        if (jhi >= 0 && jlo == jhi) {
          const int m = __reduce_max_sync(kAll, dbits);
          if (lane == 0 && m > 0) atomicMax(&A.dmax[jhi], m);
        } else if (j >= 0 && dbits > 0) {
          atomicMax(&A.dmax[j], dbits);
        }
      }
    }
    grid.sync();
    ICP_PROF(0)
    // ---- C: ranges -----------------------------------------------------------------------------------------
    pg_ranges(A.mov, tid, nth);
    grid.sync();
    ICP_PROF(1)
    // ---- D: scatter ----------------------------------------------------------------------------------------
    for (long long v = tid; v < nmv; v += nth) {
      const int j = A.mv_inst[v];
      if (A.phase[j] != PH_RUN) continue;
      const float4 p = A.mv[v];
      pg_scatter(A.mov, j, p.y, p.z, p.w, 0u, (int)v);
    }
    grid.sync();
    ICP_PROF(2)
    // ---- E: both nearest-neighbour searches + raw moments of the edge set -------------------------------------
    // One work item per lane.  A query whose cached neighbour is provably still the nearest (distance bounds, see
    // below) is settled by its own lane; the others are searched by the whole warp, one after the other, over the 27
    // cells around the query, seeded with the previous neighbour.
    // active work items of this iteration: [forward voxels | backward reference voxels] of every running instance
    // (every thread loads one instance -- the barrier before this phase invalidated L1, a serial loop of one thread
    // over the instances would be a chain of L2 round trips with the rest of the CTA waiting)
    for (int jj = threadIdx.x; jj < A.J; jj += blockDim.x) {
      int nf = 0, nb = 0;
      if (A.phase[jj] == PH_RUN) {
        nf = A.mvend[jj] - A.mvbeg[jj];
        nb = A.boff[jj + 1] - A.boff[jj];
        if (A.prof && blockIdx.x == 0) {
          atomicAdd((unsigned long long *)A.prof + 13, (unsigned long long)nf);
          atomicAdd((unsigned long long *)A.prof + 14, (unsigned long long)nb);
        }
      }
      s_aoff[jj + 1] = nf + nb;
    }
    if (threadIdx.x == 0) s_aoff[0] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {  // in-place inclusive scan of s_aoff[1 .. J] by the first warp
      int carry = 0;
      for (int base = 1; base <= A.J; base += 32) {
        const int i = base + lane;
        int v = i <= A.J ? s_aoff[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kAll, v, o);
          if (lane >= o) v += t;
        }
        v += carry;
        if (i <= A.J) s_aoff[i] = v;
        carry = __shfl_sync(kAll, v, 31);
      }
    }
    __syncthreads();
    const int n_active = s_aoff[A.J];
    // work items per warp batch (<= 32): the items of a batch are searched one after the other by the warp, so in
    // the long tail of an ICP (few instances still iterating) smaller batches spread the latency chains over all warps
    int BQ = (int)(((long long)n_active + nwarps - 1) / nwarps);
    BQ = BQ < 1 ? 1 : (BQ > A.batch ? A.batch : BQ);
    const long long nbatch = ((long long)n_active + BQ - 1) / BQ;
    // Batches are handed out by a per-iteration work counter (a query costs anything between nothing -- cached -- and
    // a multi-cell search, so static shares leave most warps waiting at the barrier for the unlucky ones); the next
    // ticket is drawn before the current batch is processed, so its latency is hidden.  Static mode (PCS_ICP_MODE bit
    // 2): consecutive batches go to different CTAs.
    const bool dyn = (A.mode & 4) == 0;
    int *wctr = A.itcnt + (A.max_iter + 2) * 2 + it;
    long long ticket = (long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    if (dyn) {
      int t0 = 0;
      if (lane == 0) t0 = atomicAdd(wctr, 1);
      ticket = __shfl_sync(kAll, t0, 0);
    }
    while (ticket < nbatch) {
      const long long batch = ticket;
      int tnext = 0;
      if (dyn) {
        if (lane == 0) tnext = atomicAdd(wctr, 1);
      } else {
        ticket += nwarps;
      }
      const int w = (int)(batch * BQ) + lane;
      bool active = lane < BQ && w < n_active;
      bool fwd = true;
      int mi = -1, ri = -1, j = 0, prev = -1, b = 0;
      float qx = 0.f, qy = 0.f, qz = 0.f;
      unsigned int skip = 0u;
      if (active) {
        j = seg_of_item(s_aoff, A.J, w);
        const int local = w - s_aoff[j];
        const int nf = A.mvend[j] - A.mvbeg[j];
        fwd = local < nf;
        if (fwd) {  // forward: moving voxel -> nearest non-stationary reference voxel
          mi = A.mvbeg[j] + local;
          const float4 q = A.mv[mi];
          qx = q.y, qy = q.z, qz = q.w;
          skip = (unsigned int)A.skipmask[j];
          prev = it > 0 ? A.nn_fwd[mi] : -1;
        } else {  // backward: non-stationary reference voxel -> nearest moving voxel of the instance
          b = A.boff[j] + (local - nf);
          ri = A.ref_off[A.ref_group[j]] + (local - nf);
          const float4 q = A.ref.sorted[ri];
          qx = q.y, qy = q.z, qz = q.w;
          if (__float_as_uint(q.x) & (unsigned int)A.skipmask[j]) {
            active = false;
            A.nn_bwd[b] = -1;
          } else {
            prev = it > 0 ? A.nn_bwd[b] : -1;
          }
        }
      }
      int res = -1;
      bool need_warp = active;
      unsigned long long pkey0 = ~0ull;  // the previous neighbour as a first candidate
      // Neighbour caching (exact, triangle inequality): `sec` bounds the distance of every target other than the
      // cached neighbour from below; it shrinks by whatever moved since (the query itself in the forward direction, at
      // most dmax[instance] for the moving targets of the backward direction).  While the cached neighbour is still
      // within the radius and strictly nearer than that bound it IS the nearest neighbour: no search.  An unmatched
      // query (no cached neighbour) stays unmatched while the bound stays above the radius.
      float sec = 0.f;
      bool cached = false;
      if (A.sec_fwd && active && it > 0) {
        sec = (fwd ? A.sec_fwd[mi] : A.sec_bwd[b]) - (fwd ? A.disp[mi] : __int_as_float(A.dmax[j]));
        if (prev < 0 && sec > A.r3skip) cached = true;
      }
      if (active && prev >= 0) {
        const float4 c = fwd ? A.ref.sorted[prev] : A.mv[prev];
        const float d2p = dist2_acc(A.acc0, c.y, c.z, c.w, qx, qy, qz);
        if (d2p <= A.r2) pkey0 = ((unsigned long long)__float_as_uint(d2p) << 32) | (unsigned int)prev;
        if (A.sec_fwd && it > 0 && d2p <= A.r2 &&
            sqrtf(fmaxf(d2p - A.acc0, 0.f)) * 1.00001f + 1e-6f < sec) {
          cached = true;
          res = prev;
        }
        if ((A.mode & 3) == 1 && d2p <= A.r2 && d2p - A.acc0 <= A.cover2) {
          const unsigned long long k = fwd ? nn_search_thread(A.ref, false, A.ref_group[j], qx, qy, qz, A.acc0, d2p, skip)
                                           : nn_search_thread(A.mov, true, j, qx, qy, qz, A.acc0, d2p, 0u);
          res = k == ~0ull ? prev : (int)(unsigned int)(k & 0xffffffffu);
          need_warp = false;
        }
      }
      if (cached) {
        need_warp = false;
        if (fwd) A.sec_fwd[mi] = sec;
        else A.sec_bwd[b] = sec;
      }
      unsigned int todo = __ballot_sync(kAll, need_warp);
      if (A.prof) {
        const unsigned int nc_ = __ballot_sync(kAll, cached);
        if (lane == 0 && nc_) atomicAdd((unsigned long long *)A.prof + 15, (unsigned long long)__popc(nc_));
        const unsigned int na = __ballot_sync(kAll, active);
        if (lane == 0 && na) {
          atomicAdd((unsigned long long *)A.prof + 10, (unsigned long long)__popc(na & ~todo));
          atomicAdd((unsigned long long *)A.prof + 11, (unsigned long long)__popc(todo));
          if (it < 96) atomicAdd((unsigned long long *)A.prof + 112 + it, (unsigned long long)__popc(na));
        }
      }
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const bool sf = __shfl_sync(kAll, (int)fwd, src) != 0;
        const int sj = __shfl_sync(kAll, j, src);
        const float sx = __shfl_sync(kAll, qx, src), sy = __shfl_sync(kAll, qy, src), sz = __shfl_sync(kAll, qz, src);
        const unsigned int ss = __shfl_sync(kAll, skip, src);
        const unsigned long long sk = (A.mode & 3) == 2 ? ~0ull : __shfl_sync(kAll, pkey0, src);
        float olb2 = 0.f;
        float *olb = A.sec_fwd ? &olb2 : nullptr;
        const float mg = olb ? A.margin : 0.f;
        const int r = sf ? nn_search_rings(A.ref, false, A.ref_group[sj], sx, sy, sz, A.acc0, A.r2, ss, lane, A.rings, sk, olb, mg)
                         : nn_search_rings(A.mov, true, sj, sx, sy, sz, A.acc0, A.r2, 0u, lane, A.rings, sk, olb, mg);
        if (lane == src) {
          if (A.prof) {  // what the searches were good for: [208] same neighbour as before, [209] changed, [210] no previous
            const int slot = it == 0 ? 211 : (prev < 0 ? 210 : (r == prev ? 208 : 209));
            atomicAdd((unsigned long long *)A.prof + slot, 1ull);
          }
          res = r;
          sec = sqrtf(fmaxf(olb2 - A.acc0, 0.f)) * 0.9999f;
        }
      }
      int c = -1;
      float4 mp = make_float4(0.f, 0.f, 0.f, 0.f), rp = mp;
      if (active) {
        if (fwd) {
          A.nn_fwd[mi] = res;
          ri = res;
          if (A.sec_fwd && !cached) A.sec_fwd[mi] = sec;
        } else {
          A.nn_bwd[b] = res;
          mi = res;
          if (A.sec_fwd && !cached) A.sec_bwd[b] = sec;
        }
        if (res >= 0) {
          mp = A.mv[mi];
          rp = A.ref.sorted[ri];
          c = A.mv_gid[mi];
        }
      }
      if (A.prof) {
        const unsigned int nu = __ballot_sync(kAll, active && res < 0);
        if (lane == 0 && nu) atomicAdd((unsigned long long *)A.prof + 12, (unsigned long long)__popc(nu));
      }
      // raw moments: segmented warp sums over runs of equal component, one set of atomics per run
      {
        const int prevc = __shfl_up_sync(kAll, c, 1);
        const bool head = lane == 0 || prevc != c;
        const unsigned int heads = __ballot_sync(kAll, head);
        if (__ballot_sync(kAll, c >= 0) != 0u) {
          const unsigned int above = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));
          const int end = above ? (__ffs(above) - 1) : 32;
          const double m[3] = {mp.y, mp.z, mp.w}, r[3] = {rp.y, rp.z, rp.w};
#pragma unroll
          for (int k = 0; k < kMomN; k++) {
            double v;
            if (k == 0) v = 1.0;
            else if (k < 4) v = m[k - 1];
            else if (k < 7) v = r[k - 4];
            else if (k < 16) v = m[(k - 7) / 3] * r[(k - 7) % 3];
            else {
              const double dx = m[0] - r[0], dy = m[1] - r[1], dz = m[2] - r[2];
              v = dx * dx + dy * dy + dz * dz;
            }
            if (c < 0) v = 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const double tv = __shfl_down_sync(kAll, v, o);
              if (lane + o < end) v += tv;
            }
            if (head && c >= 0) atomicAdd(A.mom + (long long)c * kMomN + k, v);
          }
        }
      }
      if (dyn) ticket = __shfl_sync(kAll, tnext, 0);
    }
    grid.sync();
    if (A.prof && tid == 0 && it < 96) A.prof[16 + it] += gtime() - t_prev;
    ICP_PROF(3)
    // ---- F: per-component solve (thread per component) + instance loss ; the moving grid is released ----------
    // consecutive groups of 32 components go to different CTAs: the solve is a long fp64 latency chain per thread and
    // there are only a few hundred warps of it, which should not queue up on the first SMs
    for (long long c0 = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32; c0 < A.G; c0 += nwarps * 32) {
      const long long c = c0 + lane;
      double loss_part = 0.0;
      int j = -1;
      if (c < A.G) {
        j = A.g_inst[c];
        if (A.phase[j] != PH_RUN) j = -1;
      }
      if (j >= 0) {
        double *s = A.mom + c * kMomN;
        const double n = s[0];
        double mu_m[3] = {0, 0, 0}, mu_r[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (n > 0.5) {
          for (int k = 0; k < 3; k++) {  // fp32 segment means cast to fp64 (:150-151)
            mu_m[k] = (double)(float)(s[1 + k] / n);
            mu_r[k] = (double)(float)(s[4 + k] / n);
          }
          for (int i = 0; i < 3; i++)
            for (int q = 0; q < 3; q++)
              cov[i * 3 + q] = (s[7 + i * 3 + q] - mu_m[i] * s[4 + q] - s[1 + i] * mu_r[q] + n * mu_m[i] * mu_r[q]) / n;
          double dd = 0.0, cross = 0.0;
          for (int k = 0; k < 3; k++) {
            const double d = mu_m[k] - mu_r[k];
            dd += d * d;
            cross += d * (s[1 + k] - s[4 + k]);
          }
          loss_part = s[16] - 2.0 * cross + n * dd;
        }
        double *T = A.T + c * 12;
        double Am[9], R[9], t[3];
        for (int k = 0; k < 9; k++) Am[k] = cov[k] + A.angle_reg * T[k];  // :165
        kabsch_rotation_b(Am, R);
        for (int i = 0; i < 3; i++) t[i] = mu_r[i] - (R[i * 3 + 0] * mu_m[0] + R[i * 3 + 1] * mu_m[1] + R[i * 3 + 2] * mu_m[2]);
        double *Ti = A.Ti + c * 12;
        for (int k = 0; k < 9; k++) Ti[k] = R[k];
        for (int k = 0; k < 3; k++) Ti[9 + k] = t[k];
        double Rn[9], tn[3];
        for (int i = 0; i < 3; i++) {
          for (int q = 0; q < 3; q++) Rn[i * 3 + q] = R[i * 3 + 0] * T[q] + R[i * 3 + 1] * T[3 + q] + R[i * 3 + 2] * T[6 + q];
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
        A.l1_n[c] = n;
        for (int k = 0; k < kMomN; k++) s[k] = 0.0;
      }
      // loss per instance: aggregate the lanes of the same instance (components are grouped by instance)
      const unsigned int peers = __match_any_sync(kAll, j);
      double sum = loss_part;
      if (peers == kAll) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kAll, sum, o);
        if (lane == 0 && j >= 0 && sum != 0.0) atomicAdd(A.loss + j, sum);
      } else if (j >= 0 && loss_part != 0.0) {
        atomicAdd(A.loss + j, loss_part);
      }
    }
    pg_clear_used(A.mov, tid, nth);
    grid.sync();
    ICP_PROF(4)
    // ---- G: stopping rule per instance (:180-186) ---------------------------------------------------------------
    if (tid < A.J) {
      const int j = (int)tid;
      if (A.phase[j] == PH_FINISHING) A.phase[j] = PH_FROZEN;
      if (A.sec_fwd) A.dmax[j] = 0;  // consumed by phase E of this iteration
      if (A.phase[j] == PH_RUN) {
        const double loss = A.loss[j], last = A.last[j];
        int cd = A.cd[j];
        if (last - loss < A.stopping_delta) cd -= 1;
        else cd = 3;
        const int iters = A.iters[j] + 1;
        A.iters[j] = iters;
        A.cd[j] = cd;
        A.loss[j] = 0.0;
        if (cd <= 0 || iters >= A.max_iter) {
          A.phase[j] = PH_FINISHING;
          atomicAdd(A.itcnt + it * 2 + 0, 1);
        } else {
          A.last[j] = loss;
          atomicAdd(A.itcnt + it * 2 + 1, 1);
        }
      }
    }
    if (tid == 0) A.mov.ctr[0] = A.mov.ctr[1] = 0;
    grid.sync();
    ICP_PROF(5)
    if (A.prof && tid == 0) A.prof[8] += 1;
    const int n_fin = A.itcnt[it * 2 + 0], n_run = A.itcnt[it * 2 + 1];
    if (n_fin > 0) {
      if (A.want_l1) {
        // truncated mean residual of the instance's last iteration, positions before the update (:156, :44-58)
        for (int pass = 0; pass < 2; pass++) {
          for (long long e = tid; e < total; e += nth) {
            int mi, ri;
            if (!icp_edge_of(A, (int)e, nmv, mi, ri, PH_FINISHING)) continue;
            const float4 mp = A.mv[mi], rp = A.ref.sorted[ri];
            const int c = A.mv_gid[mi];
            const double *mu = A.mu + (long long)c * 6;
            const double dx = ((double)mp.y - mu[0]) - ((double)rp.y - mu[3]);
            const double dy = ((double)mp.z - mu[1]) - ((double)rp.z - mu[4]);
            const double dz = ((double)mp.w - mu[2]) - ((double)rp.w - mu[5]);
            double d = sqrt(dx * dx + dy * dy + dz * dz);
            if (pass == 1) {
              const double n = A.l1_n[c];
              const double mean = A.l1_sum[c * 2 + 0] / (n > 0.5 ? n : 1.0);
              d = fmin(fmax(d, mean - 0.3), mean + 0.3);
            }
            atomicAdd(A.l1_sum + (long long)c * 2 + pass, d);
          }
          grid.sync();
        }
        for (long long c = tid; c < A.G; c += nth) {
          if (A.phase[A.g_inst[c]] != PH_FINISHING) continue;
          const double n = A.l1_n[c];
          A.l1_err[c] = n > 0.5 ? A.l1_sum[c * 2 + 1] / n : 0.0;
        }
      }
      // the reference moves the points before it breaks (:179)
      for (long long v = tid; v < nmv; v += nth) {
        if (A.phase[A.mv_inst[v]] != PH_FINISHING) continue;
        float4 p = A.mv[v];
        apply_T(p, A.Ti + (long long)A.mv_gid[v] * 12);
        A.mv[v] = p;
      }
      ICP_PROF(6)
    }
    if (n_run == 0) break;
  }
  if (A.prof && tid == 0) A.prof[9] += 1;
  if (A.want_ratio) {
    grid.sync();
    // matched fraction: moving voxels with ANY reference voxel (stationary included) within the radius (:189-199).
    // The neighbour of the last iteration usually still is within the radius at the final position: no search then.
    for (long long base = warp_id * 32; base < nmv; base += nwarps * 32) {
      const int w = (int)base + lane;
      bool active = w < nmv;
      int j = 0;
      float4 mp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
        j = A.mv_inst[w];
        active = A.act[j] != 0;
        if (active) mp = A.mv[w];
      }
      bool matched = false;
      if (active) {
        const int prev = A.nn_fwd[w];
        if (prev >= 0) {
          const float4 c = A.ref.sorted[prev];
          matched = dist2_acc(A.acc0, c.y, c.z, c.w, mp.y, mp.z, mp.w) <= A.r2;
        }
      }
      unsigned int todo = __ballot_sync(kAll, active && !matched);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int sj = __shfl_sync(kAll, j, src);
        const float sx = __shfl_sync(kAll, mp.y, src), sy = __shfl_sync(kAll, mp.z, src), sz = __shfl_sync(kAll, mp.w, src);
        const int r = nn_search_rings(A.ref, false, A.ref_group_all[sj], sx, sy, sz, A.acc0, A.r2, 0u, lane, A.rings);
        if (lane == src) matched = r >= 0;
      }
      if (matched) atomicAdd(A.match_cnt + A.mv_gid[w], 1);
    }
    grid.sync();
    for (long long c = tid; c < A.G; c += nth) {
      if (!A.act[A.g_inst[c]]) continue;
      A.ratio[c] = (float)A.match_cnt[c] / ((float)A.vdeg[c] + 1e-6f);  // :199
    }
    ICP_PROF(7)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// tracker state machine (track_frame, cluster_tracking.py:430-787), batched over instances
// ---------------------------------------------------------------------------------------------------------------
struct Trk {
  int J, G, M, F;
  // instances
  const int *inst_anchor, *inst_key, *inst_C, *inst_fmin, *inst_fmax, *inst_has_valid, *inst_goff;
  // sequence, frame-sorted
  const float4 *seq_sorted;
  const int *frame_off;
  // moving points, grouped by component (hence by instance)
  float4 *mp;
  const float4 *mp0;
  float4 *m_last;
  const int *m_gid, *m_inst;
  // components
  const int *g_inst, *g_deg;
  const float *g_diam;
  const unsigned char *g_valid;
  unsigned char *g_stopped, *g_moving, *g_final;
  int *g_minf, *g_maxf;
  double *transforms;  // [G][17][12]
  float *velos, *velos_b, *centers, *diffs, *cv_pre, *g_delta, *adam_m, *adam_v;  // [G][17][3] / [G][3] / [G][8][2]
  double *csum, *vsum;  // [G][3]
  double *l1_err;
  float *ratio;
  double *T;  // [G][12] level transform (written by the ICP kernel)
  int *vdeg;  // [G] voxels per component of the level-0 sampling (consumed by the level-0 ICP)
  // per-step instance state
  int *cur_act, *cur_nxt, *cur_rel, *cur_haslv, *cur_grp, *cur_grp_all, *anyns /*[18][J]*/;
  int n_keys;
  unsigned int *sb;  // [J][6] bounds for the sampler
  // parameters
  float reg_error_coeff, angle_threshold;
  int min_move_frame;
  // extraction
  PGrid eg;
  int *eoff;      // [J + 1]
  const long long *exoff;  // [J * 17 + 1]
  int *ex;        // flat (instance, slot, row) -> component id or -1
  float nn_r2;
};

__device__ __forceinline__ void step_dir(int t, int &dir, int &s) {
  dir = t <= 8 ? -1 : 1;
  s = t <= 8 ? t : t - 8;
}

__global__ void __launch_bounds__(128) trk_step_begin_kernel(Trk A, int t) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= A.J) return;
  int dir, s;
  step_dir(t, dir, s);
  const int a = A.inst_anchor[j];
  const int nxt = a + dir * s;
  const bool inr = nxt >= A.inst_fmin[j] && nxt <= A.inst_fmax[j];
  int act;
  if (s == 1) act = A.inst_has_valid[j] && inr;
  else act = A.cur_act[j] && A.anyns[(t - 1) * A.J + j] && inr;
  A.cur_act[j] = act;
  A.cur_nxt[j] = nxt;
  const int fcl = nxt < 0 ? 0 : (nxt >= A.F ? A.F - 1 : nxt);
  A.cur_grp[j] = A.inst_key[j] * A.F + fcl;
  A.cur_grp_all[j] = A.n_keys * A.F + fcl;
  A.cur_rel[j] = kAnchorRel + dir * s;
  A.cur_haslv[j] = dir < 0 ? (s >= 2) : (s >= 2 || a > 0);
  if (j == 0) {
    A.eg.ctr[0] = A.eg.ctr[1] = 0;
  }
}

// start of a tracking direction (:542-551): the anchor points return to their original position
__global__ void __launch_bounds__(256) trk_dir_init_kernel(Trk A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.M) {
    const float4 p = A.mp0[i];
    A.mp[i] = p;
    A.m_last[i] = p;
  }
  if (i < A.G) {
    const unsigned char v = A.g_valid[i];
    A.g_stopped[i] = !v;
    A.g_moving[i] = v;
  }
}

// constant-velocity prediction (:566-573) ; produces the bounds for the level-0 sampler
__global__ void __launch_bounds__(256) trk_predict_kernel(Trk A, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int dir, s;
  step_dir(t, dir, s);
  if (i < A.G) {
    const int j = A.g_inst[i];
    if (A.cur_act[j]) {
      const int k = A.cur_rel[j];
      double *dst = A.transforms + ((long long)i * kRelFrames + k) * 12;
      const double *src = A.transforms + ((long long)i * kRelFrames + (k - dir)) * 12;
      for (int q = 0; q < 12; q++) dst[q] = src[q];
      if (A.cur_haslv[j] && !A.g_stopped[i]) {
        const float *lv = A.velos + ((long long)i * kRelFrames + (k - dir)) * 3;
        for (int q = 0; q < 3; q++) dst[9 + q] += (double)lv[q] * dir;
      }
    }
  }
  const bool inb = i < A.M;
  int j = 0;
  bool valid = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (inb) {
    j = A.m_inst[i];
    valid = A.cur_act[j] != 0;
    if (valid) {
      p = A.mp[i];
      const int g = A.m_gid[i];
      if (A.cur_haslv[j] && !A.g_stopped[g]) {
        const float *lv = A.velos + ((long long)g * kRelFrames + (A.cur_rel[j] - dir)) * 3;
        p.y += lv[0] * dir;
        p.z += lv[1] * dir;
        p.w += lv[2] * dir;
        A.mp[i] = p;
      }
    }
  }
  bounds_accumulate(A.sb, j, valid, p.y, p.z, p.w, threadIdx.x & 31);
}

// segmented warp sum of 3 doubles over runs of equal `g` (lanes with g < 0 do not contribute); the run heads add
// their totals to dst[g][0..2]
__device__ __forceinline__ void seg_add3(double *dst, int g, double x, double y, double z, int lane) {
  const int prev = __shfl_up_sync(kAll, g, 1);
  const bool head = lane == 0 || prev != g;
  const unsigned int heads = __ballot_sync(kAll, head);
  // run of this lane ends before the next head above it
  const unsigned int above = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));
  const int end = above ? (__ffs(above) - 1) : 32;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double tx = __shfl_down_sync(kAll, x, o), ty = __shfl_down_sync(kAll, y, o), tz = __shfl_down_sync(kAll, z, o);
    if (lane + o < end) {
      x += tx;
      y += ty;
      z += tz;
    }
  }
  if (head && g >= 0) {
    atomicAdd(dst + (long long)g * 3 + 0, x);
    atomicAdd(dst + (long long)g * 3 + 1, y);
    atomicAdd(dst + (long long)g * 3 + 2, z);
  }
}

// apply the level's transforms to the full-resolution anchor points and compose them into transforms[:, k]
// (:626-627); after the last level also the sums for the component centre / velocity (:629-633)
__global__ void __launch_bounds__(256) trk_apply_kernel(Trk A, int t, int last) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int dir, s;
  step_dir(t, dir, s);
  if (i < A.G) {
    const int j = A.g_inst[i];
    if (A.cur_act[j]) {
      const double *T = A.T + (long long)i * 12;
      double *X = A.transforms + ((long long)i * kRelFrames + A.cur_rel[j]) * 12;
      double Rn[9], tn[3];
      for (int r = 0; r < 3; r++) {
        for (int q = 0; q < 3; q++) Rn[r * 3 + q] = T[r * 3 + 0] * X[q] + T[r * 3 + 1] * X[3 + q] + T[r * 3 + 2] * X[6 + q];
        tn[r] = T[r * 3 + 0] * X[9] + T[r * 3 + 1] * X[10] + T[r * 3 + 2] * X[11] + T[9 + r];
      }
      for (int q = 0; q < 9; q++) X[q] = Rn[q];
      for (int q = 0; q < 3; q++) X[9 + q] = tn[q];
    }
  }
  const bool inb = i < A.M;
  int j = 0, g = -1;
  bool valid = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  double vx = 0, vy = 0, vz = 0;
  if (inb) {
    j = A.m_inst[i];
    valid = A.cur_act[j] != 0;
    if (valid) {
      p = A.mp[i];
      g = A.m_gid[i];
      const double *T = A.T + (long long)g * 12;
      const double x = p.y, y = p.z, z = p.w;
      // rotation in fp64, cast to fp32, translation added in fp32 (:626)
      p.y = (float)(T[0] * x + T[1] * y + T[2] * z) + (float)T[9];
      p.z = (float)(T[3] * x + T[4] * y + T[5] * z) + (float)T[10];
      p.w = (float)(T[6] * x + T[7] * y + T[8] * z) + (float)T[11];
      A.mp[i] = p;
      if (last) {
        const float4 l = A.m_last[i];
        vx = (double)((p.y - l.y) * dir);
        vy = (double)((p.z - l.z) * dir);
        vz = (double)((p.w - l.w) * dir);
      }
    }
  }
  if (last) {
    seg_add3(A.csum, g, (double)p.y, (double)p.z, (double)p.w, threadIdx.x & 31);
    seg_add3(A.vsum, g, vx, vy, vz, threadIdx.x & 31);
  } else {
    bounds_accumulate(A.sb, j, valid, p.y, p.z, p.w, threadIdx.x & 31);
  }
}

// component centre, raw velocity estimate, centre differences (:629-635)
__global__ void __launch_bounds__(256) trk_velocity_kernel(Trk A, int t) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= A.G) return;
  A.vdeg[g] = 0;
  const int j = A.g_inst[g];
  if (!A.cur_act[j]) return;
  int dir, s;
  step_dir(t, dir, s);
  const int k = A.cur_rel[j];
  const double deg = (double)A.g_deg[g];
  float *cen = A.centers + ((long long)g * kRelFrames + k) * 3;
  const float *cprev = A.centers + ((long long)g * kRelFrames + (k - dir)) * 3;
  float *vel = A.velos + ((long long)g * kRelFrames + k) * 3;
  float *dif = A.diffs + ((long long)g * kRelFrames + k) * 3;
  for (int q = 0; q < 3; q++) {
    const float c = (float)(A.csum[(long long)g * 3 + q] / deg);
    float v = (float)(A.vsum[(long long)g * 3 + q] / deg);
    if (q == 2) v = 0.f;
    cen[q] = c;
    vel[q] = v;
    A.cv_pre[(long long)g * 3 + q] = v;
    dif[q] = (c - cprev[q]) * dir;
    A.csum[(long long)g * 3 + q] = 0.0;
    A.vsum[(long long)g * 3 + q] = 0.0;
  }
}

// AdamW velocity smoothing (smooth_velo, :162-199): one thread-block cluster per instance.
// loss = w0 * mean (v - d)^2 + w * mean |v[f] - v[f+1]| over the xy velocities of ALL C components (the empty ones
// contribute zeros but count in the means) and the frames [a + dir, nxt]; lr 1e-2, MultiStepLR [100, 200, 300],
// 3-strike rule on the fp32 loss.  torch.optim.AdamW also decays every element without gradient.
constexpr int kSmoothCluster = 8;
constexpr int kSmoothThreads = 512;

__global__ void __cluster_dims__(kSmoothCluster, 1, 1) __launch_bounds__(kSmoothThreads)
trk_smooth_kernel(Trk A, int t, float w0, float w, int num_itr, float stopping) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_red[kSmoothThreads / 32];
  __shared__ double s_part[2];
  const int j = blockIdx.x / kSmoothCluster;
  const int crank = (int)cluster.block_rank();
  int dir, s;
  step_dir(t, dir, s);
  if (!A.cur_act[j] || s < 2) return;  // uniform over the cluster
  const int k = A.cur_rel[j];
  const int ra = dir < 0 ? k : kAnchorRel + 1, rb = dir < 0 ? kAnchorRel - 1 : k;
  const int nf = rb - ra + 1;
  const int g0 = A.inst_goff[j], ng = A.inst_goff[j + 1] - g0;
  const int S = ng * nf * 2;
  const float Cn = (float)A.inst_C[j];
  const float n1 = Cn * (float)nf * 2.f, n2 = Cn * (float)(nf - 1) * 2.f;
  const int cth = crank * kSmoothThreads + threadIdx.x, cnth = kSmoothCluster * kSmoothThreads;
  float lr = 1e-2f;
  const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, wd = 1e-2f;
  double b1t = 1.0, b2t = 1.0, decay = 1.0;
  double last_loss = 1e10;
  int countdown = 3;
  float *src = A.velos, *dst = A.velos_b;
  for (int e = cth; e < S; e += cnth) {
    A.adam_m[(long long)g0 * 16 + e] = 0.f;
    A.adam_v[(long long)g0 * 16 + e] = 0.f;
  }
#define VEL(buf, c, f, q) buf[((long long)(g0 + (c)) * kRelFrames + (f)) * 3 + (q)]
  int it = 0;
  for (; it < num_itr; it++) {
    double lsum = 0.0;
    b1t *= beta1;
    b2t *= beta2;
    const float bc1 = (float)(1.0 - b1t), bc2s = (float)sqrt(1.0 - b2t);
    for (int e = cth; e < S; e += cnth) {
      const int q = e & 1, fi = (e >> 1) % nf, c = (e >> 1) / nf, f = ra + fi;
      // the neighbours' values were written by other CTAs of the cluster in the previous iteration: read them
      // through L2 (ld.global.cg), never from a possibly stale L1 line
      const float vv = __ldcg(&VEL(src, c, f, q));
      const float r = vv - VEL(A.diffs, c, f, q);
      float grad = w0 * 2.f * r / n1;
      double l = (double)w0 * r * r / n1;
      if (f < rb) {
        const float d = vv - __ldcg(&VEL(src, c, f + 1, q));
        grad += w * ((d > 0.f) - (d < 0.f)) / n2;
        l += (double)w * fabsf(d) / n2;
      }
      if (f > ra) {
        const float d = __ldcg(&VEL(src, c, f - 1, q)) - vv;
        grad -= w * ((d > 0.f) - (d < 0.f)) / n2;
      }
      lsum += l;
      float p = vv * (1.f - lr * wd);
      const long long me = (long long)g0 * 16 + e;
      const float m = beta1 * A.adam_m[me] + (1.f - beta1) * grad;
      const float v = beta2 * A.adam_v[me] + (1.f - beta2) * grad * grad;
      A.adam_m[me] = m;
      A.adam_v[me] = v;
      p -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
      VEL(dst, c, f, q) = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(kAll, lsum, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
      for (int ww = 0; ww < kSmoothThreads / 32; ww++) tt += s_red[ww];
      s_part[it & 1] = tt;
    }
    cluster.sync();  // partial losses and the new velocities are visible to the whole cluster
    double total = 0.0;
    for (int r = 0; r < kSmoothCluster; r++) total += *cluster.map_shared_rank(&s_part[it & 1], r);
    float *tmp = src;
    src = dst;
    dst = tmp;
    decay *= (double)(1.f - lr * wd);
    if (it + 1 == 100 || it + 1 == 200 || it + 1 == 300) lr *= 0.1f;
    const double loss = (double)(float)total;
    if (last_loss - loss < (double)stopping) countdown -= 1;
    else countdown = 3;
    if (countdown <= 0) {
      ++it;
      break;
    }
    last_loss = loss;
  }
  cluster.sync();  // nobody reads a remote s_part after this point; all writes of the last iteration are visible
  // result -> velos ; every other element of the instance's velocity tensor only sees the weight decay
  const float dec = (float)decay;
  const int total_e = ng * kRelFrames * 3;
  for (int e = cth; e < total_e; e += cnth) {
    const int q = e % 3, f = (e / 3) % kRelFrames, c = e / (3 * kRelFrames);
    const bool optimised = q < 2 && f >= ra && f <= rb;
    if (optimised) {
      if (src != A.velos) VEL(A.velos, c, f, q) = __ldcg(&VEL(src, c, f, q));
    } else {
      VEL(A.velos, c, f, q) *= dec;
    }
  }
#undef VEL
}

__device__ __forceinline__ float dist_compensate_f(int deg) {
  if (deg < 10) return 1.f;
  if (deg < 40) return 0.5f;
  if (deg < 100) return 0.3f;
  if (deg < 200) return 0.2f;
  if (deg < 400) return 0.1f;
  return 0.f;
}

// smoothed-velocity correction, stopping tests, bookkeeping (:636-705)
__global__ void __launch_bounds__(256) trk_update_kernel(Trk A, int t) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= A.G) return;
  const int j = A.g_inst[g];
  if (!A.cur_act[j]) return;
  int dir, s;
  step_dir(t, dir, s);
  const int k = A.cur_rel[j];
  const float *vel = A.velos + ((long long)g * kRelFrames + k) * 3;
  const float *vprev = A.velos + ((long long)g * kRelFrames + (k - dir)) * 3;
  double *X = A.transforms + ((long long)g * kRelFrames + k) * 12;
  float cv[3];
  for (int q = 0; q < 3; q++) {
    const float d = vel[q] - A.cv_pre[(long long)g * 3 + q];
    A.g_delta[(long long)g * 3 + q] = d;
    X[9 + q] += (double)(d * dir);
    cv[q] = vel[q];
  }
  const float diam = A.g_diam[g];
  bool stopped = A.g_stopped[g] != 0;
  stopped |= A.l1_err[g] > (double)(A.reg_error_coeff * diam * (1.f + dist_compensate_f(A.g_deg[g])));
  stopped |= A.ratio[g] < 0.5f;
  if (s == A.min_move_frame) {
    const float *c1 = A.centers + ((long long)g * kRelFrames + k) * 3;
    const float *c0 = A.centers + ((long long)g * kRelFrames + kAnchorRel) * 3;
    const float dx = c1[0] - c0[0], dy = c1[1] - c0[1], dz = c1[2] - c0[2];
    const float travelled = sqrtf(dx * dx + dy * dy + dz * dz);
    A.g_moving[g] = A.g_moving[g] && (travelled > 0.08f * diam);
  }
  if (A.cur_haslv[j]) {
    const float dx = cv[0] - vprev[0], dy = cv[1] - vprev[1], dz = cv[2] - vprev[2];
    stopped |= sqrtf(dx * dx + dy * dy + dz * dz) > 0.24f * diam;
    const float na = sqrtf(cv[0] * cv[0] + cv[1] * cv[1] + cv[2] * cv[2]);
    const float nb = sqrtf(vprev[0] * vprev[0] + vprev[1] * vprev[1] + vprev[2] * vprev[2]);
    const float norm = fmaxf(na * nb, 1e-6f);
    float cosv = (cv[0] * vprev[0] + cv[1] * vprev[1] + cv[2] * vprev[2]) / norm;
    cosv = fminf(fmaxf(cosv, -1.f), 1.f);
    const float angle = acosf(cosv) / 3.14159265358979323846f * 180.0f;
    stopped |= (angle > A.angle_threshold) && (sqrtf(cv[0] * cv[0] + cv[1] * cv[1]) > 0.01f);
  }
  A.g_stopped[g] = stopped;
  if (dir < 0 && s == 1) {
    float *va = A.velos + ((long long)g * kRelFrames + kAnchorRel) * 3;
    for (int q = 0; q < 3; q++) va[q] = cv[q];
  }
  if (!stopped) {
    if (dir < 0) A.g_minf[g] = A.cur_nxt[j];
    else A.g_maxf[g] = A.cur_nxt[j];
    A.anyns[t * A.J + j] = 1;
  }
}

// shift the anchor points by the velocity correction (:639), remember them (:641), and count them into the grid of
// the extraction search
__global__ void __launch_bounds__(256) trk_shift_count_kernel(Trk A, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.M) return;
  const int j = A.m_inst[i];
  if (!A.cur_act[j]) return;
  int dir, s;
  step_dir(t, dir, s);
  const int g = A.m_gid[i];
  float4 p = A.mp[i];
  const float *d = A.g_delta + (long long)g * 3;
  p.y += d[0] * dir;
  p.z += d[1] * dir;
  p.w += d[2] * dir;
  A.mp[i] = p;
  A.m_last[i] = p;
  pg_count(A.eg, j, p.y, p.z, p.w);
}

__global__ void __launch_bounds__(256) trk_eg_ranges_kernel(Trk A) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pg_ranges(A.eg, tid, (long long)gridDim.x * blockDim.x);
  if (tid == 0) {
    int o = 0;
    for (int j = 0; j < A.J; j++) {
      A.eoff[j] = o;
      if (A.cur_act[j]) o += A.frame_off[A.cur_nxt[j] + 1] - A.frame_off[A.cur_nxt[j]];
    }
    A.eoff[A.J] = o;
  }
}

__global__ void __launch_bounds__(256) trk_eg_scatter_kernel(Trk A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.M) return;
  const int j = A.m_inst[i];
  if (!A.cur_act[j]) return;
  const float4 p = A.mp[i];
  pg_scatter(A.eg, j, p.y, p.z, p.w, 0u, i);
}

// every point of the target frame takes the component of its nearest moved anchor point (r = NN_GRAPH.RADIUS) unless
// that component has stopped (:710-721)
__global__ void __launch_bounds__(256) trk_extract_kernel(Trk A, int t) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int total = A.eoff[A.J];
  for (long long w = warp_id; w < total; w += nwarps) {
    const int j = seg_of_item(A.eoff, A.J, (int)w);
    const int i = (int)w - A.eoff[j];
    const float4 q = A.seq_sorted[A.frame_off[A.cur_nxt[j]] + i];
    const int mi = nn_search(A.eg, true, j, q.y, q.z, q.w, 0.f, A.nn_r2, 0u, lane);
    if (lane == 0) {
      int out = -1;
      if (mi >= 0) {
        const int g = A.m_gid[mi];
        if (!A.g_stopped[g]) out = g;
      }
      A.ex[A.exoff[(long long)j * kRelFrames + t] + i] = out;
    }
  }
}

// the same with one THREAD per query (most points of the target frame have no moved anchor point near them)
__global__ void __launch_bounds__(256) trk_extract_thread_kernel(Trk A, int t) {
  const int total = A.eoff[A.J];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int j = seg_of_item(A.eoff, A.J, (int)w);
    const int i = (int)w - A.eoff[j];
    const float4 q = A.seq_sorted[A.frame_off[A.cur_nxt[j]] + i];
    const unsigned long long k = nn_search_thread(A.eg, true, j, q.y, q.z, q.w, 0.f, A.nn_r2, 0u);
    int out = -1;
    if (k != ~0ull) {
      const int g = A.m_gid[(int)(unsigned int)(k & 0xffffffffu)];
      if (!A.g_stopped[g]) out = g;
    }
    A.ex[A.exoff[(long long)j * kRelFrames + t] + i] = out;
  }
}

__global__ void __launch_bounds__(256) trk_eg_clear_kernel(Trk A) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pg_clear_used(A.eg, tid, (long long)gridDim.x * blockDim.x);
}

// final component filter (:753-757) and the anchor-frame entries of the extraction table (:533-538)
__global__ void __launch_bounds__(256) trk_finish_kernel(Trk A, const int *__restrict__ m_frow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.G) {
    const int a = A.inst_anchor[A.g_inst[i]];
    A.g_final[i] = A.g_valid[i] && ((A.g_maxf[i] >= a + A.min_move_frame) || (A.g_minf[i] <= a - A.min_move_frame));
  }
  if (i < A.M) {
    const int j = A.m_inst[i], g = A.m_gid[i];
    A.ex[A.exoff[(long long)j * kRelFrames] + m_frow[i]] = A.g_valid[g] ? g : -1;
  }
}

}  // namespace pcs

using namespace pcs;

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
namespace {

PGrid make_pgrid(void *table, int64_t H, void *sorted, void *sidx, void *cells, void *ctr, const double *lo, double cs) {
  PGrid g;
  g.table = (pcs_slot_t *)table;
  g.mask = (unsigned int)(H - 1);
  g.sorted = (float4 *)sorted;
  g.sidx = (int *)sidx;
  g.cells = (int *)cells;
  g.ctr = (int *)ctr;
  g.tie = nullptr;
  g.lo0 = (float)lo[0];
  g.lo1 = (float)lo[1];
  g.lo2 = (float)lo[2];
  g.cs = (float)cs;
  g.inv_cs = 1.0f / (float)cs;
  return g;
}

SampArgs make_samp(const pcs_trk_sampler_t *S) {
  SampArgs A;
  A.pts = (const float4 *)S->pts;
  A.group = (const int *)S->group;
  A.skey = (const int *)S->skey;
  A.bits = (const unsigned char *)S->bits;
  A.act = (const int *)S->act;
  A.n = (int)S->n;
  A.n_groups = (int)S->n_groups;
  A.n_keys = (int)S->n_keys;
  A.ns_only = (int)S->ns_only;
  A.s1 = (float)S->size[0];
  A.s2 = (float)S->size[1];
  A.s3 = (float)S->size[2];
  A.sb = (unsigned int *)S->sb;
  A.table = (SampSlot *)S->table;
  A.mask = (unsigned int)(S->H - 1);
  A.vsum = (double *)S->vsum;
  A.vbits = (int *)S->vbits;
  A.vk = (int *)S->vk;
  A.pnext = (int *)S->pnext;
  A.vlist = (int *)S->vlist;
  A.ctr = (int *)S->ctr;
  A.vres = (int *)S->vres;
  A.kcount = (int *)S->kcount;
  A.koff = (int *)S->koff;
  A.kcur = (int *)S->kcur;
  A.vdeg = (int *)S->vdeg;
  A.out_pts = (float4 *)S->out_pts;
  A.out_key = (int *)S->out_key;
  A.out_group = (int *)S->out_group;
  return A;
}

IcpB make_icp(const pcs_trk_icp_t *P) {
  IcpB A;
  A.J = (int)P->J;
  A.G = (int)P->G;
  A.act = (const int *)P->act;
  A.ref_group = (const int *)P->ref_group;
  A.ref_group_all = (const int *)P->ref_group_all;
  A.skipmask = (const int *)P->skipmask;
  A.ref_off = (const int *)P->ref_off;
  A.g_inst = (const int *)P->g_inst;
  A.ref = make_pgrid(P->ref_table, P->ref_H, P->ref_pts, nullptr, nullptr, nullptr, P->lo, P->cs);
  A.mov = make_pgrid(P->mov_table, P->mov_H, P->mov_sorted, P->mov_sidx, P->mov_cells, P->mov_ctr, P->lo, P->cs);
  A.mov.tie = (const int *)P->mv_gid;
  A.mv = (float4 *)P->mv;
  A.mv_gid = (const int *)P->mv_gid;
  A.mv_inst = (const int *)P->mv_inst;
  A.n_mv = (const int *)P->n_mv;
  A.vdeg = (const int *)P->vdeg;
  const float r = (float)P->radius;
  A.r2 = r * r;
  A.acc0 = (float)((double)P->df * (double)P->df);
  A.rings = (int)P->rings < 2 ? 1 : 2;
  A.r3skip = sqrtf(fmaxf(A.r2 - A.acc0, 0.f)) * 1.0005f + 1e-5f;
  {
    const char *m = getenv("PCS_ICP_MODE");
    A.mode = m ? atoi(m) : 0;
    const char *mg = getenv("PCS_ICP_MARGIN");  // fraction of the cell size
    A.margin = (float)((mg ? atof(mg) : 0.0) * P->cs);
    const char *bq = getenv("PCS_ICP_BATCH");
    A.batch = bq ? atoi(bq) : 16;
    if (A.batch < 1 || A.batch > 32) A.batch = 32;
  }
  {
    const float cover = 0.98f * (float)P->cs;
    A.cover2 = cover * cover;
  }
  A.angle_reg = P->angle_reg;
  A.stopping_delta = P->stopping_delta;
  A.max_iter = (int)P->max_iter;
  A.want_l1 = (int)P->want_l1;
  A.want_ratio = (int)P->want_ratio;
  A.nn_fwd = (int *)P->nn_fwd;
  A.nn_bwd = (int *)P->nn_bwd;
  A.boff = (int *)P->boff;
  A.mvbeg = (int *)P->mvbeg;
  A.mvend = (int *)P->mvend;
  A.sec_fwd = (float *)P->sec_fwd;
  A.sec_bwd = (float *)P->sec_bwd;
  A.disp = (float *)P->disp;
  A.dmax = (int *)P->dmax;
  if (!A.sec_fwd || !A.sec_bwd || !A.disp || !A.dmax) A.sec_fwd = nullptr;  // all or none
  A.mom = (double *)P->mom;
  A.Ti = (double *)P->Ti;
  A.T = (double *)P->T;
  A.mu = (double *)P->mu;
  A.l1_sum = (double *)P->l1_sum;
  A.l1_n = (double *)P->l1_n;
  A.phase = (int *)P->phase;
  A.cd = (int *)P->cd;
  A.iters = (int *)P->iters;
  A.itcnt = (int *)P->itcnt;
  A.last = (double *)P->last;
  A.loss = (double *)P->loss;
  A.match_cnt = (int *)P->match_cnt;
  A.l1_err = (double *)P->l1_err;
  A.ratio = (float *)P->ratio;
  A.prof = (long long *)P->prof;
  return A;
}

Trk make_trk(const pcs_trk_ctx_t *C) {
  Trk A;
  A.J = (int)C->J;
  A.G = (int)C->G;
  A.M = (int)C->M;
  A.F = (int)C->F;
  A.inst_anchor = (const int *)C->inst_anchor;
  A.inst_key = (const int *)C->inst_key;
  A.inst_C = (const int *)C->inst_C;
  A.inst_fmin = (const int *)C->inst_fmin;
  A.inst_fmax = (const int *)C->inst_fmax;
  A.inst_has_valid = (const int *)C->inst_has_valid;
  A.inst_goff = (const int *)C->inst_goff;
  A.seq_sorted = (const float4 *)C->seq_sorted;
  A.frame_off = (const int *)C->frame_off;
  A.mp = (float4 *)C->mp;
  A.mp0 = (const float4 *)C->mp0;
  A.m_last = (float4 *)C->m_last;
  A.m_gid = (const int *)C->m_gid;
  A.m_inst = (const int *)C->m_inst;
  A.g_inst = (const int *)C->g_inst;
  A.g_deg = (const int *)C->g_deg;
  A.g_diam = (const float *)C->g_diam;
  A.g_valid = (const unsigned char *)C->g_valid;
  A.g_stopped = (unsigned char *)C->g_stopped;
  A.g_moving = (unsigned char *)C->g_moving;
  A.g_final = (unsigned char *)C->g_final;
  A.g_minf = (int *)C->g_minf;
  A.g_maxf = (int *)C->g_maxf;
  A.transforms = (double *)C->transforms;
  A.velos = (float *)C->velos;
  A.velos_b = (float *)C->velos_b;
  A.centers = (float *)C->centers;
  A.diffs = (float *)C->diffs;
  A.cv_pre = (float *)C->cv_pre;
  A.g_delta = (float *)C->g_delta;
  A.adam_m = (float *)C->adam_m;
  A.adam_v = (float *)C->adam_v;
  A.csum = (double *)C->csum;
  A.vsum = (double *)C->vsum;
  A.l1_err = (double *)C->l1_err;
  A.ratio = (float *)C->ratio;
  A.T = (double *)C->T;
  A.vdeg = (int *)C->vdeg;
  A.cur_act = (int *)C->cur_act;
  A.cur_nxt = (int *)C->cur_nxt;
  A.cur_rel = (int *)C->cur_rel;
  A.cur_haslv = (int *)C->cur_haslv;
  A.cur_grp = (int *)C->cur_grp;
  A.cur_grp_all = (int *)C->cur_grp_all;
  A.n_keys = (int)C->n_keys;
  A.anyns = (int *)C->anyns;
  A.sb = (unsigned int *)C->sb;
  A.reg_error_coeff = (float)C->reg_error_coeff;
  A.angle_threshold = (float)C->angle_threshold;
  A.min_move_frame = (int)C->min_move_frame;
  const double ecs = C->nn_radius * 1.001;
  A.eg = make_pgrid(C->eg_table, C->eg_H, C->eg_sorted, C->eg_sidx, C->eg_cells, C->eg_ctr, C->lo, ecs);
  A.eoff = (int *)C->eoff;
  A.exoff = (const long long *)C->exoff;
  A.ex = (int *)C->ex;
  const float nr = (float)C->nn_radius;
  A.nn_r2 = nr * nr;
  return A;
}

inline unsigned int blocks_for(long long n, int block) { return (unsigned int)((n > 0 ? n : 1) + block - 1) / block; }

int coop_grid(const void *kernel, int threads, long long want_threads, int max_per_sm = 2) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > max_per_sm) per_sm = max_per_sm;
  long long blocks = (want_threads + threads - 1) / threads;
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" {

int pcs_trk_cell_keys(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, const double *lo, double cs,
                      int64_t *keys) {
  if (n < 0 || n >= (1LL << 31) || !lo || cs <= 0.0 || (n > 0 && (!pts || !group || !keys)) || ((uintptr_t)pts & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_cell_keys: bad args");
  if (n == 0) return 0;
  PCS_LAUNCH(trk_cell_keys_kernel, blocks_for(n, 256), 256, 0, as_stream(s), (const float4 *)pts, group, (int)n,
             (float)lo[0], (float)lo[1], (float)lo[2], 1.0f / (float)cs, (long long *)keys);
  return 0;
}

int pcs_trk_grid_fill(pcs_stream_t s, pcs_slot_t *table, int64_t H, const int64_t *keys, const int32_t *starts,
                      const int32_t *counts, int64_t n, int32_t *err) {
  if (!table || H < 2 || (H & (H - 1)) || H > (1LL << 31) || n < 0 || n > H / 2 + 1 || !err ||
      (n > 0 && (!keys || !starts || !counts)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_grid_fill: bad args (H power of two >= 2 n)");
  cudaStream_t st = as_stream(s);
  PCS_LAUNCH(trk_table_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, (int *)nullptr);
  if (n > 0)
    PCS_LAUNCH(trk_grid_fill_kernel, blocks_for(n, 256), 256, 0, st, table, (unsigned int)(H - 1),
               (const long long *)keys, starts, counts, (int)n, err);
  return 0;
}

int pcs_trk_table_clear(pcs_stream_t s, pcs_slot_t *table, int64_t H, int32_t *ctr) {
  if (!table || H < 2 || (H & (H - 1))) return set_error(PCS_ERR_BAD_ARG, "pcs_trk_table_clear: bad args");
  PCS_LAUNCH(trk_table_clear_kernel, grid_for(H, 256, 8), 256, 0, as_stream(s), (int4 *)table, (long long)H, ctr);
  return 0;
}

int pcs_trk_bounds_reset(pcs_stream_t s, uint32_t *sb, int n_groups) {
  if (!sb || n_groups < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_trk_bounds_reset: bad args");
  PCS_LAUNCH(trk_bounds_reset_kernel, blocks_for(n_groups * 6, 256), 256, 0, as_stream(s), sb, n_groups);
  return 0;
}

int pcs_trk_group_bounds(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, uint32_t *sb) {
  if (n < 0 || n >= (1LL << 31) || !sb || (n > 0 && (!pts || !group)) || ((uintptr_t)pts & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_group_bounds: bad args");
  if (n == 0) return 0;
  PCS_LAUNCH(trk_group_bounds_kernel, blocks_for(n, 256), 256, 0, as_stream(s), (const float4 *)pts, group, (int)n, sb);
  return 0;
}

int pcs_trk_sampler_init(pcs_stream_t s, const pcs_trk_sampler_t *S) {
  if (!S || !S->table || S->H < 2 || (S->H & (S->H - 1)) || S->H > (1LL << 31))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_sampler_init: bad args");
  SampArgs A = make_samp(S);
  PCS_LAUNCH(trk_samp_init_kernel, grid_for(S->H, 256, 8), 256, 0, as_stream(s), A, (long long)S->H);
  return 0;
}

int pcs_trk_sample(pcs_stream_t s, const pcs_trk_sampler_t *S) {
  if (!S || !S->table || S->H < 2 || (S->H & (S->H - 1)) || S->n < 0 || S->n >= (1LL << 31) || S->n_groups < 1 ||
      S->n_groups >= 32768 || S->n_keys < 1 || !S->sb || !S->ctr || !S->kcount || !S->koff || !S->kcur ||
      (S->n > 0 && (!S->pts || !S->group || !S->vlist || !S->vres || !S->out_pts || !S->out_key || !S->out_group)) ||
      (S->skey && !S->pnext) || ((uintptr_t)S->pts & 15) || ((uintptr_t)S->out_pts & 15) || S->H < S->n)
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_sample: bad args");
  cudaStream_t st = as_stream(s);
  SampArgs A = make_samp(S);
  const long long cover = S->n > S->n_groups * 6 ? S->n : S->n_groups * 6;
  PCS_LAUNCH(trk_samp_reset_kernel, blocks_for(S->n_keys + 1 > 4 ? S->n_keys + 1 : 4, 256), 256, 0, st, A);
  if (S->n > 0) {
    PCS_LAUNCH(trk_samp_insert_kernel, blocks_for(S->n, 256), 256, 0, st, A);
    PCS_LAUNCH(trk_samp_finalize1_kernel, blocks_for(S->n, 256), 256, 0, st, A);
  }
  PCS_LAUNCH(trk_scan_kernel, 1, 1024, 0, st, (const int *)A.kcount, A.koff, A.n_keys, A.ctr + 1);
  PCS_LAUNCH(trk_samp_finalize2_kernel, blocks_for(cover, 256), 256, 0, st, A);
  return 0;
}

// optional CUDA-event timing of the ICP launches (bench.py's roofline leg)
static bool g_icp_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_icp_events;

static int launch_icp(cudaStream_t st, const pcs_trk_icp_t *P) {
  if (!P || P->J < 1 || P->G < 1 || P->max_iter < 1 || (P->ref_H & (P->ref_H - 1)) || (P->mov_H & (P->mov_H - 1)) ||
      P->ref_H < 2 || P->mov_H < 2 || !P->mv || !P->n_mv || !P->ref_pts || ((uintptr_t)P->mv & 15) ||
      ((uintptr_t)P->ref_pts & 15) || ((uintptr_t)P->mov_sorted & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_icp: bad args");
  IcpB A = make_icp(P);
  long long cover = P->G > P->J ? P->G : P->J;
  if ((P->max_iter + 2) * 3 > cover) cover = (P->max_iter + 2) * 3;
  if (P->J > kMaxInst) return set_error(PCS_ERR_BAD_ARG, "pcs_trk_icp: more than 1024 instances");
  PCS_LAUNCH(trk_icp_setup_kernel, blocks_for(cover, 256), 256, 0, st, A);
  PCS_LAUNCH(trk_icp_ranges_kernel, blocks_for(P->mv_cap > 0 ? P->mv_cap : 1, 256), 256, 0, st, A);
  static int occ = -1;
  if (occ < 0) {
    const char *o = getenv("PCS_ICP_OCC");
    occ = o ? atoi(o) : 4;
  }
  // CTA size: the same 32 warps per SM as 4 x 256, 2 x 512 or 1 x 1024 threads -- fewer, larger CTAs make the grid-wide
  // barrier (one atomic + one spinning thread per CTA) cheaper; PCS_ICP_BLOCK selects
  static int block = -1;
  if (block < 0) {
    const char *b = getenv("PCS_ICP_BLOCK");
    block = b ? atoi(b) : 1024;  // measured (198 frames): tracker 526 / 512 / 487 ms at 256 / 512 / 1024
    if (block != 512 && block != 256) block = 1024;
  }
  const void *kern = occ >= 4 ? (const void *)trk_icp_kernel<4> : (occ == 3 ? (const void *)trk_icp_kernel<3> : (const void *)trk_icp_kernel<2>);
  int per_sm = occ >= 4 ? 4 : (occ == 3 ? 3 : 2);
  if (block == 512) {
    kern = (const void *)trk_icp_kernel<2, 512>;
    per_sm = 2;
  } else if (block == 1024) {
    kern = (const void *)trk_icp_kernel<1, 1024>;
    per_sm = 1;
  }
  const int blocks = coop_grid(kern, block, 1LL << 40, per_sm);
  void *args[] = {&A};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_icp_timing) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3((unsigned)block), args, 0, st);
  if (g_icp_timing) {
    cudaEventRecord(ev1, st);
    g_icp_events.emplace_back(ev0, ev1);
  }
  g_launches++;
  if (e != cudaSuccess) return set_error((int)e, "trk_icp_kernel (cooperative launch)");
  return check_launch("trk_icp_kernel");
}

int pcs_trk_icp(pcs_stream_t s, const pcs_trk_icp_t *P) { return launch_icp(as_stream(s), P); }

void pcs_trk_icp_timing(int enable) {
  g_icp_timing = enable != 0;
  for (auto &p : g_icp_events) {
    cudaEventDestroy(p.first);
    cudaEventDestroy(p.second);
  }
  g_icp_events.clear();
}

/* Synchronises on the recorded events; returns the number of timed ICP launches and their summed duration. */
int pcs_trk_icp_elapsed(double *total_ms) {
  double t = 0.0;
  for (auto &p : g_icp_events) {
    cudaEventSynchronize(p.second);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.first, p.second);
    t += ms;
  }
  if (total_ms) *total_ms = t;
  return (int)g_icp_events.size();
}

int pcs_trk_dir_init(pcs_stream_t s, const pcs_trk_ctx_t *C) {
  if (!C || C->J < 1 || C->G < 1 || C->M < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_trk_dir_init: bad args");
  Trk A = make_trk(C);
  PCS_LAUNCH(trk_dir_init_kernel, blocks_for(A.M > A.G ? A.M : A.G, 256), 256, 0, as_stream(s), A);
  return 0;
}

// One tracking step t in [1, 16] (t <= 8: anchor -> anchor - t, t > 8: anchor -> anchor + (t - 8)) for all instances.
int pcs_trk_step(pcs_stream_t s, const pcs_trk_ctx_t *C, const pcs_trk_sampler_t *S, const pcs_trk_icp_t *levels,
                 int n_levels, int t) {
  if (!C || !S || !levels || n_levels < 1 || n_levels > 8 || t < 1 || t > 16 || C->J < 1 || C->G < 1 || C->M < 1)
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_step: bad args");
  cudaStream_t st = as_stream(s);
  Trk A = make_trk(C);
  int dir, sdist;
  dir = t <= 8 ? -1 : 1;
  sdist = t <= 8 ? t : t - 8;
  const long long mg = A.M > A.G ? A.M : A.G;
  if (t == 1 || t == 9) PCS_LAUNCH(trk_dir_init_kernel, blocks_for(mg, 256), 256, 0, st, A);
  PCS_LAUNCH(trk_step_begin_kernel, blocks_for(A.J, 128), 128, 0, st, A, t);
  PCS_LAUNCH(trk_predict_kernel, blocks_for(mg, 256), 256, 0, st, A, t);
  for (int lv = 0; lv < n_levels; lv++) {
    pcs_trk_sampler_t Sl = *S;
    Sl.size[0] = C->voxel_size[lv * 3 + 0];
    Sl.size[1] = C->voxel_size[lv * 3 + 1];
    Sl.size[2] = C->voxel_size[lv * 3 + 2];
    Sl.vdeg = lv == 0 ? S->vdeg : nullptr;
    int rc = pcs_trk_sample(s, &Sl);
    if (rc) return rc;
    pcs_trk_icp_t P = levels[lv];
    P.df = dir * sdist;
    P.radius = (double)(float)sqrt(C->radius[lv] * C->radius[lv] + (double)(sdist * sdist));  // :112
    P.want_ratio = lv == 0;
    P.want_l1 = lv == n_levels - 1;
    rc = launch_icp(st, &P);
    if (rc) return rc;
    PCS_LAUNCH(trk_apply_kernel, blocks_for(mg, 256), 256, 0, st, A, t, lv == n_levels - 1 ? 1 : 0);
  }
  PCS_LAUNCH(trk_velocity_kernel, blocks_for(A.G, 256), 256, 0, st, A, t);
  if (sdist >= 2) {
    PCS_LAUNCH(trk_smooth_kernel, A.J * kSmoothCluster, kSmoothThreads, 0, st, A, t, 1.0f, 10.0f, 300, 1e-3f);
  }
  PCS_LAUNCH(trk_update_kernel, blocks_for(A.G, 256), 256, 0, st, A, t);
  PCS_LAUNCH(trk_shift_count_kernel, blocks_for(A.M, 256), 256, 0, st, A, t);
  PCS_LAUNCH(trk_eg_ranges_kernel, grid_for(A.M, 256, 4), 256, 0, st, A);
  PCS_LAUNCH(trk_eg_scatter_kernel, blocks_for(A.M, 256), 256, 0, st, A);
  static int warp_mode = -1;
  if (warp_mode < 0) {
    const char *e = getenv("PCS_GSEARCH_WARP");
    warp_mode = (e && atoi(e)) ? 1 : 0;
  }
  if (warp_mode)
    PCS_LAUNCH(trk_extract_kernel, 148 * 8, 256, 0, st, A, t);
  else
    PCS_LAUNCH(trk_extract_thread_kernel, 148 * 8, 256, 0, st, A, t);
  PCS_LAUNCH(trk_eg_clear_kernel, grid_for(A.M, 256, 4), 256, 0, st, A);
  return 0;
}

int pcs_trk_finish(pcs_stream_t s, const pcs_trk_ctx_t *C, const int32_t *m_frow) {
  if (!C || !m_frow || C->J < 1 || C->G < 1 || C->M < 1) return set_error(PCS_ERR_BAD_ARG, "pcs_trk_finish: bad args");
  Trk A = make_trk(C);
  PCS_LAUNCH(trk_finish_kernel, blocks_for(A.M > A.G ? A.M : A.G, 256), 256, 0, as_stream(s), A, m_frow);
  return 0;
}

int pcs_trk_run(pcs_stream_t s, const pcs_trk_ctx_t *C, const pcs_trk_sampler_t *S, const pcs_trk_icp_t *levels,
                int n_levels, const int32_t *m_frow) {
  for (int t = 1; t <= 16; t++) {
    const int rc = pcs_trk_step(s, C, S, levels, n_levels, t);
    if (rc) return rc;
  }
  return pcs_trk_finish(s, C, m_frow);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// generic grouped nearest-neighbour query (extract_traces_and_update_boxes, cluster_tracking.py:356-358): a grid over
// rows tagged with a group id, queried by segments of a (frame-sorted) point array, one group per segment
// ---------------------------------------------------------------------------------------------------------------
namespace pcs {

__global__ void __launch_bounds__(256) trk_gcount_kernel(PGrid g, const float4 *__restrict__ pts,
                                                         const int *__restrict__ group, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  pg_count(g, group[i], p.y, p.z, p.w);
}

__global__ void __launch_bounds__(256) trk_granges_kernel(PGrid g) {
  pg_ranges(g, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) trk_gscatter_kernel(PGrid g, const float4 *__restrict__ pts,
                                                           const int *__restrict__ group, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  pg_scatter(g, group[i], p.y, p.z, p.w, 0u, i);
}

__global__ void __launch_bounds__(256) trk_gsearch_kernel(PGrid g, const float4 *__restrict__ queries,
                                                          const int *__restrict__ seg_qstart,
                                                          const int *__restrict__ seg_off,
                                                          const int *__restrict__ seg_group, int nseg, float r2,
                                                          int *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int total = seg_off[nseg];
  // 32 consecutive queries per warp (coalesced loads and stores, one segment lookup per lane), searched one by one
  for (long long base = warp_id * 32; base < total; base += nwarps * 32) {
    const long long w = base + lane;
    bool need = w < total;
    int grp = 0;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (need) {
      const int sg = seg_of_item(seg_off, nseg, (int)w);
      q = queries[seg_qstart[sg] + ((int)w - seg_off[sg])];
      grp = seg_group[sg];
    }
    int res = -1;
    unsigned int todo = __ballot_sync(kAll, need);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int sgrp = __shfl_sync(kAll, grp, src);
      const float sx = __shfl_sync(kAll, q.y, src), sy = __shfl_sync(kAll, q.z, src), sz = __shfl_sync(kAll, q.w, src);
      const int r = nn_search(g, true, sgrp, sx, sy, sz, 0.f, r2, 0u, lane);
      if (lane == src) res = r;
    }
    if (w < total) out[w] = res;
  }
}

// One THREAD per query: most queries of the trace extraction have no stored point anywhere near them (27 empty lookups,
// nothing to sweep), which a whole warp per query only makes 32 times more expensive.
__global__ void __launch_bounds__(256) trk_gsearch_thread_kernel(PGrid g, const float4 *__restrict__ queries,
                                                                 const int *__restrict__ seg_qstart,
                                                                 const int *__restrict__ seg_off,
                                                                 const int *__restrict__ seg_group, int nseg, float r2,
                                                                 int *__restrict__ out) {
  const int total = seg_off[nseg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int sg = seg_of_item(seg_off, nseg, (int)w);
    const float4 q = queries[seg_qstart[sg] + ((int)w - seg_off[sg])];
    const unsigned long long k = nn_search_thread(g, true, seg_group[sg], q.y, q.z, q.w, 0.f, r2, 0u);
    out[w] = k == ~0ull ? -1 : (int)(unsigned int)(k & 0xffffffffu);
  }
}

}  // namespace pcs

extern "C" {

/* Grid over n rows float4 (., x, y, z) tagged with group[n] (< 32768).  table: 16-byte slots [H] cleared by the
 * call; sorted float4[n], sidx int32[n], cells int32[n], ctr int32[4] scratch. */
int pcs_trk_group_grid(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, const double *lo, double cs,
                       pcs_slot_t *table, int64_t H, float *sorted, int32_t *sidx, int32_t *cells, int32_t *ctr) {
  if (n < 0 || n >= (1LL << 31) || !lo || cs <= 0.0 || !table || H < 2 || (H & (H - 1)) || H < n || !ctr ||
      (n > 0 && (!pts || !group || !sorted || !sidx || !cells)) || ((uintptr_t)pts & 15) || ((uintptr_t)sorted & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_group_grid: bad args");
  cudaStream_t st = as_stream(s);
  PGrid g = make_pgrid(table, H, sorted, sidx, cells, ctr, lo, cs);
  PCS_LAUNCH(trk_table_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, ctr);
  if (n == 0) return 0;
  PCS_LAUNCH(trk_gcount_kernel, blocks_for(n, 256), 256, 0, st, g, (const float4 *)pts, group, (int)n);
  PCS_LAUNCH(trk_granges_kernel, grid_for(n, 256, 4), 256, 0, st, g);
  PCS_LAUNCH(trk_gscatter_kernel, blocks_for(n, 256), 256, 0, st, g, (const float4 *)pts, group, (int)n);
  return 0;
}

/* Nearest grid row (index into the pts given to pcs_trk_group_grid, or -1) within `radius` (3-D, d2 <= r*r in the
 * reference's fp32 FMA order) for the queries of nseg segments: segment k covers queries[seg_qstart[k] ...] with
 * seg_off[k + 1] - seg_off[k] rows and searches group seg_group[k]; out int32[seg_off[nseg]]. */
int pcs_trk_group_nn(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted, const int32_t *sidx,
                     const double *lo, double cs, const float *queries, const int32_t *seg_qstart,
                     const int32_t *seg_off, const int32_t *seg_group, int nseg, float radius, int32_t *out) {
  if (!table || H < 2 || (H & (H - 1)) || !lo || cs <= 0.0 || nseg < 0 || ((uintptr_t)queries & 15) ||
      ((uintptr_t)sorted & 15) || (nseg > 0 && (!queries || !seg_qstart || !seg_off || !seg_group || !out)))
    return set_error(PCS_ERR_BAD_ARG, "pcs_trk_group_nn: bad args");
  if (nseg == 0) return 0;
  PGrid g = make_pgrid((void *)table, H, (void *)sorted, (void *)sidx, nullptr, nullptr, lo, cs);
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("PCS_GSEARCH_WARP");
    mode = (e && atoi(e)) ? 1 : 0;
  }
  if (mode == 1)
    PCS_LAUNCH(trk_gsearch_kernel, 148 * 8, 256, 0, as_stream(s), g, (const float4 *)queries, seg_qstart, seg_off,
               seg_group, nseg, radius * radius, out);
  else
    PCS_LAUNCH(trk_gsearch_thread_kernel, 148 * 8, 256, 0, as_stream(s), g, (const float4 *)queries, seg_qstart,
               seg_off, seg_group, nseg, radius * radius, out);
  return 0;
}

}  // extern "C"
