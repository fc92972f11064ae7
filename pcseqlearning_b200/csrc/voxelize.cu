// voxelize.cu -- grid voxelization: cell key, sorted-unique numbering, inverse map, per-voxel mean /
// max-index / upper-median (sm_100a).
//
// Replaces GridSampling3D.forward (pcdet/models/model_utils/grid_sampling.py:22-46), i.e.
//   torch_cluster.grid_cluster -> torch.unique(sorted, return_inverse) -> torch_scatter.scatter(mean),
// the pick-one subsample of SimpleReg.forward (pcdet/models/registration/simple_reg.py:119-124,
// scatter(arange, inv, 'max')) and the per-voxel majority / upper median of sample_frame
// (pcdet/models/registration/preprocessors/cluster_tracking.py:39-51, registration_utils.py:60-81).
//
// One pass over the points hashes the truncation-based cell key (open addressing, unique keys),
// records the slot of every point and accumulates count / fp64 sums / max index per slot.  Only the
// V unique keys are sorted (radix sort, CUB) to obtain torch.unique's ascending-key numbering; a last
// pass writes inv and the per-voxel outputs.  fp64 accumulation makes the means independent of the
// atomic order to well below one fp32 ulp (the reference's fp32 atomics are order dependent).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pcs {

struct VoxGeom {
  float start[4];
  float size[4];
  long long stride[4];
  int ignore_dim0;
};

// start = min (start[0] -= 0.5), end = max (end[0] += 0.5); nvox_d = trunc((end-start)/size)+1 (fp32);
// strides k_0 = 1, k_{d+1} = k_d * nvox_d   (grid_sampling.py:31-37 + torch_cluster grid_cluster)
__global__ void voxelize_params_kernel(const unsigned int *__restrict__ bounds, float s0, float s1, float s2, float s3,
                                       int ignore_dim0, float *__restrict__ start_out,
                                       long long *__restrict__ stride_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float size[4] = {s0, s1, s2, s3};
  long long k = 1;
  for (int d = 0; d < 4; d++) {
    float mn = ord2f(bounds[d]), mx = ord2f(bounds[4 + d]);
    if (bounds[d] == 0xffffffffu && bounds[4 + d] == 0u) mn = mx = 0.f;
    if (d == 0) {
      if (ignore_dim0) mn = mx = 0.f;
      mn = __fsub_rn(mn, 0.5f);
      mx = __fadd_rn(mx, 0.5f);
    }
    start_out[d] = mn;
    stride_out[d] = k;
    long long nv = (long long)__fdiv_rn(__fsub_rn(mx, mn), size[d]) + 1;
    k *= nv;
  }
  stride_out[4] = k;  // total number of cells of the bounding grid
}

struct VoxSlot {
  long long key;
  int id;   // dense voxel id in claim order (-1 until the claimer has initialised the voxel's rows)
  int pad;
};

__global__ void __launch_bounds__(256) vox_clear_kernel(int4 *__restrict__ table, long long H,
                                                        int *__restrict__ counters) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const int4 e = make_int4(-1, -1, -1, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += stride) table[i] = e;
  if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
}

__device__ __forceinline__ long long vox_key(float4 p, const float *start, const long long *stride, float s0,
                                             float s1, float s2, float s3) {
  long long c0 = (long long)__fdiv_rn(__fsub_rn(p.x, start[0]), s0);
  long long c1 = (long long)__fdiv_rn(__fsub_rn(p.y, start[1]), s1);
  long long c2 = (long long)__fdiv_rn(__fsub_rn(p.z, start[2]), s2);
  long long c3 = (long long)__fdiv_rn(__fsub_rn(p.w, start[3]), s3);
  return c0 * stride[0] + c1 * stride[1] + c2 * stride[2] + c3 * stride[3];
}

// insert + accumulate.  The hash table only maps a cell key to a DENSE voxel id (claim order); everything per voxel
// (count, fp64 sums, max index) lives in arrays indexed by that id, so the random traffic of this pass and of the
// finishing passes goes to V-row arrays (L2-sized for the sequences of this path) instead of the H-slot table.
// Lanes with equal keys are aggregated: one probe sequence per distinct voxel and warp, counts added once; sums /
// max index still go through per-lane atomics (fp64 add, int max).
__global__ void __launch_bounds__(256)
vox_insert_kernel(const float4 *__restrict__ pts, long long n, const float *__restrict__ start_d,
                  const long long *__restrict__ stride_d, float s0, float s1, float s2, float s3, int ignore_dim0,
                  VoxSlot *__restrict__ table, long long mask, int *__restrict__ pt_vid, double *__restrict__ sums,
                  int *__restrict__ maxidx, int *__restrict__ counts, long long *__restrict__ ukeys,
                  int *__restrict__ uids, int *__restrict__ counters) {
  __shared__ float start[4];
  __shared__ long long stride[4];
  if (threadIdx.x < 4) {
    start[threadIdx.x] = start_d[threadIdx.x];
    stride[threadIdx.x] = stride_d[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  long long gstride = (long long)gridDim.x * blockDim.x;
  long long nround = ((n + 31) / 32) * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gstride) {
    bool valid = i < n;
    long long key = -2 - lane;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      p = ldg_stream_f4(pts + i);
      if (ignore_dim0) p.x = 0.f;
      key = vox_key(p, start, stride, s0, s1, s2, s3);
    }
    unsigned int peers = __match_any_sync(0xffffffffu, key);
    int leader = __ffs(peers) - 1;
    int vid = -1;
    if (valid && lane == leader) {
      long long slot = hash_key(key) & mask;
      bool found = false;
      for (long long probes = 0; probes <= mask; ++probes) {
        long long cur = *((volatile long long *)&table[slot].key);
        if (cur == key) {
          found = true;
          break;
        }
        if (cur == PCS_EMPTY_KEY) {
          unsigned long long prev = atomicCAS((unsigned long long *)&table[slot].key,
                                              (unsigned long long)PCS_EMPTY_KEY, (unsigned long long)key);
          if (prev == (unsigned long long)PCS_EMPTY_KEY) {
            // claimer: take the next dense id, initialise the voxel's rows, then publish the id
            vid = atomicAdd(&counters[0], 1);
            counts[vid] = 0;
            if (sums) {
              const double2 z = make_double2(0.0, 0.0);
              reinterpret_cast<double2 *>(sums)[(long long)vid * 2] = z;
              reinterpret_cast<double2 *>(sums)[(long long)vid * 2 + 1] = z;
            }
            if (maxidx) maxidx[vid] = -1;
            ukeys[vid] = key;
            uids[vid] = vid;
            __threadfence();
            *((volatile int *)&table[slot].id) = vid;
            break;
          }
          if ((long long)prev == key) {
            found = true;
            break;
          }
        }
        slot = (slot + 1) & mask;
      }
      if (found) {
        while ((vid = *((volatile int *)&table[slot].id)) < 0) {
        }
        __threadfence();
      }
      if (vid < 0)
        atomicExch(&counters[2], PCS_ERR_TABLE_FULL);
      else
        atomicAdd(&counts[vid], __popc(peers));
    }
    vid = __shfl_sync(0xffffffffu, vid, leader);
    if (valid && vid >= 0) {
      pt_vid[i] = vid;
      if (sums) {
        double *s = sums + (long long)vid * 4;
        atomicAdd(s + 0, (double)p.x);
        atomicAdd(s + 1, (double)p.y);
        atomicAdd(s + 2, (double)p.z);
        atomicAdd(s + 3, (double)p.w);
      }
      if (maxidx) {
        // the highest index of the group: only the last peer lane needs to go to memory
        int top = 31 - __clz(peers);
        if (lane == top) atomicMax(maxidx + vid, (int)i);
      }
    }
  }
}

// r-th smallest key <-> dense id: rank_of[id] = r plus the per-voxel outputs in ascending-key order
__global__ void __launch_bounds__(256) vox_rank_kernel(const int *__restrict__ ids_sorted, long long V,
                                                       const double *__restrict__ sums,
                                                       const int *__restrict__ maxidx, const int *__restrict__ counts,
                                                       int *__restrict__ rank_of, float4 *__restrict__ sampled,
                                                       long long *__restrict__ maxidx_out,
                                                       int *__restrict__ counts_out) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= V) return;
  const int id = ids_sorted[r];
  rank_of[id] = (int)r;
  const int c = counts[id];
  if (counts_out) counts_out[r] = c;
  if (sampled && sums) {
    const double2 a = reinterpret_cast<const double2 *>(sums)[(long long)id * 2];
    const double2 b = reinterpret_cast<const double2 *>(sums)[(long long)id * 2 + 1];
    double inv = 1.0 / (double)(c > 0 ? c : 1);
    sampled[r] = make_float4((float)(a.x * inv), (float)(a.y * inv), (float)(b.x * inv), (float)(b.y * inv));
  }
  if (maxidx_out && maxidx) maxidx_out[r] = maxidx[id];
}

__global__ void __launch_bounds__(256) vox_inv_kernel(const int *__restrict__ rank_of, const int *__restrict__ pt_vid,
                                                      long long n, long long *__restrict__ inv) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = rank_of[pt_vid[i]];
}

// ---- per-group upper median (robust_median) --------------------------------------------------------
// values int64[n], group inv int64[n] in [0,V); out[g] = element of rank deg//2 of the group's sorted values
// (empty groups -> -1e10).  Groups here are voxels with a handful of points: rows are first
// counting-sorted by group (offsets from an exclusive scan of the group sizes), then one thread per
// group selects the rank by counting.
__global__ void __launch_bounds__(256) group_fill_kernel(const long long *__restrict__ inv, long long n,
                                                         const long long *__restrict__ offsets,
                                                         int *__restrict__ cursor, int *__restrict__ rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long g = inv[i];
  int pos = atomicAdd(cursor + g, 1);
  rows[offsets[g] + pos] = (int)i;
}

__global__ void __launch_bounds__(256) group_median_kernel(const long long *__restrict__ values,
                                                           const long long *__restrict__ offsets, long long V,
                                                           const int *__restrict__ rows, long long *__restrict__ out) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= V) return;
  long long b = offsets[g], e = offsets[g + 1];
  int deg = (int)(e - b);
  if (deg == 0) {
    out[g] = -10000000000LL;
    return;
  }
  int target = deg / 2;
  long long ans = 0;
  for (int a = 0; a < deg; a++) {
    long long va = values[rows[b + a]];
    int less = 0, eq = 0;
    for (int c = 0; c < deg; c++) {
      long long vc = values[rows[b + c]];
      less += vc < va;
      eq += vc == va;
    }
    if (less <= target && target < less + eq) {
      ans = va;
      break;
    }
  }
  out[g] = ans;
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_voxelize_params(pcs_stream_t s, const uint32_t *bounds, const float *size, int ignore_dim0, float *start,
                        int64_t *strides) {
  if (!bounds || !size || !start || !strides) return set_error(PCS_ERR_BAD_ARG, "pcs_voxelize_params: bad args");
  PCS_LAUNCH(voxelize_params_kernel, 1, 32, 0, as_stream(s), bounds, size[0], size[1], size[2], size[3], ignore_dim0,
             start, (long long *)strides);
  return 0;
}

int pcs_voxelize_insert(pcs_stream_t s, const float *pts, int64_t n, const float *start, const int64_t *strides,
                        const float *size, int ignore_dim0, void *table, int64_t H, int32_t *pt_vid, double *sums,
                        int32_t *maxidx, int32_t *counts, int64_t *ukeys, int32_t *uids, int32_t *counters) {
  if (!table || H < 2 || (H & (H - 1)) || n < 0 || n >= (1LL << 31) || ((uintptr_t)pts & 15) || !counters ||
      (n > 0 && (!pt_vid || !counts || !ukeys || !uids)) || ((uintptr_t)sums & 15))
    return set_error(PCS_ERR_BAD_ARG, "pcs_voxelize_insert: bad args");
  cudaStream_t st = as_stream(s);
  PCS_LAUNCH(vox_clear_kernel, grid_for(H, 256, 8), 256, 0, st, (int4 *)table, (long long)H, counters);
  if (n == 0) return 0;
  PCS_LAUNCH(vox_insert_kernel, grid_for(n, 256, 8), 256, 0, st, (const float4 *)pts, (long long)n, start,
             (const long long *)strides, size[0], size[1], size[2], size[3], ignore_dim0, (VoxSlot *)table,
             (long long)(H - 1), pt_vid, sums, maxidx, counts, (long long *)ukeys, uids, counters);
  return 0;
}

int64_t pcs_sort_pairs_tmp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const int *)nullptr, (int *)nullptr, (int)n);
  return (int64_t)bytes + 256;
}

int pcs_sort_pairs(pcs_stream_t s, const int64_t *keys_in, int64_t *keys_out, const int32_t *vals_in,
                   int32_t *vals_out, int64_t n, void *tmp, int64_t tmp_bytes) {
  if (n < 0 || n >= (1LL << 31) || tmp_bytes < pcs_sort_pairs_tmp_bytes(n))
    return set_error(PCS_ERR_BAD_ARG, "pcs_sort_pairs: bad args / tmp too small");
  if (n == 0) return 0;
  size_t bytes = (size_t)tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, (const unsigned long long *)keys_in,
                                                  (unsigned long long *)keys_out, vals_in, vals_out, (int)n, 0, 64,
                                                  as_stream(s));
  g_launches += 8;
  if (e != cudaSuccess) return set_error((int)e, "cub::DeviceRadixSort::SortPairs");
  return 0;
}

int pcs_voxelize_finish(pcs_stream_t s, const int32_t *ids_sorted, int64_t V, const int32_t *pt_vid, int64_t n,
                        const double *sums, const int32_t *maxidx, const int32_t *counts, int32_t *rank_of,
                        int64_t *inv, float *sampled, int64_t *maxidx_out, int32_t *counts_out) {
  if (V < 0 || n < 0 || ((uintptr_t)sampled & 15) || ((uintptr_t)sums & 15) ||
      (V > 0 && (!ids_sorted || !counts || !rank_of)) || (n > 0 && inv && !pt_vid))
    return set_error(PCS_ERR_BAD_ARG, "pcs_voxelize_finish: bad args");
  cudaStream_t st = as_stream(s);
  if (V > 0)
    PCS_LAUNCH(vox_rank_kernel, (unsigned)((V + 255) / 256), 256, 0, st, ids_sorted, (long long)V, sums, maxidx,
               counts, rank_of, (float4 *)sampled, (long long *)maxidx_out, counts_out);
  if (n > 0 && inv)
    PCS_LAUNCH(vox_inv_kernel, (unsigned)((n + 255) / 256), 256, 0, st, rank_of, pt_vid, (long long)n,
               (long long *)inv);
  return 0;
}

int pcs_group_median(pcs_stream_t s, const int64_t *values, const int64_t *inv, int64_t n, const int64_t *offsets,
                     int64_t V, int32_t *cursor_zeroed, int32_t *rows, int64_t *out) {
  if (n < 0 || V < 0 || (n > 0 && (!values || !inv || !offsets || !cursor_zeroed || !rows)) || (V > 0 && !out))
    return set_error(PCS_ERR_BAD_ARG, "pcs_group_median: bad args");
  cudaStream_t st = as_stream(s);
  if (n > 0)
    PCS_LAUNCH(group_fill_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (const long long *)inv, (long long)n,
               (const long long *)offsets, cursor_zeroed, rows);
  if (V > 0)
    PCS_LAUNCH(group_median_kernel, (unsigned)((V + 255) / 256), 256, 0, st, (const long long *)values,
               (const long long *)offsets, (long long)V, rows, (long long *)out);
  return 0;
}

}  // extern "C"
