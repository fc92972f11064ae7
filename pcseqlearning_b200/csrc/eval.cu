// eval.cu -- GT evaluation on the device: point-in-rotated-box membership per frame (sm_100a).
//
// Replaces points_in_boxes_cpu (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168: O(boxes x points) on one
// CPU thread plus device<->host copies) and the per-component Python loops around it
// (preprocessors/cluster_proposal.py:90-114, 206-255; cluster_tracking.py:340-411).  One thread per point tests the
// boxes of the point's own frame only, records the first box that holds it (`bp_mask.argmax(0)`) and adds its
// memberships to per-(component, box) count tables (`bi_mask.sum(-1)` of assign_instances_to_boxes).
#include "common.cuh"

namespace pcs {

struct BoxRec {  // 48 bytes
  float cx, cy, cz, cosa, sina, pad;
  double hx, hy, hz;
};

// check_pt_in_box3d_cpu mixes float and double: cos / sin are evaluated in double and narrowed to float, the
// half extents and the 1 cm margin are compared in double (roiaware_pool3d.cpp:121-141)
__global__ void __launch_bounds__(256) box_prep_kernel(const float *__restrict__ boxes, int B, BoxRec *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float *b = boxes + (long long)i * 7;
  BoxRec r;
  r.cx = b[0];
  r.cy = b[1];
  r.cz = b[2];
  const float neg = -b[6];
  r.cosa = (float)cos((double)neg);
  r.sina = (float)sin((double)neg);
  r.pad = 0.f;
  const float margin = 1e-2f;
  r.hx = (double)b[3] / 2.0 + (double)margin;
  r.hy = (double)b[4] / 2.0 + (double)margin;
  r.hz = (double)b[5] / 2.0;
  out[i] = r;
}

__device__ __forceinline__ bool pt_in_box(const BoxRec &r, float x, float y, float z) {
  if ((double)fabsf(__fsub_rn(z, r.cz)) > r.hz) return false;
  const float sx = __fsub_rn(x, r.cx), sy = __fsub_rn(y, r.cy);
  // local_x = shift_x * cosa + shift_y * (-sina); local_y = shift_x * sina + shift_y * cosa  (gcc -O2, no FMA on x86-64)
  const float lx = __fadd_rn(__fmul_rn(sx, r.cosa), __fmul_rn(sy, -r.sina));
  const float ly = __fadd_rn(__fmul_rn(sx, r.sina), __fmul_rn(sy, r.cosa));
  return ((double)fabsf(lx) < r.hx) & ((double)fabsf(ly) < r.hy);
}

__global__ void __launch_bounds__(256)
points_in_boxes_kernel(const float4 *__restrict__ pts, const int *__restrict__ sel, long long n,
                       const BoxRec *__restrict__ boxes, const int *__restrict__ box_off, int F, int Bmax,
                       const long long *__restrict__ cid0, const long long *__restrict__ cid1,
                       const long long *__restrict__ cid2, int *cnt0, int *cnt1, int *cnt2, int *__restrict__ first_out,
                       int *err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[sel ? sel[i] : i];
  const int f = (int)rintf(p.x);
  int first = -1;
  if (f >= 0 && f < F) {
    const int b0 = box_off[f], b1 = box_off[f + 1];
    for (int b = b0; b < b1; b++) {
      if (!pt_in_box(boxes[b], p.y, p.z, p.w)) continue;
      const int local = b - b0;
      if (first < 0) first = local;
      if (local >= Bmax) {
        atomicExch(err, PCS_ERR_BAD_ARG);
        continue;
      }
      if (cid0) atomicAdd(cnt0 + cid0[i] * Bmax + local, 1);
      if (cid1) atomicAdd(cnt1 + cid1[i] * Bmax + local, 1);
      if (cid2) atomicAdd(cnt2 + cid2[i] * Bmax + local, 1);
    }
  }
  if (first_out) first_out[i] = first;
}

}  // namespace pcs

using namespace pcs;

extern "C" {

int pcs_box_prep(pcs_stream_t s, const float *boxes, int64_t B, void *recs) {
  if (B < 0 || (B > 0 && (!boxes || !recs)) || ((uintptr_t)recs & 7))
    return set_error(PCS_ERR_BAD_ARG, "pcs_box_prep: bad args");
  if (B == 0) return 0;
  PCS_LAUNCH(box_prep_kernel, (unsigned)((B + 255) / 256), 256, 0, as_stream(s), boxes, (int)B, (BoxRec *)recs);
  return 0;
}

int pcs_points_in_boxes(pcs_stream_t s, const float *pts, const int32_t *sel, int64_t n, const void *recs,
                        const int32_t *box_off, int F, int Bmax, const int64_t *cid0, const int64_t *cid1,
                        const int64_t *cid2, int32_t *cnt0, int32_t *cnt1, int32_t *cnt2, int32_t *first_out,
                        int32_t *err) {
  if (n < 0 || F < 1 || Bmax < 1 || !recs || !box_off || !err || (n > 0 && !pts) || ((uintptr_t)pts & 15) ||
      (cid0 && !cnt0) || (cid1 && !cnt1) || (cid2 && !cnt2))
    return set_error(PCS_ERR_BAD_ARG, "pcs_points_in_boxes: bad args");
  if (n == 0) return 0;
  PCS_LAUNCH(points_in_boxes_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(s), (const float4 *)pts, sel,
             (long long)n, (const BoxRec *)recs, box_off, F, Bmax, (const long long *)cid0, (const long long *)cid1,
             (const long long *)cid2, cnt0, cnt1, cnt2, first_out, err);
  return 0;
}

}  // extern "C"
