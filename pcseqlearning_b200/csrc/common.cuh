// common.cuh -- shared device helpers of libpcseq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcseq_b200.h"

#define PCS_EMPTY_KEY (-1LL)
#define PCS_SEG_SHIFT 48  // key' = (segment << 48) | key ; keys must stay below 2^48

namespace pcs {

extern thread_local char g_err[512];
extern long long g_launches;

int set_error(int code, const char *what);
int check_launch(const char *what);

static inline cudaStream_t as_stream(pcs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define PCS_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ::pcs::g_launches++;                                                   \
    int _e = ::pcs::check_launch(#kernel);                                 \
    if (_e) return _e;                                                     \
  } while (0)

// Geometry of up to PCS_MAX_SEGMENTS key segments, passed by value-pointer into kernels and staged
// in shared memory by the blocks that need it.
struct SegGeom {
  const float *lo;        // [n_seg][4]
  const long long *dims;  // [n_seg][4]
  float vs[4];
  int seg_div;
  int n_seg;
};

// ---- order-preserving float <-> uint32 (for atomic / redux min-max) ---------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  unsigned int b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}

// L1 prefetch of the line holding *p (no register, no dependency: hides the latency of a load issued later)
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- streaming 128-bit loads ----------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---- cell hash ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int hash_key(long long k) {
  unsigned long long x = (unsigned long long)k;
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (unsigned int)x;
}

__device__ __forceinline__ int point_segment(float frame, int seg_div, int n_seg) {
  int s = (int)frame / seg_div;
  return s < 0 ? 0 : (s >= n_seg ? n_seg - 1 : s);
}

// Reference-exact voxel coordinate of one component: rint((p - lo) / vs) + 1 in fp32 with IEEE
// division and round-half-even (graph_utils.py:174-175; torch.round == nearbyint).
__device__ __forceinline__ long long voxel_coord(float p, float lo, float vs) {
  float d = __fsub_rn(p, lo);
  float q = __fdiv_rn(d, vs);
  return (long long)rintf(q) + 1;
}

// map2key of the reference (torch_hash_kernel.cu:31-47): clamp each digit to [0, dims_i] (the upper
// clamp is dims_i, not dims_i - 1) and linearise row-major.
__device__ __forceinline__ long long map2key4(long long c0, long long c1, long long c2, long long c3,
                                              const long long *d) {
  c0 = c0 < 0 ? 0 : (c0 > d[0] ? d[0] : c0);
  c1 = c1 < 0 ? 0 : (c1 > d[1] ? d[1] : c1);
  c2 = c2 < 0 ? 0 : (c2 > d[2] ? d[2] : c2);
  c3 = c3 < 0 ? 0 : (c3 > d[3] ? d[3] : c3);
  return ((c0 * d[1] + c1) * d[2] + c2) * d[3] + c3;
}

// fp32 4-D squared distance in the reference's accumulation order (ref - query, one FMA per
// dimension, dimension 0 first; torch_hash_kernel.cu:364-368 compiled with -fmad=true).
__device__ __forceinline__ float dist2_ref(const float4 r, const float4 q) {
  float d = __fsub_rn(r.x, q.x);
  float acc = __fmaf_rn(d, d, 0.0f);
  d = __fsub_rn(r.y, q.y);
  acc = __fmaf_rn(d, d, acc);
  d = __fsub_rn(r.z, q.z);
  acc = __fmaf_rn(d, d, acc);
  d = __fsub_rn(r.w, q.w);
  acc = __fmaf_rn(d, d, acc);
  return acc;
}

// ---- lock-free union-find (hook the larger root under the smaller one) ------------------------------
__device__ __forceinline__ int uf_find(int *parent, int x) {
  volatile int *p = parent;
  while (true) {
    int px = p[x];
    if (px == x) return x;
    int gp = p[px];
    if (gp != px) p[x] = gp;  // path halving; racing writers only ever store ancestors
    x = px;
  }
}

__device__ __forceinline__ void uf_unite(int *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    // a > b : hook root a under b
    int old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}


// ---- symmetric 3x3 eigen-decomposition, cyclic Jacobi from the identity, everything in registers ----------------
// On return (d0, d1, d2) are the eigenvalues and v** the eigenvector matrix (columns), A = V diag(d) V^T.
struct Eig3 {
  double d0, d1, d2;
  double v00, v01, v02, v10, v11, v12, v20, v21, v22;
};

#define PCS_JACOBI_ROT(app, aqq, apq, arp, arq, v0p, v0q, v1p, v1q, v2p, v2q)          \
  if ((apq) != 0.0) {                                                                  \
    const double theta = ((aqq) - (app)) / (2.0 * (apq));                              \
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;                              \
    (app) -= t * (apq);                                                                \
    (aqq) += t * (apq);                                                                \
    (apq) = 0.0;                                                                       \
    { const double x = (arp), y = (arq); (arp) = c * x - sn * y; (arq) = sn * x + c * y; } \
    { const double x = (v0p), y = (v0q); (v0p) = c * x - sn * y; (v0q) = sn * x + c * y; } \
    { const double x = (v1p), y = (v1q); (v1p) = c * x - sn * y; (v1q) = sn * x + c * y; } \
    { const double x = (v2p), y = (v2q); (v2p) = c * x - sn * y; (v2q) = sn * x + c * y; } \
  }

__device__ __forceinline__ Eig3 jacobi_eig3(double a00, double a01, double a02, double a11, double a12, double a22) {
  Eig3 e;
  e.v00 = e.v11 = e.v22 = 1.0;
  e.v01 = e.v02 = e.v10 = e.v12 = e.v20 = e.v21 = 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double diag = fabs(a00) + fabs(a11) + fabs(a22);
    if (off == 0.0 || off <= 1e-18 * diag) break;
    PCS_JACOBI_ROT(a00, a11, a01, a02, a12, e.v00, e.v01, e.v10, e.v11, e.v20, e.v21)  // (p,q) = (0,1), r = 2
    PCS_JACOBI_ROT(a00, a22, a02, a01, a12, e.v00, e.v02, e.v10, e.v12, e.v20, e.v22)  // (0,2), r = 1
    PCS_JACOBI_ROT(a11, a22, a12, a01, a02, e.v01, e.v02, e.v11, e.v12, e.v21, e.v22)  // (1,2), r = 0
  }
  e.d0 = a00;
  e.d1 = a11;
  e.d2 = a22;
  return e;
}

__device__ __forceinline__ void load_geom(const SegGeom &g, float4 *s_lo, long long *s_dims) {
  for (int i = threadIdx.x; i < g.n_seg; i += blockDim.x) {
    s_lo[i] = make_float4(g.lo[i * 4 + 0], g.lo[i * 4 + 1], g.lo[i * 4 + 2], g.lo[i * 4 + 3]);
    s_dims[i * 4 + 0] = g.dims[i * 4 + 0];
    s_dims[i * 4 + 1] = g.dims[i * 4 + 1];
    s_dims[i * 4 + 2] = g.dims[i * 4 + 2];
    s_dims[i * 4 + 3] = g.dims[i * 4 + 3];
  }
}

__device__ __forceinline__ long long point_key(const float4 p, const SegGeom &g, const float4 *s_lo,
                                               const long long *s_dims, long long *c, bool *overflow) {
  int seg = point_segment(p.x, g.seg_div, g.n_seg);
  float4 lo = s_lo[seg];
  c[0] = voxel_coord(p.x, lo.x, g.vs[0]);
  c[1] = voxel_coord(p.y, lo.y, g.vs[1]);
  c[2] = voxel_coord(p.z, lo.z, g.vs[2]);
  c[3] = voxel_coord(p.w, lo.w, g.vs[3]);
  long long k = map2key4(c[0], c[1], c[2], c[3], s_dims + seg * 4);
  *overflow = (k >> PCS_SEG_SHIFT) != 0;  // key does not fit below the segment prefix (flagged by the caller)
  return k | ((long long)seg << PCS_SEG_SHIFT);
}

static inline SegGeom make_geom(const float *seg_lo, const int64_t *seg_dims, const float *vs, int seg_div, int n_seg) {
  SegGeom g;
  g.lo = seg_lo;
  g.dims = (const long long *)seg_dims;
  for (int i = 0; i < 4; i++) g.vs[i] = vs[i];
  g.seg_div = seg_div < 1 ? 1 : seg_div;
  g.n_seg = n_seg;
  return g;
}

static inline int grid_for(long long n, int block, int per_sm) {
  long long want = (n + block - 1) / block;
  long long cap = 148LL * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace pcs
