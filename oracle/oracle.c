/*
 * oracle.c -- CPU restatement of the reference's torch_hash op and points_in_boxes_cpu.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under pcseqlearning_b200/ links or calls it.
 *
 * Every function follows the cited lines of /root/reference (read-only) and keeps the reference's
 * quirks (upper clamp == dims_i, probe-until-EMPTY walks, strict '<' insertion sort, '<=' radius
 * acceptance in the graph kernels vs '<' in points_in_radius).  The one thing a CPU cannot restate
 * is the race that decides which thread claims which slot: here points are inserted sequentially
 * in index order, which is one legal outcome of the reference's atomicCAS race.
 *
 * Floating point: the reference is compiled by nvcc with its default -fmad=true, so
 * `dist2 = dist2 + di*di` (torch_hash_kernel.cu:364-368) contracts to one fused multiply-add per
 * dimension.  We call fmaf() explicitly and build with -ffp-contract=off so nothing else fuses.
 *
 * Parity pin: checked against the reference op itself (oracle/_ref, built from the reference's own
 * sources) on the B200 box -- see tests/test_ref_op_gpu.py and tests/golden/ref_op_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t Key;
#define EMPTY ((Key)-1)
#define MAXD 8

/* torch_hash_kernel.cu:31-47 */
static inline Key map2key(const Key *c, const Key *dims, int D) {
  Key ans = 0;
  for (int i = 0; i < D; i++) {
    Key k = c[i];
    if (k >= dims[i]) k = dims[i];
    if (k < 0) k = 0;
    ans = ans * dims[i] + k;
  }
  return ans;
}

/* torch_hash_kernel.cu:49-51 (rp0 = 999269, rp1 = 999437) */
static inline Key hashkey(Key key, Key H) { return ((key % H) * 999269 + 999437) % H; }

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void oracle_map2key(const Key *coords, const Key *dims, int D, int64_t N, Key *out) {
  for (int64_t i = 0; i < N; i++) out[i] = map2key(coords + i * D, dims, D);
}

/* torch_hash_kernel.cu:54-91 -- sequential realisation of the CAS race (index order). */
void oracle_hash_insert(Key *keys, float *values, Key *rev, int64_t H, const Key *dims, int D,
                        const Key *ins_keys, const float *ins_vals, int64_t N) {
  for (int64_t t = 0; t < N; t++) {
    Key k = map2key(ins_keys + t * D, dims, D);
    Key h = hashkey(k, H);
    while (keys[h] != EMPTY) h = (h + 1) % H;
    keys[h] = k;
    for (int i = 0; i < D; i++) values[h * D + i] = ins_vals[t * D + i];
    rev[h] = t;
  }
}

static inline int num_comb(const int *qmin, const int *qmax, int D) {
  int n = 1;
  for (int i = 0; i < D; i++) n *= (qmax[i] - qmin[i] + 1);
  return n;
}

/* cell offset enumeration, torch_hash_kernel.cu:254-259: digit i = c mod range_i + qmin_i, dim 0 fastest */
static inline void offset_coords(const Key *q, const int *qmin, const int *qmax, int D, int c, Key *out) {
  int temp = c;
  for (int i = 0; i < D; i++) {
    int range = qmax[i] - qmin[i] + 1;
    out[i] = q[i] + temp % range + qmin[i];
    temp /= range;
  }
}

static inline float dist2_f32(const float *a, const float *b, int D) {
  float d2 = 0.0f;
  for (int i = 0; i < D; i++) {
    float di = a[i] - b[i];
    d2 = fmaf(di, di, d2); /* nvcc -fmad=true contraction of dist2 + di*di */
  }
  return d2;
}

/* torch_hash_kernel.cu:224-288 */
void oracle_radius_graph_count(const Key *keys, const float *values, int64_t H, const Key *dims, int D,
                               const Key *qkeys, const float *qvals, int64_t M, const int *qmin,
                               const int *qmax, const float *radius, int max_nbr, int *degree) {
  int nc = num_comb(qmin, qmax, D);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t t = 0; t < M; t++) {
    Key cc[MAXD];
    int n = 0;
    float r2 = radius[t] * radius[t];
    for (int c = 0; c < nc; c++) {
      offset_coords(qkeys + t * D, qmin, qmax, D, c, cc);
      Key qk = map2key(cc, dims, D);
      Key h = hashkey(qk, H);
      while (keys[h] != EMPTY) {
        if (keys[h] == qk) {
          float d2 = dist2_f32(values + h * D, qvals + t * D, D);
          if (d2 <= r2 && (max_nbr == -1 || n < max_nbr)) n++;
        }
        h = (h + 1) % H;
      }
    }
    degree[t] = n;
  }
}

/* torch_hash_kernel.cu:290-409; `cap` is the per-query degree (the reference passes the degree array
 * as max_num_neighbors, :552), `offset` the exclusive scan of it (:534-535). */
void oracle_radius_graph_fill(const Key *keys, const float *values, const Key *rev, int64_t H,
                              const Key *dims, int D, const Key *qkeys, const float *qvals, int64_t M,
                              const int *qmin, const int *qmax, const float *radius, const int *cap,
                              const int64_t *offset, int sort_by_dist, Key *edges, float *dists) {
  int nc = num_comb(qmin, qmax, D);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t t = 0; t < M; t++) {
    Key cc[MAXD];
    Key *e = edges + offset[t] * 2;
    float *dd = dists + offset[t];
    int n = 0;
    int maxn = cap[t];
    float r2 = radius[t] * radius[t];
    for (int c = 0; c < nc; c++) {
      offset_coords(qkeys + t * D, qmin, qmax, D, c, cc);
      Key qk = map2key(cc, dims, D);
      Key h = hashkey(qk, H);
      while (keys[h] != EMPTY) {
        if (keys[h] == qk) {
          float d2 = dist2_f32(values + h * D, qvals + t * D, D);
          if (d2 <= r2) {
            int nid = n;
            if (sort_by_dist) {
              while (nid > 0 && d2 < dd[nid - 1]) {
                if (nid < maxn) {
                  dd[nid] = dd[nid - 1];
                  e[nid * 2] = e[nid * 2 - 2];
                  e[nid * 2 + 1] = e[nid * 2 - 1];
                }
                nid--;
              }
            }
            if (nid < maxn) {
              e[nid * 2] = rev[h];
              e[nid * 2 + 1] = t;
              dd[nid] = d2;
              n++;
            }
            if (n > maxn) n = maxn;
          }
        }
        h = (h + 1) % H;
      }
    }
  }
}

/* torch_hash_kernel.cu:96-155: nearest over the queried cells, no radius test, strict '<' */
void oracle_correspondence(const Key *keys, const float *values, const Key *rev, int64_t H, const Key *dims,
                           int D, const Key *qkeys, const float *qvals, int64_t M, const int *qmin,
                           const int *qmax, Key *corres) {
  int nc = num_comb(qmin, qmax, D);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t t = 0; t < M; t++) {
    Key cc[MAXD];
    float best = 1e10f;
    corres[t] = -1;
    for (int c = 0; c < nc; c++) {
      offset_coords(qkeys + t * D, qmin, qmax, D, c, cc);
      Key qk = map2key(cc, dims, D);
      Key h = hashkey(qk, H);
      while (keys[h] != EMPTY) {
        if (keys[h] == qk) {
          float d2 = dist2_f32(values + h * D, qvals + t * D, D);
          if (d2 < best) {
            best = d2;
            corres[t] = rev[h];
          }
        }
        h = (h + 1) % H;
      }
    }
  }
}

/* torch_hash_kernel.cu:160-222: strict '<' */
void oracle_points_in_radius(const Key *keys, const float *values, const Key *rev, int64_t H, const Key *dims,
                             int D, const Key *qkeys, const float *qvals, int64_t M, const int *qmin,
                             const int *qmax, float radius, Key *visited) {
  int nc = num_comb(qmin, qmax, D);
  float r2 = radius * radius;
  for (int64_t t = 0; t < M; t++) {
    Key cc[MAXD];
    for (int c = 0; c < nc; c++) {
      offset_coords(qkeys + t * D, qmin, qmax, D, c, cc);
      Key qk = map2key(cc, dims, D);
      Key h = hashkey(qk, H);
      while (keys[h] != EMPTY) {
        if (keys[h] == qk) {
          float d2 = dist2_f32(values + h * D, qvals + t * D, D);
          if (d2 < r2) visited[rev[h]] = 1;
        }
        h = (h + 1) % H;
      }
    }
  }
}

/* pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168: z half-extent exact, then rotate into the
 * box frame and test |x| < dx/2 + MARGIN, |y| < dy/2 + MARGIN with MARGIN = 1e-2. */
static inline int pt_in_box(const float *pt, const float *box) {
  float x = pt[0], y = pt[1], z = pt[2];
  float cx = box[0], cy = box[1], cz = box[2];
  float dx = box[3], dy = box[4], dz = box[5], rz = box[6];
  if (fabsf(z - cz) > dz / 2.0) return 0;
  float cosa = cosf(-rz), sina = sinf(-rz);
  float sx = x - cx, sy = y - cy;
  float lx = sx * cosa + sy * (-sina);
  float ly = sx * sina + sy * cosa;
  const float MARGIN = 1e-2f; /* float constant promoted to double in the comparison, as in the reference */
  return (fabs(lx) < dx / 2.0 + MARGIN) & (fabs(ly) < dy / 2.0 + MARGIN);
}

void oracle_points_in_boxes(const float *boxes, int64_t B, const float *pts, int64_t N, int *out) {
  for (int64_t i = 0; i < B; i++)
    for (int64_t j = 0; j < N; j++) out[i * N + j] = pt_in_box(pts + j * 3, boxes + i * 7);
}

/* Union-find connected components with scipy's numbering (weak connectivity, labels ascend with the
 * smallest member index) -- the result scipy.sparse.csgraph.connected_components gives on the edge
 * list at graph_utils.py:51-52; used as the multi-core-free CPU timing point and cross-checked
 * against scipy itself in tests/test_oracle.py. */
static int64_t uf_find(int64_t *p, int64_t x) {
  while (p[x] != x) {
    p[x] = p[p[x]];
    x = p[x];
  }
  return x;
}

int64_t oracle_connected_components(const int64_t *e0, const int64_t *e1, int64_t E, int64_t N, int64_t *label) {
  int64_t *p = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
  for (int64_t i = 0; i < N; i++) p[i] = i;
  for (int64_t k = 0; k < E; k++) {
    int64_t a = uf_find(p, e0[k]), b = uf_find(p, e1[k]);
    if (a < b) p[b] = a;
    else if (b < a) p[a] = b;
  }
  int64_t nc = 0;
  for (int64_t i = 0; i < N; i++) {
    int64_t r = uf_find(p, i);
    if (r == i) label[i] = nc++;
    else label[i] = label[r]; /* r < i always: roots are set minima */
  }
  free(p);
  return nc;
}
