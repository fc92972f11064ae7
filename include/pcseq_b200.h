/*
 * pcseq_b200.h -- C ABI of libpcseq_b200.so, the sm_100a implementation of PCSeqLearning's
 * cluster-extraction / tracking hot path.
 *
 * Conventions
 *   - every entry point takes a CUDA stream (cudaStream_t passed as void*; NULL = legacy default
 *     stream), plain device pointers and sizes; nothing here depends on torch;
 *   - all buffers are caller-allocated device memory unless marked (host);
 *   - return value: 0 on success, otherwise a cudaError_t / negative pcs error code;
 *     pcs_last_error() gives the text.  No entry point synchronises the stream unless it says so;
 *   - points are rows of 4 float32 (frame, x, y, z) = one 16-byte aligned float4 per point, the
 *     `point_fxyz` layout of the reference (pcdet/models/registration/simple_reg.py:115-117).
 *
 * Each function names the reference interface it replaces (paths relative to the reference root).
 */
#ifndef PCSEQ_B200_H
#define PCSEQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_ERR_BAD_ARG (-2)
#define PCS_ERR_TABLE_FULL (-3)
#define PCS_ERR_KEY_RANGE (-4)
#define PCS_MAX_SEGMENTS 64
#define PCS_MAX_K 32

typedef void *pcs_stream_t;

/* One slot of the open-addressing cell table (16 bytes, read with one 128-bit load). */
typedef struct {
  int64_t key;   /* linearised cell key (with the segment prefix), -1 = empty */
  int32_t start; /* first row of the cell in the cell-sorted point array */
  int32_t count; /* number of points in the cell */
} pcs_slot_t;

int pcs_version(void);
const char *pcs_last_error(void);
/* Number of kernels launched by this library since load / since the last reset (for bench.py's gpu_launches). */
int64_t pcs_launch_count(void);
void pcs_reset_launch_count(void);

/* ---- grid geometry --------------------------------------------------------------------------
 * Replaces the torch ops of RadiusGraph.build_graph that derive the voxel grid
 * (pcdet/models/model_utils/graph_utils.py:169-176): min/max over ref U query, origin
 * lo = min - 2*vs, dims = rint((max + 2*vs - lo)/vs) + 3, coordinates rint((p - lo)/vs) + 1.
 *
 * A "segment" is a group of frames that the reference processes in one RadiusGraph call (one
 * 10-frame chunk in ClusterProposal.propose_cluster, cluster_proposal.py:63-67): segment id of a point
 * = min(int(frame) / seg_div, n_seg - 1).  Every segment has its own origin/dims exactly as if the
 * reference had been called on that chunk alone; all segments share one table (key prefix).
 * bounds: uint32[n_seg][8] order-preserving encodings of (min f,x,y,z, max f,x,y,z). */
int pcs_bounds_init(pcs_stream_t s, uint32_t *bounds, int n_seg);
int pcs_bounds_update(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, uint32_t *bounds);
/* vs (host): voxel size per dimension.  Writes seg_lo f32[n_seg][4], seg_dims i64[n_seg][4].
 * pad = 0 reproduces the reference geometry bit for bit; pad > 0 adds that many cells of margin on every side
 * (used by pcs_register_icp, whose moving points leave their initial bounding box). */
int pcs_grid_params(pcs_stream_t s, const uint32_t *bounds, int n_seg, const float *vs, int pad, float *seg_lo,
                    int64_t *seg_dims);
/* Reference-exact voxel coordinates / linear keys of points (graph_utils.py:174-175 +
 * torch_hash_kernel.cu:31-47 map2key); either output may be NULL.  Used by parity tests and by the
 * torch_hash-compatible API. */
int pcs_voxel_keys(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                   const int64_t *seg_dims, const float *vs, int64_t *coords /*[n][4]*/, int64_t *keys /*[n]*/);

/* ---- voxel hash build -----------------------------------------------------------------------
 * Replaces hash_insert_gpu (pcdet/ops/torch_hash/src/torch_hash_kernel.cu:54-91, 411-442) together
 * with the key computation above.  Instead of a multimap with one slot per point, unique cell keys
 * are hashed and the points are counting-sorted by cell:
 *   table[H]        open addressing over unique cell keys (H = power of two)
 *   sorted_pts[n]   float4 points grouped by cell, sorted_idx[n] their original row index
 *   counters int32[4]: [0] number of occupied cells, [1] scatter cursor, [2] error flag, [3] unused
 *   occ        optional uint32[occ_bits / 32] occupancy bitmap (occ_bits = power of two in [32, 2^32]): one bit per
 *              occupied cell at position hash(key) >> (32 - log2 occ_bits).  A search given the same bitmap rejects
 *              most empty neighbour cells with one 4-byte load instead of a probe sequence (no false negatives).
 * The table and bitmap must NOT be pre-filled; the call clears them. */
int pcs_hash_build(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                   const int64_t *seg_dims, const float *vs, pcs_slot_t *table, int64_t H, float *sorted_pts,
                   int32_t *sorted_idx, int32_t *counters, uint32_t *occ, int64_t occ_bits);

/* Same result contract as pcs_hash_build, but the row ranges of the occupied cells are assigned in ascending KEY
 * order (frame-major, z fastest) instead of table-slot (= hash) order: the cells of a frame are contiguous in
 * sorted_pts and self-queries walk the grid coherently.  Costs a radix sort of the <= min(n, H) occupied cells
 * (no host sync: unused entries are padded with INT64_MAX).  ws: 16-byte aligned scratch of
 * pcs_hash_build_sorted_ws_bytes(n, H) bytes.  Staged for the next round: the default path does not use it yet. */
int64_t pcs_hash_build_sorted_ws_bytes(int64_t n, int64_t H);
int pcs_hash_build_sorted(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                          const int64_t *seg_dims, const float *vs, pcs_slot_t *table, int64_t H, float *sorted_pts,
                          int32_t *sorted_idx, int32_t *counters, uint32_t *occ, int64_t occ_bits, void *ws,
                          int64_t ws_bytes);

/* ---- neighbour search -----------------------------------------------------------------------
 * Replaces radius_graph_gpu = count_radius_graph_degree_kernel + radius_graph_kernel
 * (torch_hash_kernel.cu:224-409, 487-561) in ONE pass: for every query the cells
 * [c + qmin, c + qmax] are looked up, candidates are filtered by the fp32 4-D distance test
 * d2 <= r*r (same FMA order as the reference) and the K nearest are kept.
 *   queries    float4[m]; NULL = self-query mode: the grid's own points are the queries (m == n), read
 *              in cell order straight from sorted_pts / sorted_idx (outputs indexed by original row)
 *   order      optional int32[m]: the i-th work item processes query order[i] (cell-coherent order)
 *   radius     optional float[m] per-query radius; NULL -> radius_scalar
 *   K          1..32 neighbours kept (the K smallest (d2, ref index) pairs; ties by ascending index)
 *   nbr_idx    int32[m][K] reference row indices, ascending (d2, index); may be NULL when uf_parent set
 *   nbr_d2     float[m][K] optional
 *   nbr_cnt    int32[m] = min(#accepted, K)
 *   uf_parents (host) array of n_uf <= 3 device forests int32[n_ref == m]; forest k receives every (query,
 *              neighbour) pair with d2 <= uf_r2[k], united in-kernel (fused connected components, no edge list
 *              materialised); with uf_need_full[k] set it is only fed by queries whose list is full (count == K):
 *              for those the K nearest within this radius are also the K nearest within any larger radius, which
 *              is how ONE fine search serves several proposal radii (multi-radius search).
 *   skip_full_cnt optional int32[m]: counts of a previous (finer) pass; queries with count >= K are skipped.
 *   occ, occ_bits optional occupancy bitmap written by pcs_hash_build for this table (NULL / 0 = not used). */
int pcs_radius_search(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                      const int32_t *sorted_idx, int seg_div, int n_seg, const float *seg_lo,
                      const int64_t *seg_dims, const float *vs, const float *queries, int64_t m,
                      const int32_t *order, const int *qmin, const int *qmax, const float *radius,
                      float radius_scalar, int K, int32_t *nbr_idx, float *nbr_d2, int32_t *nbr_cnt,
                      int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full, int n_uf,
                      const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits);

/* Self-query search with fused connected components, ONE THREAD per query, for the cluster-proposal passes
 * (cluster_proposal.py:63-81 -> graph_utils.py:149-209 + :40-53; torch_hash_kernel.cu:224-409): every grid row is a
 * query (taken in cell order), no lists are written, only nbr_cnt int32[n] = min(#accepted, K) and the unions of the
 * query with its K nearest rows into the forests (same uf_* / skip_full_cnt semantics and the same candidate set,
 * distance test and tie rule as pcs_radius_search, whose self-query + union-find mode it replaces).  Scalar radius. */
int pcs_self_search_uf(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                       const int32_t *sorted_idx, int64_t n, int seg_div, int n_seg, const float *seg_lo,
                       const int64_t *seg_dims, const float *vs, const int *qmin, const int *qmax, float radius, int K,
                       int32_t *nbr_cnt, int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full,
                       int n_uf, const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits);

/* Exclusive scan int32 -> int64 with the grand total at out[n] (out has n+1 entries).
 * Replaces `cumsum(degree) - degree` (torch_hash_kernel.cu:534-538) without the two blocking .item(). */
int pcs_exclusive_scan(pcs_stream_t s, const int32_t *in, int64_t n, int64_t *out, void *tmp, int64_t tmp_bytes);
int64_t pcs_exclusive_scan_tmp_bytes(int64_t n);

/* Padded neighbour lists -> reference edge layout int64[E][2] rows (ref_idx, query_idx), grouped by
 * ascending query (torch_hash_kernel.cu:395-399, 540-541).  dists float[E] optional. */
int pcs_lists_to_edges(pcs_stream_t s, const int32_t *nbr_idx, const float *nbr_d2, const int32_t *nbr_cnt,
                       const int64_t *offsets, int64_t m, int K, int64_t *edges, float *dists);

/* ---- connected components -------------------------------------------------------------------
 * Replaces graph_utils.connected_components (graph_utils.py:40-53: device->host copy + single-thread
 * scipy.sparse.csgraph.connected_components) with a lock-free union-find on the device. */
int pcs_uf_init(pcs_stream_t s, int32_t *parent, int64_t n);
int pcs_uf_union_edges(pcs_stream_t s, int32_t *parent, const int64_t *e0, const int64_t *e1, int64_t E);
/* Flatten + canonical numbering.  labels int64[n]: rank of the component's smallest member index
 * among the components of the same segment (segment of node i = seg_of[i], NULL = one segment) plus
 * the number of components in all earlier segments -- scipy's numbering applied chunk by chunk
 * with a running offset (cluster_proposal.py:75-81).  n_comp int64[n_seg] per-segment counts.
 * tmp: pcs_uf_labels_tmp_bytes(n, n_seg) bytes. */
int pcs_uf_labels(pcs_stream_t s, int32_t *parent, int64_t n, const int32_t *seg_of, int n_seg, int64_t *labels,
                  int64_t *n_comp, void *tmp, int64_t tmp_bytes);
int64_t pcs_uf_labels_tmp_bytes(int64_t n, int n_seg);
/* int32 segment id per point from the frame column (seg = min(int(frame)/seg_div, n_seg-1)). */
int pcs_point_segments(pcs_stream_t s, const float *pts, int64_t n, int seg_div, int n_seg, int32_t *seg_of);

/* ---- voxelization ----------------------------------------------------------------------------
 * Replaces GridSampling3D.forward (pcdet/models/model_utils/grid_sampling.py:22-46):
 * torch_cluster.grid_cluster (truncation-based cell key, frame digit fastest) -> torch.unique(sorted,
 * return_inverse) -> torch_scatter.scatter(mean); plus scatter(arange, inv, 'max') of
 * simple_reg.py:122-124 and robust_median of registration_utils.py:60-81.
 *   bounds   uint32[8] from pcs_bounds_* with n_seg = 1
 *   size     (host) float[4] = [1, gx, gy, gz]
 *   start    float[4] (device), strides int64[5] (device; [4] = cells of the bounding grid)
 *   table    16-byte slots [H] (H power of two >= n, cleared by the call): cell key -> dense voxel id
 *   pt_vid   int32[n] dense voxel id (claim order) of every point
 *   sums     optional double[n][4], maxidx optional int32[n], counts int32[n]: rows indexed by dense id; only
 *            the first V rows are touched (the claimer of an id initialises them, no pre-clearing needed)
 *   ukeys / uids  int64[n] / int32[n]: unique keys and their dense ids in claim order (uids[i] = i)
 *   counters int32[4]: [0] = V (number of voxels), [2] = error flag
 * pcs_sort_pairs sorts the V unique (key, id) pairs by key (radix sort);
 * pcs_voxelize_finish numbers voxels by ascending key (rank_of int32[V] scratch: dense id -> rank) and writes
 * inv int64[n], sampled float4[V] (mean of all four columns), maxidx_out int64[V], counts_out int32[V]
 * (any of the outputs may be NULL). */
int pcs_voxelize_params(pcs_stream_t s, const uint32_t *bounds, const float *size, int ignore_dim0, float *start,
                        int64_t *strides);
int pcs_voxelize_insert(pcs_stream_t s, const float *pts, int64_t n, const float *start, const int64_t *strides,
                        const float *size, int ignore_dim0, void *table, int64_t H, int32_t *pt_vid, double *sums,
                        int32_t *maxidx, int32_t *counts, int64_t *ukeys, int32_t *uids, int32_t *counters);
int64_t pcs_sort_pairs_tmp_bytes(int64_t n);
int pcs_sort_pairs(pcs_stream_t s, const int64_t *keys_in, int64_t *keys_out, const int32_t *vals_in,
                   int32_t *vals_out, int64_t n, void *tmp, int64_t tmp_bytes);
int pcs_voxelize_finish(pcs_stream_t s, const int32_t *ids_sorted, int64_t V, const int32_t *pt_vid, int64_t n,
                        const double *sums, const int32_t *maxidx, const int32_t *counts, int32_t *rank_of,
                        int64_t *inv, float *sampled, int64_t *maxidx_out, int32_t *counts_out);
/* Upper median per group (rank deg/2 of the sorted values; empty groups -> -1e10).  offsets int64[V+1]
 * = exclusive scan of the group sizes, cursor int32[V] zero-filled scratch, rows int32[n] scratch. */
int pcs_group_median(pcs_stream_t s, const int64_t *values, const int64_t *inv, int64_t n, const int64_t *offsets,
                     int64_t V, int32_t *cursor_zeroed, int32_t *rows, int64_t *out);

/* ---- ground-stage solvers ---------------------------------------------------------------------
 * pcs_ground_ransac replaces iterative_reweighted_ransac and the 30-ratio loop around it
 * (pcdet/models/registration/preprocessors/preprocessor_utils.py:32-80, 147-170) with one cooperative
 * persistent launch: per super-pillar IRLS plane fits (fp64 moment accumulation, Jacobi 3x3 eigen solve),
 * the global max|dw| < stopping_delta rule and the best-plane bookkeeping all run on the device.
 *   vox float4[Nv] (unused,x,y,z) sorted by super-pillar id cidx int32[Nv]; seg_start int32[C+1]
 *   origin float[C][3] local origins; cmin_z / cmax_z float[C]; ratios float[n_ratios]
 *   scratch (zero-filled by the caller): acc double[3][R][C][10], nhit int32[3][R][C], gmax uint32[3][R],
 *   planes float[2][R][C][6], fin int32[R][2]   (R = n_ratios <= 32; all ratios are iterated concurrently)
 *   outputs: best_center float[C][3] (init 0), best_normal float[C][3] (init (0,0,1)), best_conf float[C]
 *   (init 0), iters_out int32[n_ratios] (optional).
 * pcs_l1_heightfield replaces l1_minimization (preprocessor_utils.py:313-350): AdamW (torch defaults,
 * MultiStepLR milestone decay_step, factor lr_gamma) on the X x Y pillar height grid with the reference's
 * 3-strike stopping rule, up to max_iters iterations in one single-CTA launch.  h is in/out (start point),
 * m / v zero-filled scratch, info int32[2] = (iterations run, stopped early), loss_out float[1]. */
int pcs_ground_ransac(pcs_stream_t s, const float *vox, const int32_t *cidx, const int32_t *seg_start,
                      const float *origin, const float *cmin_z, const float *cmax_z, const float *ratios, int64_t Nv,
                      int C, int n_ratios, float sigma2, float stopping_delta, int max_iter, double *acc,
                      int32_t *nhit, uint32_t *gmax, float *planes, int32_t *fin, float *best_center,
                      float *best_normal, float *best_conf, int32_t *iters_out);
int pcs_l1_heightfield(pcs_stream_t s, const float *min_z, const float *weight, float *h, float *m, float *v, int X,
                       int Y, float lr, float lr_gamma, int decay_step, float rigid_weight, int max_iters,
                       int32_t *info, float *loss_out);

/* Curvature pruning of plane centres ("Truncated Least Squares", preprocessor_utils.py:175-193): for every threshold
 * (descending), kNN (self included) mean curvature of the surviving planes; planes with curvature >= threshold are
 * dropped whenever threshold <= max curvature.  One cooperative launch (grid barriers only around rounds that remove
 * planes).  xyz / normal float[n][3], keep int32[n] out, curv float[n] scratch, state uint32[4] zero-filled scratch. */
int pcs_plane_prune(pcs_stream_t s, const float *xyz, const float *normal, int n, int K, const float *thresholds,
                    int n_thr, int32_t *keep, float *curv, uint32_t *state);

/* Velocity smoothing of the tracker, smooth_velo (pcdet/models/registration/preprocessors/cluster_tracking.py:162-199):
 * AdamW (torch defaults, lr 1e-2, MultiStepLR [100,200,300]) on velos[:, a..b, :2] with loss
 * w0 * mean (v - d)^2 + w * mean |v[f] - v[f+1]| and the reference's 3-strike stopping rule, in one single-CTA launch.
 * velos float[C][F][3] in/out (all elements receive AdamW's weight decay like the reference's parameter tensor),
 * diffs float[C][F][3], m / v zero-filled scratch float[C * (b-a+1) * 2], info int32[2] = (iterations, stopped). */
int pcs_smooth_velo(pcs_stream_t s, float *velos, const float *diffs, float *m, float *v, int C, int F, int a, int b,
                    float w0, float w, int num_itr, float stopping, int32_t *info);

/* Row gather dst[i] = src[idx[i]] (rows of 1 / 4 / 8 / 12 / 16 bytes, int64 indices): the re-ordering after the
 * subsample (simple_reg.py:126-130), the ground-mask filter (ground_plane_remover.py:238-247) and the voxel -> point
 * broadcast of the ground stage (preprocessor_utils.py:416-419). */
int pcs_gather_rows(pcs_stream_t s, const void *src, const int64_t *idx, int64_t n, int row_bytes, void *dst);

/* Per-group min and max of a float column: out_min / out_max float[C] (empty groups -> 0, the torch_scatter
 * convention), values[i * stride], ids int64[n] in [0, C) (rows sorted by group reduce per warp before the atomics),
 * tmp uint32[2 * C] scratch.  Replaces the scatter(min) / scatter(max) pair of preprocessor_utils.py:113-114. */
int pcs_group_minmax(pcs_stream_t s, const float *values, int64_t stride, const int64_t *ids, int64_t n, int64_t C,
                     uint32_t *tmp, float *out_min, float *out_max);


/* ---- batched cluster tracker ---------------------------------------------------------------------
 * Replaces the anchor / key / target-frame loops of ClusterTracking.forward and track_frame
 * (pcdet/models/registration/preprocessors/cluster_tracking.py:430-787, 853-884) together with sample_frame
 * (:39-51), register_to_next_frame (registration_utils.py:83-206), smooth_velo (:162-199) and the nn_graph
 * extraction (:710-721).  All (anchor frame, component key) "instances" of a sequence advance together: tracking
 * step t in [1, 16] moves every anchor a to frame a - t (t <= 8) or a + (t - 8); no host synchronisation between
 * steps.  The descriptor structs below hold device pointers and scalars only; every field is 8 bytes wide
 * (pointer, int64_t or double) so that any FFI can fill them without caring about padding.
 *
 * Cell grids of the tracker: 16-byte pcs_slot_t tables over packed keys (group << 48 | cx << 32 | cy << 16 | cz),
 * cell = floor((p - lo) / cs) with one sequence-global origin lo; group = instance (moving side) or frame
 * (reference side).  Rows are float4 (payload bits, x, y, z). */
typedef struct {
  const void *pts;   /* float4[n] (., x, y, z) */
  const void *group; /* int32[n] instance / frame of every point, < 32768 */
  const void *skey;  /* int32[n] sort key (component id, ascending == ascending reference id) or NULL: sort by group */
  const void *bits;  /* uint8[n] up to three flag bits (`stationary` per component key) or NULL */
  const void *act;   /* int32[n_groups] or NULL: groups with act == 0 are skipped */
  int64_t n, n_groups, n_keys, ns_only; /* ns_only: voxels whose flag bit 0 wins the majority vote are dropped */
  double size[3];    /* voxel size (GridSampling3D grid_size) */
  void *sb;          /* uint32[n_groups][6] bounds of the groups' points (pcs_trk_group_bounds); reset on return */
  void *table;       /* 16-byte slots [H] (pcs_trk_sampler_init) */
  int64_t H;
  void *vsum, *vbits, *vk, *vres; /* double[H][3], int32[H][3], int32[H][2], int32[H][2] */
  void *pnext, *vlist;            /* int32[n] each */
  void *ctr;                      /* int32[4]: [0] voxels, [1] voxels kept, [2] sticky error flag */
  void *kcount, *koff, *kcur;     /* int32[n_keys + 1] each; koff = first output row of every sort key */
  void *vdeg;                     /* int32[n_keys] += voxels per sort key (dropped ones included), or NULL */
  void *out_pts, *out_key, *out_group; /* float4[n] (flag bits, mean x, y, z), int32[n], int32[n]; grouped by sort key */
} pcs_trk_sampler_t;

typedef struct {
  int64_t J, G;
  const void *act, *ref_group, *ref_group_all, *skipmask; /* int32[J]: participates, group of its reference voxels
                                             (ICP targets / matched-fraction search), flag mask skipped */
  const void *ref_off;                    /* int32[n_groups + 1] rows of every group in ref_pts */
  const void *g_inst;                     /* int32[G] instance of every component (components grouped by instance) */
  void *ref_table; int64_t ref_H; void *ref_pts; /* static reference grid; ref_pts float4[.] sorted by cell key */
  void *mov_table; int64_t mov_H; void *mov_sorted, *mov_sidx, *mov_cells, *mov_ctr; /* moving grid scratch */
  void *mv; const void *mv_gid, *mv_inst, *n_mv; /* moving voxels float4[.] (in place), component, instance, count */
  int64_t mv_cap;                         /* capacity of the moving voxel arrays (n_mv[0] <= mv_cap) */
  const void *vdeg;                       /* int32[G] voxels per component (matched-fraction denominator) */
  double lo[3], cs;                       /* grid origin and cell size (>= 1.001 * 3-D radius / rings) */
  int64_t rings;                          /* 1: search 3x3x3 cells, 2: 5x5x5 (cell size = half the radius) */
  double radius; int64_t df; double angle_reg; int64_t max_iter; double stopping_delta;
  int64_t want_l1, want_ratio;
  void *nn_fwd, *nn_bwd, *boff;           /* int32[cap mv], int32[cap backward items], int32[J + 1] */
  void *mvbeg, *mvend;                    /* int32[J] scratch */
  void *sec_fwd, *sec_bwd, *disp, *dmax;  /* optional (all or none) neighbour caching: float[cap mv], float[cap backward
                                             items], float[cap mv], int32[J].  A query skips its search while its cached
                                             neighbour is provably still the nearest (distance bounds, exact) */
  void *mom, *Ti, *T, *mu, *l1_sum, *l1_n; /* double [G][17], [G][12], [G][12] (out), [G][6], [G][2], [G] */
  void *phase, *cd, *iters, *itcnt;       /* int32[J] x3 (iters = iterations run, out), int32[(max_iter + 2) * 3] */
  void *last, *loss;                      /* double[J] */
  void *match_cnt;                        /* int32[G] */
  void *l1_err, *ratio;                   /* double[G], float[G] outputs (want_l1 / want_ratio) */
  void *prof;                             /* optional int64[256] += ns per phase (B C D E F G H tail), [8] iterations,
                                             [9] launches, [13] / [14] moving / reference voxels searched (summed over
                                             iterations), [15] searches answered from the neighbour cache, [16..] per-iteration search ns, [112..] active queries */
} pcs_trk_icp_t;

typedef struct {
  int64_t J, G, M, F;
  const void *inst_anchor, *inst_key, *inst_C, *inst_fmin, *inst_fmax, *inst_has_valid, *inst_goff; /* int32[J(+1)] */
  const void *seq_sorted, *frame_off;     /* float4[N] points sorted by frame, int32[F + 1] */
  void *mp; const void *mp0; void *m_last; const void *m_gid, *m_inst; /* moving points grouped by component */
  const void *g_inst, *g_deg, *g_diam, *g_valid; /* int32, int32, float, uint8 [G] */
  void *g_stopped, *g_moving, *g_final, *g_minf, *g_maxf; /* uint8 x3, int32 x2 [G] */
  void *transforms;                       /* double[G][17][12] (R row-major, t), identity on entry */
  void *velos, *velos_b, *centers, *diffs; /* float[G][17][3] */
  void *cv_pre, *g_delta, *adam_m, *adam_v; /* float[G][3] x2, float[G][16] x2 */
  void *csum, *vsum, *l1_err, *ratio, *T; /* double[G][3] x2 (zero), double[G], float[G], double[G][12] */
  void *vdeg;                             /* int32[G] (zero) */
  void *cur_act, *cur_nxt, *cur_rel, *cur_haslv, *cur_grp, *cur_grp_all; /* int32[J] x6 */
  int64_t n_keys;                         /* component keys: reference group = key * F + frame, all voxels: n_keys * F + frame */
  void *anyns;                            /* int32[17][J] (zero) */
  void *sb;                               /* uint32[J][6] (pcs_trk_bounds_reset) */
  double reg_error_coeff, angle_threshold; int64_t min_move_frame;
  double radius[8], voxel_size[24];       /* per registration level */
  double lo[3], nn_radius;
  void *eg_table; int64_t eg_H; void *eg_sorted, *eg_sidx, *eg_cells, *eg_ctr; /* extraction grid scratch */
  void *eoff; const void *exoff; void *ex; /* int32[J + 1], int64[J * 17 + 1], int32[exoff[J * 17]] (filled with -1) */
} pcs_trk_ctx_t;

int pcs_trk_cell_keys(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, const double *lo, double cs,
                      int64_t *keys);
/* table <- n unique cells (key, first row, rows) of a key-sorted row array; H power of two >= 2 n; err int32[1]. */
int pcs_trk_grid_fill(pcs_stream_t s, pcs_slot_t *table, int64_t H, const int64_t *keys, const int32_t *starts,
                      const int32_t *counts, int64_t n, int32_t *err);
int pcs_trk_table_clear(pcs_stream_t s, pcs_slot_t *table, int64_t H, int32_t *ctr /* int32[4] zeroed, or NULL */);
int pcs_trk_bounds_reset(pcs_stream_t s, uint32_t *sb, int n_groups);
int pcs_trk_group_bounds(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, uint32_t *sb);
int pcs_trk_sampler_init(pcs_stream_t s, const pcs_trk_sampler_t *S);
/* sample_frame for all groups at once: per-group origin = min of its points, cell = trunc((p - origin) / size),
 * per voxel fp64 mean, majority vote per flag bit, upper-median sort key. */
int pcs_trk_sample(pcs_stream_t s, const pcs_trk_sampler_t *S);
/* register_to_next_frame for all instances in one persistent cooperative launch. */
int pcs_trk_icp(pcs_stream_t s, const pcs_trk_icp_t *P);
/* CUDA-event timing of the ICP launches: enable / reset, then read (count, summed ms; synchronises on the events). */
void pcs_trk_icp_timing(int enable);
int pcs_trk_icp_elapsed(double *total_ms);
int pcs_trk_dir_init(pcs_stream_t s, const pcs_trk_ctx_t *C);
int pcs_trk_step(pcs_stream_t s, const pcs_trk_ctx_t *C, const pcs_trk_sampler_t *S, const pcs_trk_icp_t *levels,
                 int n_levels, int t);
/* final component filter + anchor-frame rows of the extraction table; m_frow int32[M] = row of the point in its frame */
int pcs_trk_finish(pcs_stream_t s, const pcs_trk_ctx_t *C, const int32_t *m_frow);
/* steps 1..16 and pcs_trk_finish */
int pcs_trk_run(pcs_stream_t s, const pcs_trk_ctx_t *C, const pcs_trk_sampler_t *S, const pcs_trk_icp_t *levels,
                int n_levels, const int32_t *m_frow);


/* Grid over n rows float4 (., x, y, z) tagged with group[n] (< 32768).  table: 16-byte slots [H >= n, power of two]
 * cleared by the call; sorted float4[n], sidx int32[n], cells int32[n], ctr int32[4] scratch. */
int pcs_trk_group_grid(pcs_stream_t s, const float *pts, const int32_t *group, int64_t n, const double *lo, double cs,
                       pcs_slot_t *table, int64_t H, float *sorted, int32_t *sidx, int32_t *cells, int32_t *ctr);
/* Nearest grid row (index into the pts given to pcs_trk_group_grid, or -1) within `radius` (d2 <= r*r, fp32 FMA order
 * of the reference) for the queries of nseg segments: segment k covers queries[seg_qstart[k] ...] with
 * seg_off[k + 1] - seg_off[k] rows and searches group seg_group[k]; out int32[seg_off[nseg]].  Replaces the per-frame
 * nn_graph calls of extract_traces_and_update_boxes (cluster_tracking.py:356-358). */
int pcs_trk_group_nn(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted, const int32_t *sidx,
                     const double *lo, double cs, const float *queries, const int32_t *seg_qstart,
                     const int32_t *seg_off, const int32_t *seg_group, int nseg, float radius, int32_t *out);

/* ---- GT evaluation -------------------------------------------------------------------------------
 * Replaces points_in_boxes_cpu (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168) and the per-component
 * loops around it (cluster_proposal.py:90-114, 206-255; cluster_tracking.py:340-411).
 * pcs_box_prep: boxes float[B][7] (x, y, z, dx, dy, dz, heading), sorted by frame -> recs (48 bytes per box).
 * pcs_points_in_boxes: item i is point pts[sel ? sel[i] : i] (float4, column 0 = frame); it is tested against the
 * boxes box_off[f] .. box_off[f + 1] of its frame.  first_out[i] = frame-local index of the first box holding it
 * (-1: none); for every box holding it, cnt_k[cid_k[i] * Bmax + local index] += 1 (k = 0..2, optional).
 * err int32[1] is set when a frame has more than Bmax boxes. */
int pcs_box_prep(pcs_stream_t s, const float *boxes, int64_t B, void *recs);
int pcs_points_in_boxes(pcs_stream_t s, const float *pts, const int32_t *sel, int64_t n, const void *recs,
                        const int32_t *box_off, int F, int Bmax, const int64_t *cid0, const int64_t *cid1,
                        const int64_t *cid2, int32_t *cnt0, int32_t *cnt1, int32_t *cnt2, int32_t *first_out,
                        int32_t *err);


/* ---- reference torch_hash op on caller-supplied voxel coordinates -------------------------------------
 * The four entry points of pcdet/ops/torch_hash (torch_hash.h:16-32, torch_hash_api.cpp:9-15) for callers that bring
 * their own integer voxel coordinates (HashTable.find_corres*, torch_hash_modules.RadiusGraph / ChamferDistance).
 * key = map2key(coord, dims) with the reference's clamp; cells [coord + qmin, coord + qmax] (dimension 0 fastest);
 * fp32 distance over all nd <= 4 columns with one FMA per dimension.  The table holds one 16-byte slot per occupied
 * cell (H power of two, >= 2 n is always enough) and `rows` the point indices grouped by cell -- the caller's
 * (keys, values, reverse_indices) buffers of the reference signatures are opaque scratch for the binding
 * (pcseqlearning_b200/torch_hash_cuda.py).  dims / qmin / qmax are HOST arrays.
 *   pcs_compat_hash_insert     hash_insert_gpu      (torch_hash_kernel.cu:54-91, 411-442)
 *   pcs_compat_radius_degree   first pass of radius_graph_gpu (:224-288); K = -1 counts all neighbours
 *   pcs_compat_radius_fill     second pass (:290-409): edges int64[E][2] rows (ref, query), ascending query, ascending
 *                              (distance, index) inside a query; offsets = exclusive scan of degree
 *   pcs_nn_correspondence      correspondence       (:96-155): nearest row, NO radius test, -1 if no cell is occupied
 *   pcs_points_in_radius       points_in_radius_gpu (:160-222): visited[row] = 1 for d2 < r*r (strict) */
int pcs_compat_hash_insert(pcs_stream_t s, const int64_t *coords, int64_t n, int nd, const int64_t *dims_host,
                           pcs_slot_t *table, int64_t H, int32_t *rows, int32_t *ctr);
int pcs_compat_radius_degree(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows,
                             const float *values, int nd, const int64_t *dims_host, const int64_t *qcoords,
                             const float *qvalues, int64_t m, const int *qmin, const int *qmax, const float *radius,
                             int K, int32_t *degree);
int pcs_compat_radius_fill(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                           int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                           const int *qmin, const int *qmax, const float *radius, const int32_t *degree,
                           const int64_t *offsets, int64_t *edges, float *dists);
int pcs_nn_correspondence(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                          int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                          const int *qmin, const int *qmax, int64_t *corres);
int pcs_points_in_radius(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const int32_t *rows, const float *values,
                         int nd, const int64_t *dims_host, const int64_t *qcoords, const float *qvalues, int64_t m,
                         const int *qmin, const int *qmax, float radius, int64_t *visited);


/* ---- composite entry points (the names of SURVEY.md section 8b) -----------------------------------------
 * pcs_radius_graph          = pcs_radius_search (single pass, multi-radius capable)
 * pcs_connected_components  = pcs_uf_init + pcs_uf_union_edges + pcs_uf_labels (graph_utils.py:40-53)
 * pcs_voxelize              = pcs_voxelize_params + _insert + pcs_sort_pairs + _finish (grid_sampling.py:22-46);
 *                             synchronises the stream once, *num_voxels (host) receives V
 * pcs_register_icp          = pcs_trk_icp (registration_utils.py:83-206 for a batch of instances) */
int pcs_radius_graph(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                     const int32_t *sorted_idx, int seg_div, int n_seg, const float *seg_lo, const int64_t *seg_dims,
                     const float *vs, const float *queries, int64_t m, const int32_t *order, const int *qmin,
                     const int *qmax, const float *radius, float radius_scalar, int K, int32_t *nbr_idx, float *nbr_d2,
                     int32_t *nbr_cnt, int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full, int n_uf,
                     const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits);
int pcs_connected_components(pcs_stream_t s, int32_t *parent, const int64_t *e0, const int64_t *e1, int64_t E, int64_t n,
                             const int32_t *seg_of, int n_seg, int64_t *labels, int64_t *n_comp, void *tmp,
                             int64_t tmp_bytes);
int pcs_voxelize(pcs_stream_t s, const float *pts, int64_t n, const uint32_t *bounds, const float *size, int ignore_dim0,
                 float *start, int64_t *strides, void *table, int64_t H, int32_t *pt_vid, double *sums, int32_t *maxidx,
                 int32_t *counts, int64_t *ukeys, int32_t *uids, int32_t *counters, int64_t *keys_sorted,
                 int32_t *ids_sorted, void *sort_tmp, int64_t sort_tmp_bytes, int32_t *rank_of, int64_t *inv,
                 float *sampled, int64_t *maxidx_out, int32_t *counts_out, int64_t *num_voxels);
int pcs_register_icp(pcs_stream_t s, const pcs_trk_icp_t *P);

#ifdef __cplusplus
}
#endif
#endif
