// aliases.cu -- the entry-point names SURVEY.md section 8b lists for the C ABI, as thin compositions of the library's
// finer-grained calls (no kernels of their own).
#include "common.cuh"

using namespace pcs;

extern "C" {

/* radius_graph_gpu in one pass (multi-radius capable): same arguments as pcs_radius_search. */
int pcs_radius_graph(pcs_stream_t s, const pcs_slot_t *table, int64_t H, const float *sorted_pts,
                     const int32_t *sorted_idx, int seg_div, int n_seg, const float *seg_lo, const int64_t *seg_dims,
                     const float *vs, const float *queries, int64_t m, const int32_t *order, const int *qmin,
                     const int *qmax, const float *radius, float radius_scalar, int K, int32_t *nbr_idx, float *nbr_d2,
                     int32_t *nbr_cnt, int32_t *const *uf_parents, const float *uf_r2, const int *uf_need_full, int n_uf,
                     const int32_t *skip_full_cnt, const uint32_t *occ, int64_t occ_bits) {
  return pcs_radius_search(s, table, H, sorted_pts, sorted_idx, seg_div, n_seg, seg_lo, seg_dims, vs, queries, m, order,
                           qmin, qmax, radius, radius_scalar, K, nbr_idx, nbr_d2, nbr_cnt, uf_parents, uf_r2,
                           uf_need_full, n_uf, skip_full_cnt, occ, occ_bits);
}

/* graph_utils.connected_components on an edge list: union-find init + union + canonical labels in one call. */
int pcs_connected_components(pcs_stream_t s, int32_t *parent, const int64_t *e0, const int64_t *e1, int64_t E, int64_t n,
                             const int32_t *seg_of, int n_seg, int64_t *labels, int64_t *n_comp, void *tmp,
                             int64_t tmp_bytes) {
  int rc = pcs_uf_init(s, parent, n);
  if (rc) return rc;
  rc = pcs_uf_union_edges(s, parent, e0, e1, E);
  if (rc) return rc;
  return pcs_uf_labels(s, parent, n, seg_of, n_seg, labels, n_comp, tmp, tmp_bytes);
}

/* GridSampling3D.forward (grid_cluster + unique(sorted) + scatter mean) in one call.  Synchronises the stream once
 * (the number of voxels sizes the sort); returns it in *num_voxels (host).  Buffers as in pcs_voxelize_insert /
 * pcs_sort_pairs / pcs_voxelize_finish; keys_sorted int64[n], ids_sorted int32[n], rank_of int32[n] scratch. */
int pcs_voxelize(pcs_stream_t s, const float *pts, int64_t n, const uint32_t *bounds, const float *size, int ignore_dim0,
                 float *start, int64_t *strides, void *table, int64_t H, int32_t *pt_vid, double *sums, int32_t *maxidx,
                 int32_t *counts, int64_t *ukeys, int32_t *uids, int32_t *counters, int64_t *keys_sorted,
                 int32_t *ids_sorted, void *sort_tmp, int64_t sort_tmp_bytes, int32_t *rank_of, int64_t *inv,
                 float *sampled, int64_t *maxidx_out, int32_t *counts_out, int64_t *num_voxels) {
  int rc = pcs_voxelize_params(s, bounds, size, ignore_dim0, start, strides);
  if (rc) return rc;
  rc = pcs_voxelize_insert(s, pts, n, start, strides, size, ignore_dim0, table, H, pt_vid, sums, maxidx, counts, ukeys,
                           uids, counters);
  if (rc) return rc;
  int32_t host[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, as_stream(s));
  if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(s));
  if (e != cudaSuccess) return set_error((int)e, "pcs_voxelize: reading the voxel count");
  if (host[2] != 0) return set_error(host[2], "pcs_voxelize: device-side error flag");
  const int64_t V = host[0];
  if (num_voxels) *num_voxels = V;
  rc = pcs_sort_pairs(s, ukeys, keys_sorted, uids, ids_sorted, V, sort_tmp, sort_tmp_bytes);
  if (rc) return rc;
  return pcs_voxelize_finish(s, ids_sorted, V, pt_vid, n, sums, maxidx, counts, rank_of, inv, sampled, maxidx_out,
                             counts_out);
}

/* register_to_next_frame for a batch of (moving, target) instances: alias of pcs_trk_icp. */
int pcs_register_icp(pcs_stream_t s, const pcs_trk_icp_t *P) { return pcs_trk_icp(s, P); }

}  // extern "C"
