"""Host-side wrappers of the C ABI (include/pcseq_b200.h) on torch CUDA tensors.

torch is used for device memory and the current stream only; every computation on the path happens
in libpcseq_b200.so.  All functions raise on CPU tensors -- there is no fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

PCS_MAX_K = 32
PCS_MAX_SEGMENTS = 64
GRID_SORTED = int(os.environ.get("PCS_GRID_SORTED", "0"))  # staged: key-ordered cell ranges in the proposal grids
SEARCH_THREADS = int(os.environ.get("PCS_SEARCH_THREADS", "0"))  # 1: thread-per-query kernel for the proposal passes
# (measured, 198 frames: fine / mid / coarse pass 8.2 / 13.6 / 4.4 ms vs 7.8 / 9.4 / 1.6 ms for the warp kernel; with
# key-ordered cells 5.4 / 12.0 / 3.0 ms but + 0.55 ms per grid build -- serial dependent loads per thread: off)
OCC_BITS_PER_SLOT = int(os.environ.get("PCS_OCC_BITS_PER_SLOT", "16"))  # 0 disables the occupancy bitmap

# Optional per-kernel CUDA-event log (bench.py's roofline leg): name -> list of (start, end, meta)
_EVENT_LOG = None


def enable_event_log(on=True):
    global _EVENT_LOG
    _EVENT_LOG = {} if on else None


def event_log():
    return _EVENT_LOG


class _timed:
    """Records CUDA events around one kernel launch on the current stream when the event log is enabled."""

    def __init__(self, name, **meta):
        self.name, self.meta = name, meta

    def __enter__(self):
        if _EVENT_LOG is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _EVENT_LOG is not None:
            self.b.record()
            _EVENT_LOG.setdefault(self.name, []).append((self.a, self.b, self.meta))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f4(vals):
    return (ctypes.c_float * 4)(*[float(v) for v in vals])


def _i4(vals):
    return (ctypes.c_int * 4)(*[int(v) for v in vals])


def _as_points(t, name="points"):
    """float32 [N,4] contiguous CUDA rows; [N,3] inputs (batch,x,y) get a zero 4th coordinate."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PcsError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dim() != 2 or t.shape[1] not in (3, 4):
        raise _lib.PcsError(f"{name} must have shape [N,4] (or [N,3]), got {tuple(t.shape)}")
    t = t.float()
    if t.shape[1] == 3:
        t = torch.cat([t, t.new_zeros(t.shape[0], 1)], 1)
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def next_pow2(x):
    return 1 << max(1, int(x - 1).bit_length())


def radius_voxel_size(radius):
    """[1 - 1e-3, r, r, r] in fp32 (graph_utils.py:170)."""
    r = float(np.float32(radius))
    return [float(np.float32(1 - 1e-3)), r, r, r]


class CellGrid:
    """Voxel hash of a reference point set: unique-cell open-addressing table + cell-sorted points.

    Replaces the (keys, values, reverse_indices) multimap the reference allocates per RadiusGraph call
    (graph_utils.py:179-192).  `bounds_sets` are the point sets whose joint min/max define the grid
    origin (the reference uses ref U query, graph_utils.py:171-173).
    """

    def __init__(self, ref, voxel_size, bounds_sets=None, seg_div=1, n_seg=1, table_size=None, pad=0, geometry=None,
                 sorted_cells=False):
        L = _lib.lib()
        self.ref = _as_points(ref, "ref")
        dev = self.ref.device
        self.n = self.ref.shape[0]
        self.vs = [float(v) for v in voxel_size]
        self.seg_div, self.n_seg = int(seg_div), int(n_seg)
        if not (1 <= self.n_seg <= PCS_MAX_SEGMENTS):
            raise _lib.PcsError(f"n_seg must be in [1, {PCS_MAX_SEGMENTS}]")
        with torch.cuda.device(dev):
            s = _stream()
            if geometry is not None:  # share another grid's origin / dims (same cells)
                self.seg_lo, self.seg_dims = geometry
            else:
                self.bounds = torch.empty(self.n_seg * 8, dtype=torch.int32, device=dev)
                _lib.check(L.pcs_bounds_init(s, _ptr(self.bounds), self.n_seg), "pcs_bounds_init")
                sets = [self.ref] if bounds_sets is None else [_as_points(b, "bounds set") for b in bounds_sets]
                seen = set()
                for b in sets:
                    if b.data_ptr() in seen or b.shape[0] == 0:
                        continue
                    seen.add(b.data_ptr())
                    _lib.check(L.pcs_bounds_update(s, _ptr(b), b.shape[0], self.seg_div, self.n_seg,
                                                   _ptr(self.bounds)), "pcs_bounds_update")
                self.seg_lo = torch.empty(self.n_seg, 4, dtype=torch.float32, device=dev)
                self.seg_dims = torch.empty(self.n_seg, 4, dtype=torch.int64, device=dev)
                _lib.check(L.pcs_grid_params(s, _ptr(self.bounds), self.n_seg, _f4(self.vs), int(pad),
                                             _ptr(self.seg_lo), _ptr(self.seg_dims)), "pcs_grid_params")
            self.H = int(table_size) if table_size else next_pow2(max(2 * self.n, 1024))
            self.table = torch.empty(self.H, 4, dtype=torch.int32, device=dev)  # pcs_slot_t[H]
            self.sorted_pts = torch.empty(max(self.n, 1), 4, dtype=torch.float32, device=dev)
            self.sorted_idx = torch.empty(max(self.n, 1), dtype=torch.int32, device=dev)
            self.counters = torch.empty(4, dtype=torch.int32, device=dev)
            # occupancy bitmap: 16 bits per table slot (<= 2^32 bits); empty neighbour cells cost the search one load
            self.occ_bits = min(max(self.H * OCC_BITS_PER_SLOT, 32), 1 << 32) if OCC_BITS_PER_SLOT > 0 else 0
            self.occ = torch.empty(self.occ_bits // 32, dtype=torch.int32, device=dev) if self.occ_bits else None
            with _timed("hash_build", n=self.n, H=self.H):
                if sorted_cells:  # staged: cell ranges in key order (frame-major), see pcs_hash_build_sorted
                    wb = int(L.pcs_hash_build_sorted_ws_bytes(self.n, self.H))
                    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
                    _lib.check(L.pcs_hash_build_sorted(s, _ptr(self.ref), self.n, self.seg_div, self.n_seg,
                                                       _ptr(self.seg_lo), _ptr(self.seg_dims), _f4(self.vs),
                                                       _ptr(self.table), self.H, _ptr(self.sorted_pts),
                                                       _ptr(self.sorted_idx), _ptr(self.counters), _ptr(self.occ),
                                                       self.occ_bits, _ptr(ws), wb), "pcs_hash_build_sorted")
                else:
                    _lib.check(L.pcs_hash_build(s, _ptr(self.ref), self.n, self.seg_div, self.n_seg,
                                                _ptr(self.seg_lo), _ptr(self.seg_dims), _f4(self.vs), _ptr(self.table),
                                                self.H, _ptr(self.sorted_pts), _ptr(self.sorted_idx),
                                                _ptr(self.counters), _ptr(self.occ), self.occ_bits), "pcs_hash_build")

    def check(self):
        """Synchronising check of the device-side error flag (table full / key overflow)."""
        c = self.counters.tolist()
        if c[2] != 0:
            raise _lib.PcsError(f"voxel hash build failed on device (code {c[2]})")
        return c[0]  # number of occupied cells

    def voxel_keys(self, pts):
        """Reference-exact voxel coordinates int64[N,4] and linear keys int64[N] (graph_utils.py:174-175)."""
        pts = _as_points(pts)
        coords = torch.empty(pts.shape[0], 4, dtype=torch.int64, device=pts.device)
        keys = torch.empty(pts.shape[0], dtype=torch.int64, device=pts.device)
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib().pcs_voxel_keys(_stream(), _ptr(pts), pts.shape[0], self.seg_div, self.n_seg,
                                                 _ptr(self.seg_lo), _ptr(self.seg_dims), _f4(self.vs), _ptr(coords),
                                                 _ptr(keys)), "pcs_voxel_keys")
        return coords, keys

    def search(self, query, K, radius, qmin=(0, -1, -1, -1), qmax=(0, 1, 1, 1), order=None, want_d2=False,
               uf_parent=None, want_lists=True, uf_targets=None, skip_full_cnt=None, cnt_out=None):
        """K nearest reference points within `radius` of every query (padded lists).

        query=None is the self-query mode: the grid's own points are the queries and are visited in
        cell order straight from the cell-sorted array (outputs are still indexed by original row).
        radius: python float or float32 tensor [M].  uf_parent: one union-find forest fed with every list entry;
        uf_targets: list of up to 3 (forest, radius, need_full) for the multi-radius search.
        Returns (nbr_idx i32[M,K] | None, nbr_cnt i32[M], nbr_d2 f32[M,K] | None).
        """
        if not (1 <= K <= PCS_MAX_K):
            raise _lib.PcsError(f"K must be in [1, {PCS_MAX_K}] (got {K})")
        if query is None:
            m, dev = self.n, self.ref.device
        else:
            query = _as_points(query, "query")
            m, dev = query.shape[0], query.device
        nbr_idx = torch.empty(m, K, dtype=torch.int32, device=dev) if want_lists else None
        nbr_d2 = torch.empty(m, K, dtype=torch.float32, device=dev) if (want_d2 and want_lists) else None
        nbr_cnt = cnt_out if cnt_out is not None else torch.empty(m, dtype=torch.int32, device=dev)
        rad_t, rad_s = None, 0.0
        if isinstance(radius, torch.Tensor):
            rad_t = radius.float().contiguous()
        else:
            rad_s = float(np.float32(radius))
        if order is not None:
            order = order.int().contiguous()
        targets = list(uf_targets or [])
        if uf_parent is not None:
            targets = [(uf_parent, float("inf"), False)] + targets
        n_uf = len(targets)
        uf_ptrs = (ctypes.c_void_p * 3)(*[t[0].data_ptr() for t in targets] + [0] * (3 - n_uf))
        uf_r2 = (ctypes.c_float * 3)(*[float(np.float32(t[1]) * np.float32(t[1])) if np.isfinite(t[1]) else 3.0e38
                                       for t in targets] + [0.0] * (3 - n_uf))
        uf_full = (ctypes.c_int * 3)(*[int(bool(t[2])) for t in targets] + [0] * (3 - n_uf))
        if (query is None and not want_lists and n_uf >= 1 and rad_t is None and order is None and SEARCH_THREADS):
            # cluster-proposal passes: thread-per-query kernel (no lists, unions + counts only)
            with torch.cuda.device(dev), _timed("radius_search", n_ref=self.n, n_query=m, K=int(K), lists=False,
                                                fused_uf=n_uf, threads=1):
                _lib.check(_lib.lib().pcs_self_search_uf(
                    _stream(), _ptr(self.table), self.H, _ptr(self.sorted_pts), _ptr(self.sorted_idx), self.n,
                    self.seg_div, self.n_seg, _ptr(self.seg_lo), _ptr(self.seg_dims), _f4(self.vs), _i4(qmin), _i4(qmax),
                    rad_s, int(K), _ptr(nbr_cnt), uf_ptrs, uf_r2, uf_full, n_uf, _ptr(skip_full_cnt), _ptr(self.occ),
                    self.occ_bits), "pcs_self_search_uf")
            return None, nbr_cnt, None
        with torch.cuda.device(dev), _timed("radius_search", n_ref=self.n, n_query=m, K=int(K), lists=bool(want_lists),
                                            fused_uf=n_uf):
            _lib.check(_lib.lib().pcs_radius_search(
                _stream(), _ptr(self.table), self.H, _ptr(self.sorted_pts), _ptr(self.sorted_idx), self.seg_div,
                self.n_seg, _ptr(self.seg_lo), _ptr(self.seg_dims), _f4(self.vs), _ptr(query), m, _ptr(order),
                _i4(qmin), _i4(qmax), _ptr(rad_t), rad_s, int(K), _ptr(nbr_idx), _ptr(nbr_d2), _ptr(nbr_cnt),
                uf_ptrs, uf_r2, uf_full, n_uf, _ptr(skip_full_cnt), _ptr(self.occ), self.occ_bits),
                "pcs_radius_search")
        return nbr_idx, nbr_cnt, nbr_d2


def exclusive_scan(counts):
    """int32[n] -> int64[n+1] exclusive prefix sums with the total in the last entry (no host sync)."""
    counts = counts.int().contiguous()
    n = counts.shape[0]
    L = _lib.lib()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    tb = int(L.pcs_exclusive_scan_tmp_bytes(n))
    tmp = torch.empty(tb, dtype=torch.uint8, device=counts.device)
    with torch.cuda.device(counts.device):
        _lib.check(L.pcs_exclusive_scan(_stream(), _ptr(counts), n, _ptr(out), _ptr(tmp), tb), "pcs_exclusive_scan")
    return out


def lists_to_edges(nbr_idx, nbr_cnt, nbr_d2=None):
    """Padded lists -> (edges int64[E,2] rows (ref, query) by ascending query, dists | None).  One host sync (E)."""
    m, K = nbr_idx.shape
    offsets = exclusive_scan(nbr_cnt)
    E = int(offsets[-1].item())
    edges = torch.empty(E, 2, dtype=torch.int64, device=nbr_idx.device)
    dists = torch.empty(E, dtype=torch.float32, device=nbr_idx.device) if nbr_d2 is not None else None
    with torch.cuda.device(nbr_idx.device):
        _lib.check(_lib.lib().pcs_lists_to_edges(_stream(), _ptr(nbr_idx), _ptr(nbr_d2), _ptr(nbr_cnt), _ptr(offsets),
                                                 m, K, _ptr(edges), _ptr(dists)), "pcs_lists_to_edges")
    return edges, dists


def radius_graph(ref, query, radius, max_num_neighbors=32, sort_by_dist=True, qmin=(0, -1, -1, -1),
                 qmax=(0, 1, 1, 1), return_dists=False):
    """RadiusGraph.build_graph (graph_utils.py:149-209) on the new kernels -> (e_ref, e_query[, d2]).

    The K kept neighbours are always the K nearest (ties by ascending reference index); with
    sort_by_dist=False the reference keeps the first K in its race-dependent discovery order, of
    which "the K nearest" is one legal outcome.
    """
    ref = _as_points(ref, "ref")
    same = query is ref
    query = ref if same else _as_points(query, "query")
    if ref.shape[0] == 0 or query.shape[0] == 0:
        z = torch.zeros(0, dtype=torch.int64, device=ref.device)
        return (z, z.clone(), torch.zeros(0, device=ref.device)) if return_dists else (z, z.clone())
    rmax = float(radius.max().item()) if isinstance(radius, torch.Tensor) else float(np.float32(radius))
    grid = CellGrid(ref, radius_voxel_size(rmax), bounds_sets=[ref, query])
    K = int(max_num_neighbors)
    if K < 1 or K > PCS_MAX_K:
        raise _lib.PcsError(f"max_num_neighbors must be in [1, {PCS_MAX_K}] on this path (got {K})")
    nbr_idx, nbr_cnt, nbr_d2 = grid.search(None if same else query, K, radius, qmin, qmax, want_d2=return_dists)
    edges, dists = lists_to_edges(nbr_idx, nbr_cnt, nbr_d2)
    grid.check()
    if return_dists:
        return edges[:, 0], edges[:, 1], dists
    return edges[:, 0], edges[:, 1]


def connected_components(edges, num_nodes, seg_of=None, n_seg=1):
    """graph_utils.connected_components (graph_utils.py:40-53) on the device.

    edges int64[2,E] (or a pair of int64[E]).  Returns (n_comp int64[n_seg] tensor, labels int64[N]).
    """
    e0, e1 = edges[0].long().contiguous(), edges[1].long().contiguous()
    dev = e0.device
    L = _lib.lib()
    parent = torch.empty(num_nodes, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        s = _stream()
        _lib.check(L.pcs_uf_init(s, _ptr(parent), num_nodes), "pcs_uf_init")
        _lib.check(L.pcs_uf_union_edges(s, _ptr(parent), _ptr(e0), _ptr(e1), e0.shape[0]), "pcs_uf_union_edges")
    return uf_labels(parent, seg_of, n_seg)


def uf_new(n, device):
    parent = torch.empty(n, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().pcs_uf_init(_stream(), _ptr(parent), n), "pcs_uf_init")
    return parent


def uf_labels(parent, seg_of=None, n_seg=1):
    n = parent.shape[0]
    dev = parent.device
    L = _lib.lib()
    labels = torch.empty(n, dtype=torch.int64, device=dev)
    n_comp = torch.zeros(n_seg, dtype=torch.int64, device=dev)
    tb = int(L.pcs_uf_labels_tmp_bytes(n, n_seg))
    tmp = torch.empty(max(tb, 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.pcs_uf_labels(_stream(), _ptr(parent), n, _ptr(seg_of), n_seg, _ptr(labels), _ptr(n_comp),
                                   _ptr(tmp), tb), "pcs_uf_labels")
    return n_comp, labels


def point_segments(pts, seg_div, n_seg):
    pts = _as_points(pts)
    out = torch.empty(pts.shape[0], dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(_lib.lib().pcs_point_segments(_stream(), _ptr(pts), pts.shape[0], seg_div, n_seg, _ptr(out)),
                   "pcs_point_segments")
    return out


def compact_grid(points, voxel_size, **kw):
    """CellGrid with a table sized for the usual case of many points per cell (H ~ N/4 instead of 2N: 8x less table
    to clear / scan and L2-resident lookups).  The build reports overflow through its error flag; in that case the
    grid is rebuilt with the always-sufficient default size.  Costs one host sync."""
    n = points.shape[0]
    small = next_pow2(max(n // 2, 1024))  # cells / points is ~0.45 at r = 0.25 m, far less for the larger radii
    if GRID_SORTED:
        kw = dict(kw, sorted_cells=True)
    grid = CellGrid(points, voxel_size, table_size=small, **kw)
    cells, _, code, _ = grid.counters.tolist()  # host sync
    if code == _lib.PCS_ERR_TABLE_FULL:
        grid = CellGrid(points, voxel_size, **kw)
        cells, _, code, _ = grid.counters.tolist()
    elif code == 0 and 2 * cells > small:
        # more than half full: linear probing degrades quickly (0.9 load = tens of probes per lookup, measured as a
        # 30x slower search on a frame-window shard).  Rebuild at a load factor <= 1/3.
        grid = CellGrid(points, voxel_size, table_size=next_pow2(3 * cells), **kw)
        cells, _, code, _ = grid.counters.tolist()
    if code != 0:  # key range overflow etc.: a rebuild cannot help
        raise _lib.PcsError(f"voxel hash build failed on device (code {code})")
    return grid


def cluster_labels(fxyz, radius, max_num_neighbors=32, chunk=10, num_frames=None, grid=None):
    """Fused cluster proposal for one radius: intra-frame K-nearest radius graph per `chunk`-frame
    segment + connected components, without materialising edges.

    Equals ClusterProposal.propose_cluster's inner loop (cluster_proposal.py:63-81): labels are numbered
    per chunk by ascending smallest member index with a running offset.  Returns (labels int64[N],
    n_comp int64[n_seg]).
    """
    fxyz = _as_points(fxyz, "point_fxyz")
    n = fxyz.shape[0]
    if num_frames is None:
        num_frames = int(fxyz[:, 0].max().item()) + 1
    n_seg = max(1, (num_frames + chunk - 1) // chunk)
    if n_seg > PCS_MAX_SEGMENTS:
        raise _lib.PcsError(f"{n_seg} chunks exceed PCS_MAX_SEGMENTS={PCS_MAX_SEGMENTS}")
    if grid is None:
        grid = compact_grid(fxyz, radius_voxel_size(radius), seg_div=chunk, n_seg=n_seg)
    parent = uf_new(n, fxyz.device)
    grid.search(None, int(max_num_neighbors), radius, uf_parent=parent, want_lists=False)
    seg_of = point_segments(fxyz, chunk, n_seg)
    n_comp, labels = uf_labels(parent, seg_of, n_seg)
    return labels, n_comp


def voxelize(points, grid_size, ignore_dim0=False, want_mean=True, want_max=False, want_counts=False,
             want_sums=False, bounds_hook=None):
    """GridSampling3D.forward on the device (grid_sampling.py:22-46).

    points f32[N,4]; grid_size [gx,gy,gz] (the frame digit has size 1).  Voxels are numbered by ascending
    cell key (torch.unique(sorted=True)).  Returns a dict with `inv` int64[N], `num` (python int, one host
    sync) and optionally `sampled` f32[V,4] (mean of all columns), `maxidx` int64[V] (highest point index
    per voxel: simple_reg.py:122-124), `counts` int32[V].  ignore_dim0 treats column 0 as zero
    (preprocessor_utils.grid_sample :21-30).  want_sums adds the fp64 per-voxel column sums f64[V,4];
    bounds_hook(bounds) may widen the point bounds in place before the grid is derived (frame-window sharding:
    all ranks voxelize on the grid of the whole sequence).
    """
    pts = _as_points(points, "points")
    n, dev = pts.shape[0], pts.device
    L = _lib.lib()
    size = [1.0] + [float(np.float32(g)) for g in grid_size]
    out = {}
    if n == 0 and bounds_hook is None:
        out.update(inv=torch.zeros(0, dtype=torch.int64, device=dev), num=0)
        return out
    with torch.cuda.device(dev):
        s = _stream()
        bounds = torch.empty(8, dtype=torch.int32, device=dev)
        _lib.check(L.pcs_bounds_init(s, _ptr(bounds), 1), "pcs_bounds_init")
        _lib.check(L.pcs_bounds_update(s, _ptr(pts), n, 1, 1, _ptr(bounds)), "pcs_bounds_update")
        if bounds_hook is not None:
            bounds_hook(bounds)
        start = torch.empty(4, dtype=torch.float32, device=dev)
        strides = torch.empty(5, dtype=torch.int64, device=dev)
        _lib.check(L.pcs_voxelize_params(s, _ptr(bounds), _f4(size), int(ignore_dim0), _ptr(start), _ptr(strides)),
                   "pcs_voxelize_params")
        H = next_pow2(max(n + n // 4, 1024))  # load factor <= 0.8 even if every point had its own voxel
        table = torch.empty(H, 4, dtype=torch.int32, device=dev)
        pt_vid = torch.empty(n, dtype=torch.int32, device=dev)
        # per-voxel rows are indexed by dense id: n rows allocated, the first V touched
        sums = torch.empty(n, 4, dtype=torch.float64, device=dev) if (want_mean or want_sums) else None
        maxidx = torch.empty(n, dtype=torch.int32, device=dev) if want_max else None
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        ukeys = torch.empty(n, dtype=torch.int64, device=dev)
        uids = torch.empty(n, dtype=torch.int32, device=dev)
        counters = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.check(L.pcs_voxelize_insert(s, _ptr(pts), n, _ptr(start), _ptr(strides), _f4(size), int(ignore_dim0),
                                         _ptr(table), H, _ptr(pt_vid), _ptr(sums), _ptr(maxidx), _ptr(cnt),
                                         _ptr(ukeys), _ptr(uids), _ptr(counters)), "pcs_voxelize_insert")
        c = counters.tolist()  # host sync: V is needed to size the outputs
        if c[2] != 0:
            raise _lib.PcsError(f"voxelize failed on device (code {c[2]})")
        V = c[0]
        tb = int(L.pcs_sort_pairs_tmp_bytes(V))
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        keys_sorted = torch.empty(V, dtype=torch.int64, device=dev)
        ids_sorted = torch.empty(V, dtype=torch.int32, device=dev)
        _lib.check(L.pcs_sort_pairs(s, _ptr(ukeys), _ptr(keys_sorted), _ptr(uids), _ptr(ids_sorted), V, _ptr(tmp),
                                    tb), "pcs_sort_pairs")
        inv = torch.empty(n, dtype=torch.int64, device=dev)
        rank_of = torch.empty(max(V, 1), dtype=torch.int32, device=dev)
        sampled = torch.empty(V, 4, dtype=torch.float32, device=dev) if (want_mean or want_sums) else None
        maxidx_out = torch.empty(V, dtype=torch.int64, device=dev) if want_max else None
        counts = torch.empty(V, dtype=torch.int32, device=dev) if want_counts else None
        _lib.check(L.pcs_voxelize_finish(s, _ptr(ids_sorted), V, _ptr(pt_vid), n, _ptr(sums), _ptr(maxidx), _ptr(cnt),
                                         _ptr(rank_of), _ptr(inv), _ptr(sampled), _ptr(maxidx_out), _ptr(counts)),
                   "pcs_voxelize_finish")
    out.update(inv=inv, num=V, keys=keys_sorted)
    if want_sums:
        out["sums"] = sums[ids_sorted.long()]
    if want_mean or want_sums:
        out["sampled"] = sampled
    if want_max:
        out["maxidx"] = maxidx_out
    if want_counts:
        out["counts"] = counts
    return out


def group_median(values, inv, num_groups, counts=None):
    """robust_median (registration_utils.py:60-81): upper median of int64 `values` per group `inv`."""
    values = values.long().contiguous().reshape(-1)
    inv = inv.long().contiguous()
    n, dev = values.shape[0], values.device
    if counts is None:
        counts = torch.bincount(inv, minlength=num_groups).int()
    offsets = exclusive_scan(counts)
    cursor = torch.zeros(max(num_groups, 1), dtype=torch.int32, device=dev)
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    out = torch.empty(num_groups, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pcs_group_median(_stream(), _ptr(values), _ptr(inv), n, _ptr(offsets), num_groups,
                                               _ptr(cursor), _ptr(rows), _ptr(out)), "pcs_group_median")
    return out


def ground_ransac(vox_sorted, cidx_sorted, num_coarse, cmin_z, cmax_z, ratios, sigma2, stopping_delta=1e-2,
                  max_iter=50):
    """IRLS plane fits of all super-pillars for all height ratios in one cooperative launch
    (preprocessor_utils.py:32-80 + :147-170).  vox_sorted f32[Nv,4] (.,x,y,z) sorted by cidx_sorted int[Nv].
    Returns (best_center [C,3], best_normal [C,3], best_conf [C], iters int32[n_ratios])."""
    vox = _as_points(vox_sorted, "voxels")
    dev = vox.device
    Nv, C = vox.shape[0], int(num_coarse)
    cidx = cidx_sorted.int().contiguous()
    counts = torch.bincount(cidx_sorted.long(), minlength=C)
    seg_start = torch.zeros(C + 1, dtype=torch.int32, device=dev)
    seg_start[1:] = counts.cumsum(0).int()
    # local origins: first voxel of every super-pillar (empty ones get 0)
    first = seg_start[:-1].long().clamp(max=max(Nv - 1, 0))
    origin = torch.where((counts > 0)[:, None], vox[first, 1:], torch.zeros(C, 3, device=dev)).contiguous()
    ratios = ratios.float().contiguous().to(dev)
    R = int(ratios.shape[0])
    acc = torch.zeros(3 * R * C * 10, dtype=torch.float64, device=dev)
    nhit = torch.zeros(3 * R * C, dtype=torch.int32, device=dev)
    gmax = torch.zeros(3 * R, dtype=torch.int32, device=dev)
    planes = torch.zeros(2 * R * C * 6, dtype=torch.float32, device=dev)
    fin = torch.zeros(R * 2, dtype=torch.int32, device=dev)
    best_center = torch.zeros(C, 3, dtype=torch.float32, device=dev)
    best_normal = torch.zeros(C, 3, dtype=torch.float32, device=dev)
    best_normal[:, 2] = 1.0
    best_conf = torch.zeros(C, dtype=torch.float32, device=dev)
    iters = torch.zeros(ratios.shape[0], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _timed("ground_ransac", Nv=Nv, C=C):
        _lib.check(_lib.lib().pcs_ground_ransac(
            _stream(), _ptr(vox), _ptr(cidx), _ptr(seg_start), _ptr(origin), _ptr(cmin_z.float().contiguous()),
            _ptr(cmax_z.float().contiguous()), _ptr(ratios), Nv, C, int(ratios.shape[0]), float(sigma2),
            float(stopping_delta), int(max_iter), _ptr(acc), _ptr(nhit), _ptr(gmax), _ptr(planes), _ptr(fin),
            _ptr(best_center), _ptr(best_normal), _ptr(best_conf), _ptr(iters)), "pcs_ground_ransac")
    return best_center, best_normal, best_conf, iters


def plane_prune(xyz, normal, K, thresholds):
    """Curvature pruning loop of preprocessor_utils.py:175-193 in one launch -> bool keep mask [n]."""
    xyz = xyz.float().contiguous()
    normal = normal.float().contiguous()
    n, dev = xyz.shape[0], xyz.device
    thr = torch.as_tensor(np.asarray(thresholds, dtype=np.float32), device=dev)
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    curv = torch.zeros(n, dtype=torch.float32, device=dev)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _timed("plane_prune", n=n):
        _lib.check(_lib.lib().pcs_plane_prune(_stream(), _ptr(xyz), _ptr(normal), n, int(K), _ptr(thr),
                                              int(thr.shape[0]), _ptr(keep), _ptr(curv), _ptr(state)),
                   "pcs_plane_prune")
    return keep.bool()


def smooth_velo(velos, diffs, a, b, weight0=1.0, weight=10.0, num_itr=300, stopping=1e-3):
    """smooth_velo (cluster_tracking.py:162-199) in one launch; velos f32[C,F,3] is updated IN PLACE (a < b)."""
    assert velos.is_contiguous() and velos.dtype == torch.float32
    C, F, _ = velos.shape
    dev = velos.device
    diffs = diffs.float().contiguous()
    S = C * (b - a + 1) * 2
    m = torch.zeros(S, dtype=torch.float32, device=dev)
    v = torch.zeros(S, dtype=torch.float32, device=dev)
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _timed("smooth_velo", S=S):
        _lib.check(_lib.lib().pcs_smooth_velo(_stream(), _ptr(velos), _ptr(diffs), _ptr(m), _ptr(v), C, F, int(a), int(b),
                                              float(weight0), float(weight), int(num_itr), float(stopping), _ptr(info)),
                   "pcs_smooth_velo")
    return velos, info


SMOOTH_VELO_MAX = 24 * 1024
PRUNE_MAX_PLANES = 8192
L1_MAX_CELLS = 65536


def l1_heightfield(min_z, weight, lr, decay_steps, rigid_weight, max_iters, lr_gamma=0.1):
    """AdamW L1 smoothing of the pillar height grid in one launch (preprocessor_utils.py:313-350).
    Returns (height [X,Y], iterations run (device int32[2]: iterations, stopped-early))."""
    X, Y = min_z.shape
    dev = min_z.device
    mz = min_z.float().contiguous()
    wt = weight.float().contiguous().reshape(X, Y)
    h = torch.zeros(X, Y, dtype=torch.float32, device=dev)
    m = torch.zeros_like(h)
    v = torch.zeros_like(h)
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    decay = int(decay_steps[0]) if len(decay_steps) else 0
    assert len(decay_steps) <= 1, "one MultiStepLR milestone is supported on this path"
    with torch.cuda.device(dev), _timed("l1_heightfield", cells=X * Y):
        _lib.check(_lib.lib().pcs_l1_heightfield(_stream(), _ptr(mz), _ptr(wt), _ptr(h), _ptr(m), _ptr(v), X, Y,
                                                 float(lr), float(lr_gamma), decay, float(rigid_weight),
                                                 int(max_iters), _ptr(info), _ptr(loss)), "pcs_l1_heightfield")
    return h, info


def register_icp(mov_fxyz, mov_comp, mov_stationary, ref_fxyz, ref_stationary, num_components, radius, frame_offset,
                 angle_regularizer=10.0, max_iter=20, stopping_delta=5e-2):
    """register_to_next_frame (registration_utils.py:83-206) in one persistent launch (the batched tracker kernel of
    csrc/track.cu with a batch of one instance).

    mov_fxyz f32[vm,4] / mov_comp int[vm] / mov_stationary bool[vm]: the moving (down-sampled) frame;
    ref_fxyz f32[vr,4] / ref_stationary bool[vr]: the target frame; frame_offset = ref frame - moving frame.
    Returns (moved_fxyz f32[vm,4], T f64[C,4,4], l1_component_error f64[C], comp_edge_ratio f32[C], info) where
    info is an int32[4] device tensor whose entry 1 is the number of iterations run.
    """
    from .tracker import register_pair
    return register_pair(mov_fxyz, mov_comp, mov_stationary, ref_fxyz, ref_stationary, num_components, radius,
                         frame_offset, angle_regularizer, max_iter, stopping_delta)


def cluster_labels_multi(fxyz, radii, max_num_neighbors=32, chunk=10, num_frames=None):
    """Multi-radius fused cluster proposals: a cascade of searches, finest radius first.

    Pass i searches radius r_i on its own (reference-identical) grid, but only for the queries whose list was not
    full after pass i-1.  A query whose list is full (K entries) already holds its K nearest points overall, so the
    same list is united into the forests of ALL larger radii and the query drops out of the cascade: the dense
    regions -- the expensive ones for large cells -- are settled by the cheap fine search.
    (Top-K within r is the prefix of top-K within R >= r cut at d <= r.)
    Returns ([labels per radius, in the order of `radii`], [n_comp per radius]).
    """
    fxyz = _as_points(fxyz, "point_fxyz")
    n = fxyz.shape[0]
    K = int(max_num_neighbors)
    if num_frames is None:
        num_frames = int(fxyz[:, 0].max().item()) + 1
    n_seg = max(1, (num_frames + chunk - 1) // chunk)
    order = sorted(range(len(radii)), key=lambda i: radii[i])
    r_sorted = [float(radii[i]) for i in order]
    if not (1 <= len(r_sorted) <= 3):
        raise _lib.PcsError("cluster_labels_multi expects 1 to 3 radii")
    # Edge sets grow with the radius (full lists are identical at every larger radius, the others are prefixes),
    # so the forest of radius r_{i+1} is the forest of r_i plus the edges found by pass i+1: every pass feeds ONE
    # forest and the next one starts from a copy of it.
    parents = []
    cnt = None
    dbg = os.environ.get("PCS_STAGE_TIMING")
    for i, r in enumerate(r_sorted):
        if dbg:
            import time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        parent = uf_new(n, fxyz.device) if i == 0 else parents[-1].clone()
        grid = compact_grid(fxyz, radius_voxel_size(r), seg_div=chunk, n_seg=n_seg)
        if dbg:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        _, cnt, _ = grid.search(None, K, r, uf_parent=parent, want_lists=False, skip_full_cnt=cnt, cnt_out=cnt)
        parents.append(parent)
        if dbg:
            torch.cuda.synchronize()
            print(f"[labels_multi] r={r} grid {1e3 * (t1 - t0):.1f} ms (H={grid.H}, cells={int(grid.counters[0])}), "
                  f"search {1e3 * (time.perf_counter() - t1):.1f} ms", flush=True)
    seg_of = point_segments(fxyz, chunk, n_seg)
    labels, n_comp = [None] * len(radii), [None] * len(radii)
    for pos, i in enumerate(order):
        n_comp[i], labels[i] = uf_labels(parents[pos], seg_of, n_seg)
    return labels, n_comp


def group_minmax(values, ids, num_groups):
    """(min, max) of float `values` [n] per group `ids` int64[n] -> two f32[num_groups] (empty groups 0): the
    scatter(min) / scatter(max) pair of the ground stage in one pass.  `values` may be a strided column view."""
    assert values.dim() == 1 and values.dtype == torch.float32 and values.is_cuda
    ids = ids.long().contiguous()
    n, dev, C = values.shape[0], values.device, int(num_groups)
    stride = values.stride(0) if n > 1 else 1
    tmp = torch.empty(2 * C, dtype=torch.int32, device=dev)
    out_min = torch.empty(C, dtype=torch.float32, device=dev)
    out_max = torch.empty(C, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pcs_group_minmax(_stream(), _ptr(values), stride, _ptr(ids), n, C, _ptr(tmp),
                                               _ptr(out_min), _ptr(out_max)), "pcs_group_minmax")
    return out_min, out_max


def gather_rows(src, idx):
    """src[idx] for a contiguous tensor whose rows are 1, 4, 8, 12 or 16 bytes (anything else falls back to torch
    indexing).  idx: int64 row indices."""
    if not src.is_cuda or not src.is_contiguous() or src.dim() == 0 or src.shape[0] == 0 or idx.numel() == 0:
        return src[idx]
    row_bytes = src.element_size() * int(np.prod(src.shape[1:])) if src.dim() > 1 else src.element_size()
    if row_bytes not in (1, 4, 8, 12, 16) or (row_bytes == 16 and src.data_ptr() % 16):
        return src[idx]
    idx = idx.long().contiguous()
    n = idx.shape[0]
    out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().pcs_gather_rows(_stream(), _ptr(src), _ptr(idx), n, int(row_bytes), _ptr(out)),
                   "pcs_gather_rows")
    return out


def launch_count():
    return int(_lib.lib().pcs_launch_count())


def reset_launch_count():
    _lib.lib().pcs_reset_launch_count()
