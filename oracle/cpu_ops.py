"""numpy + C (oracle.c) restatement of the reference's voxel-hash / radius-graph / voxelization / CC ops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function cites the reference lines (relative to /root/reference) it follows.  All arrays are
numpy; float32 arithmetic is done on float32 arrays so every intermediate rounds exactly as the
reference's fp32 torch kernels do (true division, round-half-even).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")
_lib = None


def build_oracle_lib(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off oracle.c -> oracle/_build/liboracle.so"""
    src = os.path.join(HERE, "oracle.c")
    if (not force) and os.path.exists(_SO) and (
            not os.path.exists(src) or os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle_lib())
        _lib.oracle_connected_components.restype = ctypes.c_int64
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return int(lib().oracle_num_threads())


def set_threads(n):
    lib().oracle_set_threads(ctypes.c_int(int(n)))


# --------------------------------------------------------------------------------------------
# RadiusGraph keys -- pcdet/models/model_utils/graph_utils.py:169-183
# --------------------------------------------------------------------------------------------
def radius_graph_keys(ref, query, radius):
    """Voxel coordinates, dims and table size of RadiusGraph.build_graph.

    ref f32[N,D], query f32[M,D], radius: python float or f32[M].
    Returns (coors_ref i64[N,D], coors_query i64[M,D], dims i64[D], radius f32[M], H).
    """
    ref = _c(ref, np.float32)
    query = _c(query, np.float32)
    D = ref.shape[1]
    rq = (np.zeros(query.shape[0], np.float32) + np.asarray(radius, np.float32)).astype(np.float32)  # :169
    rmax = float(rq.max())
    vs = np.array([1 - 1e-3] + [rmax] * (D - 1), dtype=np.float32)  # :170
    allp = np.concatenate([ref, query], 0)
    lo = allp.min(0) - vs * np.float32(2)  # :172
    hi = allp.max(0) + vs * np.float32(2)  # :173
    cr = np.rint((ref - lo) / vs).astype(np.int64) + 1  # :174 (torch.round = half-to-even)
    cq = np.rint((query - lo) / vs).astype(np.int64) + 1  # :175
    dims = np.rint((hi - lo) / vs).astype(np.int64) + 3  # :176
    H = int(ref.shape[0] / 0.5)  # :179 util_ratio
    return cr, cq, dims, rq, H


# --------------------------------------------------------------------------------------------
# torch_hash op -- pcdet/ops/torch_hash/src/torch_hash_kernel.cu
# --------------------------------------------------------------------------------------------
def hash_insert(keys, values, rev, dims, ins_keys, ins_vals):
    """hash_insert_gpu (:411-442), in place on keys i64[H], values f32[H,D], rev i64[H]."""
    D = ins_vals.shape[1]
    lib().oracle_hash_insert(_p(keys), _p(values), _p(rev), ctypes.c_int64(keys.shape[0]), _p(_c(dims, np.int64)),
                             ctypes.c_int(D), _p(_c(ins_keys, np.int64)), _p(_c(ins_vals, np.float32)),
                             ctypes.c_int64(ins_keys.shape[0]))


def radius_graph(keys, values, rev, dims, qkeys, qvals, qmin, qmax, radius, max_nbr, sort_by_dist,
                 return_dists=False):
    """radius_graph_gpu (:487-561): count kernel, exclusive scan, fill kernel -> edges i64[E,2]."""
    D = qvals.shape[1]
    M = qkeys.shape[0]
    dims = _c(dims, np.int64)
    qkeys = _c(qkeys, np.int64)
    qvals = _c(qvals, np.float32)
    qmin = _c(qmin, np.int32)
    qmax = _c(qmax, np.int32)
    radius = _c(radius, np.float32)
    degree = np.zeros(M, np.int32)
    L = lib()
    L.oracle_radius_graph_count(_p(keys), _p(values), ctypes.c_int64(keys.shape[0]), _p(dims), ctypes.c_int(D),
                                _p(qkeys), _p(qvals), ctypes.c_int64(M), _p(qmin), _p(qmax), _p(radius),
                                ctypes.c_int(int(max_nbr)), _p(degree))
    offset = (np.cumsum(degree, dtype=np.int64) - degree).astype(np.int64)  # :534-535
    E = int(degree.sum())
    edges = np.zeros((E, 2), np.int64)
    dists = np.zeros(E, np.float32)
    L.oracle_radius_graph_fill(_p(keys), _p(values), _p(rev), ctypes.c_int64(keys.shape[0]), _p(dims),
                               ctypes.c_int(D), _p(qkeys), _p(qvals), ctypes.c_int64(M), _p(qmin), _p(qmax),
                               _p(radius), _p(degree), _p(offset), ctypes.c_int(int(bool(sort_by_dist))),
                               _p(edges), _p(dists))
    if return_dists:
        return edges, dists
    return edges


def correspondence(keys, values, rev, dims, qkeys, qvals, qmin, qmax):
    """correspondence (:444-485) -> corres i64[M] (-1 when no candidate)."""
    D = qvals.shape[1]
    M = qkeys.shape[0]
    out = np.zeros(M, np.int64)
    lib().oracle_correspondence(_p(keys), _p(values), _p(rev), ctypes.c_int64(keys.shape[0]),
                                _p(_c(dims, np.int64)), ctypes.c_int(D), _p(_c(qkeys, np.int64)),
                                _p(_c(qvals, np.float32)), ctypes.c_int64(M), _p(_c(qmin, np.int32)),
                                _p(_c(qmax, np.int32)), _p(out))
    return out


def points_in_radius(keys, values, rev, dims, qkeys, qvals, qmin, qmax, radius, num_ref):
    """points_in_radius_gpu (:563-605) -> visited i64[num_ref] in {0,1}."""
    D = qvals.shape[1]
    M = qkeys.shape[0]
    visited = np.zeros(num_ref, np.int64)
    lib().oracle_points_in_radius(_p(keys), _p(values), _p(rev), ctypes.c_int64(keys.shape[0]),
                                  _p(_c(dims, np.int64)), ctypes.c_int(D), _p(_c(qkeys, np.int64)),
                                  _p(_c(qvals, np.float32)), ctypes.c_int64(M), _p(_c(qmin, np.int32)),
                                  _p(_c(qmax, np.int32)), ctypes.c_float(float(radius)), _p(visited))
    return visited


def new_table(H, D):
    """Table allocation of graph_utils.py:181-183 (keys = -1)."""
    return (np.full(H, -1, np.int64), np.zeros((H, D), np.float32), np.zeros(H, np.int64))


def radius_graph_build(ref, query, radius, max_nbr=32, sort_by_dist=False, qmin=None, qmax=None,
                       return_dists=False):
    """RadiusGraph.build_graph (graph_utils.py:149-209) -> (e_ref i64[E], e_query i64[E])."""
    ref = _c(ref, np.float32)
    query = _c(query, np.float32)
    D = ref.shape[1]
    if qmin is None:
        qmin = [0] + [-1] * (D - 1)  # :143
    if qmax is None:
        qmax = [0] + [1] * (D - 1)  # :144
    cr, cq, dims, rq, H = radius_graph_keys(ref, query, radius)
    keys, values, rev = new_table(H, D)
    hash_insert(keys, values, rev, dims, cr, ref)
    out = radius_graph(keys, values, rev, dims, cq, query, qmin, qmax, rq, max_nbr, sort_by_dist,
                       return_dists=return_dists)
    if return_dists:
        edges, d2 = out
        return edges[:, 0].copy(), edges[:, 1].copy(), d2
    return out[:, 0].copy(), out[:, 1].copy()


def points_in_boxes(points, boxes):
    """points_in_boxes_cpu (roiaware_pool3d.cpp:143-168) -> i32[B, N]."""
    points = _c(points, np.float32)
    boxes = _c(boxes, np.float32)
    out = np.zeros((boxes.shape[0], points.shape[0]), np.int32)
    lib().oracle_points_in_boxes(_p(boxes), ctypes.c_int64(boxes.shape[0]), _p(points),
                                 ctypes.c_int64(points.shape[0]), _p(out))
    return out


# --------------------------------------------------------------------------------------------
# torch_cluster.grid_cluster / GridSampling3D -- pcdet/models/model_utils/grid_sampling.py:22-46
# (third-party torch_cluster is NOT vendored in /root/reference and no version is pinned; the
#  arithmetic below restates upstream pytorch_cluster csrc/cuda/grid_cuda.cu: per-dimension
#  c_d = int64((pos_d - start_d) / size_d) [fp32 division, truncation], key = sum c_d * k_d with
#  k_0 = 1, k_{d+1} = k_d * (int64((end_d - start_d) / size_d) + 1).)
# --------------------------------------------------------------------------------------------
def grid_cluster(pos, size, start, end):
    pos = _c(pos, np.float32)
    size = _c(size, np.float32)
    start = _c(start, np.float32)
    end = _c(end, np.float32)
    c = ((pos - start) / size).astype(np.int64)  # trunc toward zero (values are >= 0)
    nvox = ((end - start) / size).astype(np.int64) + 1
    k = np.ones(pos.shape[1], np.int64)
    for d in range(1, pos.shape[1]):
        k[d] = k[d - 1] * nvox[d - 1]
    return (c * k).sum(1)


def scatter(src, index, dim_size, reduce):
    """torch_scatter.scatter(dim=0) semantics relied on: empty groups -> 0; mean = sum/max(cnt,1)."""
    src = np.asarray(src)
    index = np.asarray(index, np.int64)
    shape = (int(dim_size),) + src.shape[1:]
    if reduce in ("sum", "mean"):
        out = np.zeros(shape, src.dtype)
        np.add.at(out, index, src)
        if reduce == "mean":
            cnt = np.bincount(index, minlength=dim_size).astype(src.dtype if src.dtype.kind == "f" else np.int64)
            cnt = np.maximum(cnt, 1).reshape((-1,) + (1,) * (src.ndim - 1))
            out = out / cnt if src.dtype.kind == "f" else out // cnt
        return out.astype(src.dtype)
    if reduce in ("max", "min"):
        big = np.finfo(src.dtype).max if src.dtype.kind == "f" else np.iinfo(src.dtype).max
        init = -big if reduce == "max" else big
        out = np.full(shape, init, src.dtype)
        (np.maximum if reduce == "max" else np.minimum).at(out, index, src)
        has = np.bincount(index, minlength=dim_size) > 0
        out[~has] = 0
        return out
    raise ValueError(reduce)


def grid_sampling(points, grid_size):
    """GridSampling3D.forward(points, return_inverse=True) -> (sampled f32[V,4], inv i64[N]).

    Float sums: the reference's scatter-mean accumulates with atomics in fp32 (order undefined); the
    oracle accumulates in index order in fp32 -- comparisons use a small tolerance on the means.
    """
    points = _c(points, np.float32)
    size = np.array([1.0] + list(grid_size), np.float32)  # :16-18
    start = points.min(0).copy()
    start[0] -= np.float32(0.5)  # :32
    end = points.max(0).copy()
    end[0] += np.float32(0.5)  # :34
    cluster = grid_cluster(points, size, start, end)  # :37
    uniq, inv = np.unique(cluster, return_inverse=True)  # :38 sorted=True
    inv = inv.reshape(-1).astype(np.int64)
    sampled = scatter(points, inv, uniq.shape[0], "mean")  # :42
    return sampled, inv


def subsample_pick(points, grid_size=(0.08, 0.08, 0.08)):
    """simple_reg.py:119-124: index of the last (highest-index) point of every cell, cells by ascending key."""
    _, inv = grid_sampling(points, grid_size)
    n = int(inv.max()) + 1
    return scatter(np.arange(points.shape[0], dtype=np.int64), inv, n, "max")


# --------------------------------------------------------------------------------------------
# connected components -- graph_utils.py:40-53 (scipy IS the reference implementation here)
# --------------------------------------------------------------------------------------------
def connected_components(e0, e1, num_nodes):
    import scipy.sparse as sp
    e0 = np.asarray(e0, np.int64)
    e1 = np.asarray(e1, np.int64)
    adj = sp.coo_matrix((np.ones(e0.shape[0]), (e0, e1)), shape=(num_nodes, num_nodes))  # to_scipy_sparse_matrix
    n, lab = sp.csgraph.connected_components(adj)
    return int(n), lab.astype(np.int64)


def connected_components_c(e0, e1, num_nodes):
    """Union-find restatement (oracle.c) with scipy's numbering; cross-checked against scipy in tests."""
    e0 = _c(e0, np.int64)
    e1 = _c(e1, np.int64)
    lab = np.zeros(num_nodes, np.int64)
    n = lib().oracle_connected_components(_p(e0), _p(e1), ctypes.c_int64(e0.shape[0]), ctypes.c_int64(num_nodes),
                                          _p(lab))
    return int(n), lab


def propose_clusters(fxyz, radius, chunk=10, max_nbr=32):
    """ClusterProposal.propose_cluster for one radius (cluster_proposal.py:59-83) -> component i64[N]."""
    fxyz = _c(fxyz, np.float32)
    frame = np.rint(fxyz[:, 0]).astype(np.int64)
    num_frames = int(frame.max()) + 1
    comp = np.zeros(fxyz.shape[0], np.int64)
    total = 0
    for f0 in range(0, num_frames, chunk):  # :63
        mask = (frame >= f0) & (frame < f0 + chunk)
        if not mask.any():
            continue
        pts = fxyz[mask]
        e0, e1 = radius_graph_build(pts, pts, radius, max_nbr, True)  # :72
        n, lab = connected_components_c(e0, e1, pts.shape[0])  # :75-77
        comp[mask] = lab + total  # :80
        total += n
    return comp, total
