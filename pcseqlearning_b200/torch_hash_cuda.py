"""Drop-in for the reference's native module ``pcdet.ops.torch_hash.torch_hash_cuda``
(pcdet/ops/torch_hash/src/torch_hash.h:16-32, torch_hash_api.cpp:9-15): the same four names, argument lists and result
conventions, on the sm_100a kernels of csrc/compat.cu.

    hash_insert_gpu(keys, values, reverse_indices, dims, insert_keys, insert_values) -> None
    radius_graph_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius,
                     max_num_neighbors, sort_by_dist) -> edges int64[E, 2]  rows (ref index, query index)
    correspondence(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, corres_indices) -> None
    points_in_radius_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius,
                         visited) -> None

The caller still allocates the reference's table buffers (`keys` pre-filled with -1, `values`, `reverse_indices`;
graph_utils.py:179-183) and passes them to every call.  They are opaque scratch here: hash_insert_gpu copies the points
into `values[:N]` and the cell-grouped point indices into `reverse_indices[:N]`, and keeps its unique-cell table (one
16-byte slot per occupied cell -- it does not fit the reference's 8-byte key array in the worst case) in a side buffer
associated with the `keys` tensor.  Differences, all on the side of determinism: ties at equal distance are resolved by
ascending point index (a race in the reference); errors raise PcsError instead of exit(-1); max_num_neighbors = -1
returns every neighbour within the radius (the reference's fill kernel writes nothing for -1)."""
import ctypes

import torch

from . import _lib
from .ops import _ptr, _stream, exclusive_scan


def _host_i64(t):
    v = [int(x) for x in t.tolist()]
    return (ctypes.c_int64 * len(v))(*v)


def _host_i32(t):
    v = [int(x) for x in t.tolist()]
    return (ctypes.c_int * len(v))(*v)


def _check(*tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.PcsError("torch_hash_cuda: all tensors must be CUDA tensors (no CPU path exists)")
        if not t.is_contiguous():
            raise _lib.PcsError("torch_hash_cuda: tensors must be contiguous")


_TABLES = {}  # keys.data_ptr() -> dict(table, H, n, nd, shape): the cell table built by hash_insert_gpu for that buffer


def hash_insert_gpu(keys, values, reverse_indices, dims, insert_keys, insert_values):
    _check(keys, values, reverse_indices, dims, insert_keys, insert_values)
    n, nd = insert_keys.shape
    if keys.dtype != torch.int64 or reverse_indices.dtype != torch.int64:
        raise _lib.PcsError("torch_hash_cuda.hash_insert_gpu: keys / reverse_indices must be int64")
    if values.shape[0] < n or reverse_indices.shape[0] < n or values.shape[1] != nd:
        raise _lib.PcsError("torch_hash_cuda.hash_insert_gpu: table smaller than the number of inserted points")
    H = 1 << max(10, int(2 * n - 1).bit_length())
    table = torch.empty(H, 4, dtype=torch.int32, device=keys.device)
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=keys.device)
    ctr = torch.zeros(4, dtype=torch.int32, device=keys.device)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().pcs_compat_hash_insert(_stream(), _ptr(insert_keys.long().contiguous()), n, nd,
                                                     _host_i64(dims), _ptr(table), H, _ptr(rows), _ptr(ctr)),
                   "pcs_compat_hash_insert")
    values[:n] = insert_values.to(values.dtype)
    reverse_indices[:n] = rows[:n].long()
    if len(_TABLES) > 64:  # tables of buffers that went away
        _TABLES.clear()
    _TABLES[keys.data_ptr()] = dict(table=table, H=H, n=n, nd=nd, shape=tuple(keys.shape), rows=rows)


def _state(keys, values, reverse_indices):
    st = _TABLES.get(keys.data_ptr())
    if st is None or st["shape"] != tuple(keys.shape):
        raise _lib.PcsError("torch_hash_cuda: this `keys` buffer was not filled by hash_insert_gpu")
    n = st["n"]
    return st["table"], st["H"], st["rows"], values[:max(n, 1)].float().contiguous()


def radius_graph_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius,
                     max_num_neighbors, sort_by_dist):
    _check(keys, values, reverse_indices, dims, query_keys, query_values, radius)
    table, slots, rows, vals = _state(keys, values, reverse_indices)
    m, nd = query_keys.shape
    dev = keys.device
    L = _lib.lib()
    dims_h, qmin_h, qmax_h = _host_i64(dims), _host_i32(qmin), _host_i32(qmax)
    qk, qv = query_keys.long().contiguous(), query_values.float().contiguous()
    rad = radius.float().contiguous()
    degree = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        s = _stream()
        _lib.check(L.pcs_compat_radius_degree(s, _ptr(table), slots, _ptr(rows), _ptr(vals), nd, dims_h, _ptr(qk), _ptr(qv),
                                              m, qmin_h, qmax_h, _ptr(rad), int(max_num_neighbors), _ptr(degree)),
                   "pcs_compat_radius_degree")
        offsets = exclusive_scan(degree[:m])
        E = int(offsets[-1].item())
        edges = torch.empty(E, 2, dtype=torch.int64, device=dev)
        dists = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
        _lib.check(L.pcs_compat_radius_fill(s, _ptr(table), slots, _ptr(rows), _ptr(vals), nd, dims_h, _ptr(qk), _ptr(qv), m,
                                            qmin_h, qmax_h, _ptr(rad), _ptr(degree), _ptr(offsets), _ptr(edges),
                                            _ptr(dists)), "pcs_compat_radius_fill")
    return edges


def correspondence(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, corres_indices):
    _check(keys, values, reverse_indices, dims, query_keys, query_values, corres_indices)
    table, slots, rows, vals = _state(keys, values, reverse_indices)
    m, nd = query_keys.shape
    out = torch.empty(max(m, 1), dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().pcs_nn_correspondence(
            _stream(), _ptr(table), slots, _ptr(rows), _ptr(vals), nd, _host_i64(dims), _ptr(query_keys.long().contiguous()),
            _ptr(query_values.float().contiguous()), m, _host_i32(qmin), _host_i32(qmax), _ptr(out)),
            "pcs_nn_correspondence")
    corres_indices[:m] = out[:m].to(corres_indices.dtype)


def points_in_radius_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius, visited):
    _check(keys, values, reverse_indices, dims, query_keys, query_values, visited)
    table, slots, rows, vals = _state(keys, values, reverse_indices)
    m, nd = query_keys.shape
    vis = visited if visited.dtype == torch.int64 else visited.long()
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().pcs_points_in_radius(
            _stream(), _ptr(table), slots, _ptr(rows), _ptr(vals), nd, _host_i64(dims), _ptr(query_keys.long().contiguous()),
            _ptr(query_values.float().contiguous()), m, _host_i32(qmin), _host_i32(qmax), float(radius), _ptr(vis)),
            "pcs_points_in_radius")
    if vis is not visited:
        visited.copy_(vis.to(visited.dtype))
