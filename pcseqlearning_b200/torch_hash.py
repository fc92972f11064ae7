"""Module-style API of the reference's torch_hash op (mirror of pcdet/ops/torch_hash/torch_hash_modules.py:10-126)
on the new kernels: ``RadiusGraph(max_num_points, ndim)`` and ``ChamferDistance``.

The four raw functions of the native module (``hash_insert_gpu`` / ``radius_graph_gpu`` / ``correspondence`` /
``points_in_radius_gpu``, torch_hash_api.cpp:9-15) live in ``pcseqlearning_b200.torch_hash_cuda`` with the reference's
signatures; ``correspondence`` / ``points_in_radius`` below are the point-level conveniences of
``HashTable.find_corres`` / ``points_in_radius_step2`` (torch_hash_utils.py:32-75, 116-150) on top of them.
"""
import torch
from torch import nn

from . import ops


class RadiusGraph(nn.Module):
    """``forward(ref, query, radius, num_neighbors, sort_by_dist) -> edges int64[2, E]`` rows (idx_of_ref,
    idx_of_query); points are [N, 1 + ndim] with the batch index first (torch_hash_modules.py:10-90)."""

    def __init__(self, max_num_points=400000, ndim=3):
        super().__init__()
        self.ndim = ndim
        self.max_num_points = max_num_points
        self.qmin = torch.tensor([0] + [-1] * ndim, dtype=torch.int32)
        self.qmax = torch.tensor([0] + [1] * ndim, dtype=torch.int32)

    @torch.no_grad()
    def forward(self, ref, query, radius, num_neighbors, sort_by_dist=False):
        assert ref.shape[1] == self.ndim + 1, f"points must have {self.ndim + 1} dimensions"
        if isinstance(radius, (int, float)):
            radius = float(radius)
        qmin = self.qmin.tolist() + [0] * (3 - self.ndim)
        qmax = self.qmax.tolist() + [0] * (3 - self.ndim)
        e_ref, e_query = ops.radius_graph(ref, ref if query is ref else query, radius, int(num_neighbors),
                                          bool(sort_by_dist), qmin=qmin, qmax=qmax)
        return torch.stack([e_ref, e_query], dim=0)

    def extra_repr(self):
        return f"ndim={self.ndim}"


class ChamferDistance(nn.Module):
    """Two-way nearest-neighbour squared distance within `radius` (torch_hash_modules.py:96-126)."""

    def __init__(self, max_num_points=400000, ndim=3, radius_graph=None):
        super().__init__()
        self.radius_graph = radius_graph if radius_graph is not None else RadiusGraph(max_num_points, ndim=ndim)
        self.ndim = self.radius_graph.ndim
        self.max_num_points = self.radius_graph.max_num_points

    def forward(self, src_bxyz, target_bxyz, radius):
        fwd_src, fwd_target = self.radius_graph(src_bxyz, target_bxyz, radius, 1, sort_by_dist=True)
        bwd_target, bwd_src = self.radius_graph(target_bxyz, src_bxyz, radius, 1, sort_by_dist=True)
        dist_fwd = (src_bxyz[fwd_src] - target_bxyz[fwd_target]).square().sum(-1).mean()
        dist_bwd = (src_bxyz[bwd_src] - target_bxyz[bwd_target]).square().sum(-1).mean()
        return dist_fwd + dist_bwd


def _voxelization(ref_p, query_p, voxel_size):
    """graph_utils.py:170-176 geometry: coordinates of both sets and dims for a [1 - 1e-3, v, v, v] grid."""
    vs = torch.tensor(ops.radius_voxel_size(voxel_size), device=ref_p.device)
    allp = torch.cat([ref_p, query_p], 0)
    lo = allp.min(0)[0] - vs * 2
    hi = allp.max(0)[0] + vs * 2
    coord = lambda x: torch.round((x - lo) / vs).long() + 1
    return coord(ref_p), coord(query_p), torch.round((hi - lo) / vs).long() + 3


def _table(ref_p, coords, dims):
    from . import torch_hash_cuda as op
    H = max(int(ref_p.shape[0] / 0.5), 8)
    keys = torch.full((H,), -1, dtype=torch.int64, device=ref_p.device)
    values = torch.empty(H, 4, dtype=torch.float32, device=ref_p.device)
    rev = torch.zeros(H, dtype=torch.int64, device=ref_p.device)
    op.hash_insert_gpu(keys, values, rev, dims, coords, ref_p)
    return keys, values, rev


def correspondence(ref, query, voxel_size):
    """Nearest reference point of every query over the 27 cells (size `voxel_size`) around it, with NO radius test
    (torch_hash_kernel.cu:96-155): -1 only when all 27 cells are empty."""
    from . import torch_hash_cuda as op
    ref_p, query_p = ops._as_points(ref), ops._as_points(query)
    cr, cq, dims = _voxelization(ref_p, query_p, voxel_size)
    keys, values, rev = _table(ref_p, cr, dims)
    qmin = torch.tensor([0, -1, -1, -1], dtype=torch.int32, device=ref_p.device)
    out = torch.empty(query_p.shape[0], dtype=torch.int64, device=ref_p.device)
    op.correspondence(keys, values, rev, dims, cq, query_p, qmin, -qmin, out)
    return out


def points_in_radius(ref, query, radius):
    """bool[N_ref]: reference points with a query strictly closer than `radius` (torch_hash_kernel.cu:160-222)."""
    from . import torch_hash_cuda as op
    ref_p, query_p = ops._as_points(ref), ops._as_points(query)
    cr, cq, dims = _voxelization(ref_p, query_p, radius)
    keys, values, rev = _table(ref_p, cr, dims)
    qmin = torch.tensor([0, -1, -1, -1], dtype=torch.int32, device=ref_p.device)
    visited = torch.zeros(ref_p.shape[0], dtype=torch.int64, device=ref_p.device)
    op.points_in_radius_gpu(keys, values, rev, dims, cq, query_p, qmin, -qmin, float(radius), visited)
    return visited.bool()
