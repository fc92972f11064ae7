"""Module-style API of the reference's torch_hash op (mirror of pcdet/ops/torch_hash/torch_hash_modules.py:10-126)
on the new kernels: ``RadiusGraph(max_num_points, ndim)`` and ``ChamferDistance``.

The four raw functions of ``torch_hash_cuda`` that operate on caller-allocated multimap buffers
(``hash_insert_gpu`` / ``radius_graph_gpu`` / ``correspondence`` / ``points_in_radius_gpu``,
torch_hash_api.cpp:9-15) are not re-exported: the table layout behind them is gone (unique-cell table + cell-sorted
points); their only caller on the cluster-tracking path, ``graph_utils.RadiusGraph``, is served by
``pcseqlearning_b200.graph_utils``.  ``correspondence`` / ``points_in_radius`` are provided as functions on points.
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


def correspondence(ref, query, radius):
    """Nearest reference point of every query within the 27 cells of size `radius` (-1 when the cells are empty);
    the reference's `correspondence` (torch_hash_kernel.cu:96-155) has no radius test, so pass the cell size and
    accept matches up to the cell diagonal."""
    ref_p = ops._as_points(ref)
    query_p = ops._as_points(query)
    grid = ops.CellGrid(ref_p, ops.radius_voxel_size(radius), bounds_sets=[ref_p, query_p])
    idx, cnt, _ = grid.search(query_p, 1, float(radius) * 3.5)
    out = idx[:, 0].long()
    out[cnt == 0] = -1
    return out
