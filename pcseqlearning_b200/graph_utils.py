"""Graph plugin layer (mirror of pcdet/models/model_utils/graph_utils.py for the registration path).

``build_graph(cfg, runtime_cfg)`` -> ``GRAPHS[cfg['TYPE']]``; ``graph(ref_dict, query_dict)`` ->
``(e_ref, e_query, None)``.  Callers mutate ``graph.radius``, ``graph.qmin[0]`` and ``graph.qmax[0]``
between calls (registration_utils.py:107-112,131-137), so those stay plain live attributes; they are
host tensors here, which removes the device round trip the reference pays to read them.
"""
import torch
from torch import nn

from . import ops


def build_graph(graph_cfg, runtime_cfg=None):
    graph = GRAPHS[graph_cfg["TYPE"]]
    return graph(model_cfg=graph_cfg, runtime_cfg=runtime_cfg)


def connected_components(edges, num_nodes=None):
    """Weakly connected components of a directed edge list [2, E] (graph_utils.py:40-53).

    Returns (num_components, component[N]); components are numbered by ascending smallest member index,
    which is the numbering scipy.sparse.csgraph.connected_components produces.
    """
    if num_nodes is None:
        num_nodes = int(edges.max().long().item()) + 1
    n_comp, labels = ops.connected_components(edges, int(num_nodes))
    return int(n_comp.sum().item()), labels.to(edges.dtype)


class RadiusGraph(nn.Module):
    """Radius graph with at most MAX_NUM_NEIGHBORS nearest reference points per query
    (graph_utils.py:131-212), on the voxel-hash + warp-per-query search kernels."""

    def __init__(self, runtime_cfg, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.radius = model_cfg.get("RADIUS", None)
        self.max_num_neighbors = model_cfg.get("MAX_NUM_NEIGHBORS", 32)
        self.sort_by_dist = model_cfg.get("SORT_BY_DIST", False)
        self.util_ratio = 0.5
        self.relative_key = model_cfg.get("RELATIVE_KEY", "bxyz")
        if model_cfg.get("DYNAMIC_RADIUS", False):
            raise NotImplementedError("DYNAMIC_RADIUS is not used on the cluster-tracking path")
        # live, caller-mutable query ranges (host tensors)
        self.qmin = torch.tensor([0, -1, -1, -1], dtype=torch.int32)
        self.qmax = torch.tensor([0, 1, 1, 1], dtype=torch.int32)

    def forward(self, ref, query):
        return self.build_graph(ref, query)

    def build_graph(self, ref, query):
        ref_pts = ref[self.relative_key]
        query_pts = query[self.relative_key]
        assert ref_pts.shape[-1] == 4
        same = (ref_pts.data_ptr() == query_pts.data_ptr()) and (ref_pts.shape == query_pts.shape)
        e_ref, e_query = ops.radius_graph(ref_pts, ref_pts if same else query_pts, float(self.radius),
                                          self.max_num_neighbors, self.sort_by_dist,
                                          qmin=self.qmin.tolist(), qmax=self.qmax.tolist())
        return e_ref, e_query, None

    def neighbor_lists(self, ref, query, want_d2=False):
        """Padded K-nearest lists (nbr_idx i32[M,K], nbr_cnt i32[M], d2) without building the int64 edge list."""
        ref_pts = ops._as_points(ref[self.relative_key])
        query_pts = ops._as_points(query[self.relative_key])
        grid = ops.CellGrid(ref_pts, ops.radius_voxel_size(float(self.radius)), bounds_sets=[ref_pts, query_pts])
        return grid.search(query_pts, int(self.max_num_neighbors), float(self.radius), self.qmin.tolist(),
                           self.qmax.tolist(), want_d2=want_d2)

    def extra_repr(self):
        return f"radius={self.radius}, max_ngbrs={self.max_num_neighbors}, sort={self.sort_by_dist}"


GRAPHS = dict(RadiusGraph=RadiusGraph)
