"""CPU stand-in for torch_geometric.utils.to_scipy_sparse_matrix (not installed here)."""
import numpy as np
import scipy.sparse as sp


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    row, col = edge_index.detach().cpu().numpy()
    if num_nodes is None:
        num_nodes = int(max(row.max(), col.max())) + 1
    return sp.coo_matrix((np.ones(row.shape[0]), (row, col)), (num_nodes, num_nodes))
