"""CPU stand-in for torch_cluster (not installed here).  grid_cluster restates upstream
pytorch_cluster csrc/cpu/grid_cpu.cpp; knn is brute force."""
import torch


def grid_cluster(pos, size, start=None, end=None):
    pos = pos.reshape(pos.shape[0], -1)
    size = size.to(pos)
    start = pos.min(0)[0] if start is None else start.to(pos)
    end = pos.max(0)[0] if end is None else end.to(pos)
    c = ((pos - start) / size).long()
    num = ((end - start) / size).long() + 1
    k = torch.ones(pos.shape[1], dtype=torch.long)
    for d in range(1, pos.shape[1]):
        k[d] = k[d - 1] * num[d - 1]
    return (c * k.to(c.device)).sum(1)


def knn(x, y, k, batch_x=None, batch_y=None):
    """Rows (y_idx, x_idx), k nearest x for every y, ascending distance (self included when x is y)."""
    d = torch.cdist(y.double(), x.double())
    if batch_x is not None and batch_y is not None:
        d = d + (batch_y[:, None] != batch_x[None, :]).double() * 1e30
    kk = min(k, x.shape[0])
    idx = d.topk(kk, dim=1, largest=False).indices
    row = torch.arange(y.shape[0])[:, None].expand(-1, kk)
    return torch.stack([row.reshape(-1), idx.reshape(-1)], 0)
