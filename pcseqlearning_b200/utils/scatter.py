"""Segmented reductions on torch tensors (the torch_scatter conventions the reference relies on, SURVEY.md
A.5): reduction over dim 0, empty groups -> 0, mean = sum / max(count, 1).  Small host-side plumbing for the
pillar / component bookkeeping; the per-point heavy lifting is done by the CUDA kernels."""
import torch


def _index(index, src):
    index = index.long().reshape(-1)
    if src.dim() == 1:
        return index
    shape = [-1] + [1] * (src.dim() - 1)
    return index.reshape(shape).expand_as(src)


def scatter_sum(src, index, dim_size):
    out = torch.zeros((int(dim_size),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index.long().reshape(-1), src)


def scatter_count(index, dim_size, dtype=torch.float32):
    return torch.bincount(index.long().reshape(-1), minlength=int(dim_size)).to(dtype)


def scatter_mean(src, index, dim_size):
    s = scatter_sum(src, index, dim_size)
    cnt = scatter_count(index, dim_size, src.dtype).clamp(min=1)
    return s / cnt.reshape([-1] + [1] * (src.dim() - 1))


def scatter_min(src, index, dim_size):
    out = torch.zeros((int(dim_size),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(0, _index(index, src), src, "amin", include_self=False)


def scatter_max(src, index, dim_size):
    out = torch.zeros((int(dim_size),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(0, _index(index, src), src, "amax", include_self=False)


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """torch_scatter.scatter-compatible entry (dim 0 only)."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    return dict(sum=scatter_sum, add=scatter_sum, mean=scatter_mean, min=scatter_min, max=scatter_max)[reduce](
        src, index, dim_size)
