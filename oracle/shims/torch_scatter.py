"""CPU stand-in for torch_scatter (not installed here).  Semantics per SURVEY.md Appendix A.5."""
import torch


def _expand_index(index, src, dim):
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.reshape(shape).expand_as(src)
    return index


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    index = index.long()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    dim_size = int(dim_size)
    idx = _expand_index(index, src, dim)
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
    if reduce == "mean":
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
        cnt = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, torch.ones_like(src))
        cnt = cnt.clamp(min=1)
        if src.is_floating_point():
            return s / cnt
        return torch.div(s, cnt, rounding_mode="floor")
    if reduce in ("max", "min"):
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return out.scatter_reduce_(dim, idx, src, "amax" if reduce == "max" else "amin", include_self=False)
    raise ValueError(reduce)


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    """index must be sorted (callers argsort first); reduction over dim 0."""
    return scatter(src, index, dim=0, dim_size=dim_size, reduce=reduce)
