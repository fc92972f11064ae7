"""GridSampling3D on the voxelize kernels (mirror of pcdet/models/model_utils/grid_sampling.py)."""
import torch
from torch import nn

from . import ops


class GridSampling3D(nn.Module):
    """Voxel-grid mean of (frame, x, y, z) rows; voxels numbered by ascending cell key.

    Same constructor and call convention as the reference class (grid_sampling.py:7-46):
    ``GridSampling3D(grid_size)(points, return_inverse=False)``.
    """

    def __init__(self, grid_size):
        super().__init__()
        self._grid_size = grid_size
        if isinstance(grid_size, (list, tuple)):
            size = torch.tensor([1] + list(grid_size)).float()
        else:
            size = torch.tensor([1] + [grid_size for _ in range(3)]).float()
        assert size.shape[0] == 4, "Expecting 4D grid size."
        self.register_buffer("grid_size", size)
        self._size_list = [float(v) for v in size[1:].tolist()]

    def forward(self, points, return_inverse=False):
        res = ops.voxelize(points, self._size_list, want_mean=True)
        if return_inverse:
            return res["sampled"], res["inv"]
        return res["sampled"]

    def voxelize(self, points, **kw):
        """Full result dict of ops.voxelize (inv, num, sampled, maxidx, counts) for callers on the fast path."""
        return ops.voxelize(points, self._size_list, **kw)

    def extra_repr(self):
        return "grid size {}".format(self._grid_size)
