"""Per-cluster registration helpers (mirror of pcdet/models/registration/preprocessors/registration_utils.py).

`register_to_next_frame` keeps the reference's signature and return values but runs the whole ICP loop in one
persistent kernel (ops.register_icp); the small segmented reductions used by the tracking state machine are torch
index ops on a few thousand components."""
import torch

from .. import _lib, ops
from ..utils.scatter import scatter_count, scatter_mean, scatter_sum


def robust_mean(data, index, max_index):
    """Per-group mean, groups without members stay 0 (registration_utils.py:12-23)."""
    return scatter_mean(data, index, max_index)


def efficient_robust_mean(data, index, max_index, sort_index=None):
    """registration_utils.py:25-34 (the reference sorts first to use segment_coo; the result is the same mean)."""
    return scatter_mean(data, index, max_index)


def efficient_robust_sum(data, index, max_index, sort_index=None):
    """registration_utils.py:36-45."""
    return scatter_sum(data, index, max_index)


def truncated_robust_mean(data, index, max_index, trunc_dist=0.3):
    """Mean of the values clamped to +-trunc_dist around the plain group mean (registration_utils.py:44-58)."""
    mean = scatter_mean(data, index, max_index)
    clamped = torch.minimum(torch.maximum(data, mean[index] - trunc_dist), mean[index] + trunc_dist)
    return scatter_mean(clamped, index, max_index)


def robust_median(data, index, max_index):
    """Upper median per group, empty groups -> -1e10 (registration_utils.py:60-81)."""
    return ops.group_median(data, index, int(max_index))


def register_to_next_frame(graph, moving, ref, num_components, angle_regularizer=10, max_iter=20,
                           stopping_delta=5e-2, frame_offset=None):
    """Trimmed two-way ICP of every component of `moving` onto `ref` (registration_utils.py:83-206).

    moving: dict with component [vm], frame, fxyz [vm,4], stationary [vm]; ref: frame, fxyz, stationary.
    graph: the level's RadiusGraph (only its radius is used; qmin/qmax are left untouched).
    Returns (moving with fxyz moved, T f64[C,4,4], l1_component_error f64[C], comp_edge_ratio f32[C]).
    """
    if int(getattr(graph, "max_num_neighbors", 1)) != 1:
        raise _lib.PcsError("register_to_next_frame: the registration graph must have MAX_NUM_NEIGHBORS == 1 "
                            "(the ICP kernel keeps the single nearest neighbour, like the reference's config)")
    if frame_offset is None:
        frame_offset = int((ref.frame.reshape(-1)[0] - moving.frame.reshape(-1)[0]).long().item())
    moved, T, l1, ratio, info = ops.register_icp(
        moving.fxyz, moving.component, moving.stationary, ref.fxyz, ref.stationary, num_components,
        float(graph.radius), frame_offset, angle_regularizer=float(angle_regularizer), max_iter=int(max_iter),
        stopping_delta=float(stopping_delta))
    if moving.fxyz.shape == moved.shape and moving.fxyz.dtype == moved.dtype:
        moving.fxyz.copy_(moved)  # the reference updates the tensor in place (:179): holders of it see the move
    else:
        moving.fxyz = moved
    moving["icp_info"] = info
    return moving, T, l1, ratio
