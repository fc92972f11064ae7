"""GT-evaluation helpers: point-in-rotated-box membership on the device.

Mirror of points_in_boxes_cpu (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:128-168; z half extent exact,
x/y with a 1 cm margin, rotation by -heading) as a dense [B, N] torch expression instead of an O(B*N) single
thread CPU loop with device<->host copies.  Evaluation only -- not part of the cluster algorithm."""
import torch


def points_in_boxes(points_xyz, boxes):
    """points_xyz f32[N,3], boxes f32[B,7] (x,y,z,dx,dy,dz,heading) -> int64[B,N] in {0,1}."""
    p = points_xyz.float()
    b = boxes.float()
    dz_ok = (p[None, :, 2] - b[:, None, 2]).abs() <= (b[:, None, 5] / 2.0)
    cosa, sina = torch.cos(-b[:, 6])[:, None], torch.sin(-b[:, 6])[:, None]
    sx = p[None, :, 0] - b[:, None, 0]
    sy = p[None, :, 1] - b[:, None, 1]
    lx = sx * cosa + sy * (-sina)
    ly = sx * sina + sy * cosa
    inside = dz_ok & (lx.abs() < b[:, None, 3] / 2.0 + 1e-2) & (ly.abs() < b[:, None, 4] / 2.0 + 1e-2)
    return inside.long()
