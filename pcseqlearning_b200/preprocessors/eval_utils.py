"""GT-evaluation helpers on the device.

`points_in_boxes` is the dense mirror of points_in_boxes_cpu (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:128-168;
z half extent exact, x/y with a 1 cm margin, rotation by -heading) as a [B, N] torch expression (small inputs, tests).
`FrameBoxes` is what the preprocessors use: the GT boxes of a whole sequence sorted by frame, tested by the
pcs_points_in_boxes kernel (one thread per point against the boxes of its own frame) together with the
per-(component, box) membership counts of assign_instances_to_boxes (cluster_proposal.py:90-114).
Evaluation only -- not part of the cluster algorithm."""
import ctypes

import torch

from .. import _lib
from ..ops import _ptr, _stream


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


class FrameBoxes:
    """GT boxes of a sequence, grouped by frame (stable: the frame-local index of a box is its rank among the boxes
    of its frame in the original order, i.e. the row of the reference's `filter_dict(seq_boxes, frame_box_mask)`)."""

    def __init__(self, attr, frame, num_frames):
        dev = attr.device
        self.B, self.F = int(attr.shape[0]), int(num_frames)
        frame = frame.reshape(-1).long()
        self.order = torch.argsort(frame, stable=True)  # sorted position -> original box row
        cnt = torch.bincount(frame, minlength=self.F)
        off = torch.zeros(self.F + 1, dtype=torch.int64, device=dev)
        off[1:] = cnt.cumsum(0)
        self.off = off
        self.off32 = off.int().contiguous()
        self.Bmax = max(int(cnt.max().item()) if self.B else 1, 1)  # host sync (table widths)
        self.frame_sorted = frame[self.order]
        self.local = torch.arange(self.B, device=dev) - off[self.frame_sorted]  # frame-local index of sorted box
        self.recs = torch.empty(max(self.B, 1) * 6, dtype=torch.float64, device=dev)  # 48 bytes per box
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pcs_box_prep(_stream(), _ptr(attr[self.order].float().contiguous()), self.B,
                                               _ptr(self.recs)), "pcs_box_prep")

    def query(self, pts, sel=None, cids=(), want_first=True):
        """pts f32[N,4] (frame, x, y, z); item i = pts[sel[i]] (sel int32, optional).  cids: up to three int64[n]
        component ids; returns (first int32[n] frame-local box index or -1, [counts int32[C_k, Bmax] per cid])."""
        dev = pts.device
        n = int(sel.shape[0]) if sel is not None else int(pts.shape[0])
        first = torch.empty(n, dtype=torch.int32, device=dev) if want_first else None
        cid_t, cnt_t = [], []
        for c, size in cids:
            cid_t.append(c.long().contiguous())
            cnt_t.append(torch.zeros(int(size), self.Bmax, dtype=torch.int32, device=dev))
        pad = [None] * (3 - len(cid_t))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pcs_points_in_boxes(
                _stream(), _ptr(pts), _ptr(sel), n, _ptr(self.recs), _ptr(self.off32), self.F, self.Bmax,
                *[_ptr(x) for x in cid_t + pad], *[_ptr(x) for x in cnt_t + pad], _ptr(first), _ptr(self.err)),
                "pcs_points_in_boxes")
        return first, cnt_t
