"""GPU: the batched tracker (tracker.TrackBatch on csrc/track.cu) against
  (a) the golden recorded from the reference's OWN track_frame (tests/golden/tracking.npz), and
  (b) the one-pair-at-a-time mirror of the reference control flow (ClusterTracking.track_frame) on the same inputs.
"""
import os

import numpy as np
import pytest
import torch

from helpers import component_centers, transform_errors

pytestmark = pytest.mark.gpu


def _cfg():
    from pcseqlearning_b200.config import cluster_tracking_cfg
    cfg = [p for p in cluster_tracking_cfg().PREPROCESSORS if p.NAME == "ClusterTracking"][0]
    cfg.VERBOSE = False
    return cfg


def _golden_inputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "tracking.npz"))
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return g, cuda(g["points"]), cuda(g["sweep"]).reshape(-1), cuda(g["component"]), cuda(g["seg"])


def _pairs(orig, comp):
    return set((np.asarray(orig, dtype=np.int64) * 100000 + np.asarray(comp, dtype=np.int64)).tolist())


def test_batched_tracker_vs_reference_golden(golden_dir):
    from pcseqlearning_b200.tracker import TrackBatch
    g, pts, sweep, comp, seg = _golden_inputs(golden_dir)
    anchor = int(g["anchor"])
    tb = TrackBatch(pts, sweep, [comp], _cfg(), anchors=[anchor]).run()
    tb.check()
    j, ex = tb.results(seg_label=seg)[(0, anchor)]
    T = tb.transforms(j).cpu().numpy()
    Tw = g["transforms"]
    assert T.shape == Tw.shape
    got_c = set(ex.component.unique().tolist())
    want_c = set(np.unique(g["ex_component"]).tolist())
    jac_c = len(got_c & want_c) / max(len(got_c | want_c), 1)
    got_p = _pairs(ex.original_indices.cpu().numpy(), ex.component.cpu().numpy())
    want_p = _pairs(g["ex_original_indices"], g["ex_component"])
    jac_p = len(got_p & want_p) / max(len(got_p | want_p), 1)
    both = sorted(got_c & want_c)
    am = np.rint(g["points"][:, 0]) == anchor
    anchor_comp = g["component"][am] - g["component"][am].min()
    ctr = component_centers(g["points"][am], anchor_comp, T.shape[0])[both]
    ang, dt = transform_errors(T[both], Tw[both], np.repeat(ctr[:, None, :], T.shape[1], axis=1))
    dt = dt * np.maximum(1.0, np.linalg.norm(ctr, axis=-1))[:, None]
    msg = dict(jac_components=jac_c, jac_points=jac_p, ang_med=float(np.median(ang)), ang_p90=float(np.quantile(ang, 0.9)),
               dt_med=float(np.median(dt)), dt_p90=float(np.quantile(dt, 0.9)), n_both=len(both))
    print(msg)
    assert jac_c > 0.95 and jac_p > 0.97, msg
    assert np.median(ang) < 1e-3 and np.median(dt) < 2e-2, msg
    assert np.quantile(ang, 0.9) < 2e-2 and np.quantile(dt, 0.9) < 0.25, msg
    # layout of the extracted dict: anchor points first, then target frames in tracking order
    f = ex.fxyz[:, 0].round().long().cpu().numpy()
    first_other = np.argmax(f != anchor) if (f != anchor).any() else len(f)
    assert (f[:first_other] == anchor).all()
    assert torch.equal(pts[ex.original_indices], ex.fxyz)


def test_batched_tracker_vs_sequential_mirror(golden_dir):
    """Same inputs through the batched kernels and through the reference-shaped sequential control flow."""
    from pcseqlearning_b200.preprocessors.cluster_tracking import ClusterTracking, component_diameter
    from pcseqlearning_b200.tracker import TrackBatch
    from pcseqlearning_b200.utils import EasyDict, filter_dict
    g, pts, sweep, comp, seg = _golden_inputs(golden_dir)
    anchor = int(g["anchor"])
    cfg = _cfg()
    tb = TrackBatch(pts, sweep, [comp], cfg, anchors=[anchor]).run()
    j, ex = tb.results(seg_label=seg)[(0, anchor)]
    T = tb.transforms(j)

    mod = ClusterTracking(cfg, {}).cuda()
    seq_points = EasyDict(fxyz=pts.clone(), frame=sweep.reshape(-1, 1).clone(), gt_box_id=torch.zeros_like(comp) - 1,
                          segmentation_label=seg, component=comp)
    diam = component_diameter(seq_points)[seq_points.component]
    seq_points.component_diameter = diam
    seq_points.stationary = diam > 12.5
    seq_points.extracted = torch.zeros_like(seq_points.fxyz[:, 0]).bool()
    frame_mask = (seq_points.fxyz[:, 0] == anchor).reshape(-1)
    frame_points = EasyDict(filter_dict(seq_points, frame_mask))
    frame_points.component = frame_points.component - frame_points.component.min()
    ex2 = mod.track_frame(seq_points, frame_points, None)

    got = _pairs(ex.original_indices.cpu().numpy(), ex.component.cpu().numpy())
    want = _pairs(ex2.original_indices.cpu().numpy(), ex2.component.cpu().numpy())
    jac = len(got & want) / max(len(got | want), 1)
    kept = sorted(set(ex.component.unique().tolist()) & set(ex2.component.unique().tolist()))
    dT = (T[kept] - ex2.transforms[kept]).abs()
    msg = dict(jac=jac, n_kept=len(kept), dR_med=float(dT[..., :3, :3].amax((-1, -2)).median()),
               dR_max=float(dT[..., :3, :3].max()))
    print(msg)
    assert jac > 0.97, msg
    assert float(dT[..., :3, :3].amax((-1, -2)).median()) < 1e-3, msg


def test_voxel_sampler_matches_sample_frame():
    """pcs_trk_sample (batched sample_frame) against the single-cloud voxelize / group_median kernels."""
    import ctypes
    from pcseqlearning_b200 import _lib, ops
    from pcseqlearning_b200.ops import _ptr, _stream
    from pcseqlearning_b200.tracker import VoxelSampler
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = 20000
    pts = torch.rand(n, 4, generator=gen, device="cuda") * torch.tensor([0, 30, 30, 4], device="cuda")
    comp = torch.randint(0, 50, (n,), generator=gen, device="cuda").int()
    comp, o = torch.sort(comp)
    pts = pts[o].contiguous()
    stat = (torch.rand(n, generator=gen, device="cuda") < 0.3).to(torch.uint8)
    group = torch.zeros(n, dtype=torch.int32, device="cuda")
    L = _lib.lib()
    sp = VoxelSampler(n, 64, 4, pts.device)
    sb = torch.empty(6, dtype=torch.int32, device="cuda")
    out_pts = torch.empty(n, 4, device="cuda")
    out_key = torch.empty(n, dtype=torch.int32, device="cuda")
    out_group = torch.empty(n, dtype=torch.int32, device="cuda")
    vdeg = torch.zeros(64, dtype=torch.int32, device="cuda")
    size = [0.4, 0.4, 0.6]
    s = _stream()
    _lib.check(L.pcs_trk_bounds_reset(s, _ptr(sb), 1), "reset")
    _lib.check(L.pcs_trk_group_bounds(s, _ptr(pts), _ptr(group), n, _ptr(sb)), "bounds")
    st = sp.struct(pts, group, comp, stat, None, n, 1, 50, 0, size, sb, vdeg, out_pts, out_key, out_group)
    _lib.check(L.pcs_trk_sample(s, ctypes.byref(st)), "sample")
    V = int(sp.t["ctr"][1].item())
    res = ops.voxelize(pts, size, want_mean=True, want_counts=True)
    assert V == res["num"]
    med = ops.group_median(comp.long(), res["inv"], V, res["counts"])
    cnt = res["counts"].float()
    st_major = (torch.zeros(V, device="cuda").index_add_(0, res["inv"], stat.float()) / cnt) > 0.5
    # match voxels through the grid cell of their means
    start = pts[:, 1:].min(0)[0]
    sz = torch.tensor(size, device="cuda")

    def cell_key(xyz):
        c = ((xyz - start) / sz).floor().long()
        return (c[:, 0] * 100000 + c[:, 1]) * 100000 + c[:, 2]

    a = torch.cat([out_pts[:V, 1:], out_key[:V, None].float(), (out_pts[:V, 0].view(torch.int32) & 1)[:, None].float()], 1)
    b = torch.cat([res["sampled"][:, 1:], med[:, None].float(), st_major[:, None].float()], 1)
    ka, kb = cell_key(a[:, :3]), cell_key(b[:, :3])
    assert ka.unique().numel() == V
    a, b = a[torch.argsort(ka)], b[torch.argsort(kb)]
    assert torch.allclose(a, b, atol=2e-5), float((a - b).abs().max())
    assert int(vdeg.sum()) == V
    assert torch.equal(vdeg[:50].long(), torch.bincount(med, minlength=50))
    # the scratch cleans itself: a second pass gives the same voxel count
    _lib.check(L.pcs_trk_group_bounds(s, _ptr(pts), _ptr(group), n, _ptr(sb)), "bounds")
    _lib.check(L.pcs_trk_sample(s, ctypes.byref(st)), "sample")
    assert int(sp.t["ctr"][1].item()) == V
    assert int(sp.t["ctr"][2].item()) == 0
