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


def _small_sequence(frames=17, beams=24, az=600):
    """A small synthetic sequence pushed through subsample + ground removal + proposals (with GT evaluation)."""
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    from pcseqlearning_b200.synthetic import generate_sequence
    dev = torch.device("cuda", 0)
    batch = generate_sequence(3, num_frames=frames, num_beams=beams, num_azimuth=az, device=dev)
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_test_out")
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
        p.SAVE = False
    cfg.SAVE_DIR = None
    return batch, cfg, dev


def test_extract_traces_batched_vs_sequential():
    """Batched re-association + IoU bookkeeping against the per-instance mirror of the reference loop."""
    from pcseqlearning_b200.simple_reg import SimpleReg
    from pcseqlearning_b200.utils import EasyDict
    batch, cfg, dev = _small_sequence()
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    model(batch)
    seq = model.forward_dict["sequences"][0]
    trk = [m for m in model.preprocessors if type(m).__name__ == "ClusterTracking"][0]
    tb = seq["tracking_batch"]
    res = seq["tracking_results"]
    boxes_b = seq["tracking_boxes"]
    assert len(res) == tb.J
    # sequential mirror on the same tracked points
    all_points = EasyDict(fxyz=seq["full_point_fxyz"], frame=seq["full_point_sweep"], height=seq["full_point_height"],
                          full_instance_label=seq["full_instance_label"],
                          full_segmentation_label=seq["full_segmentation_label"])
    keep = seq["full_point_height"] > 0.0
    all_points = EasyDict({k: v[keep] for k, v in all_points.items()})
    seq_boxes = trk.format_boxes(seq, tb.F)
    seq_boxes.best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
    per_inst = tb.results(seg_label=seq["segmentation_label"])
    n_checked = 0
    for (ki, a), (j, ex) in per_inst.items():
        key = f"{a:03d}_{trk.component_keys[ki]}"
        got = res[key]
        if ex.fxyz.shape[0] == 0:
            continue
        ex = EasyDict(dict(ex))
        ex.transforms = tb.transforms(j)
        want, seq_boxes = trk.extract_traces_and_update_boxes(all_points, ex, seq_boxes)
        for k in ("fxyz", "component", "frame_indices", "original_indices", "moving", "segmentation_label",
                  "instance_label", "component_hit", "component_size"):
            assert torch.equal(got[k], want[k]), (key, k, got[k].shape, want[k].shape)
        assert torch.equal(got["transforms"], want["transforms"])
        n_checked += 1
    assert n_checked > 0
    assert torch.allclose(boxes_b.best_iou, seq_boxes.best_iou, atol=1e-6)
    assert float(boxes_b.best_iou.max()) > 0.3


def test_evaluate_proposal_vs_oracle_loops():
    """Sequence-wide evaluate_proposal against a per-frame / per-component restatement of the reference loops
    (cluster_proposal.py:142-285) built on the oracle's points_in_boxes."""
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200.simple_reg import SimpleReg
    batch, cfg, dev = _small_sequence(frames=6)
    cfg.PREPROCESSORS = [p for p in cfg.PREPROCESSORS if p.NAME != "ClusterTracking"]
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    model(batch)
    seq = model.forward_dict["sequences"][0]
    fxyz = seq["point_fxyz"].cpu().numpy()
    frame = np.rint(fxyz[:, 0]).astype(np.int64)
    attr = seq["gt_box_attr"].reshape(-1, 7).cpu().numpy()
    bframe = seq["gt_box_frame"].reshape(-1).cpu().numpy()
    trace = seq["gt_box_track_label"].reshape(-1).cpu().numpy()
    best = np.zeros(attr.shape[0], np.float32)
    tbest = np.zeros(int(trace.max()) + 1, np.float32)
    gt_box = np.full(fxyz.shape[0], -1, np.int64)
    pred_box = None
    for key in ("component_rad1x25", "component_rad0x75", "component_rad0x25"):
        comp = seq[f"point_{key}"].cpu().numpy()
        pred_box = np.full(fxyz.shape[0], -1, np.int64)
        for f in range(int(frame.max()) + 1):
            pm, bm = frame == f, bframe == f
            if not pm.any() or not bm.any():
                continue
            bp = oracle.points_in_boxes(np.ascontiguousarray(fxyz[pm, 1:]), np.ascontiguousarray(attr[bm]))
            inb = (bp == 1).any(0)
            g = np.full(pm.sum(), -1, np.int64)
            g[inb] = bp[:, inb].argmax(0)
            gt_box[pm] = g
            c_f = comp[pm]
            pb = np.full(pm.sum(), -1, np.int64)
            bidx = np.nonzero(bm)[0]
            for c in np.unique(c_f):
                cm = c_f == c
                if not bp[:, cm].any():
                    continue
                b = int(bp[:, cm].sum(-1).argmax())
                pb[cm] = b
                m1 = g == b
                iou = float((m1 & cm).sum()) / (float((m1 | cm).sum()) + 1e-6)
                best[bidx[b]] = max(best[bidx[b]], np.float32(iou))
                tbest[trace[bidx[b]]] = max(tbest[trace[bidx[b]]], np.float32(iou))
            pred_box[pm] = pb
    np.testing.assert_array_equal(seq["point_gt_box_id"].cpu().numpy(), gt_box)
    np.testing.assert_array_equal(seq["point_pred_box_id"].cpu().numpy(), pred_box)
    np.testing.assert_allclose(seq["gt_box_best_iou"].cpu().numpy(), best, atol=1e-6)
    np.testing.assert_allclose(seq["gt_trace_best_iou"].cpu().numpy(), tbest, atol=1e-6)
    assert best.max() > 0.5
