"""Goldens recorded from the reference's OWN Python (oracle/gen_golden.py::gen_eval_tracking, CPU run of the unmodified
simple_reg.format_boxes, ClusterProposal.evaluate_proposal, ClusterTracking.track_frame with every
register_to_next_frame call recorded, and extract_traces_and_update_boxes):

  * GT formatting (CPU test)                                   simple_reg.py:35-101
  * evaluate_proposal                                          cluster_proposal.py:142-285
  * every recorded ICP solve, teacher-forced, at 1e-4          registration_utils.py:83-206
  * extract_traces_and_update_boxes on the reference's tracks  cluster_tracking.py:287-428
"""
import os

import numpy as np
import pytest
import torch

from helpers import component_centers, transform_errors


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "eval_tracking.npz"))


def _seq_with_boxes(g, dev):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    from pcseqlearning_b200.utils import EasyDict
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seq = EasyDict(dict(point_sweep=t(g["sweep_all"]), gt_box_attr=t(g["raw_gt_box_attr"]),
                        gt_box_cls_label=t(g["raw_gt_box_cls_label"]), augmented=t(g["raw_augmented"]),
                        num_points_in_gt=t(g["raw_num_points_in_gt"]), obj_ids=g["raw_obj_ids"]))
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_test_out")
    cfg.PREPROCESSORS = []
    model = SimpleReg(cfg, {}, None)
    return model.format_boxes(seq)


def test_format_boxes_vs_reference(golden_dir):
    g = _g(golden_dir)
    seq = _seq_with_boxes(g, "cpu")
    for k in ["gt_box_attr", "gt_box_cls_label", "gt_box_frame", "gt_box_track_label", "moving"]:
        np.testing.assert_array_equal(seq[k].numpy(), g[f"box_{k}"], err_msg=k)
    np.testing.assert_allclose(seq["gt_box_velo"].numpy(), g["box_gt_box_velo"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_evaluate_proposal_vs_reference(golden_dir):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.preprocessors.cluster_proposal import ClusterProposal
    g = _g(golden_dir)
    dev = torch.device("cuda", 0)
    seq = _seq_with_boxes(g, dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    keep = g["seg_all"] < 17
    seq["point_fxyz"] = t(g["points_all"][keep])
    seq["point_sweep"] = t(g["sweep_all"][keep])
    seq["segmentation_label"] = t(g["seg_all"][keep])
    for k in ("rad1x25", "rad0x75", "rad0x25"):
        seq[f"point_component_{k}"] = t(g[f"comp_{k}"])
    cfg = [p for p in cluster_tracking_cfg().PREPROCESSORS if p.NAME == "ClusterProposal"][0]
    mod = ClusterProposal(cfg, {}).to(dev)
    seq = mod.evaluate_proposal(seq)
    for k in ["point_gt_box_id", "point_gt_trace_id", "point_pred_trace_id", "point_pred_box_id"]:
        np.testing.assert_array_equal(seq[k].cpu().numpy(), g[f"eval_{k}"], err_msg=k)
    np.testing.assert_allclose(seq["gt_box_best_iou"].cpu().numpy(), g["eval_gt_box_best_iou"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(seq["gt_trace_best_iou"].cpu().numpy(), g["eval_gt_trace_best_iou"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_icp_teacher_forced(golden_dir):
    """Every recorded register_to_next_frame call of the reference's track_frame run, fed with the reference's own
    inputs of that call (so errors cannot accumulate along the chain): transforms within 1e-4."""
    from pcseqlearning_b200 import ops
    s = np.load(os.path.join(golden_dir, "tracking_steps.npz"))
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    n = len(s["picked"])
    assert n >= 16
    worst = (0.0, 0.0)
    for i in range(n):
        p = f"c{i}_"
        mov, ref = s[p + "mov"], s[p + "ref"]
        C = int(s[p + "C"])
        df = int(ref[0, 0]) - int(mov[0, 0])
        moved, T, l1, ratio, info = ops.register_icp(
            cuda(mov), cuda(s[p + "mov_comp"]), cuda(s[p + "mov_stat"]), cuda(ref), cuda(s[p + "ref_stat"]), C,
            float(s[p + "radius"]), df, angle_regularizer=float(s[p + "reg"]), max_iter=int(s[p + "max_iter"]),
            stopping_delta=float(s[p + "delta"]))
        comp = s[p + "mov_comp"]
        ok = comp >= 0
        ang, dt = transform_errors(T.cpu().numpy(), s[p + "T"], component_centers(mov[ok], comp[ok], C))
        worst = (max(worst[0], float(ang.max())), max(worst[1], float(dt.max())))
        assert ang.max() < 1e-4 and dt.max() < 1e-4, (i, int(s["picked"][i]), df, float(ang.max()), float(dt.max()))
        np.testing.assert_allclose(l1.cpu().numpy(), s[p + "l1"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(ratio.cpu().numpy(), s[p + "ratio"], rtol=0, atol=1e-6)
    print("teacher-forced ICP: worst rotation / translation error", worst)


@pytest.mark.gpu
def test_extract_traces_vs_reference(golden_dir):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.preprocessors.cluster_tracking import ClusterTracking
    from pcseqlearning_b200.utils import EasyDict
    g = _g(golden_dir)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seq = _seq_with_boxes(g, dev)
    cfg = [p for p in cluster_tracking_cfg().PREPROCESSORS if p.NAME == "ClusterTracking"][0]
    cfg.VERBOSE = False
    mod = ClusterTracking(cfg, {}).to(dev)
    seq_boxes = mod.format_boxes(seq, 17)
    seq_boxes.best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
    am = g["height_all"] > 0
    all_points = EasyDict(fxyz=t(g["points_all"][am]), frame=t(g["sweep_all"][am]), height=t(g["height_all"][am]),
                          full_instance_label=t(g["inst_all"][am]), full_segmentation_label=t(g["seg_all"][am]))
    ex = EasyDict({k: t(g[f"ex_{k}"]) for k in ["fxyz", "component", "segmentation_label", "frame_indices",
                                                "original_indices", "moving", "transforms"]})
    full, seq_boxes = mod.extract_traces_and_update_boxes(all_points, ex, seq_boxes)
    for k in ["fxyz", "component", "segmentation_label", "instance_label", "original_indices", "frame_indices",
              "moving", "component_hit", "component_size"]:
        np.testing.assert_array_equal(full[k].cpu().numpy(), g[f"full_{k}"], err_msg=k)
    np.testing.assert_allclose(seq_boxes.best_iou.cpu().numpy(), g["best_iou_after_tracking"], rtol=0, atol=1e-6)
