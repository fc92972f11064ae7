"""ClusterProposal preprocessor (mirror of pcdet/models/registration/preprocessors/cluster_proposal.py).

propose_cluster runs, per radius, ONE voxel-hash build and ONE fused search + union-find launch over the
whole sequence; the reference's 10-frame chunks become key segments of that launch, so the grids, the
neighbour sets and the chunk-wise component numbering (running offset, cluster_proposal.py:63-81) are the
same as the reference's per-chunk calls."""
from collections import defaultdict

import torch
from torch import nn

from .. import graph_utils, ops
from ..utils import EasyDict, Timer, filter_dict
from .eval_utils import points_in_boxes

CHUNK_FRAMES = 10  # cluster_proposal.py:63


class ClusterProposal(nn.Module):
    def __init__(self, model_cfg, runtime_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.forward_dict = EasyDict()
        self.fake_params = nn.Parameter(torch.zeros(1, dtype=torch.float32), requires_grad=True)
        self.component_keys = list(model_cfg["COMPONENT_KEYS"])
        for i, key in enumerate(self.component_keys):
            graph_cfg = dict(model_cfg["GRAPH"])
            graph_cfg["RADIUS"] = graph_cfg["RADIUS"][i]
            self.add_module(f"graph_{key}", graph_utils.build_graph(graph_cfg, runtime_cfg=runtime_cfg))

    def propose_cluster(self, seq_dict):
        """point_fxyz [N,4] (after ground removal) -> seq_dict['point_<comp_key>'] int64[N]."""
        fxyz = seq_dict["point_fxyz"]
        num_frames = int(seq_dict["point_sweep"].max().long().item()) + 1
        verbose = self.model_cfg.get("VERBOSE", True)
        graphs = [getattr(self, f"graph_{k}") for k in self.component_keys]
        same_k = len({int(g.max_num_neighbors) for g in graphs}) == 1
        if self.model_cfg.get("FUSE_RADII", True) and same_k and len(graphs) in (2, 3):
            # multi-radius search: one fine search + one coarse search over the sparse remainder serve all radii
            with Timer("Propose Cluster (multi-radius)", verbose=verbose):
                labels, n_comp = ops.cluster_labels_multi(fxyz, [float(g.radius) for g in graphs],
                                                          int(graphs[0].max_num_neighbors), chunk=CHUNK_FRAMES,
                                                          num_frames=num_frames)
            for comp_key, lab, nc in zip(self.component_keys, labels, n_comp):
                seq_dict[f"point_{comp_key}"] = lab
                seq_dict[f"num_{comp_key}"] = nc
            return seq_dict
        for comp_key in self.component_keys:
            with Timer(f"Propose Cluster for {comp_key}", verbose=verbose):
                graph = getattr(self, f"graph_{comp_key}")
                labels, n_comp = ops.cluster_labels(fxyz, float(graph.radius), int(graph.max_num_neighbors),
                                                    chunk=CHUNK_FRAMES, num_frames=num_frames)
                seq_dict[f"point_{comp_key}"] = labels
                seq_dict[f"num_{comp_key}"] = n_comp
        return seq_dict

    # ---- GT bookkeeping (evaluation, cluster_proposal.py:90-285) --------------------------------------------
    def format_boxes(self, seq_dict, num_frames):
        return EasyDict(dict(attr=seq_dict["gt_box_attr"].reshape(-1, 7),
                             cls_label=seq_dict["gt_box_cls_label"].reshape(-1),
                             frame=seq_dict["gt_box_frame"].reshape(-1),
                             trace_id=seq_dict["gt_box_track_label"].reshape(-1)))

    def assign_instances_to_boxes(self, point_instance_label, bp_mask):
        """Majority box of every instance (cluster_proposal.py:90-114), vectorised: one [B, I] count matrix."""
        uniq, inst = torch.unique(point_instance_label, return_inverse=True)
        counts = torch.zeros(bp_mask.shape[0], uniq.shape[0], dtype=torch.long, device=bp_mask.device)
        counts.index_add_(1, inst, bp_mask.long())
        has = counts.sum(0) > 0
        box_of = counts.argmax(0)
        box_of[~has] = -1
        instance2box = defaultdict(lambda: -1)
        for k, b in zip(uniq.tolist(), box_of.tolist()):
            if b >= 0:
                instance2box[k] = b
        return instance2box, box_of[inst], uniq, box_of, inst

    def evaluate_proposal(self, seq_dict):
        """Point-wise IoU of proposed clusters against GT boxes; emits point_gt_box_id etc. which
        ClusterTracking requires (cluster_tracking.py:800)."""
        num_frames = int(seq_dict["point_sweep"].max().long().item()) + 1
        seq_boxes = self.format_boxes(seq_dict, num_frames)
        num_boxes = seq_boxes.attr.shape[0]
        num_points = seq_dict[f"point_{self.component_keys[0]}"].numel()
        if num_boxes == 0:
            for key in ["gt_box_id", "gt_trace_id", "pred_trace_id", "pred_box_id"]:
                seq_dict[f"point_{key}"] = seq_dict["segmentation_label"].new_zeros(num_points) - 1
            return seq_dict
        seq_boxes.best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
        num_traces = int(seq_boxes.trace_id.max().long().item()) + 1
        traces = EasyDict(dict(best_iou=seq_boxes.attr.new_zeros(num_traces),
                               min_frame=seq_boxes.trace_id.new_zeros(num_traces),
                               max_frame=seq_boxes.trace_id.new_zeros(num_traces)))
        big = num_frames + 1
        traces.min_frame = torch.full_like(traces.min_frame, big).scatter_reduce_(
            0, seq_boxes.trace_id.long(), seq_boxes.frame.long(), "amin")
        traces.max_frame = torch.zeros_like(traces.max_frame).scatter_reduce_(
            0, seq_boxes.trace_id.long(), seq_boxes.frame.long(), "amax")
        fxyz = seq_dict["point_fxyz"]
        frame_of_point = fxyz[:, 0].round().long()
        seq_points = None
        for comp_key in self.component_keys:
            seq_points = EasyDict(component=seq_dict[f"point_{comp_key}"])
            for key in ["gt_box_id", "pred_box_id", "gt_trace_id", "pred_trace_id"]:
                seq_points[key] = torch.zeros_like(seq_points.component) - 1
            for frame_id in range(num_frames):
                frame_mask = frame_of_point == frame_id
                frame_box_mask = (seq_boxes.frame == frame_id).reshape(-1)
                if not frame_mask.any() or not frame_box_mask.any():
                    continue
                comp = seq_points.component[frame_mask]
                boxes = EasyDict(filter_dict(seq_boxes, frame_box_mask))
                bp_mask = points_in_boxes(fxyz[frame_mask, 1:], boxes.attr)  # [B, n] int
                in_box = (bp_mask == 1).any(0)
                gt_box_id = torch.zeros_like(comp) - 1
                gt_box_id[in_box] = bp_mask[:, in_box].argmax(0)
                gt_trace_id = torch.zeros_like(comp) - 1
                gt_trace_id[in_box] = boxes.trace_id[gt_box_id[in_box]].to(gt_trace_id)
                _, pred_box_id, uniq, box_of, inst = self.assign_instances_to_boxes(comp, bp_mask)
                pred_trace_id = torch.zeros_like(comp) - 1
                valid = pred_box_id >= 0
                pred_trace_id[valid] = boxes.trace_id[pred_box_id[valid]].to(pred_trace_id)
                # IoU of every (component, assigned box) pair, all at once (cluster_proposal.py:237-255)
                sel = box_of >= 0
                if sel.any():
                    comp_size = torch.bincount(inst, minlength=uniq.shape[0])
                    gt_size = torch.bincount(gt_box_id[gt_box_id >= 0], minlength=bp_mask.shape[0])
                    # intersection = points of the component whose gt_box_id is the assigned box
                    hit = (gt_box_id == box_of[inst]) & (box_of[inst] >= 0)
                    inter = torch.bincount(inst[hit], minlength=uniq.shape[0]).float()
                    union = (comp_size + gt_size[box_of.clamp(min=0)]).float() - inter
                    iou = torch.where(sel, inter / (union + 1e-6), torch.zeros_like(inter))
                    best = boxes.best_iou.clone()
                    best.scatter_reduce_(0, box_of[sel], iou[sel].to(best), "amax")
                    seq_boxes.best_iou[frame_box_mask] = best
                    tr = boxes.trace_id[box_of[sel]].long()
                    traces.best_iou.scatter_reduce_(0, tr, iou[sel].to(traces.best_iou), "amax")
                for key, val in (("gt_box_id", gt_box_id), ("gt_trace_id", gt_trace_id),
                                 ("pred_trace_id", pred_trace_id), ("pred_box_id", pred_box_id)):
                    seq_points[key][frame_mask] = val
            seq_boxes[f"best_iou_after_{comp_key}"] = seq_boxes["best_iou"].clone()
        seq_dict["gt_box_best_iou"] = seq_boxes.best_iou
        seq_dict["gt_trace_best_iou"] = traces.best_iou
        for key in ["gt_box_id", "gt_trace_id", "pred_trace_id", "pred_box_id"]:
            seq_dict[f"point_{key}"] = seq_points[key]
        return seq_dict

    def forward(self, seq_dict):
        seq_dict = self.propose_cluster(seq_dict)
        if self.model_cfg.get("EVALUATE", True):
            with Timer("Evaluate Proposal", verbose=self.model_cfg.get("VERBOSE", True)):
                seq_dict = self.evaluate_proposal(seq_dict)
        return seq_dict

    def get_output_feature_dim(self):
        return 0
