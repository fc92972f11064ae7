"""ClusterProposal preprocessor (mirror of pcdet/models/registration/preprocessors/cluster_proposal.py).

propose_cluster runs, per radius, ONE voxel-hash build and ONE fused search + union-find launch over the
whole sequence; the reference's 10-frame chunks become key segments of that launch, so the grids, the
neighbour sets and the chunk-wise component numbering (running offset, cluster_proposal.py:63-81) are the
same as the reference's per-chunk calls."""
import torch
from torch import nn

from .. import graph_utils, ops, parallel
from ..utils import EasyDict, Timer
from .eval_utils import FrameBoxes

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
        shard = parallel.SHARD
        num_frames = shard.F if shard is not None else int(seq_dict["point_sweep"].max().long().item()) + 1
        seq_dict["num_frames"] = num_frames
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
                seq_dict[f"point_{comp_key}"], seq_dict[f"num_{comp_key}"] = self._globalize(lab, nc, fxyz)
            return seq_dict
        for comp_key in self.component_keys:
            with Timer(f"Propose Cluster for {comp_key}", verbose=verbose):
                graph = getattr(self, f"graph_{comp_key}")
                labels, n_comp = ops.cluster_labels(fxyz, float(graph.radius), int(graph.max_num_neighbors),
                                                    chunk=CHUNK_FRAMES, num_frames=num_frames)
                seq_dict[f"point_{comp_key}"], seq_dict[f"num_{comp_key}"] = self._globalize(labels, n_comp, fxyz)
        return seq_dict

    @staticmethod
    def _globalize(labels, n_comp, fxyz):
        """Frame-window sharding: this rank numbered the components of ITS chunks with a running offset; the
        per-chunk counts of all ranks (disjoint chunks: sum == gather) give the sequence-wide running offset of
        cluster_proposal.py:80-81."""
        shard = parallel.SHARD
        if shard is None:
            return labels, n_comp
        counts = shard.all_reduce_sum(n_comp.clone())
        chunk = torch.div(fxyz[:, 0].round().long(), CHUNK_FRAMES, rounding_mode="floor")
        local_off = torch.cumsum(n_comp, 0) - n_comp
        global_off = torch.cumsum(counts, 0) - counts
        if labels.numel():
            labels = labels - local_off[chunk] + global_off[chunk]
        return labels, counts

    # ---- GT bookkeeping (evaluation, cluster_proposal.py:90-285) --------------------------------------------
    def format_boxes(self, seq_dict, num_frames):
        return EasyDict(dict(attr=seq_dict["gt_box_attr"].reshape(-1, 7),
                             cls_label=seq_dict["gt_box_cls_label"].reshape(-1),
                             frame=seq_dict["gt_box_frame"].reshape(-1),
                             trace_id=seq_dict["gt_box_track_label"].reshape(-1)))

    def evaluate_proposal(self, seq_dict):
        """Point-wise IoU of proposed clusters against GT boxes; emits point_gt_box_id etc. which
        ClusterTracking requires (cluster_tracking.py:800).

        The reference loops over component keys x frames x components with a CPU point-in-box test per frame
        (cluster_proposal.py:142-285).  A component lives in exactly one frame, so the whole sequence is evaluated at
        once: one pcs_points_in_boxes launch gives every point its first box and the (component, box) membership
        counts of all keys; the majority box, the IoU and the per-box / per-trace maxima are segment reductions."""
        fxyz = seq_dict["point_fxyz"]
        num_points = fxyz.shape[0]
        num_frames = int(seq_dict["num_frames"]) if "num_frames" in seq_dict else \
            int(seq_dict["point_sweep"].max().long().item()) + 1
        seq_boxes = self.format_boxes(seq_dict, num_frames)
        num_boxes = seq_boxes.attr.shape[0]
        if num_boxes == 0:
            for key in ["gt_box_id", "gt_trace_id", "pred_trace_id", "pred_box_id"]:
                seq_dict[f"point_{key}"] = seq_dict["segmentation_label"].new_zeros(num_points) - 1
            return seq_dict
        dev = fxyz.device
        comps = [seq_dict[f"point_{k}"].reshape(-1).long() for k in self.component_keys]
        sizes = [int(v) for v in torch.stack([c.max() for c in comps]).tolist()]  # one host sync
        fb = FrameBoxes(seq_boxes.attr, seq_boxes.frame, num_frames)
        frame_of_point = seq_dict["point_sweep"].reshape(-1).long()
        best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
        trace_id = seq_boxes.trace_id.long()
        num_traces = int(trace_id.max().item()) + 1
        trace_best = seq_boxes.attr.new_zeros(num_traces)
        first, counts = [], []
        for i0 in range(0, len(comps), 3):  # three count tables per launch
            f, c = fb.query(fxyz, cids=[(cc, sz + 1) for cc, sz in zip(comps[i0:i0 + 3], sizes[i0:i0 + 3])])
            first, counts = f, counts + c
        gt_local = first.long()  # frame-local index of the first box holding the point (-1: none)
        in_box = gt_local >= 0
        gt_sorted = torch.where(in_box, fb.off[frame_of_point] + gt_local, gt_local)  # row among frame-sorted boxes
        trace_sorted = trace_id[fb.order]
        gt_trace = torch.where(in_box, trace_sorted[gt_sorted.clamp(min=0)], gt_local)
        gt_size = torch.bincount(gt_sorted[in_box], minlength=num_boxes)
        best_sorted = torch.zeros(num_boxes, dtype=torch.float64, device=dev)
        trace_best64 = torch.zeros(num_traces, dtype=torch.float64, device=dev)
        pred_box = pred_trace = None
        for comp_key, c, cnt in zip(self.component_keys, comps, counts):
            Ck = cnt.shape[0]
            has = cnt.sum(1) > 0
            box_of = cnt.argmax(1)  # first maximum, like bi_mask.sum(-1).argmax()
            cframe = torch.zeros(Ck, dtype=torch.int64, device=dev).scatter_(0, c, frame_of_point)
            comp_size = torch.bincount(c, minlength=Ck)
            hit = in_box & has[c] & (gt_local == box_of[c])
            inter = torch.bincount(c[hit], minlength=Ck).double()
            assigned = (fb.off[cframe] + box_of).clamp(max=num_boxes - 1)
            union = (comp_size + gt_size[assigned]).double() - inter
            iou = torch.where(has, inter / (union + 1e-6), torch.zeros_like(inter))
            best_sorted.scatter_reduce_(0, assigned[has], iou[has], "amax")
            trace_best64.scatter_reduce_(0, trace_sorted[assigned[has]], iou[has], "amax")
            pred_box = torch.where(has[c], box_of[c], torch.full_like(c, -1))
            pred_trace = torch.where(has[c], trace_sorted[assigned[c]], torch.full_like(c, -1))
            after = torch.zeros_like(best_iou)
            after[fb.order] = best_sorted.to(best_iou)
            seq_boxes[f"best_iou_after_{comp_key}"] = after
        best_iou[fb.order] = best_sorted.to(best_iou)
        trace_best = trace_best64.to(trace_best)
        if parallel.SHARD is not None:  # every rank evaluated its own frames
            parallel.SHARD.all_reduce_max(best_iou)
            parallel.SHARD.all_reduce_max(trace_best)
        seq_dict["gt_box_best_iou"] = best_iou
        seq_dict["gt_trace_best_iou"] = trace_best
        seq_dict["point_gt_box_id"] = gt_local.to(comps[-1])
        seq_dict["point_gt_trace_id"] = gt_trace.to(comps[-1])
        seq_dict["point_pred_trace_id"] = pred_trace
        seq_dict["point_pred_box_id"] = pred_box
        seq_dict["frame_boxes"] = fb
        return seq_dict

    def forward(self, seq_dict):
        import os
        import time
        timing = os.environ.get("PCS_STAGE_TIMING")
        if timing:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        seq_dict = self.propose_cluster(seq_dict)
        if timing:
            torch.cuda.synchronize()
            print(f"[proposal] propose_cluster {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
        if self.model_cfg.get("EVALUATE", True):
            with Timer("Evaluate Proposal", verbose=self.model_cfg.get("VERBOSE", True)):
                seq_dict = self.evaluate_proposal(seq_dict)
        return seq_dict

    def get_output_feature_dim(self):
        return 0
