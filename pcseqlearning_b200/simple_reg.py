"""SimpleReg model plugin (mirror of pcdet/models/registration/simple_reg.py and
registration_module_template.py): no learnable computation -- it splits the collated batch into sequences,
sub-samples each to one point per 0.08 m cell, runs the preprocessors and returns a zero loss so that the
reference's tools/train.py loop (optimizer, DDP, checkpointing) runs unchanged around it."""
import os

import numpy as np
import torch
from torch import nn

from . import ops, parallel
from . import preprocessors as preprocessor
from .utils import EasyDict, filter_dict


class RegistrationTemplate(nn.Module):
    def __init__(self, model_cfg, runtime_cfg, dataset):
        super().__init__()
        self.model_cfg = model_cfg
        self.runtime_cfg = runtime_cfg
        self.dataset = dataset
        self.register_buffer("global_step", torch.LongTensor(1).zero_())
        self.scale = 1 if "SCALE" not in model_cfg else model_cfg.pop("SCALE")
        self.module_topology = ["preprocessors"]
        self.visualizer = None

    def update_ema(self):
        pass

    @property
    def mode(self):
        return "TRAIN" if self.training else "TEST"

    def update_global_step(self):
        self.global_step += 1

    def build_networks(self):
        info = {"module_list": [], "scale": self.scale}
        if self.dataset is not None:
            info["num_point_features"] = getattr(self.dataset, "num_point_features", 0)
            info["max_num_points"] = getattr(self.dataset, "max_num_points", 0) * getattr(self.dataset, "num_sweeps", 1) * 2
            info.update(getattr(self.dataset, "runtime_cfg", {}))
        for name in self.module_topology:
            module, info = getattr(self, f"build_{name}")(model_info_dict=info)
            self.add_module(name, module)
        return info["module_list"]

    def build_preprocessors(self, model_info_dict):
        if self.model_cfg.get("PREPROCESSORS", None) is None:
            return None, model_info_dict
        mods = nn.ModuleList()
        for cfg in self.model_cfg.PREPROCESSORS:
            mods.append(preprocessor.__all__[cfg.NAME](runtime_cfg=model_info_dict, model_cfg=cfg))
        model_info_dict["module_list"].append(mods)
        return mods, model_info_dict

    def load_params_from_file(self, filename, logger=None, to_cpu=False):
        return None

    def load_params_with_optimizer(self, filename, to_cpu=False, optimizer=None, logger=None):
        return 0, 0


def boxes_to_corners_3d(boxes):
    """[N,7] (x,y,z,dx,dy,dz,heading) -> [N,8,3] corners (pcdet/utils/box_utils.py boxes_to_corners_3d)."""
    t = boxes.new_tensor([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1],
                          [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1]]) / 2
    corners = boxes[:, None, 3:6] * t[None]
    c, s = torch.cos(boxes[:, 6]), torch.sin(boxes[:, 6])
    x = corners[..., 0] * c[:, None] - corners[..., 1] * s[:, None]
    y = corners[..., 0] * s[:, None] + corners[..., 1] * c[:, None]
    return torch.stack([x, y, corners[..., 2]], -1) + boxes[:, None, :3]


class SimpleReg(RegistrationTemplate):
    def __init__(self, model_cfg, runtime_cfg, dataset):
        super().__init__(model_cfg, runtime_cfg, dataset)
        self.module_list = self.build_networks()
        self.fake_param = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.forward_dict = EasyDict()
        self.subsample = model_cfg.get("SUBSAMPLE", False)
        self.subsample_grid = [0.08, 0.08, 0.08]  # simple_reg.py:24

    def process_sequence(self, seq_dict):
        timing = os.environ.get("PCS_STAGE_TIMING")
        if self.preprocessors:
            for module in self.preprocessors:
                if timing:
                    import time
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                seq_dict = module(seq_dict)
                if timing:
                    torch.cuda.synchronize()
                    print(f"[stage] {type(module).__name__} {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
        return seq_dict

    def format_boxes(self, seq_dict):
        """Flatten the per-frame GT boxes, drop empty slots, derive track labels and per-box speed
        (simple_reg.py:35-101)."""
        sweep = seq_dict["point_sweep"]
        if parallel.SHARD is not None:  # frame-window sharding: the GT arrays cover the whole sequence
            num_frames = parallel.SHARD.F
        else:
            num_frames = int(sweep.max().long().item()) - int(sweep.min().long().item()) + 1
        attr = seq_dict["gt_box_attr"].reshape(-1, 7)
        cls_label = seq_dict["gt_box_cls_label"].reshape(-1)
        assert attr.shape[0] % num_frames == 0, "gt boxes must be padded to a fixed count per frame"
        per_frame = cls_label.numel() // num_frames
        frame = torch.repeat_interleave(torch.arange(0, num_frames), per_frame, dim=-1).to(cls_label)
        boxes = EasyDict(dict(gt_box_attr=attr, gt_box_cls_label=cls_label, gt_box_frame=frame))
        for key in ["augmented", "num_points_in_gt"]:
            if key in seq_dict:
                boxes[key] = seq_dict[key].reshape(-1)
        keep = boxes.gt_box_attr[:, 3:6].norm(p=2, dim=-1) > 1e-5
        boxes = EasyDict(filter_dict(boxes, keep))
        obj_ids = np.asarray(seq_dict["obj_ids"]).reshape(-1)[keep.cpu().numpy()].astype(str)
        track = np.unique(obj_ids, return_inverse=True)[1]
        boxes.gt_box_track_label = torch.from_numpy(track).to(cls_label).long()
        seq_dict["obj_ids"] = obj_ids
        # per-box speed = mean corner displacement to the previous box of the same trace (the first box of a trace
        # copies the second, single-box traces get 0) -- the reference's per-trace loop (:79-92), all traces at once
        velo = torch.zeros_like(boxes.gt_box_attr[:, 0])
        nb = velo.shape[0]
        if nb > 1:
            key = boxes.gt_box_track_label.long() * (int(boxes.gt_box_frame.max().item()) + 2) + boxes.gt_box_frame.long()
            order = torch.argsort(key, stable=True)
            trk = boxes.gt_box_track_label[order]
            corners = boxes_to_corners_3d(boxes.gt_box_attr[order])
            step = (corners[1:] - corners[:-1]).norm(p=2, dim=-1).mean(dim=-1)
            same_prev = torch.zeros(nb, dtype=torch.bool, device=velo.device)
            same_prev[1:] = trk[1:] == trk[:-1]
            v_sorted = torch.zeros_like(velo)
            v_sorted[1:] = torch.where(same_prev[1:], step, torch.zeros_like(step))
            first = ~same_prev
            has_next = torch.zeros_like(first)
            has_next[:-1] = same_prev[1:]
            nxt = torch.roll(v_sorted, -1)
            v_sorted = torch.where(first & has_next, nxt, v_sorted)
            velo[order] = v_sorted
        boxes.gt_box_velo = velo
        boxes.moving = velo > 5e-2
        for key in boxes.keys():
            seq_dict[key] = boxes[key]
        return seq_dict

    def forward(self, batch_dict):
        batch_size = batch_dict["batch_size"]
        results = []
        for b in range(batch_size):
            seq_dict = EasyDict(dict())
            if batch_size == 1:
                sel = None
            else:
                sel = (batch_dict["point_bxyz"][:, 0] == b).reshape(-1)
            for key in ["point_bxyz", "point_feat", "segmentation_label", "instance_label", "is_foreground",
                        "point_sweep"]:
                if key in batch_dict:
                    seq_dict[key] = batch_dict[key] if sel is None else batch_dict[key][sel]
            seq_dict["point_fxyz"] = torch.cat([seq_dict["point_sweep"].reshape(-1, 1).float(),
                                                seq_dict["point_bxyz"][:, 1:]], dim=-1)
            seq_dict.pop("point_bxyz")
            if self.subsample:
                # one (highest-index) point per 0.08 m cell, arrays re-ordered by ascending cell key
                hook = parallel.SHARD.reduce_bounds if parallel.SHARD is not None else None
                res = ops.voxelize(seq_dict["point_fxyz"], self.subsample_grid, want_mean=False, want_max=True,
                                   bounds_hook=hook)
                pick = res["maxidx"]
                for key in ["point_fxyz", "point_feat", "segmentation_label", "instance_label", "is_foreground",
                            "point_sweep"]:
                    if key in seq_dict:
                        seq_dict[key] = ops.gather_rows(seq_dict[key], pick)
            for key in ["gt_box_cls_label", "gt_box_attr", "augmented", "num_points_in_gt", "gt_boxes", "obj_ids",
                        "frame_id", "pose", "top_lidar_origin", "num_sweeps", "gt_box_corners_3d", "gt_box_velo"]:
                if key in batch_dict:
                    seq_dict[key] = batch_dict[key][b]
            if "gt_box_attr" in seq_dict:
                seq_dict = self.format_boxes(seq_dict)
            sequence_id = seq_dict["frame_id"][0][:-4]
            done = self.model_cfg.get("SAVE_DIR", None) and os.path.exists(
                f"{self.model_cfg.SAVE_DIR}/{sequence_id}/all.pth")
            if not done:
                seq_dict = self.process_sequence(seq_dict)
            results.append(seq_dict)
        self.forward_dict["sequences"] = results
        if self.training:
            loss = torch.zeros(1, device=batch_dict["point_bxyz"].device, requires_grad=True)
            return dict(loss=loss), {}, {}
        return {}, None
