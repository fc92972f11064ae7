"""GroundPlaneRemover preprocessor (mirror of pcdet/models/registration/preprocessors/ground_plane_remover.py:152-255)."""
import os

import torch
from torch import nn

from .. import ops
from ..utils import EasyDict, Timer
from .ground_utils import ground_plane_removal


class GroundPlaneRemover(nn.Module):
    """Estimates the sequence-level ground height field, caches it in ``<DIR>/<seq>/pillar_height.pth`` and
    drops every point lower than TRUNCATE_HEIGHT above it; the unfiltered arrays are kept as ``full_*``."""

    def __init__(self, model_cfg, runtime_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        # the only leaf parameter the reference's optimizer sees
        self.fake_params = nn.Parameter(torch.zeros(1, dtype=torch.float32), requires_grad=True)
        self.forward_dict = EasyDict()

    def output_stats(self, segmentation_label, ground_mask, sequence_id, log_dir):
        """Removed-point statistics against the segmentation labels (ground_plane_remover.py:159-183)."""
        os.makedirs(log_dir, exist_ok=True)
        removed = segmentation_label[ground_mask]
        n_removed_fg = ((removed > 0) & (removed <= 7)).sum().item()
        n_removed_ground = (removed >= 17).sum().item()
        n_removed = ground_mask.long().sum().item()
        n_fg = ((segmentation_label > 0) & (segmentation_label <= 7)).sum().item()
        n_ground = (segmentation_label >= 17).sum().item()
        with open(f"{log_dir}/{sequence_id}.txt", "w") as fout:
            fout.write(f"{self.model_cfg}\n")
            fout.write(f"#removed_points={n_removed}\n")
            fout.write(f"#removed_foreground={n_removed_fg}\n")
            fout.write(f"#removed_ground={n_removed_ground}\n")
            fout.write(f"ground_precision={n_removed_ground / (n_removed + 1e-6):.6f}\n")
            fout.write(f"ground_coverage={n_removed_ground / (n_ground + 1e-6):.6f}\n")
            fout.write(f"foreground_precision={n_removed_fg / (n_removed + 1e-6):.6f}\n")
            fout.write(f"foreground_coverage={n_removed_fg / (n_fg + 1e-6):.6f}\n")

    def forward(self, seq_dict):
        sequence_id = seq_dict["frame_id"][0][:-4]
        point_fxyz = seq_dict["point_fxyz"]
        path = f"{self.model_cfg.DIR}/{sequence_id}"
        cache = f"{path}/pillar_height.pth"
        use_cache = self.model_cfg.get("USE_CACHE", True)
        if use_cache and os.path.exists(cache):
            saved = torch.load(cache, map_location=point_fxyz.device)
            height, horizon, error, _, _ = ground_plane_removal(point_fxyz, self.model_cfg, warmup=saved)
        else:
            with Timer("Ground Removal", verbose=self.model_cfg.get("VERBOSE", True)):
                height, horizon, error, pillar_height, pillar_min_z = ground_plane_removal(point_fxyz, self.model_cfg)
            if use_cache:
                os.makedirs(path, exist_ok=True)
                torch.save(dict(pillar_height=pillar_height, pillar_min_z=pillar_min_z), cache)
        seq_dict["point_horizon"] = horizon
        seq_dict["point_error"] = error
        ground_mask = None
        for h in self.model_cfg.TRUNCATE_HEIGHT:
            ground_mask = height < h
            if "segmentation_label" in seq_dict and self.model_cfg.get("LOG_DIR", None):
                self.output_stats(seq_dict["segmentation_label"], ground_mask, sequence_id,
                                  self.model_cfg.LOG_DIR + f"/height{h}")
        seq_dict["point_height"] = height
        keep_rows = (~ground_mask).nonzero().reshape(-1)  # one index list shared by every filtered array
        for key in ["point_fxyz", "segmentation_label", "point_sweep", "point_height", "instance_label",
                    "point_horizon"]:
            if key in seq_dict:
                seq_dict[f"full_{key}"] = seq_dict[key]  # the filtered copy below never aliases the full array
                seq_dict[key] = ops.gather_rows(seq_dict[key], keep_rows)
        return seq_dict

    def extra_repr(self):
        return f"{self.model_cfg}"

    def get_output_feature_dim(self):
        return 0
