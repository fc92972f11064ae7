"""Dataset plugin that feeds synthetic Waymo-shaped sequences through the reference's unmodified driver.

Boundary: `pcdet.datasets.__all__[DATA_CONFIG.DATASET](dataset_cfg, root_path, training,
logger)` (pcdet/datasets/__init__.py:13-16,73-78).  The driver / model read `runtime_cfg`, `num_point_features`,
`max_num_points`, `num_sweeps` (registration_module_template.py:38-45, tools/train.py:151),
`data_augmentor.set_epoch(epoch)` (train_utils.py:144), `collate_batch` (datasets/__init__.py:98) and optionally
`use_shared_memory` (tools/train.py:217).  Items are `{point_wise, object_wise, scene_wise}` dicts of numpy arrays
(pcdet/datasets/dataset.py:194-200); `collate_batch` follows the key rules of dataset.py:203-298 for the keys of this
path.  One item = one whole sequence (the registration path runs with batch size 1 per rank).
"""
from collections import defaultdict

import numpy as np
import torch
from torch.utils.data import Dataset

from .synthetic import generate_sequence

_POINT_CONCAT = ("point_sweep", "point_feat", "segmentation_label", "instance_label", "is_foreground")
_BOX_INT = ("gt_box_cls_label", "num_points_in_gt")
_BOX_PADDED = ("gt_boxes", "gt_box_attr", "gt_box_corners_3d") + _BOX_INT + ("augmented",)


class _NoAugmentor:
    def set_epoch(self, epoch):
        self.epoch = epoch


class SyntheticSequenceDataset(Dataset):
    """dataset_cfg keys (all optional): NUM_SEQUENCES (1), NUM_SWEEPS = frames per sequence (198), NUM_BEAMS (64),
    NUM_AZIMUTH (2650), DENSE (False), DEVICE ('cuda' when available: the ray casting runs there, items are numpy)."""

    def __init__(self, dataset_cfg=None, root_path=None, training=True, logger=None, class_names=None):
        cfg = dict(dataset_cfg or {})
        self.dataset_cfg = cfg
        self.root_path, self.training, self.logger, self.class_names = root_path, training, logger, class_names
        self.num_sequences = int(cfg.get("NUM_SEQUENCES", 1))
        self.num_sweeps = int(cfg.get("NUM_SWEEPS", 198))
        self.num_beams = int(cfg.get("NUM_BEAMS", 64))
        self.num_azimuth = int(cfg.get("NUM_AZIMUTH", 2650))
        self.dense = bool(cfg.get("DENSE", False))
        self.device = cfg.get("DEVICE", "cuda" if torch.cuda.is_available() else "cpu")
        self.num_point_features = 3  # intensity, elongation, range (waymo_dataset.py:334-343)
        self.max_num_points = self.num_beams * self.num_azimuth * self.num_sweeps
        self.runtime_cfg = dict(num_point_features=self.num_point_features, max_num_points=self.max_num_points,
                                num_sweeps=self.num_sweeps)
        self.data_augmentor = _NoAugmentor()
        self.use_shared_memory = False
        # SequenceSampler reads index_matrix [num_sequences, frames]: one item per sequence here
        self.index_matrix = np.arange(self.num_sequences).reshape(-1, 1)

    def __len__(self):
        return self.num_sequences

    def __getitem__(self, index):
        b = generate_sequence(int(index), num_frames=self.num_sweeps, num_beams=self.num_beams,
                              num_azimuth=self.num_azimuth, dense=self.dense, device=self.device)
        npy = lambda t: t.detach().cpu().numpy()  # noqa: E731
        point_wise = dict(point_xyz=npy(b["point_bxyz"][:, 1:]), point_sweep=npy(b["point_sweep"]),
                          point_feat=npy(b["point_feat"]), segmentation_label=npy(b["segmentation_label"]),
                          instance_label=npy(b["instance_label"]), is_foreground=npy(b["is_foreground"]))
        object_wise = dict(gt_box_attr=npy(b["gt_box_attr"][0]), gt_boxes=npy(b["gt_boxes"][0]),
                           gt_box_cls_label=npy(b["gt_box_cls_label"][0]),
                           gt_box_corners_3d=npy(b["gt_box_corners_3d"][0]), augmented=npy(b["augmented"][0]),
                           num_points_in_gt=npy(b["num_points_in_gt"][0]), obj_ids=b["obj_ids"][0])
        scene_wise = dict(frame_id=b["frame_id"][0], pose=b["pose"][0], num_sweeps=np.int64(b["num_sweeps"][0]))
        return dict(point_wise=point_wise, object_wise=object_wise, scene_wise=scene_wise)

    @staticmethod
    def collate_batch(batch_list, _unused=False, num_mix3d_samples=1):
        merged = defaultdict(list)
        for sample in batch_list:
            for group in sample.values():
                for key, val in group.items():
                    merged[key].append(val)
        batch_size = len(batch_list)
        ret = {}
        for key, val in merged.items():
            if key in _POINT_CONCAT:
                ret[key] = np.concatenate(val, axis=0)
            elif key == "point_xyz":  # -> point_bxyz with the sample index in column 0
                ret["point_bxyz"] = np.concatenate(
                    [np.pad(v, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, v in enumerate(val)], axis=0)
            elif key in _BOX_PADDED:
                if key in _BOX_INT:
                    val, dtype = [v.reshape(-1, 1) for v in val], np.int32
                elif key == "augmented":
                    val, dtype = [v.reshape(-1, 1) for v in val], bool
                else:
                    dtype = np.float32
                max_gt = max(len(v) for v in val)
                out = np.zeros((batch_size, max_gt) + tuple(val[0].shape[1:]), dtype=dtype)
                for k in range(batch_size):
                    out[k, :len(val[k])] = val[k]
                ret[key] = out
            elif key == "obj_ids":
                ret["obj_ids"] = val
            else:
                ret[key] = np.stack(val, axis=0)
        ret["batch_size"] = batch_size
        return ret


__all__ = {"SyntheticSequenceDataset": SyntheticSequenceDataset}
