"""Developer: sizes that shape the batched tracker (components per anchor frame, voxels per level, ...)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcseqlearning_b200.config import cluster_tracking_cfg
from pcseqlearning_b200.simple_reg import SimpleReg
from pcseqlearning_b200.synthetic import generate_sequence
from pcseqlearning_b200 import ops


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 198
    dev = torch.device("cuda", 0)
    batch = generate_sequence(0, num_frames=frames, device=dev)
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_stats_out")
    cfg.PREPROCESSORS = [p for p in cfg.PREPROCESSORS if p.NAME in ("GroundPlaneRemover", "ClusterProposal")]
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
    cfg.SAVE_DIR = None
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model(batch)
        torch.cuda.synchronize()
        print("pipeline (with GT evaluation) %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    seq = model.forward_dict["sequences"][0]
    fxyz = seq["point_fxyz"]
    frame = fxyz[:, 0].long()
    print("N_g", fxyz.shape[0], "N_s", seq["full_point_fxyz"].shape[0], "gt boxes", seq["gt_box_attr"].reshape(-1, 7).shape)
    cnt = torch.bincount(frame, minlength=frames)
    print("points/frame min/mean/max", int(cnt.min()), float(cnt.float().mean()), int(cnt.max()))
    full_h = seq["full_point_height"]
    print("all_points (height>0)", int((full_h > 0).sum()))
    anchors = list(range(0, frames, 8))
    for key in ("component_rad1x25", "component_rad0x75", "component_rad0x25"):
        c = seq[f"point_{key}"]
        ntot = int(c.max()) + 1
        rng, nonempty = [], []
        for a in anchors:
            ca = c[frame == a]
            rng.append(int(ca.max() - ca.min()) + 1)
            nonempty.append(int(ca.unique().numel()))
        print(key, "total comps", ntot, "anchor range sum", sum(rng), "max", max(rng), "nonempty sum", sum(nonempty),
              "max", max(nonempty))
    for vs in ([0.4, 0.4, 0.6], [0.2, 0.2, 0.3], [0.1, 0.1, 0.15]):
        tot, mx = 0, 0
        for a in anchors[:6]:
            r = ops.voxelize(fxyz[frame == a].contiguous(), vs, want_mean=False)
            tot += r["num"]
            mx = max(mx, r["num"])
        print("voxel", vs, "per-frame voxels mean", tot / 6, "max", mx)


if __name__ == "__main__":
    main()
