"""Developer: wall time of the full pipeline incl. tracking on a synthetic sequence."""
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
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda", 0)
    batch = generate_sequence(0, num_frames=frames, device=dev)
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_track_out")
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
        p.SAVE = False
    cfg.SAVE_DIR = None
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    times = {}
    for mod in model.preprocessors:
        orig = mod.forward

        def wrapped(seq, _orig=orig, _name=type(mod).__name__):
            torch.cuda.synchronize()
            t = time.perf_counter()
            out = _orig(seq)
            torch.cuda.synchronize()
            times[_name] = times.get(_name, 0.0) + time.perf_counter() - t
            return out

        mod.forward = wrapped
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for _ in range(reps):
        times.clear()
        ops.reset_launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model(batch)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        print(f"total={total:.3f}s stages={ {k: round(v, 3) for k, v in times.items()} }", flush=True)
    seq = model.forward_dict["sequences"][0]
    res = seq.get("tracking_results", {})
    n_ex = sum(int(v["fxyz"].shape[0]) for v in res.values())
    print(f"frames={frames} points={batch['point_bxyz'].shape[0]} total={total:.2f}s "
          f"stages={ {k: round(v, 2) for k, v in times.items()} } anchors*keys={len(res)} extracted_points={n_ex} "
          f"launches={ops.launch_count()} full-pipeline frames/s={frames / total:.2f}")
    boxes = seq.get("tracking_boxes", None)
    if boxes is not None:
        print("GT boxes:", boxes["best_iou"].shape[0], "mean best IoU", float(boxes["best_iou"].mean()),
              "coverage(IoU>0.7)", float((boxes["best_iou"] > 0.7).float().mean()))


if __name__ == "__main__":
    main()
