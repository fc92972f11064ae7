"""Developer: CUDA-event times of the logged kernels (ops.enable_event_log) over a few pipeline steps.
    python tools/time_kernels.py [frames] [steps]      (no tracking: ground removal + proposals only)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcseqlearning_b200 import ops
from pcseqlearning_b200.config import cluster_tracking_cfg
from pcseqlearning_b200.simple_reg import SimpleReg
from pcseqlearning_b200.synthetic import generate_sequence


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 198
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    batch = generate_sequence(0, num_frames=frames, device=dev)
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_time_out")
    cfg.PREPROCESSORS = [p for p in cfg.PREPROCESSORS if p.NAME != "ClusterTracking"]
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
    cfg.SAVE_DIR = None
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    for _ in range(2):
        model(batch)
    ops.enable_event_log(True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        model(batch)
    b.record()
    torch.cuda.synchronize()
    print(f"step {a.elapsed_time(b) / steps:.2f} ms (ground removal + proposals + evaluation)")
    for name, ev in ops.event_log().items():
        d = [x.elapsed_time(y) for x, y, _ in ev]
        print(f"  {name:18s} n/step={len(d) // steps:3d} mean {sum(d) / len(d):8.3f} ms  per step {sum(d) / steps:8.3f} ms  "
              f"first launches {[round(v, 2) for v in d[:6]]}")


if __name__ == "__main__":
    main()
