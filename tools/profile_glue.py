"""Developer: which torch ops (by Python call site) cost device time in one full-pipeline step.

    python tools/profile_glue.py 198 > gpurun_out/glue_profile.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from pcseqlearning_b200.config import cluster_tracking_cfg
from pcseqlearning_b200.simple_reg import SimpleReg
from pcseqlearning_b200.synthetic import generate_sequence


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 198
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
    model(batch)
    model(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        model(batch)
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages(group_by_stack_n=6):
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        self_t = getattr(ev, "self_device_time_total", None)
        if self_t is None:
            self_t = getattr(ev, "self_cuda_time_total", 0.0)
        if self_t <= 0:
            continue
        stack = [s for s in ev.stack if "pcseqlearning_b200" in s or "bench.py" in s][:3]
        rows.append((self_t, ev.count, ev.key, " <- ".join(s.split("/")[-1] for s in stack)))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total self device time {tot / 1e3:.1f} ms")
    for t, n, key, st in rows[:70]:
        print(f"{t / 1e3:8.2f} ms  n={n:5d}  {key[:60]:60s} {st}")


if __name__ == "__main__":
    main()
