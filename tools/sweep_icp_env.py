"""Developer: one process, several ICP kernel settings (the C side reads its PCS_ICP_* knobs at every launch).

    python tools/sweep_icp_env.py 198 "PCS_ICP_MODE=0" "PCS_ICP_MODE=4" "PCS_ICP_MARGIN=0.1" ...
Prints the tracker stage time and the per-level ICP phase profile of the second of two runs per setting."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcseqlearning_b200.config import cluster_tracking_cfg
from pcseqlearning_b200.simple_reg import SimpleReg
from pcseqlearning_b200.synthetic import generate_sequence

KNOBS = ["PCS_ICP_MODE", "PCS_ICP_MARGIN", "PCS_ICP_BATCH", "PCS_ICP_NOCACHE"]


def main():
    frames = int(sys.argv[1])
    settings = sys.argv[2:] or [""]
    os.environ["PCS_TRACK_TIMING"] = "1"
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
    model(batch)  # warm-up
    for st in settings:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in st.split(","):
            if "=" in kv:
                k, v = kv.split("=")
                os.environ[k] = v
        print(f"== {st or 'default'}", flush=True)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model(batch)
            torch.cuda.synchronize()
            print(f"   total {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)


if __name__ == "__main__":
    main()
