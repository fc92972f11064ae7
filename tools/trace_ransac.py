"""Debug: per-block timeline of one IRLS iteration of ground_ransac_kernel (needs a -DPCS_RANSAC_TRACE build).

    nvcc $(python -c "from pcseqlearning_b200.build import NVCC_FLAGS; print(' '.join(NVCC_FLAGS))") -DPCS_RANSAC_TRACE \
         -Iinclude -o pcseqlearning_b200/libpcseq_b200.so pcseqlearning_b200/csrc/*.cu
    python tools/trace_ransac.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pcseqlearning_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    from pcseqlearning_b200.synthetic import generate_sequence
    batch = generate_sequence(0, num_frames=198, device=dev)
    model = bench.build_model(dev)
    for _ in range(2):
        model(batch)
    torch.cuda.synchronize()
    L = _lib.lib()
    out = np.zeros(2048 * 4, dtype=np.uint64)
    L.pcs_debug_ransac_trace.argtypes = [ctypes.c_void_p]
    rc = L.pcs_debug_ransac_trace(out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    t = out.reshape(2048, 4)
    nb = int((t[:, 0] > 0).sum())
    t = t[:nb]
    t0 = t[:, 0].min()
    start = (t[:, 0] - t0).astype(np.float64) / 1e3
    end = (t[:, 1] - t0).astype(np.float64) / 1e3
    rel = (t[:, 3] - t0).astype(np.float64) / 1e3
    sm = (t[:, 2] >> np.uint64(32)).astype(np.int64)
    nr = (t[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
    dur = end - start
    print(f"blocks {nb}; step start spread {start.max():.1f} us; work duration us: min {dur.min():.1f} "
          f"p10 {np.percentile(dur, 10):.1f} median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f}")
    print(f"end-of-work us: min {end.min():.1f} median {np.median(end):.1f} max {end.max():.1f}; barrier release {np.median(rel):.1f}")
    per_sm = np.bincount(sm, minlength=148)
    print("blocks per SM histogram:", np.bincount(per_sm))
    for k in np.unique(per_sm):
        sel = np.isin(sm, np.where(per_sm == k)[0])
        print(f"  SMs with {k} blocks: mean duration {dur[sel].mean():.1f} us, max end {end[sel].max():.1f}")
    order = np.argsort(dur)
    print("slowest blocks (id, sm, nr, dur):", [(int(i), int(sm[i]), int(nr[i]), round(float(dur[i]), 1)) for i in order[-8:]])
    print("fastest blocks (id, sm, nr, dur):", [(int(i), int(sm[i]), int(nr[i]), round(float(dur[i]), 1)) for i in order[:8]])
    # duration by position inside the team (44 blocks per team in the steady state)
    bpt = nb // 10
    pos = np.arange(nb) % max(bpt, 1)
    print("mean duration by rank inside team:", [round(float(dur[pos == k].mean()), 1) for k in range(0, bpt, max(bpt // 11, 1))])


if __name__ == "__main__":
    main()
