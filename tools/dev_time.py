"""Developer timing of the individual kernels on a synthetic sequence (not the bench contract)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import torch

from pcseqlearning_b200 import ops
from pcseqlearning_b200.synthetic import generate_sequence, sequence_fxyz


def ev_time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    t = time.time()
    b = generate_sequence(0, num_frames=frames, device="cuda")
    torch.cuda.synchronize()
    print(f"generated {frames} frames in {time.time() - t:.1f}s, points={b['point_bxyz'].shape[0]}")
    f = sequence_fxyz(b)
    keep = b["segmentation_label"] < 17
    f = f[keep].contiguous()
    n = f.shape[0]
    print("non-ground points", n)
    n_seg = (frames + 9) // 10
    for r in (1.25, 0.75, 0.25):
        vs = ops.radius_voxel_size(r)
        tb = ev_time(lambda: ops.CellGrid(f, vs, seg_div=10, n_seg=n_seg))
        grid = ops.CellGrid(f, vs, seg_div=10, n_seg=n_seg)
        cells = grid.check()
        parent = ops.uf_new(n, f.device)
        ts = ev_time(lambda: grid.search(None, 32, r, uf_parent=parent, want_lists=False), 3)
        tl = ev_time(lambda: grid.search(None, 32, r), 3)
        tn = ev_time(lambda: grid.search(f, 32, r), 3)
        seg_of = ops.point_segments(f, 10, n_seg)
        tc = ev_time(lambda: ops.uf_labels(parent, seg_of, n_seg))
        nb, cnt, _ = grid.search(None, 32, r)
        E = int(cnt.sum().item())
        print(f"r={r}: cells={cells} H={grid.H} build={tb[0]:.3f}ms  search+uf={ts[0]:.3f}ms  search(lists)={tl[0]:.3f}ms "
              f"search(unordered)={tn[0]:.3f}ms labels={tc[0]:.3f}ms  E={E} "
              f"| build GB/s={n * 28 / tb[0] / 1e6:.1f} search GB/s(36B/q)={n * 36 / ts[0] / 1e6:.1f}")


if __name__ == "__main__":
    main()
