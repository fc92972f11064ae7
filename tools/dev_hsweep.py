"""Developer: search time vs table size (load factor) at one radius."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcseqlearning_b200 import ops
from pcseqlearning_b200.synthetic import generate_sequence, sequence_fxyz
from tools.dev_time import ev_time


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    b = generate_sequence(0, num_frames=frames, device="cuda")
    f = sequence_fxyz(b)
    f = f[b["segmentation_label"] < 17].contiguous()
    n = f.shape[0]
    n_seg = (frames + 9) // 10
    for r in (0.25, 1.25):
        for div in (0.5, 1, 2, 4, 8):
            H = ops.next_pow2(int(max(n / div, 1024)))
            grid = ops.CellGrid(f, ops.radius_voxel_size(r), seg_div=10, n_seg=n_seg, table_size=H)
            cells = grid.check()
            parent = ops.uf_new(n, f.device)
            ts = ev_time(lambda: grid.search(None, 32, r, uf_parent=parent, want_lists=False), 3)
            tb = ev_time(lambda: ops.CellGrid(f, ops.radius_voxel_size(r), seg_div=10, n_seg=n_seg, table_size=H), 3)
            print(f"r={r} n={n} H={H} load={cells / H:.3f} search={ts[0]:.3f}ms build={tb[0]:.3f}ms")


if __name__ == "__main__":
    main()
