"""Developer profile of one bench step: per-stage wall time (synchronised) and the top CUDA kernels."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pcseqlearning_b200 import ops
from pcseqlearning_b200.preprocessors import ground_utils as gu
from pcseqlearning_b200.synthetic import generate_sequence


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 198
    dev = torch.device("cuda", 0)
    batch = generate_sequence(0, num_frames=frames, device=dev)
    model = bench.build_model(dev)
    for _ in range(2):
        model(batch)
    torch.cuda.synchronize()

    # stage timings by wrapping the main calls
    def timeit(name, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        print(f"  {name:<32s} {1e3 * (time.perf_counter() - t):8.2f} ms")
        return r

    fxyz = torch.cat([batch["point_sweep"].reshape(-1, 1).float(), batch["point_bxyz"][:, 1:]], -1)
    res = timeit("subsample voxelize (24M)", lambda: ops.voxelize(fxyz, [0.08] * 3, want_mean=False, want_max=True))
    pick = res["maxidx"]
    sub = timeit("gather 6 arrays by pick", lambda: [batch[k][pick] for k in ["point_bxyz", "point_feat", "segmentation_label", "instance_label", "is_foreground", "point_sweep"]])
    f2 = fxyz[pick].contiguous()
    cfg = model.preprocessors[0].model_cfg
    vox = timeit("ground: grid_sample", lambda: gu.grid_sample(f2, [0.10, 0.10, 0.03]))
    voxels, inv = vox
    pillar_size = torch.tensor(cfg.PILLAR_SIZE).to(f2)
    fp = timeit("ground: format_pillars", lambda: gu.format_pillars(voxels, pillar_size, f2[:, 1:3].min(0)[0] - 0.05))
    dims, P, voxels, pillars = fp
    print("   voxels", voxels.bxyz.shape[0], "pillars", dims)
    order = ((voxels.pillar_coords // 4)[:, 0] * 1000 + (voxels.pillar_coords // 4)[:, 1]).argsort()
    cc = (voxels.pillar_coords // 4)
    cd = cc.max(0)[0] + 1
    cidx = (cc[:, 0] * cd[1] + cc[:, 1])
    order = cidx.argsort()
    from pcseqlearning_b200.utils.scatter import scatter_max, scatter_min
    C = int(cd[0] * cd[1])
    zz = voxels.bxyz[order, 3]
    r = timeit("ground: ransac kernel only", lambda: ops.ground_ransac(voxels.bxyz[order], cidx[order], C, scatter_min(zz, cidx[order], C), scatter_max(zz, cidx[order], C), torch.linspace(0.3, 1, 30), cfg.SIGMA2))
    print("   IRLS iterations per ratio", r[3].tolist(), "total", int(r[3].sum()), "C", C)
    timeit("ground: ransac (all)", lambda: gu.compute_min_height_from_ransac(dims, P, voxels, pillars, cfg))
    timeit("ground: l1", lambda: gu.l1_minimization(pillars, dims, cfg))
    timeit("ground: total", lambda: gu.ground_plane_removal(f2, cfg))
    timeit("full step", lambda: model(batch))
    seq = model.forward_dict["sequences"][0]
    pts = seq["point_fxyz"]
    for r in (1.25, 0.75, 0.25):
        timeit(f"cluster_labels r={r}", lambda: ops.cluster_labels(pts, r, 32, chunk=10, num_frames=frames))

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
