"""Developer experiment (CPU, oracle arithmetic): how fast do the per-super-pillar IRLS problems of the ground stage
converge individually?  The reference's stopping rule is global (max |dw| over ALL voxels < 1e-2), but every
super-pillar is an independent fixed-point iteration; this prints, per height ratio, when each pillar's plane stops
moving.  Usage: python tools/dev_irls_convergence.py [frames]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu_ops as ops  # noqa: E402
from oracle.ground_np import _scatter  # noqa: E402
from pcseqlearning_b200.synthetic import generate_sequence  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    batch = generate_sequence(0, num_frames=frames, device="cpu")
    fx = torch.cat([batch["point_sweep"].reshape(-1, 1).float(), batch["point_bxyz"][:, 1:]], -1).numpy()
    pick = ops.subsample_pick(fx)
    pts = np.ascontiguousarray(fx[pick])
    sigma2 = 0.0025
    pc_min = torch.from_numpy(pts[:, 1:3].min(0) - np.float32(0.05))
    z0 = pts.copy()
    z0[:, 0] = 0
    vox, _ = ops.grid_sampling(z0, [0.10, 0.10, 0.03])
    vox = torch.from_numpy(vox)
    pc = torch.div(vox[:, 1:3] - pc_min, torch.tensor([2.0, 2.0]), rounding_mode="floor").round().long()
    cd = (pc // 4).max(0)[0] + 1
    CY, C = int(cd[1]), int(cd[0] * cd[1])
    cidx = (pc // 4)[:, 0] * CY + (pc // 4)[:, 1]
    order = cidx.argsort()
    xyz, cidx = vox[order, 1:].contiguous(), cidx[order]
    z = xyz[:, -1].contiguous()
    cmin, cmax = _scatter(z, cidx, C, "min"), _scatter(z, cidx, C, "max")
    occupied = torch.bincount(cidx, minlength=C) > 0
    nvox = torch.bincount(cidx, minlength=C).float()
    print(f"{frames} frames: {vox.shape[0]} voxels, {int(occupied.sum())} occupied super-pillars of {C}")
    work_total = work_frozen = 0.0
    for ratio in torch.linspace(0.3, 1, 30)[::3]:
        cur = cmin * ratio + cmax * (1 - ratio)
        w = (sigma2 / ((cur[cidx] - z).square() + sigma2)).reshape(-1, 1)
        prev = None
        frozen_at = torch.full((C,), -1, dtype=torch.long)
        iters = 50
        for it in range(50):
            center = _scatter(xyz * w, cidx, C, "sum") / (_scatter(w, cidx, C, "sum") + 1e-6)
            d = xyz - center[cidx]
            ddT = (w[:, :, None] * d[:, :, None]) * d[:, None, :]
            cov = _scatter(ddT.reshape(-1, 9), cidx, C, "mean").reshape(C, 3, 3)
            _, Q = torch.linalg.eigh(cov)
            normal = Q[:, :, 0]
            plane = torch.cat([center, normal], -1)
            if prev is not None:
                # sign-insensitive plane change
                dn = torch.minimum((normal - prev[:, 3:]).abs().max(-1).values, (normal + prev[:, 3:]).abs().max(-1).values)
                change = torch.maximum((center - prev[:, :3]).abs().max(-1).values, dn)
                newly = (change < 1e-6) & (frozen_at < 0) & occupied
                frozen_at[newly] = it
            prev = plane
            err = (d * normal[cidx]).sum(-1).abs()
            nw = (sigma2 / (err.square() + sigma2) * 0.25 / (d.square().sum(-1) + 0.25)).reshape(-1, 1)
            gmax = (nw - w).abs().max()
            w = nw
            if gmax < 1e-2:
                iters = it + 1
                break
        fa = frozen_at[occupied]
        never = int((fa < 0).sum())
        eff = torch.where(frozen_at >= 0, frozen_at, torch.full_like(frozen_at, iters)).float()
        wt, wf = float((nvox * iters)[occupied].sum()), float((nvox * eff)[occupied].sum())
        work_total += wt
        work_frozen += wf
        q = np.percentile(np.where(fa.numpy() < 0, iters, fa.numpy()), [10, 50, 90])
        print(f"ratio {float(ratio):.2f}: global rule stops after {iters:2d} its; pillar freeze iteration p10/p50/p90 = "
              f"{q[0]:.0f}/{q[1]:.0f}/{q[2]:.0f}, never frozen {never}/{fa.numel()}; voxel-iterations with freeze "
              f"{100 * wf / wt:.0f} % of without")
    print(f"all sampled ratios: voxel-iterations with per-pillar freeze = {100 * work_frozen / work_total:.0f} % of the full sweeps")


if __name__ == "__main__":
    main()
