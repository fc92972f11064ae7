"""Sequence-level ground height field (mirror of pcdet/models/registration/preprocessors/preprocessor_utils.py).

Pipeline: time-agnostic 0.10 x 0.10 x 0.03 voxel de-duplication (voxelize kernels) -> 2 m
pillars -> IRLS plane fits on 8 m super-pillars for 30 height ratios -> kNN-curvature pruning of the planes
("truncated least squares") -> nearest-plane propagation to the pillars -> AdamW L1 smoothing of the pillar
height grid -> per-point height above ground.

The per-point work (voxelization, gathers) runs in the CUDA kernels; the pillar-level solvers are small dense
torch programs on a few thousand cells (SURVEY.md section 8f ranks moving them into fused kernels next).
"""
import numpy as np
import torch

from .. import _lib, ops, parallel
from ..utils import EasyDict
from ..utils.scatter import scatter_count, scatter_max, scatter_mean, scatter_min, scatter_sum


def grid_sample(point_fxyz, grid_size):
    """Voxel means with the frame column zeroed + point->voxel map (preprocessor_utils.py:21-30).

    Frame-window sharding: the de-duplication ignores time, so the voxels are sequence-global.  Every rank voxelizes
    its own frames on the grid of the whole sequence, the partial (column sums, count) per cell are all-gathered and
    merged by cell key; all ranks end up with the identical, ascending-key voxel list of a single-GPU run."""
    shard = parallel.SHARD
    if shard is None:
        res = ops.voxelize(point_fxyz, grid_size, ignore_dim0=True, want_mean=True)
        return EasyDict(bxyz=res["sampled"]), res["inv"]
    res = ops.voxelize(point_fxyz, grid_size, ignore_dim0=True, want_sums=True, want_counts=True,
                       bounds_hook=shard.reduce_bounds)
    keys, _ = shard.all_gather_v(res["keys"])
    sums, _ = shard.all_gather_v(res["sums"])
    cnts, _ = shard.all_gather_v(res["counts"].long())
    uk, inv = torch.unique(keys, return_inverse=True)
    gs = torch.zeros(uk.shape[0], 4, dtype=torch.float64, device=keys.device).index_add_(0, inv, sums)
    gc = torch.zeros(uk.shape[0], dtype=torch.int64, device=keys.device).index_add_(0, inv, cnts)
    bxyz = (gs / gc.clamp(min=1)[:, None].double()).float()
    point_voxel = torch.searchsorted(uk, res["keys"][res["inv"]])
    return EasyDict(bxyz=bxyz), point_voxel


def format_pillars(points, pillar_size, pc_range_min):
    """Per-pillar density / min z / mean xyz (preprocessor_utils.py:274-311)."""
    pillars = EasyDict()
    coords = torch.div(points.bxyz[:, 1:3] - pc_range_min, pillar_size, rounding_mode="floor").round().long()
    points["pillar_coords"] = coords
    pillar_dims = coords.max(0)[0].long() + 1
    num_pillars = int((pillar_dims[1] * pillar_dims[0]).item())
    X, Y = int(pillar_dims[0].item()), int(pillar_dims[1].item())
    points["pillar_idx"] = coords[:, 0] * pillar_dims[1] + coords[:, 1]
    pillars["density"] = scatter_count(points.pillar_idx, num_pillars).reshape(X, Y)
    pillars["min_z"] = scatter_min(points.bxyz[:, -1], points.pillar_idx, num_pillars).reshape(X, Y)
    pillars["xyz"] = scatter_mean(points.bxyz[:, 1:], points.pillar_idx, num_pillars).reshape(-1, 3)
    pillars["weight"] = (pillars.density > 0.5).float().reshape(-1)
    return (X, Y), num_pillars, points, pillars


def iterative_reweighted_plane_fit(xyz, pillar_idx, w0, num_pillars, sigma2, stopping_delta=1e-2, max_iter=50):
    """IRLS plane fit per super-pillar (preprocessor_utils.py:32-80).

    xyz [N,3], pillar_idx [N] (sorted by the caller as in the reference), w0 [N,1].
    Returns (plane_fitting_error [N], center [P,3], normal [P,3]).
    """
    w = w0
    cnt = scatter_count(pillar_idx, num_pillars).clamp(min=1)
    for _ in range(max_iter):
        center = scatter_sum(xyz * w, pillar_idx, num_pillars) / (scatter_sum(w, pillar_idx, num_pillars) + 1e-6)
        d = xyz - center[pillar_idx]
        ddT = (w[:, :, None] * d[:, :, None]) * d[:, None, :]
        cov = scatter_sum(ddT.reshape(-1, 9), pillar_idx, num_pillars).reshape(-1, 3, 3) / cnt[:, None, None]
        _, Q = torch.linalg.eigh(cov)
        normal = Q[:, :, 0]
        err = (d * normal[pillar_idx]).sum(-1).abs()
        new_w = sigma2 / (err.square() + sigma2)
        dist_w = (0.5 ** 2) / (d.square().sum(dim=-1) + 0.5 ** 2)
        new_w = (new_w * dist_w).reshape(-1, 1)
        if (new_w - w).abs().max() < stopping_delta:
            break
        w = new_w
    return err, center, normal


def _knn_self(xyz, k):
    """torch_cluster.knn(x, x, k) rows (query, neighbour), ascending distance, self included
    (preprocessor_utils.py:180; a few hundred super-pillars, brute force)."""
    d = torch.cdist(xyz, xyz)
    idx = d.topk(min(k, xyz.shape[0]), dim=1, largest=False).indices
    e0 = torch.arange(xyz.shape[0], device=xyz.device)[:, None].expand_as(idx)
    return e0.reshape(-1), idx.reshape(-1)


@torch.no_grad()
def compute_min_height_from_ransac(pillar_dims, num_pillars, voxels, pillars, cfg, window_size=4, use_kernels=True):
    """Plane-based estimate of the ground height of every pillar (preprocessor_utils.py:83-272)."""
    X, Y = pillar_dims
    dev = voxels.bxyz.device
    ar = torch.arange(num_pillars, device=dev)
    coarse_of_pillar = torch.stack([ar // Y, ar % Y], dim=-1) // window_size
    cdims = coarse_of_pillar.max(0)[0].long() + 1
    CY = int(cdims[1].item())
    num_coarse = int((cdims[1] * cdims[0]).item())
    pillars.best_confidence = torch.zeros_like(pillars.density).reshape(-1)
    pillars.best_normal = pillars.min_z.new_zeros(num_pillars, 3)
    pillars.best_center = pillars.min_z.new_zeros(num_pillars, 3)

    # voxels regrouped into super-pillars (sorted by super-pillar id like the reference, :107-112)
    ccoords = voxels.pillar_coords // window_size
    cidx = ccoords[:, 0] * CY + ccoords[:, 1]
    order = cidx.argsort()
    cxyz = voxels.bxyz[order, 1:]
    cidx = cidx[order]
    z = cxyz[:, -1]
    if use_kernels:
        c_min_z, c_max_z = ops.group_minmax(z, cidx, num_coarse)  # one pass, warp-reduced (rows are sorted by cidx)
    else:
        c_min_z = scatter_min(z, cidx, num_coarse)
        c_max_z = scatter_max(z, cidx, num_coarse)
    best_conf = torch.zeros_like(c_min_z)
    best_normal = c_min_z.new_zeros(num_coarse, 3)
    best_normal[:, -1] = 1.0
    best_center = c_min_z.new_zeros(num_coarse, 3)
    ratios = torch.linspace(0.3, 1, 30)
    shard = parallel.SHARD
    if use_kernels and shard is not None and shard.world > 1:
        # the 30 height ratios are independent IRLS problems: every rank solves a contiguous slice of them and the
        # per-super-pillar winners are merged in ratio order with the reference's strict '<' (first maximum wins)
        r0, r1 = (30 * shard.rank) // shard.world, (30 * (shard.rank + 1)) // shard.world
        vox_sorted = ops.gather_rows(voxels.bxyz, order)
        if r1 > r0:
            bc, bn, bf, _ = ops.ground_ransac(vox_sorted, cidx, num_coarse, c_min_z, c_max_z, ratios[r0:r1], cfg.SIGMA2)
        else:
            bc, bn, bf = best_center, best_normal, best_conf
        packed = torch.cat([bc, bn, bf[:, None]], 1).contiguous()
        for part in shard.all_gather(packed):
            better = best_conf < part[:, 6]
            best_normal = torch.where(better[:, None], part[:, 3:6], best_normal)
            best_center = torch.where(better[:, None], part[:, :3], best_center)
            best_conf = torch.where(better, part[:, 6], best_conf)
    elif use_kernels:
        # all 30 x <=50 IRLS iterations inside one persistent cooperative launch
        best_center, best_normal, best_conf, _ = ops.ground_ransac(ops.gather_rows(voxels.bxyz, order), cidx, num_coarse, c_min_z,
                                                                   c_max_z, ratios, cfg.SIGMA2)
    else:
        sigma = cfg.SIGMA2 ** 0.5
        for ratio in ratios:
            cur_z = c_min_z * ratio + c_max_z * (1 - ratio)
            z_diff = cur_z[cidx] - z
            w0 = (cfg.SIGMA2 / (z_diff.square() + cfg.SIGMA2)).reshape(-1, 1)
            err, center, normal = iterative_reweighted_plane_fit(cxyz, cidx, w0, num_coarse, cfg.SIGMA2)
            num_hit = scatter_sum((err < sigma).float(), cidx, num_coarse)
            better = best_conf < num_hit
            best_normal = torch.where(better[:, None], normal, best_normal)
            best_center = torch.where(better[:, None], center, best_center)
            best_conf = torch.where(better, num_hit, best_conf)

    # prune planes whose neighbourhood is curved ("Truncated Least Squares", :175-193)
    xyz, normal = best_center, best_normal
    K = cfg.K
    thresholds = np.logspace(np.log(5) / np.log(10), np.log(0.01) / np.log(10), 100)
    if use_kernels:
        # the whole 100-threshold loop in one launch; there is no silent eager fallback: sizes outside the kernel's
        # limits are an error (use_kernels=False is the torch mirror kept for the tests)
        if not (K <= xyz.shape[0] <= ops.PRUNE_MAX_PLANES and K <= 16):
            raise _lib.PcsError(f"plane pruning kernel limits exceeded: {xyz.shape[0]} planes (K <= n <= "
                                f"{ops.PRUNE_MAX_PLANES}), K = {K} (<= 16)")
        keep = ops.plane_prune(xyz, normal, K, thresholds)
        xyz, normal = xyz[keep], normal[keep]
        thresholds = []
    for threshold in thresholds:
        e0, e1 = _knn_self(xyz, K)
        diff = xyz[e1] - xyz[e0]
        p2p = (diff * normal[e0]).sum(dim=-1).abs()
        curvature = p2p / (diff.norm(p=2, dim=-1) + 1e-4)
        mean_curv = curvature.reshape(-1, K).mean(-1)
        if threshold > mean_curv.max():
            continue
        keep = mean_curv < threshold
        xyz, normal = xyz[keep], normal[keep]

    # every pillar takes the plane with the largest 1 / (dist + 1); sequential first-wins == first arg-max (:216-225)
    dist = (pillars.xyz[:, None, :2] - xyz[None, :, :2]).norm(p=2, dim=-1)
    conf_ind = 1.0 / (dist.pow(1.0) + 1)
    best = conf_ind.argmax(dim=1)
    pillars.best_center = xyz[best]
    pillars.best_normal = normal[best]
    pillars.best_confidence = conf_ind.gather(1, best[:, None])[:, 0]

    v_normal = pillars.best_normal[voxels.pillar_idx]
    v_center = pillars.best_center[voxels.pillar_idx]
    v_diff = voxels.bxyz[:, 1:] - v_center
    nz = v_normal[:, -1]
    v_normal_z = nz.abs().clamp(min=0.01) * ((nz >= 0).float() + 1) / 2
    v_height = (v_diff * v_normal).sum(-1) / v_normal_z
    pillars.min_z = scatter_mean(voxels.bxyz[:, -1] - v_height, voxels.pillar_idx, num_pillars).reshape(X, Y)
    pillars.height = pillars.min_z.clone()
    return voxels, pillars


def l1_minimization(pillars, pillar_dims, cfg, max_countdown=3, use_kernels=True):
    """AdamW L1 smoothing of the pillar height grid (preprocessor_utils.py:313-350)."""
    X, Y = pillar_dims
    if use_kernels:
        if not (X * Y <= ops.L1_MAX_CELLS and X >= 3 and Y >= 3 and len(cfg.DECAY_STEPS) <= 1):
            raise _lib.PcsError(f"L1 height-field kernel limits exceeded: grid {X} x {Y} (3 <= X, Y; X * Y <= "
                                f"{ops.L1_MAX_CELLS}), {len(cfg.DECAY_STEPS)} LR milestones (<= 1)")
        # the whole optimisation (<= MAX_NUM_ITERS AdamW steps + stopping rule) in one single-CTA launch
        pillars["height"], pillars["l1_info"] = ops.l1_heightfield(pillars.min_z, pillars.weight, cfg.LR,
                                                                   list(cfg.DECAY_STEPS), cfg.RIGID_WEIGHT,
                                                                   cfg.MAX_NUM_ITERS)
        return pillars
    weight = pillars.weight.reshape(X, Y)
    min_z = pillars.min_z
    h = torch.nn.Parameter(torch.zeros(X, Y, device=min_z.device), requires_grad=True)
    opt = torch.optim.AdamW([h], lr=cfg.LR)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, cfg.DECAY_STEPS)
    last_loss = 1e10
    countdown = max_countdown
    wi = weight[1:-1] + 1e-2
    wj = weight[:, 1:-1] + 1e-2
    wc = weight[1:-1, 1:-1] + 1e-2
    with torch.enable_grad():
        for _ in range(cfg.MAX_NUM_ITERS):
            opt.zero_grad()
            l1 = ((h - min_z) * weight).abs().mean()
            left = ((h[:-2] - 2 * h[1:-1] + h[2:]) * wi).abs().mean()
            up = ((h[:, :-2] - 2 * h[:, 1:-1] + h[:, 2:]) * wj).abs().mean()
            t1 = ((h[:-2, :-2] - 2 * h[1:-1, 1:-1] + h[2:, 2:]) * wc).abs().mean()
            t2 = ((h[2:, :-2] - 2 * h[1:-1, 1:-1] + h[:-2, 2:]) * wc).abs().mean()
            loss = l1 + (left + up + t1 + t2) * cfg.RIGID_WEIGHT
            loss.backward()
            opt.step()
            sched.step()
            cur = loss.item()
            if (last_loss - cur) < 1e-4:
                countdown -= 1
            else:
                countdown = 3
            if countdown == 0:
                break
            last_loss = cur
    pillars["height"] = h.data.clone()
    return pillars


def ground_plane_removal(point_fxyz, cfg, warmup=None, use_kernels=True):
    """Per-point height above the estimated ground (preprocessor_utils.py:352-419).

    Returns (height [N], horizon [N] bool, fitting_error [N], pillar_height [X,Y], pillar_min_z [X,Y]).
    """
    import os
    import time
    timing = os.environ.get("PCS_STAGE_TIMING")
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    pillar_size = torch.tensor(cfg.PILLAR_SIZE).to(point_fxyz)
    pc_range_min = point_fxyz[:, 1:3].min(0)[0] if point_fxyz.shape[0] else point_fxyz.new_full((2,), 1e30)
    if parallel.SHARD is not None:
        pc_range_min = -parallel.SHARD.all_reduce_max(-pc_range_min)
    pc_range_min = pc_range_min - 0.05
    voxels, point_voxel_index = grid_sample(point_fxyz, [0.10, 0.10, 0.03])
    mark("grid_sample")
    pillar_dims, num_pillars, voxels, pillars = format_pillars(voxels, pillar_size, pc_range_min)
    mark("format_pillars")
    if warmup is not None:
        pillars.height = warmup["pillar_height"]
        pillars.min_z = warmup["pillar_min_z"]
    else:
        if cfg.get("RANSAC", False):
            voxels, pillars = compute_min_height_from_ransac(pillar_dims, num_pillars, voxels, pillars, cfg,
                                                             use_kernels=use_kernels)
        mark("ransac")
        if cfg.get("JointOpt", False):
            pillars = l1_minimization(pillars, pillar_dims, cfg, use_kernels=use_kernels)
        mark("l1")
        if "height" not in pillars:
            pillars.height = pillars.min_z.clone()
    cx, cy = voxels.pillar_coords[:, 0], voxels.pillar_coords[:, 1]
    v_height = pillars.height[cx, cy]
    v_min_z = pillars.min_z[cx, cy]
    v_horizon = voxels.bxyz[:, -1] > v_min_z
    v_height = voxels.bxyz[:, -1] - v_height
    fitting_error = v_height - v_min_z
    if timing:
        mark("end")
        print("[ground] " + ", ".join(f"{n} {1e3 * (t - marks[i][1]):.1f} ms" for i, (n, t) in enumerate(marks[1:])),
              flush=True)
    return (ops.gather_rows(v_height, point_voxel_index), ops.gather_rows(v_horizon, point_voxel_index),
            ops.gather_rows(fitting_error, point_voxel_index), pillars.height, pillars.min_z)
