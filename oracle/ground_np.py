"""CPU restatement of the reference's sequence-level ground height estimation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
pcdet/models/registration/preprocessors/preprocessor_utils.py:21-419 of /root/reference line by line on CPU
tensors (numpy voxelization from oracle/cpu_ops.py, torch CPU for eigh / AdamW exactly as the reference uses
them).  Pinned by tests/golden/ground.npz, recorded from the reference's own ground_plane_removal.
"""
import numpy as np
import torch

from . import cpu_ops as ops


def _scatter(src, idx, n, reduce):
    return torch.from_numpy(ops.scatter(src.numpy(), idx.numpy(), n, reduce))


def _irls(xyz, pidx, w0, P, sigma2, delta=1e-2):
    """iterative_reweighted_ransac (:32-80)."""
    w = w0
    err = center = normal = None
    for _ in range(50):
        center = _scatter(xyz * w, pidx, P, "sum") / (_scatter(w, pidx, P, "sum") + 1e-6)  # :47-59
        d = xyz - center[pidx]  # :61
        ddT = (w[:, :, None] * d[:, :, None]) * d[:, None, :]  # :62
        cov = _scatter(ddT.reshape(-1, 9), pidx, P, "mean").reshape(P, 3, 3)  # :63-68
        _, Q = torch.linalg.eigh(cov)  # :70
        normal = Q[:, :, 0]
        err = (d * normal[pidx]).sum(-1).abs()  # :72
        nw = sigma2 / (err.square() + sigma2)  # :73
        dw = 0.25 / (d.square().sum(-1) + 0.25)  # :74
        nw = (nw * dw).reshape(-1, 1)
        if (nw - w).abs().max() < delta:  # :76-78
            break
        w = nw
    return err, center, normal


def ground_plane_removal(point_fxyz, cfg):
    """ground_plane_removal (:352-419) without warm start.  point_fxyz: numpy f32[N,4].
    Returns numpy (height[N], horizon[N], error[N], pillar_height[X,Y], pillar_min_z[X,Y])."""
    pts = np.ascontiguousarray(point_fxyz, np.float32)
    pc_min = torch.from_numpy(pts[:, 1:3].min(0) - np.float32(0.05))  # :367
    z0 = pts.copy()
    z0[:, 0] = 0  # :24
    vox, inv = ops.grid_sampling(z0, [0.10, 0.10, 0.03])  # :369
    vox = torch.from_numpy(vox)
    psize = torch.tensor(cfg["PILLAR_SIZE"], dtype=torch.float32)
    # format_pillars (:274-311)
    pc = torch.div(vox[:, 1:3] - pc_min, psize, rounding_mode="floor").round().long()
    dims = pc.max(0)[0] + 1
    X, Y = int(dims[0]), int(dims[1])
    P = X * Y
    pidx = pc[:, 0] * Y + pc[:, 1]
    density = _scatter(torch.ones(vox.shape[0]), pidx, P, "sum").reshape(X, Y)
    min_z = _scatter(vox[:, -1].contiguous(), pidx, P, "min").reshape(X, Y)
    pxyz = _scatter(vox[:, 1:].contiguous(), pidx, P, "mean").reshape(-1, 3)
    weight = (density > 0.5).float().reshape(-1)
    height = None
    if cfg.get("RANSAC", False):  # compute_min_height_from_ransac (:83-272)
        ws = 4
        ar = torch.arange(P)
        ccoord = torch.stack([ar // Y, ar % Y], -1) // ws  # :92-94
        cd = ccoord.max(0)[0] + 1
        CY = int(cd[1])
        C = int(cd[0] * cd[1])
        cidx = (pc // ws)[:, 0] * CY + (pc // ws)[:, 1]  # :110-111
        order = cidx.argsort()  # :112
        cxyz, cidx = vox[order, 1:].contiguous(), cidx[order]
        z = cxyz[:, -1].contiguous()
        cmin, cmax = _scatter(z, cidx, C, "min"), _scatter(z, cidx, C, "max")  # :114-117
        bconf = torch.zeros(C)
        bnormal = torch.zeros(C, 3)
        bnormal[:, -1] = 1.0
        bcenter = torch.zeros(C, 3)
        for ratio in torch.linspace(0.3, 1, 30):  # :147
            cur = cmin * ratio + cmax * (1 - ratio)
            w0 = (cfg["SIGMA2"] / ((cur[cidx] - z).square() + cfg["SIGMA2"])).reshape(-1, 1)  # :151-152
            err, center, normal = _irls(cxyz, cidx, w0, C, cfg["SIGMA2"])
            nhit = _scatter((err < cfg["SIGMA2"] ** 0.5).float(), cidx, C, "sum")  # :163-165
            m = bconf < nhit
            bnormal[m], bcenter[m], bconf[m] = normal[m], center[m], nhit[m]
        xyz, normal = bcenter, bnormal
        K = cfg["K"]
        for thr in np.logspace(np.log(5) / np.log(10), np.log(0.01) / np.log(10), 100):  # :179
            dmat = torch.cdist(xyz.double(), xyz.double())
            nb = dmat.topk(min(K, xyz.shape[0]), dim=1, largest=False).indices
            e0 = torch.arange(xyz.shape[0])[:, None].expand_as(nb).reshape(-1)
            e1 = nb.reshape(-1)
            diff = xyz[e1] - xyz[e0]
            curv = (diff * normal[e0]).sum(-1).abs() / (diff.norm(dim=-1) + 1e-4)
            mc = curv.reshape(-1, K).mean(-1)
            if thr > mc.max():
                continue
            keep = mc < thr
            xyz, normal = xyz[keep], normal[keep]
        bc = torch.zeros(P)
        pcenter = torch.zeros(P, 3)
        pnormal = torch.zeros(P, 3)
        for i in range(xyz.shape[0]):  # :216-225
            ci = 1.0 / ((pxyz[:, :2] - xyz[i, :2]).norm(dim=-1) + 1)
            m = ci > bc
            pcenter[m], pnormal[m], bc[m] = xyz[i], normal[i], ci[m]
        vn, vc = pnormal[pidx], pcenter[pidx]
        vd = vox[:, 1:] - vc
        nz = vn[:, -1]
        vnz = nz.abs().clamp(min=0.01) * ((nz >= 0).float() + 1) / 2  # :241
        vh = (vd * vn).sum(-1) / vnz
        min_z = _scatter((vox[:, -1] - vh).contiguous(), pidx, P, "mean").reshape(X, Y)  # :253-254
        height = min_z.clone()
    if cfg.get("JointOpt", False):  # l1_minimization (:313-350)
        wt = weight.reshape(X, Y)
        h = torch.nn.Parameter(torch.zeros(X, Y))
        opt = torch.optim.AdamW([h], lr=cfg["LR"])
        sch = torch.optim.lr_scheduler.MultiStepLR(opt, cfg["DECAY_STEPS"])
        last, cd = 1e10, 3
        for _ in range(cfg["MAX_NUM_ITERS"]):
            opt.zero_grad()
            l1 = ((h - min_z) * wt).abs().mean()
            a = ((h[:-2] - 2 * h[1:-1] + h[2:]) * (wt[1:-1] + 1e-2)).abs().mean()
            b = ((h[:, :-2] - 2 * h[:, 1:-1] + h[:, 2:]) * (wt[:, 1:-1] + 1e-2)).abs().mean()
            c = ((h[:-2, :-2] - 2 * h[1:-1, 1:-1] + h[2:, 2:]) * (wt[1:-1, 1:-1] + 1e-2)).abs().mean()
            d = ((h[2:, :-2] - 2 * h[1:-1, 1:-1] + h[:-2, 2:]) * (wt[1:-1, 1:-1] + 1e-2)).abs().mean()
            loss = l1 + (a + b + c + d) * cfg["RIGID_WEIGHT"]
            loss.backward()
            opt.step()
            sch.step()
            if last - loss.item() < 1e-4:
                cd -= 1
            else:
                cd = 3
            if cd == 0:
                break
            last = loss.item()
        height = h.data.clone()
    vh = height[pc[:, 0], pc[:, 1]]
    vmin = min_z[pc[:, 0], pc[:, 1]]
    horizon = vox[:, -1] > vmin  # :412
    vh = vox[:, -1] - vh
    ferr = vh - vmin
    return (vh.numpy()[inv], horizon.numpy()[inv], ferr.numpy()[inv], height.numpy(), min_z.numpy())
