"""Synthetic Waymo-shaped lidar sequences.

A spinning 64-beam lidar is ray cast against a ground height field, static boxes (parked cars,
walls, poles), moving vehicles and pedestrians while the ego vehicle drives a gently curving path.
Every frame is expressed in the pose of the LAST frame, like the reference's sequence loader
(pcdet/datasets/waymo/waymo_dataset.py:575-598), and the result is returned as the collated
``batch_dict`` the model plugin receives (pcdet/datasets/dataset.py:194-298).

Scene parameters come from ``numpy.random.Generator(PCG64(seed))`` with
``seed = 20221000 + sequence_index`` (+500 for the dense variant); per-ray noise/dropout from a
``torch.Generator`` seeded the same way on the target device.  The ray casting is written in torch
so the 198-frame sequences of the benchmark are generated on the GPU in seconds; it is input
plumbing, not part of the measured path.
"""
import math

import numpy as np
import torch

# Waymo segmentation ids used on the path (ground_plane_remover.py:161-168): 1..7 foreground,
# >= 17 ground-like.
SEG_VEHICLE, SEG_PEDESTRIAN, SEG_BUILDING, SEG_POLE, SEG_GROUND = 1, 7, 14, 11, 18
CLS_VEHICLE, CLS_PEDESTRIAN = 1, 2


def beam_elevations(num_beams=64):
    """Waymo-like non-uniform elevations in [-17.6, +2.4] degrees (denser near the horizon)."""
    u = np.linspace(0.0, 1.0, num_beams)
    deg = -17.6 + 20.0 * (1.0 - (1.0 - u) ** 1.6)
    return np.deg2rad(deg)


def _height_field(x, y, hf):
    h = torch.zeros_like(x)
    for ax, ay, ph, amp in hf:
        h = h + amp * torch.sin(ax * x + ay * y + ph)
    return h


def make_scene(seed, num_frames, dense=False):
    """Scene description (numpy, world frame; z up, ground near z=0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sc = {}
    # ego path: speed 8..12 m/s, small constant yaw rate
    speed = rng.uniform(8.0, 12.0)
    yaw_rate = rng.uniform(-0.03, 0.03)
    t = np.arange(num_frames) * 0.1
    yaw = yaw_rate * t
    if abs(yaw_rate) > 1e-6:
        ex = speed / yaw_rate * np.sin(yaw)
        ey = speed / yaw_rate * (1 - np.cos(yaw))
    else:
        ex, ey = speed * t, np.zeros_like(t)
    sc["ego"] = np.stack([ex, ey, yaw], 1)
    # low frequency ground height field, +-0.3 m, wavelength ~40 m
    hf = []
    for _ in range(4):
        lam = rng.uniform(30.0, 60.0)
        th = rng.uniform(0, 2 * np.pi)
        hf.append((2 * np.pi / lam * np.cos(th), 2 * np.pi / lam * np.sin(th), rng.uniform(0, 2 * np.pi),
                   rng.uniform(0.03, 0.09)))
    sc["hf"] = hf
    path_len = speed * t[-1]
    boxes = []  # [x, y, z, dx, dy, dz, yaw, vx, vy, yawrate, seg, cls, moving]
    # static furniture is spread along the driven corridor (+-80 m beyond both ends) at a fixed density,
    # which gives 60..120 static boxes for the 198-frame sequences of the benchmark
    region = path_len + 160.0
    n_static = int(region / rng.uniform(3.2, 5.5)) * (2 if dense else 1)
    for i in range(n_static):
        kind = rng.choice(3, p=[0.5, 0.25, 0.25])
        s = rng.uniform(-80.0, path_len + 80.0)
        side = rng.choice([-1.0, 1.0])
        if kind == 0:  # parked car
            lat = side * rng.uniform(4.0, 9.0)
            boxes.append([s, lat, 0.8, 4.5, 1.9, 1.6, rng.normal(0, 0.05), 0, 0, 0, SEG_VEHICLE, CLS_VEHICLE, 0])
        elif kind == 1:  # wall / building
            lat = side * rng.uniform(15.0, 50.0)
            boxes.append([s, lat, 3.0, rng.uniform(8, 25), rng.uniform(4, 12), 6.0, rng.normal(0, 0.1), 0, 0, 0,
                          SEG_BUILDING, 0, 0])
        else:  # pole
            lat = side * rng.uniform(5.0, 12.0)
            boxes.append([s, lat, 2.5, 0.3, 0.3, 5.0, 0.0, 0, 0, 0, SEG_POLE, 0, 0])
    n_veh = int(rng.integers(10, 31)) * (2 if dense else 1)
    for i in range(n_veh):
        s = rng.uniform(-60.0, path_len + 60.0)
        lane = rng.choice([-3.5, 3.5, 7.0])
        v = rng.uniform(0.0, 15.0) * (1.0 if lane > 0 else -1.0)
        hd = 0.0 if v >= 0 else np.pi
        boxes.append([s, lane, 0.8, 4.6, 1.9, 1.6, hd, v, 0.0, rng.uniform(-0.02, 0.02), SEG_VEHICLE, CLS_VEHICLE,
                      1 if abs(v) > 0.05 else 0])
    n_ped = int(rng.integers(10, 31)) * (2 if dense else 1)
    for i in range(n_ped):
        s = rng.uniform(-40.0, path_len + 40.0)
        lat = rng.choice([-1.0, 1.0]) * rng.uniform(3.0, 14.0)
        sp = rng.uniform(0.0, 2.0)
        th = rng.uniform(0, 2 * np.pi)
        boxes.append([s, lat, 0.85, 0.6, 0.6, 1.7, th, sp * np.cos(th), sp * np.sin(th), 0.0, SEG_PEDESTRIAN,
                      CLS_PEDESTRIAN, 1 if sp > 0.05 else 0])
    sc["boxes"] = np.asarray(boxes, np.float64)
    return sc


def _boxes_at(sc, f):
    """Box states at frame f: position advanced by velocity, heading by yaw rate."""
    b = sc["boxes"].copy()
    t = 0.1 * f
    b[:, 0] += b[:, 7] * t
    b[:, 1] += b[:, 8] * t
    b[:, 6] += b[:, 9] * t
    return b


def _raycast_frame(sc, f, dirs_v, gen, device, max_range, sensor_h, noise, dropout):
    """Cast the sensor rays of frame f.  Returns world-frame hits + labels (torch, on device)."""
    ex, ey, eyaw = sc["ego"][f]
    c, s = math.cos(eyaw), math.sin(eyaw)
    Rw = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64, device=device)
    d = dirs_v @ Rw.T  # world directions [R,3]
    o = torch.tensor([ex, ey, 0.0], dtype=torch.float64, device=device)
    oz = _height_field(o[0:1], o[1:2], sc["hf"])[0] + sensor_h
    o = torch.stack([o[0], o[1], oz])
    R = d.shape[0]
    # ground: solve o_z + t d_z = h(x(t), y(t)) by fixed point iterations from the flat-plane hit
    tg = torch.full((R,), float("inf"), dtype=torch.float64, device=device)
    down = d[:, 2] < -1e-3
    t0 = (-oz) / d[:, 2].clamp(max=-1e-3)
    for _ in range(3):
        hx = o[0] + t0 * d[:, 0]
        hy = o[1] + t0 * d[:, 1]
        t0 = (_height_field(hx, hy, sc["hf"]) - oz) / d[:, 2].clamp(max=-1e-3)
    tg = torch.where(down, t0, tg)
    best_t = tg
    best_id = torch.full((R,), -1, dtype=torch.int64, device=device)
    boxes = torch.from_numpy(_boxes_at(sc, f)).to(device)
    # only boxes that can be in range
    near = ((boxes[:, 0] - ex) ** 2 + (boxes[:, 1] - ey) ** 2).sqrt() < max_range + 30.0
    idx_near = near.nonzero().reshape(-1)
    for j0 in range(0, idx_near.numel(), 32):
        ids = idx_near[j0:j0 + 32]
        bb = boxes[ids]
        cb, sb = torch.cos(bb[:, 6]), torch.sin(bb[:, 6])
        # ray in box frame: rotate by -yaw about z
        rel = o[None, :] - bb[:, :3]  # [B,3]
        ox = cb * rel[:, 0] + sb * rel[:, 1]
        oy = -sb * rel[:, 0] + cb * rel[:, 1]
        ozb = rel[:, 2]
        dx = d[:, None, 0] * cb[None] + d[:, None, 1] * sb[None]  # [R,B]
        dy = -d[:, None, 0] * sb[None] + d[:, None, 1] * cb[None]
        dz = d[:, None, 2].expand(-1, ids.numel())
        tmin = torch.full_like(dx, 0.0)
        tmax = torch.full_like(dx, float("inf"))
        for oo, dd, hh in ((ox, dx, bb[:, 3] / 2), (oy, dy, bb[:, 4] / 2), (ozb, dz, bb[:, 5] / 2)):
            inv = 1.0 / torch.where(dd.abs() < 1e-12, torch.full_like(dd, 1e-12), dd)
            t1 = (-hh[None] - oo[None]) * inv
            t2 = (hh[None] - oo[None]) * inv
            tmin = torch.maximum(tmin, torch.minimum(t1, t2))
            tmax = torch.minimum(tmax, torch.maximum(t1, t2))
        hit = (tmax >= tmin) & (tmin > 0.5)
        tb = torch.where(hit, tmin, torch.full_like(tmin, float("inf")))
        tbest, jbest = tb.min(1)
        upd = tbest < best_t
        best_t = torch.where(upd, tbest, best_t)
        best_id = torch.where(upd, ids[jbest], best_id)
    rng_noise = torch.randn(R, generator=gen, device=device, dtype=torch.float32).double() * noise
    keep = torch.rand(R, generator=gen, device=device, dtype=torch.float32) >= dropout
    t = best_t + rng_noise
    valid = torch.isfinite(best_t) & (best_t < max_range) & (best_t > 1.0) & keep
    pts = o[None, :] + t[:, None] * d
    return pts[valid], best_id[valid], t[valid], boxes


def generate_sequence(sequence_index=0, num_frames=16, num_beams=64, num_azimuth=2650, dense=False,
                      device="cpu", max_range=75.0, sensor_height=2.0, noise=0.02, dropout=0.05,
                      max_objects=None):
    """Return the collated ``batch_dict`` (batch_size = 1) for one synthetic sequence.

    Tensors live on ``device``; shapes/dtypes follow SURVEY Appendix B.
    """
    device = torch.device(device)
    seed = 20221000 + sequence_index + (500 if dense else 0)
    if dense:
        num_azimuth = int(num_azimuth * 3900 / 2650)
    sc = make_scene(seed, num_frames, dense=dense)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    el = torch.from_numpy(beam_elevations(num_beams)).to(device)
    az = torch.arange(num_azimuth, device=device, dtype=torch.float64) * (2 * math.pi / num_azimuth)
    ce = torch.cos(el)
    dirs = torch.stack([ce[:, None] * torch.cos(az)[None], ce[:, None] * torch.sin(az)[None],
                        torch.sin(el)[:, None].expand(-1, num_azimuth)], -1).reshape(-1, 3)
    # last-frame vehicle pose (origin on the ground under the sensor)
    lx, ly, lyaw = sc["ego"][-1]
    lz = float(_height_field(torch.tensor([lx], dtype=torch.float64), torch.tensor([ly], dtype=torch.float64),
                             sc["hf"])[0])
    cl, sl = math.cos(lyaw), math.sin(lyaw)
    Rl = torch.tensor([[cl, -sl, 0.0], [sl, cl, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64, device=device)
    tl = torch.tensor([lx, ly, lz], dtype=torch.float64, device=device)
    nb = sc["boxes"].shape[0]
    xyz_l, sweep_l, feat_l, seg_l, inst_l = [], [], [], [], []
    box_attr = np.zeros((num_frames, nb, 7), np.float32)
    box_cls = np.zeros((num_frames, nb), np.int32)
    box_npts = np.zeros((num_frames, nb), np.int64)
    poses = np.zeros((num_frames, 4, 4), np.float64)
    for f in range(num_frames):
        pts, bid, rng_t, boxes = _raycast_frame(sc, f, dirs, gen, device, max_range, sensor_height, noise, dropout)
        local = (pts - tl[None]) @ Rl  # world -> last-frame vehicle pose
        seg = torch.full((pts.shape[0],), SEG_GROUND, dtype=torch.int64, device=device)
        inst = torch.zeros(pts.shape[0], dtype=torch.int64, device=device)
        hitb = bid >= 0
        seg[hitb] = boxes[bid[hitb], 10].long()
        fg = hitb & (boxes[bid.clamp(min=0), 11] > 0)
        inst[fg] = bid[fg] + 1
        inten = torch.tanh(torch.rand(pts.shape[0], generator=gen, device=device))
        elong = torch.zeros_like(inten)
        feat = torch.stack([inten, elong, (rng_t / 75.0).float()], -1)  # waymo_dataset.py:334-343
        xyz_l.append(local.float())
        sweep_l.append(torch.full((pts.shape[0], 1), f, dtype=torch.int32, device=device))
        feat_l.append(feat.float())
        seg_l.append(seg)
        inst_l.append(inst)
        # GT boxes of this frame in the last-frame pose (objects only), empty boxes zeroed
        b = boxes.cpu().numpy()
        ctr = (b[:, :3] - np.array([lx, ly, lz])) @ Rl.cpu().numpy()
        cnt = torch.bincount(bid[hitb], minlength=nb).cpu().numpy()
        is_obj = b[:, 11] > 0
        vis = is_obj & (cnt >= 5)
        box_attr[f, vis, :3] = ctr[vis]
        box_attr[f, vis, 3:6] = b[vis, 3:6]
        box_attr[f, vis, 6] = b[vis, 6] - lyaw
        box_cls[f, vis] = b[vis, 11].astype(np.int32)
        box_npts[f] = np.where(vis, cnt, 0)
        ex, ey, eyaw = sc["ego"][f]
        poses[f] = np.eye(4)
        poses[f, :2, :2] = [[math.cos(eyaw), -math.sin(eyaw)], [math.sin(eyaw), math.cos(eyaw)]]
        poses[f, :2, 3] = [ex, ey]
    # keep only object slots that are visible at least once (reference pads to a fixed max per frame)
    obj_slots = np.where((box_cls > 0).any(0))[0]
    if max_objects is not None:
        obj_slots = obj_slots[:max_objects]
    box_attr, box_cls, box_npts = box_attr[:, obj_slots], box_cls[:, obj_slots], box_npts[:, obj_slots]
    nobj = obj_slots.shape[0]
    xyz = torch.cat(xyz_l, 0)
    seqname = f"segment-synthetic{sequence_index:05d}{'d' if dense else ''}"
    obj_ids = np.array([f"obj{int(s):04d}" for s in obj_slots] * num_frames).astype(str)
    attr_t = torch.from_numpy(box_attr.reshape(1, num_frames * nobj, 7)).to(device)
    batch = dict(
        batch_size=1,
        point_bxyz=torch.cat([torch.zeros(xyz.shape[0], 1, device=device), xyz], -1).contiguous(),
        point_sweep=torch.cat(sweep_l, 0),
        point_feat=torch.cat(feat_l, 0),
        segmentation_label=torch.cat(seg_l, 0),
        instance_label=torch.cat(inst_l, 0),
        is_foreground=(torch.cat(inst_l, 0) > 0),
        gt_box_attr=attr_t,
        gt_boxes=attr_t,
        gt_box_cls_label=torch.from_numpy(box_cls.reshape(1, -1)).to(device),
        gt_box_corners_3d=torch.zeros(1, num_frames * nobj, 8, 3, device=device),
        augmented=torch.zeros(1, num_frames * nobj, dtype=torch.bool, device=device),
        num_points_in_gt=torch.from_numpy(box_npts.reshape(1, -1)).to(device),
        obj_ids=[obj_ids],
        frame_id=[np.array([f"{seqname}_{f:03d}" for f in range(num_frames)])],
        pose=[poses],
        num_sweeps=[num_frames],
    )
    return batch


def sequence_fxyz(batch):
    """(frame, x, y, z) float32 rows the preprocessors work on (simple_reg.py:115-117)."""
    return torch.cat([batch["point_sweep"].reshape(-1, 1).float(), batch["point_bxyz"][:, 1:]], -1).contiguous()
