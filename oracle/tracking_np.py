"""numpy restatement of the reference's cluster tracker, GT evaluation and trace extraction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ and by bench.py's CPU baseline / --impl reference.

Follows, line by line, /root/reference/pcdet/models/registration/preprocessors/cluster_tracking.py
  sample_frame :39-51, dist_compensate :80-87, component_diameter / component_center :89-121,
  filter_components :123-148, smooth_velo :162-199 (torch.optim.AdamW + MultiStepLR restated in numpy),
  track_frame :430-787, extract_traces_and_update_boxes :287-428, forward :789-915
and cluster_proposal.py evaluate_proposal :142-285.  Pinned by tests/golden/tracking.npz and
tests/golden/eval_tracking.npz (recorded from the reference's own Python, oracle/gen_golden.py).
"""
import numpy as np

from . import cpu_ops as ops
from . import registration_np as reg

STATIONARY_DIAMETER = 12.5


def dist_compensate(comp_deg):
    thresholds = [0, 10, 40, 100, 200, 400, 10000000]
    bonus = [1, 0.5, 0.3, 0.2, 0.1, 0.0]
    out = np.zeros(comp_deg.shape, np.float32)
    for lo, hi, b in zip(thresholds[:-1], thresholds[1:], bonus):
        out[(comp_deg >= lo) & (comp_deg < hi)] = b
    return out


def component_center(xyz, comp, n):
    deg = np.bincount(comp, minlength=n).astype(np.float32)
    s = ops.scatter(np.asarray(xyz, np.float32), comp, n, "sum")
    ok = deg > 0.5
    s[ok] = s[ok] / deg[ok, None]
    return s


def component_diameter(xyz, comp, n):
    c = component_center(xyz, comp, n)
    d = np.linalg.norm(np.asarray(xyz, np.float32) - c[comp], axis=-1).astype(np.float32)
    return ops.scatter(d, comp, n, "max") * 2


def smooth_velo(velos, diffs, frame_id, next_frame_id, weight0=1.0, weight=10.0, num_itr=300, stopping=1e-3):
    """cluster_tracking.py:162-199 -- AdamW (lr 1e-2, betas 0.9/0.999, eps 1e-8, weight decay 1e-2, decoupled, applied
    to EVERY element of the parameter tensor) with MultiStepLR([100, 200, 300]); updates `velos` in place."""
    if frame_id == next_frame_id:
        return velos
    a, b = (frame_id, next_frame_id) if frame_id < next_frame_id else (next_frame_id, frame_id)
    lr, beta1, beta2, eps, wd = 1e-2, 0.9, 0.999, 1e-8, 1e-2
    m = np.zeros_like(velos)
    v = np.zeros_like(velos)
    last, countdown = 1e10, 3
    n1 = np.float32(velos.shape[0] * (b - a + 1) * 2)
    n2 = np.float32(velos.shape[0] * (b - a) * 2)
    for it in range(num_itr):
        g = np.zeros_like(velos)
        r = velos[:, a:b + 1, :2] - diffs[:, a:b + 1, :2]
        loss_fit = np.float32(np.square(r, dtype=np.float32).mean(dtype=np.float32))
        g[:, a:b + 1, :2] += np.float32(weight0) * 2 * r / n1
        d = velos[:, a:b, :2] - velos[:, a + 1:b + 1, :2]
        loss_smooth = np.float32(np.abs(d).mean(dtype=np.float32))
        sg = np.sign(d).astype(np.float32) * np.float32(weight) / n2
        g[:, a:b, :2] += sg
        g[:, a + 1:b + 1, :2] -= sg
        loss = float(np.float32(loss_fit * np.float32(weight0) + loss_smooth * np.float32(weight)))
        velos *= np.float32(1 - lr * wd)
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        bc1, bc2 = 1 - beta1 ** (it + 1), 1 - beta2 ** (it + 1)
        velos -= (np.float32(lr / bc1) * (m / (np.sqrt(v) / np.float32(np.sqrt(bc2)) + np.float32(eps)))).astype(np.float32)
        if it + 1 in (100, 200, 300):
            lr *= 0.1
        if last - loss < stopping:
            countdown -= 1
        else:
            countdown = 3
        if countdown <= 0:
            break
        last = loss
    return velos


def track_frame(seq_fxyz, seq_frame, seq_comp, seq_stationary, anchor, cfg, seq_seg=None, trace=None):
    """cluster_tracking.py:430-787 for one anchor frame.  seq_comp: component ids of the whole sequence (one key).
    Returns dict(fxyz, component, frame_indices, original_indices, moving, transforms[, segmentation_label])."""
    radius_list = cfg["radius"]
    voxel_list = cfg["voxel_size"]
    deltas = cfg["stopping_delta"]
    interval, min_move = cfg["track_interval"], cfg["min_move_frame"]
    coeff, angle_thr, angle_reg = cfg["reg_error_coeff"], cfg["angle_threshold"], cfg["angle_regularizer"]
    nn_radius = cfg["nn_radius"]
    seq_frame = np.asarray(seq_frame).reshape(-1).astype(np.int64)
    fmask = seq_frame == anchor
    frows = np.nonzero(fmask)[0]
    fxyz = np.array(seq_fxyz[fmask], np.float32, copy=True)
    comp = np.asarray(seq_comp[fmask], np.int64)
    comp = comp - comp.min()
    stat = np.asarray(seq_stationary[fmask], bool)
    C = int(comp.max()) + 1
    fmin = max(int(seq_frame.min()), anchor - interval)
    fmax = min(int(seq_frame.max()), anchor + interval)
    comp_deg = np.bincount(comp, minlength=C)
    comp_diam = component_diameter(fxyz[:, 1:], comp, C)
    transforms = np.tile(np.eye(4), (C, fmax - fmin + 1, 1, 1))
    comp_min = np.full(C, anchor, np.int64)
    comp_max = np.full(C, anchor, np.int64)
    velos = np.zeros((C, fmax + 1, 3), np.float32)
    centers = np.zeros((C, fmax + 1, 3), np.float32)
    centers[:, anchor] = component_center(fxyz[:, 1:], comp, C)
    diffs = np.zeros((C, fmax + 1, 3), np.float32)
    valid = (comp_deg > 0) & (comp_diam < STATIONARY_DIAMETER)
    vp = valid[comp]
    ex = dict(fxyz=[fxyz[vp].copy()], component=[comp[vp]], frame_indices=[np.nonzero(vp)[0]],
              original_indices=[frows[vp]])
    last_velo_idx = None  # the reference's last_velo is a VIEW of velos[:, idx]
    moving = valid.copy()
    orig = fxyz.copy()
    for d in (-1, 1):
        nxt = anchor + d
        stopped = ~valid
        moving = valid.copy()
        last_xyz = fxyz[:, 1:].copy()
        if d == 1 and anchor > 0:
            last_velo_idx = anchor
        while fmin <= nxt <= fmax and (~stopped).any():
            nmask = seq_frame == nxt
            nrows = np.nonzero(nmask)[0]
            nf = np.asarray(seq_fxyz[nmask], np.float32)
            nstat = np.asarray(seq_stationary[nmask], bool)
            ncomp = np.asarray(seq_comp[nmask], np.int64)
            k = nxt - fmin
            transforms[:, k] = transforms[:, k - d]
            if last_velo_idx is not None:
                trans = velos[:, last_velo_idx].copy()
                trans[stopped] = 0
                fxyz[:, 1:] += trans[comp] * d
                transforms[:, k, :3, 3] += trans.astype(np.float64) * d
            ratio0 = l1_last = None
            for i, radius in enumerate(radius_list):
                sub = reg.sample_frame(fxyz, stat, comp, np.full(fxyz.shape[0], anchor), voxel_list[i])
                subn = reg.sample_frame(nf, nstat, ncomp, np.full(nf.shape[0], nxt), voxel_list[i])
                _, T, l1, ratio, n_it = reg.register_to_next_frame(
                    sub["fxyz"], sub["component"], sub["stationary"], subn["fxyz"], subn["stationary"], C, radius,
                    angle_reg, 80, deltas[i])
                if trace is not None:
                    trace.append(n_it)
                if i == 0:
                    ratio0 = ratio
                if i == len(radius_list) - 1:
                    l1_last = l1
                rot = np.einsum("nij,nj->ni", T[comp, :3, :3], fxyz[:, 1:].astype(np.float64)).astype(np.float32)
                fxyz[:, 1:] = rot + T[comp, :3, 3].astype(np.float32)
                transforms[:, k] = T @ transforms[:, k]
            centers[:, nxt] = component_center(fxyz[:, 1:], comp, C)
            pv = (fxyz[:, 1:] - last_xyz) * d
            cv = ops.scatter(pv.astype(np.float32), comp, C, "mean")
            cv[:, 2] = 0
            velos[:, nxt] = cv
            diffs[:, nxt] = (centers[:, nxt] - centers[:, nxt - d]) * d
            smooth_velo(velos, diffs, anchor + d, nxt)
            delta = velos[:, nxt] - cv
            cv = velos[:, nxt]  # a view, like the reference's comp_velo
            fxyz[:, 1:] += delta[comp] * d
            transforms[:, k, :3, 3] += (delta * d).astype(np.float64)
            last_xyz = fxyz[:, 1:].copy()
            stopped = stopped | (l1_last > (np.float32(coeff) * comp_diam * (1 + dist_compensate(comp_deg))).astype(np.float64))
            stopped = stopped | (ratio0 < 0.5)
            if (nxt - anchor) * d == min_move:
                moving = moving & (np.linalg.norm(centers[:, nxt] - centers[:, anchor], axis=-1) > np.float32(0.08) * comp_diam)
            if last_velo_idx is not None:
                lv = velos[:, last_velo_idx]
                stopped = stopped | (np.linalg.norm(cv - lv, axis=-1) > np.float32(0.24) * comp_diam)
                pvv = velos[:, nxt - d]
                norm = np.maximum(np.linalg.norm(cv, axis=-1) * np.linalg.norm(pvv, axis=-1), np.float32(1e-6))
                ang = np.arccos(np.clip((cv * pvv).sum(-1) / norm, -1, 1)) / np.float32(np.pi) * np.float32(180.0)
                stopped = stopped | ((ang > angle_thr) & (np.linalg.norm(velos[:, nxt, :2], axis=-1) > 0.01))
            last_velo_idx = nxt
            if nxt == anchor - 1:
                velos[:, anchor] = cv
            if d == -1:
                comp_min[~stopped] = nxt
            else:
                comp_max[~stopped] = nxt
            q = fxyz.copy()
            q[:, 0] = nxt
            f_this, f_next = ops.radius_graph_build(q, nf, nn_radius, 1, True)
            keep = (~stopped)[comp[f_this]]
            f_this, f_next = f_this[keep], f_next[keep]
            ex["fxyz"].append(nf[f_next])
            ex["component"].append(comp[f_this])
            ex["frame_indices"].append(f_next)
            ex["original_indices"].append(nrows[f_next])
            nxt += d
        fxyz = orig.copy()
    out = {k: np.concatenate(v) for k, v in ex.items()}
    out["moving"] = moving[out["component"]]
    valid = valid & ((comp_max >= anchor + min_move) | (comp_min <= anchor - min_move))
    sel = valid[out["component"]]
    out = {k: v[sel] for k, v in out.items()}
    if seq_seg is not None:
        out["segmentation_label"] = np.asarray(seq_seg)[out["original_indices"]]
    out["transforms"] = transforms
    return out


def evaluate_proposal(fxyz, comps, box_attr, box_frame, box_trace):
    """cluster_proposal.py:142-285 (per frame, per component loops) -> dict of the emitted arrays."""
    frame = np.rint(fxyz[:, 0]).astype(np.int64)
    best = np.zeros(box_attr.shape[0], np.float32)
    tbest = np.zeros(int(box_trace.max()) + 1, np.float32)
    n = fxyz.shape[0]
    gt_box = np.full(n, -1, np.int64)
    gt_trace = np.full(n, -1, np.int64)
    pred_box = pred_trace = None
    for comp in comps:
        pred_box = np.full(n, -1, np.int64)
        pred_trace = np.full(n, -1, np.int64)
        for f in range(int(frame.max()) + 1):
            pm, bm = frame == f, box_frame == f
            if not pm.any() or not bm.any():
                continue
            bidx = np.nonzero(bm)[0]
            bp = ops.points_in_boxes(np.ascontiguousarray(fxyz[pm, 1:]), np.ascontiguousarray(box_attr[bm]))
            inb = (bp == 1).any(0)
            g = np.full(int(pm.sum()), -1, np.int64)
            g[inb] = bp[:, inb].argmax(0)
            gt_box[pm] = g
            gt = np.full(int(pm.sum()), -1, np.int64)
            gt[inb] = box_trace[bidx[g[inb]]]
            gt_trace[pm] = gt
            c_f = comp[pm]
            pb = np.full(int(pm.sum()), -1, np.int64)
            order = np.argsort(c_f, kind="stable")
            cs = c_f[order]
            starts = np.nonzero(np.r_[True, cs[1:] != cs[:-1]])[0]
            ends = np.r_[starts[1:], cs.shape[0]]
            for s0, s1 in zip(starts, ends):
                rows = order[s0:s1]
                cnt = bp[:, rows].sum(-1)
                if not cnt.any():
                    continue
                b = int(cnt.argmax())
                pb[rows] = b
                m1 = g == b
                inter = int(m1[rows].sum())
                union = int(m1.sum()) + rows.shape[0] - inter
                iou = np.float32(inter / (union + 1e-6))
                best[bidx[b]] = max(best[bidx[b]], iou)
                tbest[box_trace[bidx[b]]] = max(tbest[box_trace[bidx[b]]], iou)
            pred_box[pm] = pb
            pt = np.full(int(pm.sum()), -1, np.int64)
            pt[pb >= 0] = box_trace[bidx[pb[pb >= 0]]]
            pred_trace[pm] = pt
    return dict(gt_box_best_iou=best, gt_trace_best_iou=tbest, point_gt_box_id=gt_box, point_gt_trace_id=gt_trace,
                point_pred_box_id=pred_box, point_pred_trace_id=pred_trace)


def extract_traces(all_fxyz, all_frame, ex, box_attr, box_frame, best_iou, nn_radius):
    """cluster_tracking.py:287-428 -> (full dict, best_iou updated in place)."""
    all_frame = np.asarray(all_frame).reshape(-1).astype(np.int64)
    C = int(ex["component"].max()) + 1
    fcol = ex["fxyz"][:, 0]
    hit = np.zeros(C, np.int64)
    size = (np.rint(ops.scatter(fcol, ex["component"], C, "max")) - np.rint(ops.scatter(fcol, ex["component"], C, "min"))
            ).astype(np.int64) + 1
    exf = np.rint(fcol).astype(np.int64)
    out = dict(fxyz=[], component=[], frame_indices=[], original_indices=[], moving=[])
    radius = nn_radius * 1.732
    for fid in np.unique(exf):
        bm = box_frame == fid
        bidx = np.nonzero(bm)[0]
        rm = all_frame == fid
        rrows = np.nonzero(rm)[0]
        ref = np.asarray(all_fxyz[rm], np.float32)
        if bm.any():
            bp = ops.points_in_boxes(np.ascontiguousarray(ref[:, 1:]), np.ascontiguousarray(box_attr[bm]))
            gt = bp.argmax(0)
            gt[bp.max(0) == 0] = -1
        om = exf == fid
        one = ex["fxyz"][om]
        oc = ex["component"][om]
        e_ext, e_ref = ops.radius_graph_build(one, ref, radius, 1, True)
        ctr = ops.scatter(np.ascontiguousarray(one[:, 1:3]), oc, C, "mean")
        diam = ops.scatter(np.linalg.norm(one[:, 1:3] - ctr[oc], axis=-1).astype(np.float32), oc, C, "max")
        dz = one[e_ext, 3] - ref[e_ref, 3]
        ok = dz < 0.5
        ok &= np.linalg.norm(ref[e_ref, 1:3] - ctr[oc[e_ext]], axis=-1) < diam[oc[e_ext]] + np.float32(0.05)
        ok &= dz > -0.05
        e_ext, e_ref = e_ext[ok], e_ref[ok]
        cur_c = oc[e_ext]
        out["fxyz"].append(ref[e_ref])
        out["component"].append(cur_c)
        out["frame_indices"].append(e_ref)
        out["original_indices"].append(rrows[e_ref].reshape(-1, 1))
        out["moving"].append(ex["moving"][om][e_ext])
        if bm.any() and e_ref.size:
            for c in np.unique(cur_c):
                cm = cur_c == c
                cnt = bp[:, e_ref[cm]].sum(-1)
                if not cnt.any():
                    continue
                b = int(cnt.argmax())
                m1 = gt == b
                mask = np.zeros(ref.shape[0], bool)
                mask[e_ref[cm]] = True
                iou = np.float32((mask & m1).sum() / ((mask | m1).sum() + 1e-6))
                if iou > 0.7:
                    hit[c] += 1
                best_iou[bidx[b]] = max(best_iou[bidx[b]], iou)
    full = {k: np.concatenate(v) for k, v in out.items()}
    full["component_hit"], full["component_size"] = hit, size
    return full


def tracking_cfg():
    """The TRACKING / REGISTRATION block of cluster_tracking_TLS_multiradius_every8.yaml."""
    return dict(radius=[2.5, 1.25, 1.0], voxel_size=[[0.4, 0.4, 0.6], [0.2, 0.2, 0.3], [0.1, 0.1, 0.15]],
                stopping_delta=[0.05, 0.05, 0.05], track_interval=8, min_move_frame=6, reg_error_coeff=0.13,
                angle_threshold=45, angle_regularizer=10, nn_radius=0.5)
