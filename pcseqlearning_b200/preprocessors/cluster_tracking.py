"""ClusterTracking preprocessor (mirror of pcdet/models/registration/preprocessors/cluster_tracking.py).

For every component key and every TRACK_INTERVAL-th frame, all clusters of that anchor frame are tracked
+-TRACK_INTERVAL frames by 3-level coarse-to-fine ICP (one persistent kernel launch per level), constant
velocity prediction with AdamW velocity smoothing, the reference's stopping tests, and nearest-neighbour point
extraction; results are written in the reference's .pth layout.
"""
import os

import numpy as np
import torch
from torch import nn

from .. import _lib, graph_utils, ops, parallel
from ..grid_sampling import GridSampling3D
from ..utils import EasyDict, Timer, filter_dict
from ..utils.scatter import scatter_max, scatter_min
from .eval_utils import points_in_boxes
from .registration_utils import efficient_robust_sum, register_to_next_frame, robust_mean

STATIONARY_DIAMETER = 12.5  # cluster_tracking.py:861 / filter_components default


def sample_frame(grid_sampler, frame):
    """Voxel down-sample of one frame: mean fxyz, majority `stationary`, upper-median component and frame id
    (cluster_tracking.py:39-51)."""
    res = grid_sampler.voxelize(frame.fxyz, want_mean=True, want_counts=True)
    inv, n = res["inv"], res["num"]
    out = EasyDict(dict())
    out["fxyz"] = res["sampled"]
    cnt = res["counts"].float().clamp(min=1)
    stat = torch.zeros(n, device=inv.device).index_add_(0, inv, frame.stationary.float())
    out["stationary"] = (stat / cnt) > 0.5
    out["component"] = ops.group_median(frame.component.long().reshape(-1), inv, n, res["counts"])
    out["frame"] = ops.group_median(frame.frame.long().reshape(-1), inv, n, res["counts"])
    return out


def dist_compensate(comp_deg):
    """Tolerance bonus for components with few points (cluster_tracking.py:80-87)."""
    thresholds = [0, 10, 40, 100, 200, 400, 10000000]
    bonus = [1, 0.5, 0.3, 0.2, 0.1, 0.0]
    out = torch.zeros_like(comp_deg).float()
    for lo, hi, b in zip(thresholds[:-1], thresholds[1:], bonus):
        out[(comp_deg >= lo) & (comp_deg < hi)] = b
    return out


def component_center(frame_points):
    """Mean xyz per component, empty components stay 0 (cluster_tracking.py:109-121)."""
    xyz = frame_points.fxyz[:, 1:]
    n = int(frame_points.component.max().long().item()) + 1
    deg = efficient_robust_sum(torch.ones_like(xyz[:, 0]), frame_points.component, n)
    center = efficient_robust_sum(xyz, frame_points.component, n)
    ok = deg > 0.5
    center[ok] = center[ok] / deg[ok, None]
    return center


def component_diameter(frame_points):
    """Twice the largest distance of a member to its component centre (cluster_tracking.py:89-107)."""
    xyz = frame_points.fxyz[:, 1:]
    n = int(frame_points.component.max().long().item()) + 1
    center = component_center(frame_points)
    dist = (xyz - center[frame_points.component]).norm(p=2, dim=-1)
    return scatter_max(dist, frame_points.component, n) * 2


def filter_components(frame_points, max_diameter=STATIONARY_DIAMETER):
    """valid = non-empty and diameter < max_diameter (cluster_tracking.py:123-148)."""
    n = int(frame_points.component.max().long().item()) + 1
    deg = torch.bincount(frame_points.component, minlength=n)
    valid = deg > 0
    if max_diameter > 0:
        valid = valid & (component_diameter(frame_points) < max_diameter)
    return valid


def smooth_velo(_comp_velos, comp_center_diffs, frame_id, next_frame_id, weight0=1, weight=10, num_itr=300,
                stopping=1e-3, use_kernels=True):
    """AdamW smoothing of the per-component xy velocities over [frame_id, next_frame_id]
    (cluster_tracking.py:162-199)."""
    if frame_id == next_frame_id:
        return _comp_velos
    if frame_id > next_frame_id:
        frame_id, next_frame_id = next_frame_id, frame_id
    n_opt = _comp_velos.shape[0] * (next_frame_id - frame_id + 1) * 2
    if use_kernels and _comp_velos.is_cuda:
        if not (_comp_velos.is_contiguous() and n_opt <= ops.SMOOTH_VELO_MAX):
            # sequential mirror only (the product path is tracker.TrackBatch, whose smoothing kernel has no such
            # limit); no silent eager fallback
            raise _lib.PcsError(f"smooth_velo kernel limit exceeded: {n_opt} optimised values (<= {ops.SMOOTH_VELO_MAX})")
        # the whole optimisation (<= 300 AdamW steps + stopping rule) in one single-CTA launch, in place like the
        # reference's nn.Parameter that shares storage with the tensor
        ops.smooth_velo(_comp_velos, comp_center_diffs, frame_id, next_frame_id, weight0, weight, num_itr, stopping)
        return _comp_velos
    velos = nn.Parameter(_comp_velos, requires_grad=True)
    opt = torch.optim.AdamW([velos], lr=1e-2)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [100, 200, 300])
    last_loss, countdown = 1e10, 3
    a, b = frame_id, next_frame_id
    with torch.enable_grad():
        for _ in range(num_itr):
            opt.zero_grad()
            fit = (velos[:, a:(b + 1), :2] - comp_center_diffs[:, a:(b + 1), :2]).square().mean()
            smooth = (velos[:, a:b, :2] - velos[:, (a + 1):(b + 1), :2]).abs().mean()
            loss = fit * weight0 + smooth * weight
            loss.backward()
            opt.step()
            sched.step()
            cur = loss.item()
            if last_loss - cur < stopping:
                countdown -= 1
            else:
                countdown = 3
            if countdown <= 0:
                break
            last_loss = cur
    return velos.data


class _stage_profile:
    """Developer aid: PCS_PROFILE_STAGE=<stage> prints the torch ops of that tracker stage by device time and shapes."""

    def __init__(self, name):
        self.on = os.environ.get("PCS_PROFILE_STAGE") == name
        self.name = name

    def __enter__(self):
        if self.on:
            from torch.profiler import ProfilerActivity, profile
            torch.cuda.synchronize()
            self.prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True)
            self.prof.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.synchronize()
            self.prof.__exit__(*exc)
            rows = []
            for ev in self.prof.key_averages(group_by_input_shape=True):
                t = getattr(ev, "self_device_time_total", None)
                if t is None:
                    t = getattr(ev, "self_cuda_time_total", 0.0)
                if t > 0:
                    rows.append((t, ev.count, ev.key, str(ev.input_shapes)[:110]))
            rows.sort(reverse=True)
            print(f"[stage profile: {self.name}] total {sum(r[0] for r in rows) / 1e3:.1f} ms device", flush=True)
            for t, n, key, shp in rows[:28]:
                print(f"  {t / 1e3:7.2f} ms n={n:4d} {key[:48]:48s} {shp}", flush=True)
        return False


class ClusterTracking(nn.Module):
    def __init__(self, model_cfg, runtime_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.forward_dict = EasyDict()
        reg = model_cfg.REGISTRATION
        self.stopping_delta = reg["STOPPING_DELTA"]
        graph_cfg = reg["GRAPH"]
        self.radius_list = graph_cfg["RADIUS"]
        self.voxel_size_list = reg["VOXEL_SIZE"]
        for i, (radius, voxel_size) in enumerate(zip(self.radius_list, self.voxel_size_list)):
            cfg_i = dict(graph_cfg)
            cfg_i["RADIUS"] = radius
            self.add_module(f"registration_graph_{i}", graph_utils.build_graph(cfg_i, {}))
            self.add_module(f"sampler_{i}", GridSampling3D(list(voxel_size)))
        self.nn_graph = graph_utils.build_graph(model_cfg["NN_GRAPH"], runtime_cfg=runtime_cfg)
        params = model_cfg.get("TRACKING_PARAMS", {})
        self.reg_error_coeff = params.get("REGISTRATION_ERROR_COEFFICIENT", 0.13)
        self.track_interval = params.get("TRACK_INTERVAL", 10)
        self.angle_threshold = params.get("ANGLE_THRESHOLD", 45)
        self.min_move_frame = params.get("MIN_MOVE_FRAME", 6)
        self.component_keys = model_cfg["COMPONENT_KEYS"]
        self.verbose = model_cfg.get("VERBOSE", True)
        # BATCHED (default): all (component key, anchor) pairs advance together on the device (tracker.TrackBatch);
        # False runs the reference's one-pair-at-a-time control flow (kept for the parity tests)
        self.batched = bool(model_cfg.get("BATCHED", True))

    def format_boxes(self, seq_dict, num_frames):
        return EasyDict(dict(attr=seq_dict["gt_box_attr"].reshape(-1, 7),
                             cls_label=seq_dict["gt_box_cls_label"].reshape(-1),
                             frame=seq_dict["gt_box_frame"].reshape(-1),
                             trace_id=seq_dict["gt_box_track_label"].reshape(-1),
                             velo=seq_dict["gt_box_velo"].reshape(-1),
                             moving=seq_dict["moving"].reshape(-1)))

    # ------------------------------------------------------------------------------------------------------
    def _nn_edges(self, ref, query, radius_scale=1.0):
        """K=1 radius graph of the NN_GRAPH block; returns (ref_idx, query_idx) of the matched queries."""
        radius = float(self.nn_graph.radius) * radius_scale
        return ops.radius_graph(ref.fxyz, query.fxyz, radius, 1, True)

    def track_frame(self, seq_points, frame, seq_boxes):
        """Track every component of `frame` through [frame - interval, frame + interval]
        (cluster_tracking.py:430-787).  Returns the `extracted` dict (+ transforms f64[C, F, 4, 4])."""
        C = int(frame.component.max().long().item()) + 1
        frame_id = int(frame.frame.reshape(-1)[0].long().item())
        frame_mask = (seq_points.frame == frame_id).reshape(-1)
        fmin = max(int(seq_points.frame.min().long().item()), frame_id - self.track_interval)
        fmax = min(int(seq_points.frame.max().long().item()), frame_id + self.track_interval)
        comp_deg = efficient_robust_sum(torch.ones_like(frame.component), frame.component, C)
        comp_diameter = component_diameter(frame)
        dev = frame.fxyz.device

        transforms = torch.diag_embed(frame.fxyz.new_ones(C, fmax - fmin + 1, 4).double())
        comp_min_frame = frame.component.new_zeros(C) + frame_id
        comp_max_frame = frame.component.new_zeros(C) + frame_id
        comp_velos = frame.fxyz.new_zeros(C, fmax + 1, 3)
        comp_centers = frame.fxyz.new_zeros(C, fmax + 1, 3)
        comp_centers[:, frame_id] = component_center(frame)
        comp_center_diffs = frame.fxyz.new_zeros(C, fmax + 1, 3)

        valid_comp = filter_components(frame)
        valid_point = valid_comp[frame.component]
        frame_rows = frame_mask.nonzero().reshape(-1)
        ex_fxyz = [frame.fxyz[valid_point]]
        ex_comp = [frame.component[valid_point]]
        ex_seg = [frame.segmentation_label[valid_point]]
        ex_idx = [valid_point.nonzero().reshape(-1)]
        ex_orig = [frame_rows[valid_point]]

        last_velo = None
        moving = valid_comp.clone()
        for track_dir in (-1, 1):
            nxt = frame_id + track_dir
            stopped = ~valid_comp
            moving = valid_comp.clone()  # re-initialised per direction, as in the reference (:545-548)
            last_xyz = frame.fxyz[:, 1:].clone()
            if track_dir == 1 and frame_id > 0:
                last_velo = comp_velos[:, frame_id]
            while (fmin <= nxt <= fmax) and bool((~stopped).any()):
                next_mask = (seq_points.frame == nxt).reshape(-1)
                next_frame = EasyDict(filter_dict(seq_points, next_mask))
                next_rows = next_mask.nonzero().reshape(-1)
                k = nxt - fmin
                transforms[:, k] = transforms[:, k - track_dir]
                if last_velo is not None:  # constant-velocity prediction (:569-573)
                    trans = last_velo.clone()
                    trans[stopped] = 0
                    frame.fxyz[:, 1:] += trans[frame.component] * track_dir
                    transforms[:, k, :3, 3] += trans.double() * track_dir
                comp_edge_ratio = l1_reg_error = None
                for i in range(len(self.radius_list)):  # coarse-to-fine registration (:574-627)
                    sampler = getattr(self, f"sampler_{i}")
                    graph = getattr(self, f"registration_graph_{i}")
                    sub = sample_frame(sampler, frame)
                    sub_next = sample_frame(sampler, next_frame)
                    sub, T, l1_i, ratio_i = register_to_next_frame(
                        graph, sub, sub_next, C, self.model_cfg.ANGLE_REGULARIZER, max_iter=80,
                        stopping_delta=self.stopping_delta[i], frame_offset=nxt - frame_id)
                    if i == 0:
                        comp_edge_ratio = ratio_i
                    if i == len(self.radius_list) - 1:
                        l1_reg_error = l1_i
                    Rp = T[frame.component, :3, :3]
                    frame.fxyz[:, 1:] = (Rp @ frame.fxyz[:, 1:, None].double()).squeeze(-1).float() \
                        + T[frame.component, :3, 3].float()
                    transforms[:, k] = T @ transforms[:, k]
                comp_centers[:, nxt] = component_center(frame)

                # velocity estimate + smoothing (:631-642)
                point_velo = (frame.fxyz[:, 1:] - last_xyz) * track_dir
                comp_velo = robust_mean(point_velo, frame.component, C)
                comp_velo[:, 2] = 0
                comp_velos[:, nxt] = comp_velo
                comp_center_diffs[:, nxt] = (comp_centers[:, nxt] - comp_centers[:, nxt - track_dir]) * track_dir
                comp_velos = smooth_velo(comp_velos, comp_center_diffs, frame_id + track_dir, nxt)
                delta_velo = comp_velos[:, nxt] - comp_velo
                comp_velo = comp_velos[:, nxt]
                frame.fxyz[:, 1:] += delta_velo[frame.component] * track_dir
                transforms[:, k, :3, 3] += delta_velo * track_dir
                last_xyz = frame.fxyz[:, 1:].clone()

                # stopping tests (:675-691)
                stopped = stopped | (l1_reg_error > self.reg_error_coeff * comp_diameter * (1 + dist_compensate(comp_deg)))
                stopped = stopped | (comp_edge_ratio < 0.5)
                if (nxt - frame_id) * track_dir == self.min_move_frame:
                    travelled = (comp_centers[:, nxt] - comp_centers[:, frame_id]).norm(p=2, dim=-1)
                    moving = moving & (travelled > 0.08 * comp_diameter)
                if last_velo is not None:
                    dev_velo = (comp_velo - last_velo).norm(p=2, dim=-1)
                    stopped = stopped | (dev_velo > 0.24 * comp_diameter)
                    prev_v = comp_velos[:, nxt - track_dir]
                    norm = (comp_velo.norm(p=2, dim=-1) * prev_v.norm(p=2, dim=-1)).clamp(min=1e-6)
                    angle = ((comp_velo * prev_v).sum(-1) / norm).clamp(min=-1, max=1).arccos() / np.pi * 180.0
                    stopped = stopped | (angle > self.angle_threshold) & (comp_velos[:, nxt, :2].norm(p=2, dim=-1) > 0.01)
                last_velo = comp_velo
                if nxt == frame_id - 1:
                    comp_velos[:, frame_id] = comp_velo
                if track_dir == -1:
                    comp_min_frame[~stopped] = nxt
                else:
                    comp_max_frame[~stopped] = nxt

                # extraction: next-frame points take the component of their nearest moved anchor point (:710-721)
                frame.fxyz[:, 0] = nxt
                f_this, f_next = self._nn_edges(frame, next_frame)
                keep = (~stopped)[frame.component[f_this]]
                f_this, f_next = f_this[keep], f_next[keep]
                ex_fxyz.append(next_frame.fxyz[f_next])
                ex_comp.append(frame.component[f_this])
                ex_seg.append(next_frame.segmentation_label[f_next])
                ex_idx.append(f_next.reshape(-1))
                ex_orig.append(next_rows[f_next].reshape(-1))
                frame.fxyz[:, 0] = frame_id
                nxt += track_dir
            frame.fxyz = seq_points.fxyz[frame_mask]

        comp_cat = torch.cat(ex_comp, dim=0)
        extracted = EasyDict(dict(
            fxyz=torch.cat(ex_fxyz, dim=0), component=comp_cat, segmentation_label=torch.cat(ex_seg, dim=0),
            frame_indices=torch.cat(ex_idx, dim=0), original_indices=torch.cat(ex_orig, dim=0),
            moving=moving[comp_cat], valid_comp_mask=moving[comp_cat], gt_box_label=torch.zeros_like(comp_cat)))
        long_enough = (comp_max_frame >= frame_id + self.min_move_frame) | (comp_min_frame <= frame_id - self.min_move_frame)
        valid_comp = valid_comp & long_enough
        extracted = EasyDict(filter_dict(extracted, valid_comp[extracted.component]))
        extracted.transforms = transforms
        seq_points.extracted[extracted.original_indices] = True
        return extracted

    # ------------------------------------------------------------------------------------------------------
    def extract_traces_and_update_boxes(self, all_points, extracted, seq_boxes):
        """Re-associate the extracted points with ALL (above-ground) points of their frames and update the GT-box
        coverage bookkeeping (cluster_tracking.py:287-428)."""
        out = {k: [] for k in ["fxyz", "component", "segmentation_label", "instance_label", "original_indices",
                               "frame_indices", "moving"]}
        transforms = extracted.pop("transforms")
        dev = all_points.fxyz.device
        C = int(extracted.component.max().long().item()) + 1
        component_hit = extracted.component.new_zeros(C)
        fcol = extracted.fxyz[:, 0]
        size_min = scatter_min(fcol, extracted.component, C).round().long()
        size_max = scatter_max(fcol, extracted.component, C).round().long()
        component_size = size_max - size_min + 1
        ex_frame = fcol.round().long()
        for fid in ex_frame.unique().tolist():
            box_mask = (seq_boxes.frame == fid).reshape(-1)
            boxes = EasyDict(filter_dict(seq_boxes, box_mask))
            ref_mask = (all_points.frame == fid).reshape(-1)
            ref_pts = EasyDict(filter_dict(all_points, ref_mask))
            n_ref = int(ref_mask.sum().item())
            ref_rows = ref_mask.nonzero().reshape(-1)
            has_boxes = bool(box_mask.any())
            if has_boxes:
                bp = points_in_boxes(ref_pts.fxyz[:, 1:], boxes.attr)
                gt_box_id = bp.argmax(0)
                gt_box_id[bp.max(0)[0] == 0] = -1
            one = EasyDict(filter_dict(extracted, ex_frame == fid))
            e_ext, e_ref = self._nn_edges(one, ref_pts, radius_scale=1.732)  # :356-358
            ctr = robust_mean(one.fxyz[:, 1:3], one.component, C)
            diam = scatter_max((one.fxyz[:, 1:3] - ctr[one.component]).norm(p=2, dim=-1), one.component, C)
            dz = one.fxyz[e_ext, -1] - ref_pts.fxyz[e_ref, -1]
            ok = dz < 0.5
            dist = (ref_pts.fxyz[e_ref, 1:3] - ctr[one.component[e_ext]]).norm(p=2, dim=-1)
            ok &= dist < diam[one.component[e_ext]] + 0.05
            ok &= dz > -0.05
            e_ext, e_ref = e_ext[ok], e_ref[ok]
            cur = dict(fxyz=ref_pts.fxyz[e_ref], component=one.component[e_ext],
                       segmentation_label=ref_pts.full_segmentation_label[e_ref],
                       instance_label=ref_pts.full_instance_label[e_ref], frame_indices=e_ref,
                       original_indices=ref_rows[e_ref].reshape(-1, 1), moving=one.moving[e_ext])
            for key in out:
                out[key].append(cur[key])
            if has_boxes and e_ref.numel() > 0:
                # IoU of every extracted component with its majority box, all at once (:388-411)
                uniq, inst = torch.unique(cur["component"], return_inverse=True)
                bp_e = bp[:, e_ref]  # [B, n_e] membership of the extracted points
                counts = torch.zeros(bp.shape[0], uniq.shape[0], dtype=torch.long, device=dev)
                counts.index_add_(1, inst, bp_e)
                has = counts.sum(0) > 0
                assigned = counts.argmax(0)
                # mask = unique reference rows hit by the component (frame_indices may repeat)
                pair = torch.unique(inst * n_ref + e_ref)
                p_inst, p_row = pair // n_ref, pair % n_ref
                mask_size = torch.bincount(p_inst, minlength=uniq.shape[0])
                inter = torch.bincount(p_inst[gt_box_id[p_row] == assigned[p_inst]], minlength=uniq.shape[0])
                gt_size = torch.bincount(gt_box_id[gt_box_id >= 0], minlength=bp.shape[0])
                union = mask_size + gt_size[assigned] - inter
                iou = inter.float() / (union.float() + 1e-6)
                iou = torch.where(has, iou, torch.zeros_like(iou))
                component_hit[uniq[has & (iou > 0.7)]] += 1
                best = boxes.best_iou.clone()
                best.scatter_reduce_(0, assigned[has], iou[has].to(best), "amax")
                seq_boxes.best_iou[box_mask] = best
        full = EasyDict({k: torch.cat(v, dim=0) for k, v in out.items()})
        full.component_hit = component_hit
        full.component_size = component_size
        full.transforms = transforms
        return full, seq_boxes

    # ------------------------------------------------------------------------------------------------------
    def extract_traces_batched(self, tb, all_points, seq_boxes):
        """extract_traces_and_update_boxes (cluster_tracking.py:287-428) for ALL instances of a TrackBatch at once.

        The reference loops over instances x extracted frames with one nn_graph call, one CPU point-in-box test and a
        Python loop over components per iteration.  Here one grid holds the extracted points of every (instance,
        frame) group, one search gives every above-ground point of those frames its nearest extracted point
        (r = 1.732 * NN_GRAPH.RADIUS), and the gating / IoU bookkeeping are segment reductions over
        (component, frame) ids.  Returns {instance: full_extracted EasyDict} (instances without extracted points
        are absent) and updates seq_boxes.best_iou in place."""
        import ctypes
        from .. import _lib
        from ..ops import _ptr, _stream, next_pow2
        from ..tracker import REL, ANCHOR_REL
        L = _lib.lib()
        dev = all_points.fxyz.device
        fl = tb.flat_results()
        sb = fl["slot_bounds"]
        J, G, F = tb.J, tb.G, tb.F
        n_ext = int(fl.gid.shape[0])
        out = {}
        if n_ext == 0:
            return out
        # ---- (instance, relative frame) groups of the extracted points -----------------------------------------
        anchor_of_inst = torch.tensor(tb.inst_anchor_h, device=dev)
        k_ext = fl.frame - (anchor_of_inst[fl.inst] - ANCHOR_REL)  # relative frame index 0..16, ascending with frame
        grp_ext = (fl.inst * REL + k_ext).int().contiguous()
        gs_ext = fl.gid * REL + k_ext  # (component, frame) id
        GS = G * REL
        xy = fl.fxyz[:, 1:3]
        cnt_gs = torch.bincount(gs_ext, minlength=GS).clamp(min=1).float()
        ctr = torch.zeros(GS, 2, device=dev).index_add_(0, gs_ext, xy) / cnt_gs[:, None]  # robust_mean (:361)
        diam = torch.zeros(GS, device=dev).scatter_reduce_(0, gs_ext, (xy - ctr[gs_ext]).norm(p=2, dim=-1), "amax",
                                                           include_self=False)
        # ---- all above-ground points, grouped by frame --------------------------------------------------------
        a_frame = all_points.frame.reshape(-1).long()
        a_order = torch.argsort(a_frame, stable=True)
        a_sorted = ops.gather_rows(all_points.fxyz.contiguous(), a_order)  # (torch's row gather: 1 block per row)
        a_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
        a_off[1:] = torch.bincount(a_frame, minlength=F).cumsum(0)
        a_off_h = a_off.tolist()  # host sync
        # segments: one per (instance, frame) group that holds extracted points, ascending frame inside an instance
        seg_group, seg_qstart, seg_cnt = [], [], []
        slot_of_rel = [8 - k if k < 8 else (0 if k == 8 else k) for k in range(REL)]
        for j in range(J):
            a = tb.inst_anchor_h[j]
            for k in range(REL):
                t = slot_of_rel[k]
                if sb[j * REL + t + 1] - sb[j * REL + t] == 0:
                    continue
                f = a - ANCHOR_REL + k
                seg_group.append(j * REL + k)
                seg_qstart.append(a_off_h[f])
                seg_cnt.append(a_off_h[f + 1] - a_off_h[f])
        nseg = len(seg_group)
        seg_off_h = np.zeros(nseg + 1, dtype=np.int64)
        seg_off_h[1:] = np.cumsum(seg_cnt)
        total_q = int(seg_off_h[-1])
        if total_q >= 2 ** 31:
            raise _lib.PcsError("extract_traces_batched: more than 2^31 queries")
        seg_group_t = torch.tensor(seg_group, dtype=torch.int32, device=dev)
        seg_qstart_t = torch.tensor(seg_qstart, dtype=torch.int32, device=dev)
        seg_off_t = torch.from_numpy(seg_off_h).to(dev)
        seg_off32 = seg_off_t.int().contiguous()
        radius = float(np.float32(float(self.nn_graph.radius) * 1.732))  # :356-358
        cs = radius * 1.001
        lo = (ctypes.c_double * 3)(*tb.lo)
        H = next_pow2(max(2 * n_ext, 1024))
        table = torch.empty(H, 4, dtype=torch.int32, device=dev)
        g_sorted = torch.empty(n_ext, 4, dtype=torch.float32, device=dev)
        g_sidx = torch.empty(n_ext, dtype=torch.int32, device=dev)
        g_cells = torch.empty(n_ext, dtype=torch.int32, device=dev)
        g_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
        nn = torch.empty(max(total_q, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            s = _stream()
            _lib.check(L.pcs_trk_group_grid(s, _ptr(fl.fxyz.contiguous()), _ptr(grp_ext), n_ext, lo, cs, _ptr(table), H,
                                            _ptr(g_sorted), _ptr(g_sidx), _ptr(g_cells), _ptr(g_ctr)),
                       "pcs_trk_group_grid")
            _lib.check(L.pcs_trk_group_nn(s, _ptr(table), H, _ptr(g_sorted), _ptr(g_sidx), lo, cs, _ptr(a_sorted),
                                          _ptr(seg_qstart_t), _ptr(seg_off32), _ptr(seg_group_t), nseg, radius,
                                          _ptr(nn)), "pcs_trk_group_nn")
        nn = nn[:total_q]
        w = (nn >= 0).nonzero().reshape(-1)  # host sync; ascending (instance, frame, row)
        e = nn[w].long()  # extracted entry of the hit
        sg = torch.searchsorted(seg_off_t, w, right=True) - 1
        i_all = w - seg_off_t[sg]
        srow = seg_qstart_t.long()[sg] + i_all  # row among the frame-sorted all_points
        gs = gs_ext[e]
        ref_xyz = ops.gather_rows(a_sorted, srow)
        dz = fl.fxyz[e, 3] - ref_xyz[:, 3]
        ok = (dz < 0.5) & (dz > -0.05)
        ok &= (ref_xyz[:, 1:3] - ctr[gs]).norm(p=2, dim=-1) < diam[gs] + 0.05  # :365-370
        ok = ok.nonzero().reshape(-1)
        e, sg, i_all, srow, gs = e[ok], sg[ok], i_all[ok], srow[ok], gs[ok]
        rows = a_order[srow]  # rows of all_points
        inst_k = fl.inst[e]
        inst_bounds = torch.searchsorted(inst_k, torch.arange(J + 1, device=dev)).tolist()  # host sync
        full = EasyDict(dict(fxyz=ops.gather_rows(ref_xyz, ok), component=fl.component[e], frame_indices=i_all,
                             original_indices=rows.reshape(-1, 1), moving=fl.moving[e]))
        for key in ("segmentation_label", "instance_label"):
            if f"full_{key}" in all_points:
                full[key] = all_points[f"full_{key}"][rows]
        # ---- per-component bookkeeping --------------------------------------------------------------------------
        fcol = fl.fxyz[:, 0]
        g_fmin = torch.full((G,), 1e9, device=dev).scatter_reduce_(0, fl.gid, fcol, "amin")
        g_fmax = torch.full((G,), -1e9, device=dev).scatter_reduce_(0, fl.gid, fcol, "amax")
        g_has = g_fmax >= g_fmin
        g_size = torch.where(g_has, (g_fmax.round() - g_fmin.round() + 1), torch.ones_like(g_fmax)).long()  # :295-297
        g_hit = torch.zeros(G, dtype=torch.long, device=dev)
        if seq_boxes.attr.shape[0] > 0 and e.numel() > 0:
            from .eval_utils import FrameBoxes
            fb = FrameBoxes(seq_boxes.attr, seq_boxes.frame, F)
            nb = fb.B
            first_all, _ = fb.query(a_sorted)  # first box of every above-ground point
            first_all = first_all.long()
            a_frame_sorted = a_frame[a_order]
            in_box = first_all >= 0
            gt_sorted = fb.off[a_frame_sorted] + first_all
            gt_size = torch.bincount(gt_sorted[in_box], minlength=nb)
            _, (counts,) = fb.query(a_sorted, sel=srow.int().contiguous(), cids=[(gs, GS)], want_first=False)
            has = counts.sum(1) > 0
            box_of = counts.argmax(1)
            k_gs = torch.arange(GS, device=dev) % REL
            gid_gs = torch.div(torch.arange(GS, device=dev), REL, rounding_mode="floor")
            f_gs = (anchor_of_inst[tb.g_inst.long()][gid_gs] - ANCHOR_REL + k_gs).clamp(0, F - 1)
            assigned = (fb.off[f_gs] + box_of).clamp(max=nb - 1)
            mask_size = torch.bincount(gs, minlength=GS)
            hit_pt = has[gs] & (first_all[srow] == box_of[gs])
            inter = torch.bincount(gs[hit_pt], minlength=GS)
            union = mask_size + gt_size[assigned] - inter
            iou = torch.where(has, inter.float() / (union.float() + 1e-6), torch.zeros(GS, device=dev))  # :400-403
            g_hit = (has & (iou > 0.7)).reshape(G, REL).sum(1)
            best_sorted = seq_boxes.best_iou[fb.order].clone()
            best_sorted.scatter_reduce_(0, assigned[has], iou[has].to(best_sorted), "amax")
            seq_boxes.best_iou[fb.order] = best_sorted
        for j in range(J):
            b0, b1 = sb[j * REL], sb[(j + 1) * REL]
            if b1 == b0:
                continue
            i0, i1 = inst_bounds[j], inst_bounds[j + 1]
            fe = EasyDict({k: v[i0:i1] for k, v in full.items()})
            g0, g1 = tb.inst_goff_h[j], tb.inst_goff_h[j + 1]
            loc = tb.g_local[g0:g1]
            Cx = int(tb.flat_cmax[j]) + 1
            # components beyond the largest extracted id go to a dump slot (no boolean-mask indexing: that is a
            # device->host sync per instance)
            hit = torch.zeros(Cx + 1, dtype=torch.long, device=dev)
            size = torch.ones(Cx + 1, dtype=torch.long, device=dev)
            li = torch.clamp(loc, max=Cx)
            hit[li] = torch.where(loc < Cx, g_hit[g0:g1], torch.zeros_like(li))
            size[li] = torch.where(loc < Cx, g_size[g0:g1], torch.ones_like(li))
            fe.component_hit, fe.component_size = hit[:Cx], size[:Cx]
            out[j] = fe
        return out

    # ------------------------------------------------------------------------------------------------------
    def forward(self, seq_dict):
        seq_points = EasyDict(fxyz=seq_dict["point_fxyz"], frame=seq_dict["point_sweep"],
                              gt_box_id=seq_dict["point_gt_box_id"])
        for key in ["instance_label", "segmentation_label"]:
            if key in seq_dict:
                seq_points[key] = seq_dict[key]
        all_points = EasyDict(fxyz=seq_dict["full_point_fxyz"], frame=seq_dict["full_point_sweep"],
                              height=seq_dict["full_point_height"])
        for key in ["full_instance_label", "full_segmentation_label"]:
            if key in seq_dict:
                all_points[key] = seq_dict[key]
        all_points = EasyDict(filter_dict(all_points, seq_dict["full_point_height"] > 0.0))
        shard = parallel.SHARD
        num_frames = shard.F if shard is not None else int(seq_points.frame.max().long().item()) + 1
        sequence_id = seq_dict["frame_id"][0][:-4]
        outfolder = f"{self.model_cfg.DIR}/{sequence_id}"
        outpath = f"{outfolder}/all.pth"
        save = self.model_cfg.get("SAVE", True)
        if save and os.path.exists(outpath):
            return seq_dict
        if save:
            os.makedirs(outfolder, exist_ok=True)
        seq_boxes = self.format_boxes(seq_dict, num_frames)
        if seq_boxes.attr.shape[0] == 0:  # the reference tracks nothing without GT boxes (:835-836)
            return seq_dict
        seq_boxes.best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
        results = {}
        if self.batched:
            from ..tracker import TrackBatch
            import time as _time
            timing = os.environ.get("PCS_TRACK_TIMING")
            marks = []

            def mark(name):
                if timing:
                    torch.cuda.synchronize()
                    marks.append((name, _time.perf_counter()))

            mark("start")
            comps = [seq_dict[f"point_{k}"] for k in self.component_keys]
            anchors = None
            if shard is not None:
                # halo exchange (NCCL all-to-all over NVLink): every rank receives the frames
                # [first anchor - interval, last anchor + interval] of its block of tracking anchors
                pack = dict(fxyz=seq_points.fxyz)
                for i, c in enumerate(comps):
                    pack[f"c{i}"] = c
                for key in ["instance_label", "segmentation_label"]:
                    if key in seq_points:
                        pack[key] = seq_points[key]
                got, fr = shard.exchange_frames(pack, seq_points.frame.reshape(-1))
                seq_points = EasyDict(fxyz=got["fxyz"], frame=fr.reshape(-1, 1).to(seq_points.frame.dtype))
                for key in ["instance_label", "segmentation_label"]:
                    if key in got:
                        seq_points[key] = got[key]
                comps = [got[f"c{i}"] for i in range(len(comps))]
                apack = {k: v for k, v in all_points.items() if k != "frame"}
                agot, afr = shard.exchange_frames(apack, all_points.frame.reshape(-1))
                all_points = EasyDict(agot)
                all_points["frame"] = afr.reshape(-1, 1).to(seq_dict["full_point_sweep"].dtype)
                anchors = list(shard.anchors)
                mark("halo")
            if anchors is not None and (len(anchors) == 0 or seq_points.fxyz.shape[0] == 0):
                seq_dict["tracking_results"], seq_dict["tracking_boxes"] = {}, seq_boxes
                shard.all_reduce_max(seq_boxes.best_iou)
                return seq_dict
            with _stage_profile("setup"):
                tb = TrackBatch(seq_points.fxyz, seq_points.frame, comps, self.model_cfg, num_frames=num_frames,
                                anchors=anchors)
            mark("setup")
            tb.run()
            mark("run")
            with _stage_profile("results"):
                per_inst = tb.results(seg_label=seq_points.get("segmentation_label"))
                tb.check()
            mark("results")
            with _stage_profile("extract_traces"):
                full = self.extract_traces_batched(tb, all_points, seq_boxes)
            mark("extract_traces")
            lazy = self.model_cfg.get("LAZY_TRANSFORMS", False) and not save
            for ki, comp_key in enumerate(self.component_keys):
                for frame_id in tb.anchors:
                    j, extracted = per_inst[(ki, frame_id)]
                    if j in full:
                        extracted = full[j]
                    if not lazy:  # reference layout f64[C, frames, 4, 4]; the compact form stays in tracking_batch
                        extracted.transforms = tb.transforms(j)
                    if save:
                        torch.save(extracted, f"{outfolder}/{frame_id:03d}_{comp_key}.pth")
                    results[f"{frame_id:03d}_{comp_key}"] = extracted
            seq_dict["tracking_batch"] = tb
            if shard is not None:
                # every rank updated the boxes of the frames its anchors reach; max-merge like the reference's
                # sequential `if iou > best_iou` updates (:410-414).  Track ids (anchor frame, key, local component)
                # are globally unique, so merging them is a gather of the per-instance index.
                shard.all_reduce_max(seq_boxes.best_iou)
                idx = torch.tensor([[tb.inst_anchor_h[j], tb.inst_key_h[j], shard.rank,
                                     int(results[f"{tb.inst_anchor_h[j]:03d}_{self.component_keys[tb.inst_key_h[j]]}"]
                                         ["fxyz"].shape[0])] for j in range(tb.J)], dtype=torch.int64,
                                   device=seq_boxes.best_iou.device).reshape(-1, 4)
                seq_dict["tracking_index"], _ = shard.all_gather_v(idx)
            mark("assemble")
            if timing:
                names = ["count", "ranges", "scatter", "search", "solve", "stop", "final", "ratio"]
                for lv, pr in enumerate(tb.prof):
                    pr = pr.tolist()
                    print(f"[icp level {lv}] launches {pr[9]} iterations {pr[8]} thread-path {pr[10]} warp-path {pr[11]} "
                          f"unmatched {pr[12]} cached {pr[15]} " +
                          ", ".join(f"{n} {pr[i] / 1e6:.1f} ms" for i, n in enumerate(names)) +
                          f" | searches: same {pr[208]} changed {pr[209]} no-prev {pr[210]} first-it {pr[211]}", flush=True)
                    if os.environ.get("PCS_TRACK_TIMING") == "2":
                        print("   per-iteration search ms (summed over launches): " +
                              " ".join(f"{pr[16 + i] / 1e6:.1f}" for i in range(80)))
                        print("   per-iteration active queries (k): " +
                              " ".join(f"{pr[112 + i] // 1000}" for i in range(80)))
                print("[track timing] " + ", ".join(f"{n} {1e3 * (t - marks[i][1]):.1f} ms"
                                                    for i, (n, t) in enumerate(marks[1:])), flush=True)
        else:
            for comp_key in self.component_keys:
                seq_points.component = seq_dict[f"point_{comp_key}"]
                diam = component_diameter(seq_points)[seq_points.component]
                seq_points.component_diameter = diam
                seq_points.stationary = diam > STATIONARY_DIAMETER
                seq_points.extracted = torch.zeros_like(seq_points.fxyz[:, 0]).bool()
                for frame_id in range(0, num_frames, self.track_interval):
                    frame_mask = (seq_points.fxyz[:, 0] == frame_id).reshape(-1)
                    if not bool(frame_mask.any()):
                        continue
                    frame_points = EasyDict(filter_dict(seq_points, frame_mask))
                    frame_points.component = frame_points.component - frame_points.component.min()
                    with Timer(f"Tracking Frame {frame_id}", verbose=self.verbose):
                        extracted = self.track_frame(seq_points, frame_points, seq_boxes)
                    if extracted.fxyz.shape[0] > 0:
                        extracted, seq_boxes = self.extract_traces_and_update_boxes(all_points, extracted, seq_boxes)
                    if save:
                        torch.save(extracted, f"{outfolder}/{frame_id:03d}_{comp_key}.pth")
                    results[f"{frame_id:03d}_{comp_key}"] = extracted
        if save:
            torch.save(seq_boxes, outpath)
        seq_dict["tracking_results"] = results
        seq_dict["tracking_boxes"] = seq_boxes
        return seq_dict

    def get_output_feature_dim(self):
        return 0

    def extra_repr(self):
        return (f"reg_error_coeff={self.reg_error_coeff}, min_move_frame={self.min_move_frame}, "
                f"angle_threshold={self.angle_threshold}, track_interval={self.track_interval}")
