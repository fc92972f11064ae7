"""Batched cluster tracker: host side of csrc/track.cu (C ABI: pcs_trk_* in include/pcseq_b200.h).

The reference tracks one (component key, anchor frame) pair at a time (ClusterTracking.forward,
pcdet/models/registration/preprocessors/cluster_tracking.py:853-884 -> track_frame :430-787).  The pairs are
independent, so `TrackBatch` lays ALL of them out as "instances" of one batch and advances them together: tracking
step t moves every anchor a to frame a - t (t <= 8) or a + (t - 8).  The set-up below (torch, a handful of host
syncs per sequence) splits the sequence by frame once, numbers the non-empty components of the anchor frames, builds
the static per-level voxel clouds / cell grids of every frame (sample_frame of the TARGET frame does not depend on
the anchor), and `run()` enqueues the 16 steps without any host synchronisation.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream, next_pow2
from .utils import EasyDict

STATIONARY_DIAMETER = 12.5  # cluster_tracking.py:861 / filter_components default
REL = 17                    # relative frame window [anchor - 8, anchor + 8]
ANCHOR_REL = 8
ICP_RINGS = int(__import__('os').environ.get('PCS_ICP_RINGS', '1'))  # 1: cell = radius, 3x3x3 search (measured
# faster than 2: cells of half the radius searched over 5x5x5 -- the shell lookups of unmatched queries dominate)


def _mk_struct(name, fields):
    return type(name, (ctypes.Structure,), {"_fields_": fields})


_P, _I, _D = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double

SamplerStruct = _mk_struct("pcs_trk_sampler_t", [
    ("pts", _P), ("group", _P), ("skey", _P), ("bits", _P), ("act", _P),
    ("n", _I), ("n_groups", _I), ("n_keys", _I), ("ns_only", _I),
    ("size", _D * 3), ("sb", _P), ("table", _P), ("H", _I),
    ("vsum", _P), ("vbits", _P), ("vk", _P), ("vres", _P), ("pnext", _P), ("vlist", _P), ("ctr", _P),
    ("kcount", _P), ("koff", _P), ("kcur", _P), ("vdeg", _P),
    ("out_pts", _P), ("out_key", _P), ("out_group", _P)])

IcpStruct = _mk_struct("pcs_trk_icp_t", [
    ("J", _I), ("G", _I), ("act", _P), ("ref_group", _P), ("ref_group_all", _P), ("skipmask", _P), ("ref_off", _P),
    ("g_inst", _P),
    ("ref_table", _P), ("ref_H", _I), ("ref_pts", _P),
    ("mov_table", _P), ("mov_H", _I), ("mov_sorted", _P), ("mov_sidx", _P), ("mov_cells", _P), ("mov_ctr", _P),
    ("mv", _P), ("mv_gid", _P), ("mv_inst", _P), ("n_mv", _P), ("mv_cap", _I), ("vdeg", _P),
    ("lo", _D * 3), ("cs", _D), ("rings", _I), ("radius", _D), ("df", _I), ("angle_reg", _D), ("max_iter", _I),
    ("stopping_delta", _D), ("want_l1", _I), ("want_ratio", _I),
    ("nn_fwd", _P), ("nn_bwd", _P), ("boff", _P), ("mvbeg", _P), ("mvend", _P),
    ("sec_fwd", _P), ("sec_bwd", _P), ("disp", _P), ("dmax", _P),
    ("mom", _P), ("Ti", _P), ("T", _P), ("mu", _P), ("l1_sum", _P), ("l1_n", _P),
    ("phase", _P), ("cd", _P), ("iters", _P), ("itcnt", _P), ("last", _P), ("loss", _P), ("match_cnt", _P),
    ("l1_err", _P), ("ratio", _P), ("prof", _P)])

CtxStruct = _mk_struct("pcs_trk_ctx_t", [
    ("J", _I), ("G", _I), ("M", _I), ("F", _I),
    ("inst_anchor", _P), ("inst_key", _P), ("inst_C", _P), ("inst_fmin", _P), ("inst_fmax", _P),
    ("inst_has_valid", _P), ("inst_goff", _P),
    ("seq_sorted", _P), ("frame_off", _P),
    ("mp", _P), ("mp0", _P), ("m_last", _P), ("m_gid", _P), ("m_inst", _P),
    ("g_inst", _P), ("g_deg", _P), ("g_diam", _P), ("g_valid", _P),
    ("g_stopped", _P), ("g_moving", _P), ("g_final", _P), ("g_minf", _P), ("g_maxf", _P),
    ("transforms", _P), ("velos", _P), ("velos_b", _P), ("centers", _P), ("diffs", _P),
    ("cv_pre", _P), ("g_delta", _P), ("adam_m", _P), ("adam_v", _P),
    ("csum", _P), ("vsum", _P), ("l1_err", _P), ("ratio", _P), ("T", _P), ("vdeg", _P),
    ("cur_act", _P), ("cur_nxt", _P), ("cur_rel", _P), ("cur_haslv", _P), ("cur_grp", _P), ("cur_grp_all", _P),
    ("n_keys", _I), ("anyns", _P), ("sb", _P),
    ("reg_error_coeff", _D), ("angle_threshold", _D), ("min_move_frame", _I),
    ("radius", _D * 8), ("voxel_size", _D * 24), ("lo", _D * 3), ("nn_radius", _D),
    ("eg_table", _P), ("eg_H", _I), ("eg_sorted", _P), ("eg_sidx", _P), ("eg_cells", _P), ("eg_ctr", _P),
    ("eoff", _P), ("exoff", _P), ("ex", _P)])


def _fill(struct, tensors, **kw):
    """Set struct fields from tensors (device pointers) / python scalars; `tensors` keeps the buffers alive."""
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            tensors[len(tensors)] = v
            setattr(struct, k, v.data_ptr())
        elif v is None:
            setattr(struct, k, None)
        elif isinstance(v, (list, tuple)):
            arr = getattr(struct, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(struct, k, v)


def _rows(src, idx):
    """src[idx] for point rows: torch's indexing kernel handles one 16-byte row per thread BLOCK (4 ms for 8 M rows);
    pcs_gather_rows moves them at memory speed."""
    from .ops import gather_rows
    return gather_rows(src.contiguous(), idx)


def _seg_minmax(values, ids, n_groups, empty_min, empty_max):
    """(min, max) int64[n_groups] of integer `values` per group (warp-aggregated kernel: ids are mostly sorted; the
    torch scatter_reduce serialises on the few hot bins).  Exact for |values| < 2^24; empty groups get the defaults."""
    from .ops import group_minmax
    if values.numel() == 0 or int(n_groups) < 1:
        return (torch.full((int(n_groups),), empty_min, dtype=torch.int64, device=values.device),
                torch.full((int(n_groups),), empty_max, dtype=torch.int64, device=values.device))
    mn, mx = group_minmax(values.float().contiguous(), ids, n_groups)
    cnt = torch.bincount(ids, minlength=n_groups)
    has = cnt > 0
    return (torch.where(has, mn.round().long(), torch.full_like(cnt, empty_min)),
            torch.where(has, mx.round().long(), torch.full_like(cnt, empty_max)))


def _z(n, dtype, dev):
    return torch.zeros(n, dtype=dtype, device=dev)


def _e(n, dtype, dev):
    return torch.empty(n, dtype=dtype, device=dev)


class VoxelSampler:
    """Scratch of pcs_trk_sample for up to `cap` points."""

    def __init__(self, cap, n_keys_max, n_groups_max, dev):
        self.cap, self.dev = int(cap), dev
        self.H = next_pow2(max(2 * self.cap, 1024))
        H = self.H
        self.t = dict(
            table=_e((H, 4), torch.int32, dev), vsum=_e((H, 3), torch.float64, dev), vbits=_e((H, 3), torch.int32, dev),
            vk=_e((H, 2), torch.int32, dev), vres=_e((H, 2), torch.int32, dev),
            pnext=_e(self.cap, torch.int32, dev), vlist=_e(self.cap, torch.int32, dev), ctr=_z(4, torch.int32, dev),
            kcount=_z(n_keys_max + 1, torch.int32, dev), koff=_z(n_keys_max + 1, torch.int32, dev),
            kcur=_z(n_keys_max + 1, torch.int32, dev))
        self.keep = {}
        s = SamplerStruct()
        _fill(s, self.keep, H=H, **self.t)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pcs_trk_sampler_init(_stream(), ctypes.byref(s)), "pcs_trk_sampler_init")

    def struct(self, pts, group, skey, bits, act, n, n_groups, n_keys, ns_only, size, sb, vdeg, out_pts, out_key,
               out_group):
        assert n <= self.cap and n_keys + 1 <= self.t["kcount"].shape[0]
        s = SamplerStruct()
        _fill(s, self.keep, H=self.H, **self.t)
        _fill(s, self.keep, pts=pts, group=group, skey=skey, bits=bits, act=act, n=int(n), n_groups=int(n_groups),
              n_keys=int(n_keys), ns_only=int(ns_only), size=list(size), sb=sb, vdeg=vdeg, out_pts=out_pts,
              out_key=out_key, out_group=out_group)
        return s


class TrackBatch:
    """All (component key, anchor frame) tracking instances of one sequence.

    fxyz f32[N,4] (frame, x, y, z) after ground removal, frame int[N], comps: list of int64[N] component ids (one
    per component key, ids dense and unique per frame as ClusterProposal produces them), cfg: the ClusterTracking
    model_cfg (REGISTRATION / NN_GRAPH / TRACKING_PARAMS / ANGLE_REGULARIZER).
    """

    def __init__(self, fxyz, frame, comps, cfg, num_frames=None, anchors=None):
        L = _lib.lib()
        dev = fxyz.device
        self.dev = dev
        self.keep = {}
        reg = cfg["REGISTRATION"]
        self.radius = [float(r) for r in reg["GRAPH"]["RADIUS"]]
        self.voxel_size = [[float(x) for x in v] for v in reg["VOXEL_SIZE"]]
        self.stopping_delta = [float(x) for x in reg["STOPPING_DELTA"]]
        self.n_levels = len(self.radius)
        if int(reg["GRAPH"].get("MAX_NUM_NEIGHBORS", 1)) != 1:
            raise _lib.PcsError("REGISTRATION.GRAPH.MAX_NUM_NEIGHBORS must be 1 on this path (K=1 ICP search)")
        params = cfg.get("TRACKING_PARAMS", {})
        self.interval = int(params.get("TRACK_INTERVAL", 10))
        if not (1 <= self.interval <= ANCHOR_REL):
            raise _lib.PcsError(f"TRACK_INTERVAL must be in [1, {ANCHOR_REL}] on this path (got {self.interval})")
        self.min_move = int(params.get("MIN_MOVE_FRAME", 6))
        self.nn_radius = float(cfg["NN_GRAPH"]["RADIUS"])
        self.angle_reg = float(cfg["ANGLE_REGULARIZER"])
        nK = len(comps)
        assert 1 <= nK <= 3, "up to three component keys (one stationary flag bit each)"
        self.nK = nK

        fxyz = fxyz.float().contiguous()
        self.fxyz = fxyz
        frame = frame.reshape(-1).long()
        N = fxyz.shape[0]
        F = int(num_frames) if num_frames is not None else int(frame.max().item()) + 1
        self.N, self.F = N, F
        with torch.cuda.device(dev):
            # ---- frame split (one stable sort instead of one boolean mask per anchor / target frame) ----------
            order = torch.argsort(frame, stable=True)
            fcnt = torch.bincount(frame, minlength=F)
            frame_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
            frame_off[1:] = fcnt.cumsum(0)
            off_h = frame_off.tolist()  # host sync: sizes of everything below
            self.order, self.frame_off_h = order, off_h
            seq_sorted = _rows(fxyz, order)
            frame_sorted = frame[order].int().contiguous()
            pos_in_sorted = torch.empty(N, dtype=torch.int64, device=dev)
            pos_in_sorted[order] = torch.arange(N, device=dev)
            if anchors is None:
                anchors = list(range(0, F, self.interval))
            anchors = [a for a in anchors if off_h[a + 1] > off_h[a]]
            self.anchors = anchors
            A = len(anchors)
            if A == 0:
                raise _lib.PcsError("no anchor frame holds points")
            J = nK * A
            self.J, self.A = J, A
            fmin_seq = min(f for f in range(F) if off_h[f + 1] > off_h[f])
            fmax_seq = max(f for f in range(F) if off_h[f + 1] > off_h[f])
            self.fmin_seq, self.fmax_seq = fmin_seq, fmax_seq
            anchor_index = torch.full((F,), -1, dtype=torch.int64, device=dev)
            anchor_index[torch.tensor(anchors, device=dev)] = torch.arange(A, device=dev)
            arow_mask = anchor_index[frame] >= 0
            arows = arow_mask.nonzero().reshape(-1)  # anchor-frame points (original rows, ascending)
            xyz = fxyz[:, 1:]

            g_inst, g_deg, g_diam, g_valid, g_center, g_local, inst_C, inst_goff = [], [], [], [], [], [], [], [0]
            m_rows, m_gid, m_stat = [], [], []
            statbits = torch.zeros(N, dtype=torch.uint8, device=dev)
            goff = 0
            for ki, c in enumerate(comps):
                c = c.reshape(-1).long()
                Ck = int(c.max().item()) + 1
                deg = torch.bincount(c, minlength=Ck)
                csum = torch.zeros(Ck, 3, dtype=torch.float64, device=dev).index_add_(0, c, xyz.double())
                center = (csum / deg.clamp(min=1)[:, None].double()).float()
                dist = (xyz - center[c]).norm(p=2, dim=-1)
                diam = torch.zeros(Ck, device=dev).scatter_reduce_(0, c, dist, "amax", include_self=True) * 2
                cframe = torch.zeros(Ck, dtype=torch.int64, device=dev).scatter_(0, c, frame)
                valid = (deg > 0) & (diam < STATIONARY_DIAMETER)
                statbits |= ((diam > STATIONARY_DIAMETER)[c].to(torch.uint8) << ki)
                if Ck < (1 << 24):
                    fminc, fmaxc = _seg_minmax(c, frame, F, Ck, -1)
                else:
                    fminc = torch.full((F,), Ck, dtype=torch.int64, device=dev).scatter_reduce_(0, frame, c, "amin")
                    fmaxc = torch.full((F,), -1, dtype=torch.int64, device=dev).scatter_reduce_(0, frame, c, "amax")
                # non-empty components of the anchor frames, grouped by anchor (instance), ascending id inside
                sel = ((deg > 0) & (anchor_index[cframe] >= 0)).nonzero().reshape(-1)
                ai = anchor_index[cframe[sel]]
                perm = torch.argsort(ai * Ck + sel)
                sel, ai = sel[perm], ai[perm]
                gid_of = torch.full((Ck,), -1, dtype=torch.int64, device=dev)
                gid_of[sel] = torch.arange(sel.shape[0], device=dev) + goff
                per_inst = torch.bincount(ai, minlength=A)
                g_inst.append(ai + ki * A)
                g_deg.append(deg[sel])
                g_diam.append(diam[sel])
                g_valid.append(valid[sel])
                g_center.append(center[sel])
                g_local.append(sel - fminc[cframe[sel]])
                at = torch.tensor(anchors, device=dev)
                inst_C.append(fmaxc[at] - fminc[at] + 1)
                inst_goff.append(per_inst)
                # moving points of this key: the anchor rows, grouped by component
                gp = gid_of[c[arows]]
                o2 = torch.argsort(gp, stable=True)
                m_rows.append(arows[o2])
                m_gid.append(gp[o2])
                m_stat.append(((statbits[arows[o2]] >> ki) & 1))
                goff += int(sel.shape[0])
            G = goff
            self.G = G
            g_inst = torch.cat(g_inst).int().contiguous()
            self.g_inst = g_inst
            self.g_deg = torch.cat(g_deg).int().contiguous()
            self.g_diam = torch.cat(g_diam).float().contiguous()
            self.g_valid = torch.cat(g_valid).to(torch.uint8).contiguous()
            self.g_local = torch.cat(g_local)
            g_center = torch.cat(g_center).float()
            self.inst_C = torch.cat(inst_C).int().contiguous()
            goffs = torch.zeros(J + 1, dtype=torch.int64, device=dev)
            goffs[1:] = torch.cat(inst_goff[1:]).cumsum(0)
            self.inst_goff = goffs.int().contiguous()
            m_rows = torch.cat(m_rows)
            self.m_rows = m_rows
            self.m_gid = torch.cat(m_gid).int().contiguous()
            self.m_inst = g_inst[self.m_gid.long()].contiguous()
            m_stat = torch.cat(m_stat).to(torch.uint8).contiguous()
            M = int(m_rows.shape[0])
            self.M = M
            mp0 = _rows(fxyz, m_rows)
            self.m_frow = (pos_in_sorted[m_rows] - frame_off[frame[m_rows]]).int().contiguous()
            inst_anchor = torch.tensor(anchors * nK, dtype=torch.int32, device=dev)
            inst_key = torch.arange(nK, device=dev).repeat_interleave(A).int()
            inst_fmin = (inst_anchor - self.interval).clamp(min=fmin_seq).int()
            inst_fmax = (inst_anchor + self.interval).clamp(max=fmax_seq).int()
            has_valid = torch.zeros(J, dtype=torch.int32, device=dev)
            has_valid.index_put_((g_inst.long(),), self.g_valid.int(), accumulate=True)
            has_valid = (has_valid > 0).int()
            self.inst_anchor_h = anchors * nK
            self.inst_C_h = self.inst_C.tolist()
            self.inst_goff_h = self.inst_goff.tolist()
            self.inst_key_h = [k for k in range(nK) for _ in range(A)]

            # ---- sequence-global grid origin (cells never go negative; 32 m of slack for the moving side) ---------
            lo = (xyz.min(0)[0] - 32.0).tolist()
            self.lo = lo

            # ---- sampler scratch shared by the set-up (whole sequence) and the steps (moving points) ----------------
            self.sampler = VoxelSampler(max(N, M), max(G, F), max(J, F), dev)
            s = _stream()
            sb_f = _e((F, 6), torch.int32, dev)
            _lib.check(L.pcs_trk_bounds_reset(s, _ptr(sb_f), F), "pcs_trk_bounds_reset")
            # ---- static per-level voxel clouds + cell grids of every frame (the ICP targets) ---------------------
            self.levels = []
            out_pts = _e((N, 4), torch.float32, dev)
            out_key = _e(N, torch.int32, dev)
            out_group = _e(N, torch.int32, dev)
            for lv in range(self.n_levels):
                _lib.check(L.pcs_trk_group_bounds(s, _ptr(seq_sorted), _ptr(frame_sorted), N, _ptr(sb_f)),
                           "pcs_trk_group_bounds")
                st = self.sampler.struct(seq_sorted, frame_sorted, None, statbits[order].contiguous(), None, N, F, F, 0,
                                         self.voxel_size[lv], sb_f, None, out_pts, out_key, out_group)
                _lib.check(L.pcs_trk_sample(s, ctypes.byref(st)), "pcs_trk_sample (static)")
                V = int(self.sampler.t["ctr"][1].item())
                cs = self.radius[lv] * 1.001 / ICP_RINGS
                lo_c = (ctypes.c_double * 3)(*lo)
                # One grid per level holds, for every component key k, the NON-stationary voxels of every frame
                # (group = k * F + frame: the ICP never has to step over the stationary bulk -- buildings, walls --
                # of a cell) and, as pseudo-key nK, all voxels (group = nK * F + frame: the matched-fraction search).
                vp, vf = out_pts[:V], out_group[:V].long()
                bits = vp[:, 0].contiguous().view(torch.int32)
                sel_pts, sel_grp = [], []
                for ki in range(nK):
                    ns = (((bits >> ki) & 1) == 0).nonzero().reshape(-1)
                    sel_pts.append(_rows(vp, ns))
                    sel_grp.append(vf[ns] + ki * F)
                sel_pts.append(vp)
                sel_grp.append(vf + nK * F)
                cat_pts = torch.cat(sel_pts).contiguous()
                cat_grp = torch.cat(sel_grp)
                Vc = int(cat_pts.shape[0])
                keys = _e(Vc, torch.int64, dev)
                _lib.check(L.pcs_trk_cell_keys(s, _ptr(cat_pts), _ptr(cat_grp.int().contiguous()), Vc, lo_c, cs,
                                               _ptr(keys)), "pcs_trk_cell_keys")
                # canonical row order (cell key, then x, y, z): equal-distance ties of the searches are resolved by row,
                # and the sampler emits voxels in a run-dependent order
                bits = cat_pts[:, 1:].contiguous().view(torch.int32).long() & 0xffffffff
                pos_key = ((bits[:, 0] << 32) | bits[:, 1]) ^ (bits[:, 2] << 13)  # any deterministic function of xyz
                perm = torch.sort(pos_key)[1]
                ks, o2 = torch.sort(keys[perm], stable=True)
                perm = perm[o2]
                rv = _rows(cat_pts, perm)
                rv_grp = cat_grp[perm]
                uk, cnt = torch.unique_consecutive(ks, return_counts=True)
                starts = (cnt.cumsum(0) - cnt).int().contiguous()
                ncell = int(uk.shape[0])
                Hc = next_pow2(max(2 * ncell, 1024))
                table = _e((Hc, 4), torch.int32, dev)
                err = _z(1, torch.int32, dev)
                _lib.check(L.pcs_trk_grid_fill(s, _ptr(table), Hc, _ptr(uk), _ptr(starts), _ptr(cnt.int().contiguous()),
                                               ncell, _ptr(err)), "pcs_trk_grid_fill")
                n_grp = (nK + 1) * F
                rv_off = torch.zeros(n_grp + 1, dtype=torch.int64, device=dev)
                rv_off[1:] = torch.bincount(rv_grp, minlength=n_grp).cumsum(0)
                rv_off_h = rv_off.tolist()
                max_cnt = max(rv_off_h[g + 1] - rv_off_h[g] for g in range(nK * F))
                self.levels.append(dict(rv=rv, rv_off=rv_off.int().contiguous(), table=table, H=Hc, cs=cs, V=V,
                                        bwd_cap=max(J * max_cnt, 1), err=err))

            # ---- tracker state --------------------------------------------------------------------------------
            t = {}
            t["seq_sorted"], t["frame_off"] = seq_sorted, frame_off.int().contiguous()
            t["inst_anchor"], t["inst_key"], t["inst_C"] = inst_anchor, inst_key, self.inst_C
            t["inst_fmin"], t["inst_fmax"], t["inst_has_valid"], t["inst_goff"] = inst_fmin, inst_fmax, has_valid, self.inst_goff
            t["mp"], t["mp0"], t["m_last"] = mp0.clone(), mp0, mp0.clone()
            t["m_gid"], t["m_inst"] = self.m_gid, self.m_inst
            t["g_inst"], t["g_deg"], t["g_diam"], t["g_valid"] = g_inst, self.g_deg, self.g_diam, self.g_valid
            t["g_stopped"], t["g_moving"], t["g_final"] = _z(G, torch.uint8, dev), _z(G, torch.uint8, dev), _z(G, torch.uint8, dev)
            ga = inst_anchor[g_inst.long()].contiguous()
            t["g_minf"], t["g_maxf"] = ga.clone(), ga.clone()
            tr = _z((G, REL, 12), torch.float64, dev)
            tr[:, :, 0] = tr[:, :, 4] = tr[:, :, 8] = 1.0
            t["transforms"] = tr
            t["velos"], t["velos_b"] = _z((G, REL, 3), torch.float32, dev), _z((G, REL, 3), torch.float32, dev)
            cen = _z((G, REL, 3), torch.float32, dev)
            cen[:, ANCHOR_REL] = g_center
            t["centers"], t["diffs"] = cen, _z((G, REL, 3), torch.float32, dev)
            t["cv_pre"], t["g_delta"] = _z((G, 3), torch.float32, dev), _z((G, 3), torch.float32, dev)
            t["adam_m"], t["adam_v"] = _e((G, 16), torch.float32, dev), _e((G, 16), torch.float32, dev)
            t["csum"], t["vsum"] = _z((G, 3), torch.float64, dev), _z((G, 3), torch.float64, dev)
            t["l1_err"], t["ratio"] = _z(G, torch.float64, dev), _z(G, torch.float32, dev)
            t["T"], t["vdeg"] = _z((G, 12), torch.float64, dev), _z(G, torch.int32, dev)
            for k in ("cur_act", "cur_nxt", "cur_rel", "cur_haslv", "cur_grp", "cur_grp_all"):
                t[k] = _z(J, torch.int32, dev)
            t["anyns"] = _z((REL + 1, J), torch.int32, dev)
            t["sb"] = _e((J, 6), torch.int32, dev)
            _lib.check(L.pcs_trk_bounds_reset(s, _ptr(t["sb"]), J), "pcs_trk_bounds_reset")
            He = next_pow2(max(2 * M, 1024))
            t["eg_table"], t["eg_sorted"] = _e((He, 4), torch.int32, dev), _e((M, 4), torch.float32, dev)
            t["eg_sidx"], t["eg_cells"], t["eg_ctr"] = _e(M, torch.int32, dev), _e(M, torch.int32, dev), _z(4, torch.int32, dev)
            _lib.check(L.pcs_trk_table_clear(s, _ptr(t["eg_table"]), He, _ptr(t["eg_ctr"])), "pcs_trk_table_clear")
            t["eoff"] = _z(J + 1, torch.int32, dev)
            # extraction table: slot 0 = the anchor frame, slot t = the target frame of step t
            sizes = np.zeros((J, REL), dtype=np.int64)
            self.slot_frame = np.full((J, REL), -1, dtype=np.int64)
            for j in range(J):
                a = self.inst_anchor_h[j]
                lo_f, hi_f = max(fmin_seq, a - self.interval), min(fmax_seq, a + self.interval)
                self.slot_frame[j, 0] = a
                for tt in range(1, REL):
                    d, sdist = (-1, tt) if tt <= 8 else (1, tt - 8)
                    f = a + d * sdist
                    if sdist <= self.interval and lo_f <= f <= hi_f:
                        self.slot_frame[j, tt] = f
                for tt in range(REL):
                    f = self.slot_frame[j, tt]
                    if f >= 0:
                        sizes[j, tt] = off_h[f + 1] - off_h[f]
            exoff = np.zeros(J * REL + 1, dtype=np.int64)
            exoff[1:] = np.cumsum(sizes.reshape(-1))
            self.exoff_h = exoff
            t["exoff"] = torch.from_numpy(exoff).to(dev)
            t["ex"] = torch.full((max(int(exoff[-1]), 1),), -1, dtype=torch.int32, device=dev)
            self.t = t

            ctx = CtxStruct()
            _fill(ctx, self.keep, J=J, G=G, M=M, F=F, n_keys=nK, reg_error_coeff=float(params.get("REGISTRATION_ERROR_COEFFICIENT", 0.13)),
                  angle_threshold=float(params.get("ANGLE_THRESHOLD", 45)), min_move_frame=self.min_move,
                  nn_radius=self.nn_radius, eg_H=He, lo=lo, **t)
            for lv in range(self.n_levels):
                ctx.radius[lv] = self.radius[lv]
                for q in range(3):
                    ctx.voxel_size[lv * 3 + q] = self.voxel_size[lv][q]
            self.ctx = ctx

            # ---- sampler / ICP descriptors of the steps ---------------------------------------------------------
            self.mv = _e((M, 4), torch.float32, dev)
            self.mv_gid, self.mv_inst = _e(M, torch.int32, dev), _e(M, torch.int32, dev)
            self.samp = self.sampler.struct(t["mp"], self.m_inst, self.m_gid, m_stat, t["cur_act"], M, J, G, 1,
                                            self.voxel_size[0], t["sb"], t["vdeg"], self.mv, self.mv_gid, self.mv_inst)
            self.m_stat = m_stat
            Hm = next_pow2(max(2 * M, 1024))
            sc = dict(mov_table=_e((Hm, 4), torch.int32, dev), mov_sorted=_e((M, 4), torch.float32, dev),
                      mov_sidx=_e(M, torch.int32, dev), mov_cells=_e(M, torch.int32, dev), mov_ctr=_z(4, torch.int32, dev),
                      nn_fwd=_e(M, torch.int32, dev), boff=_z(J + 1, torch.int32, dev),
                      mvbeg=_z(J, torch.int32, dev), mvend=_z(J, torch.int32, dev),
                      sec_fwd=_z(M, torch.float32, dev), disp=_z(M, torch.float32, dev), dmax=_z(J, torch.int32, dev),
                      mom=_z((G, 17), torch.float64, dev), Ti=_z((G, 12), torch.float64, dev), mu=_z((G, 6), torch.float64, dev),
                      l1_sum=_z((G, 2), torch.float64, dev), l1_n=_z(G, torch.float64, dev),
                      phase=_z(J, torch.int32, dev), cd=_z(J, torch.int32, dev), iters=_z(J, torch.int32, dev),
                      itcnt=_z((80 + 2) * 3, torch.int32, dev), last=_z(J, torch.float64, dev),
                      loss=_z(J, torch.float64, dev), match_cnt=_z(G, torch.int32, dev))
            _lib.check(L.pcs_trk_table_clear(s, _ptr(sc["mov_table"]), Hm, _ptr(sc["mov_ctr"])), "pcs_trk_table_clear")
            self.sc = sc
            nn_bwd = _e(max(lv["bwd_cap"] for lv in self.levels), torch.int32, dev)
            sec_bwd = _z(max(lv["bwd_cap"] for lv in self.levels), torch.float32, dev)
            if __import__('os').environ.get('PCS_ICP_NOCACHE'):
                sc['sec_fwd'] = None
            self.skipmask = torch.zeros(J, dtype=torch.int32, device=dev)  # the grids hold non-stationary voxels only
            self.prof = [_z(256, torch.int64, dev) for _ in range(self.n_levels)]
            icp_arr = (IcpStruct * self.n_levels)()
            for lv in range(self.n_levels):
                d = self.levels[lv]
                _fill(icp_arr[lv], self.keep, J=J, G=G, act=t["cur_act"], ref_group=t["cur_grp"],
                      ref_group_all=t["cur_grp_all"], skipmask=self.skipmask,
                      ref_off=d["rv_off"], g_inst=g_inst, ref_table=d["table"], ref_H=d["H"], ref_pts=d["rv"],
                      mov_H=Hm, mv=self.mv, mv_gid=self.mv_gid, mv_inst=self.mv_inst, mv_cap=M,
                      n_mv=self.sampler.t["ctr"][1:], vdeg=t["vdeg"], lo=lo, cs=d["cs"], rings=ICP_RINGS,
                      radius=self.radius[lv], df=0,
                      angle_reg=self.angle_reg, max_iter=80, stopping_delta=self.stopping_delta[lv], want_l1=0,
                      want_ratio=0, nn_bwd=nn_bwd, sec_bwd=sec_bwd, T=t["T"], l1_err=t["l1_err"], ratio=t["ratio"], prof=self.prof[lv], **sc)
            self.icp_arr = icp_arr

    # ----------------------------------------------------------------------------------------------------------
    def steps(self):
        """Global step numbers in execution order (t <= 8 backwards, t > 8 forwards)."""
        return [t for t in range(1, 17) if (t if t <= 8 else t - 8) <= self.interval]

    def step(self, t):
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().pcs_trk_step(_stream(), ctypes.byref(self.ctx), ctypes.byref(self.samp), self.icp_arr,
                                               self.n_levels, int(t)), "pcs_trk_step")

    def run(self):
        """All tracking steps + the final component filter; no host synchronisation."""
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            s = _stream()
            if self.interval == ANCHOR_REL:
                _lib.check(L.pcs_trk_run(s, ctypes.byref(self.ctx), ctypes.byref(self.samp), self.icp_arr, self.n_levels,
                                         _ptr(self.m_frow)), "pcs_trk_run")
            else:
                for t in self.steps():
                    self.step(t)
                _lib.check(L.pcs_trk_finish(s, ctypes.byref(self.ctx), _ptr(self.m_frow)), "pcs_trk_finish")
        return self

    def check(self):
        """Synchronising check of the device-side error flags."""
        bad = [int(self.sampler.t["ctr"][2].item()), int(self.sc["mov_ctr"][2].item()), int(self.t["eg_ctr"][2].item())]
        if any(bad):
            raise _lib.PcsError(f"tracker: device-side error flags (sampler, moving grid, extraction grid) = {bad}")

    # ----------------------------------------------------------------------------------------------------------
    def flat_results(self):
        """All kept (instance, slot, frame row) entries of the extraction table, concatenated in the reference's
        order (per instance: anchor frame, then the target frames in tracking order, ascending row inside a frame).
        One host sync (sizes).  Returns an EasyDict of flat tensors plus `slot_bounds` (host list, J * 17 + 1)."""
        if getattr(self, "_flat", None) is not None:
            return self._flat
        dev, t = self.dev, self.t
        J = self.J
        ex = t["ex"]
        gfin = t["g_final"].bool()
        keep = ex >= 0
        keep &= gfin[ex.clamp(min=0).long()]
        idx = keep.nonzero().reshape(-1)
        exoff = t["exoff"]
        slot = torch.searchsorted(exoff, idx, right=True) - 1  # (j * 17 + t)
        i_in = idx - exoff[slot]
        slot_frame = torch.from_numpy(self.slot_frame.reshape(-1)).to(dev)
        f_of = slot_frame[slot]
        frame_off = t["frame_off"].long()
        rows = self.order[frame_off[f_of] + i_in]
        gid = ex[idx].long()
        slot_bounds = torch.searchsorted(idx, exoff).tolist()  # host sync
        self._flat = EasyDict(dict(gid=gid, rows=rows, frame_rows=i_in, frame=f_of,
                                   inst=torch.div(slot, REL, rounding_mode="floor"),
                                   component=self.g_local[gid], moving=t["g_moving"].bool()[gid],
                                   fxyz=_rows(self.fxyz, rows)))
        self._flat["slot_bounds"] = slot_bounds
        if self.g_local.numel() and int(self.g_local.max()) < (1 << 24):
            cmax = _seg_minmax(self._flat.component, self._flat.inst, J, 0, -1)[1]
        else:
            cmax = torch.full((J,), -1, dtype=torch.long, device=dev).scatter_reduce_(0, self._flat.inst, self._flat.component, "amax")
        self.flat_cmax = cmax.tolist()
        return self._flat

    def results(self, seg_label=None):
        """Per-instance `extracted` dicts in the reference's layout (cluster_tracking.py:727-752).
        Returns {(key index, anchor frame): (instance, EasyDict)}."""
        fl = self.flat_results()
        sb = fl["slot_bounds"]
        seg = seg_label[fl.rows] if seg_label is not None else None
        out = {}
        for j in range(self.J):
            a, ki = self.inst_anchor_h[j], self.inst_key_h[j]
            b0, b1 = sb[j * REL], sb[(j + 1) * REL]
            e = EasyDict(dict(fxyz=fl.fxyz[b0:b1], component=fl.component[b0:b1], frame_indices=fl.frame_rows[b0:b1],
                              original_indices=fl.rows[b0:b1], moving=fl.moving[b0:b1],
                              valid_comp_mask=fl.moving[b0:b1],
                              gt_box_label=torch.zeros_like(fl.component[b0:b1])))
            if seg is not None:
                e["segmentation_label"] = seg[b0:b1]
            out[(ki, a)] = (j, e)
        return out

    def transforms(self, j):
        """transforms f64[C, F_j, 4, 4] of instance j in the reference layout (identity for empty components)."""
        dev = self.dev
        a = self.inst_anchor_h[j]
        C = self.inst_C_h[j]
        g0, g1 = self.inst_goff_h[j], self.inst_goff_h[j + 1]
        fmin, fmax = max(self.fmin_seq, a - self.interval), min(self.fmax_seq, a + self.interval)
        k0, k1 = fmin - (a - ANCHOR_REL), fmax - (a - ANCHOR_REL)
        nf = k1 - k0 + 1
        T = torch.zeros(C, nf, 4, 4, dtype=torch.float64, device=dev)
        T[:, :, 0, 0] = T[:, :, 1, 1] = T[:, :, 2, 2] = T[:, :, 3, 3] = 1.0
        tr = self.t["transforms"][g0:g1, k0:k1 + 1]
        loc = self.g_local[g0:g1]
        T[loc, :, :3, :3] = tr[:, :, :9].reshape(-1, nf, 3, 3)
        T[loc, :, :3, 3] = tr[:, :, 9:]
        return T


def register_pair(mov_fxyz, mov_comp, mov_stationary, ref_fxyz, ref_stationary, num_components, radius, frame_offset,
                  angle_regularizer=10.0, max_iter=20, stopping_delta=5e-2):
    """register_to_next_frame (registration_utils.py:83-206) for ONE (moving, target) pair on the batched ICP kernel
    (a batch of one instance).  Same contract as ops.register_icp."""
    L = _lib.lib()
    mov_fxyz = mov_fxyz.float().contiguous()
    ref_fxyz = ref_fxyz.float().contiguous()
    dev = mov_fxyz.device
    if not mov_fxyz.is_cuda:
        raise _lib.PcsError("register_pair: CUDA tensors required (no CPU path exists)")
    C = int(num_components)
    df = int(frame_offset)
    mov_comp = mov_comp.long().reshape(-1)
    ns_m = ~mov_stationary.bool().reshape(-1)
    ns_r = ~ref_stationary.bool().reshape(-1)
    keep = {}
    T = torch.zeros(C, 12, dtype=torch.float64, device=dev)
    T[:, 0] = T[:, 4] = T[:, 8] = 1.0
    l1 = _z(C, torch.float64, dev)
    ratio = _z(C, torch.float32, dev)
    iters = _z(1, torch.int32, dev)
    moved = mov_fxyz.clone()
    nm = int(ns_m.sum().item())
    if nm > 0 and ref_fxyz.shape[0] > 0:
        with torch.cuda.device(dev):
            s = _stream()
            # moving voxels grouped by component
            idx_m = ns_m.nonzero().reshape(-1)
            o = torch.argsort(mov_comp[idx_m], stable=True)
            idx_m = idx_m[o]
            mv = mov_fxyz[idx_m].contiguous()
            mv_gid = mov_comp[idx_m].int().contiguous()
            mv_inst = _z(nm, torch.int32, dev)
            n_mv = torch.tensor([nm], dtype=torch.int32, device=dev)
            vdeg = torch.bincount(mov_comp, minlength=C).int().contiguous()
            # reference grid: group 0 = non-stationary voxels (ICP targets), group 1 = all voxels (matched fraction)
            both = torch.cat([ref_fxyz[ns_r], ref_fxyz])
            lo = (torch.cat([ref_fxyz[:, 1:], mov_fxyz[:, 1:]]).min(0)[0] - 32.0).tolist()
            grp = torch.cat([_z(int(ns_r.sum().item()), torch.int32, dev),
                             torch.ones(ref_fxyz.shape[0], dtype=torch.int32, device=dev)])
            cs = float(radius) * 1.001
            lo_c = (ctypes.c_double * 3)(*lo)
            n_ref = int(both.shape[0])
            keys = _e(n_ref, torch.int64, dev)
            _lib.check(L.pcs_trk_cell_keys(s, _ptr(both.contiguous()), _ptr(grp), n_ref, lo_c, cs, _ptr(keys)),
                       "pcs_trk_cell_keys")
            ks, perm = torch.sort(keys)
            rv = both[perm].contiguous()
            rv[:, 0] = 0.0  # payload bits: nothing to skip
            uk, cnt = torch.unique_consecutive(ks, return_counts=True)
            starts = (cnt.cumsum(0) - cnt).int().contiguous()
            Hc = next_pow2(max(2 * int(uk.shape[0]), 1024))
            table = _e((Hc, 4), torch.int32, dev)
            err = _z(1, torch.int32, dev)
            _lib.check(L.pcs_trk_grid_fill(s, _ptr(table), Hc, _ptr(uk), _ptr(starts), _ptr(cnt.int().contiguous()),
                                           int(uk.shape[0]), _ptr(err)), "pcs_trk_grid_fill")
            n_ns = int(ns_r.sum().item())
            ref_off = torch.tensor([0, n_ns, n_ref], dtype=torch.int32, device=dev)
            Hm = next_pow2(max(2 * nm, 1024))
            sc = dict(mov_table=_e((Hm, 4), torch.int32, dev), mov_sorted=_e((nm, 4), torch.float32, dev),
                      mov_sidx=_e(nm, torch.int32, dev), mov_cells=_e(nm, torch.int32, dev), mov_ctr=_z(4, torch.int32, dev),
                      nn_fwd=_e(nm, torch.int32, dev), nn_bwd=_e(max(n_ns, 1), torch.int32, dev),
                      boff=_z(2, torch.int32, dev), mvbeg=_z(1, torch.int32, dev), mvend=_z(1, torch.int32, dev),
                      sec_fwd=_z(nm, torch.float32, dev), sec_bwd=_z(max(n_ns, 1), torch.float32, dev),
                      disp=_z(nm, torch.float32, dev), dmax=_z(1, torch.int32, dev),
                      mom=_z((C, 17), torch.float64, dev), Ti=_z((C, 12), torch.float64, dev),
                      mu=_z((C, 6), torch.float64, dev), l1_sum=_z((C, 2), torch.float64, dev), l1_n=_z(C, torch.float64, dev),
                      phase=_z(1, torch.int32, dev), cd=_z(1, torch.int32, dev), iters=iters,
                      itcnt=_z((int(max_iter) + 2) * 3, torch.int32, dev), last=_z(1, torch.float64, dev),
                      loss=_z(1, torch.float64, dev), match_cnt=_z(C, torch.int32, dev))
            _lib.check(L.pcs_trk_table_clear(s, _ptr(sc["mov_table"]), Hm, _ptr(sc["mov_ctr"])), "pcs_trk_table_clear")
            st = IcpStruct()
            r_eff = float(np.float32((float(radius) ** 2 + df ** 2) ** 0.5))  # :111-112
            _fill(st, keep, J=1, G=C, act=torch.ones(1, dtype=torch.int32, device=dev), ref_group=_z(1, torch.int32, dev),
                  ref_group_all=torch.ones(1, dtype=torch.int32, device=dev), skipmask=_z(1, torch.int32, dev),
                  ref_off=ref_off, g_inst=_z(C, torch.int32, dev), ref_table=table, ref_H=Hc, ref_pts=rv, mov_H=Hm, mv=mv,
                  mv_gid=mv_gid, mv_inst=mv_inst, n_mv=n_mv, mv_cap=nm, vdeg=vdeg, lo=lo, cs=cs, rings=1, radius=r_eff,
                  df=df, angle_reg=float(angle_regularizer), max_iter=int(max_iter),
                  stopping_delta=float(stopping_delta), want_l1=1, want_ratio=1, T=T, l1_err=l1, ratio=ratio,
                  prof=None, **sc)
            _lib.check(L.pcs_trk_icp(s, ctypes.byref(st)), "pcs_trk_icp")
            moved[idx_m] = mv
    T44 = torch.zeros(C, 4, 4, dtype=torch.float64, device=dev)
    T44[:, :3, :3] = T[:, :9].reshape(C, 3, 3)
    T44[:, :3, 3] = T[:, 9:]
    T44[:, 3, 3] = 1.0
    info = torch.stack([_z(1, torch.int32, dev)[0], iters[0], _z(1, torch.int32, dev)[0], _z(1, torch.int32, dev)[0]])
    return moved, T44, l1, ratio, info
