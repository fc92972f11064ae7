"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python from /root/reference on the CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    python -m oracle.gen_golden

See oracle/ref_import.py for exactly what is real reference code and what is substituted.  The
fixtures pin oracle/cpu_ops.py + oracle/registration_np.py (tests/test_oracle.py) and are the
reference answers the CUDA path is compared with on the GPU box (tests/test_*_gpu.py).
"""
import os
import tempfile

import numpy as np
import torch

from . import ref_import as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _scene(seq, frames, beams, az):
    from pcseqlearning_b200.synthetic import generate_sequence, sequence_fxyz
    b = generate_sequence(seq, num_frames=frames, num_beams=beams, num_azimuth=az)
    f = sequence_fxyz(b)
    seg = b["segmentation_label"]
    return b, f, seg


def gen_radius_graph():
    gu = R.load("pcdet.models.model_utils.graph_utils")
    _, f, seg = _scene(1, 3, 24, 700)
    f = f[seg < 17]  # ground removed by label
    cases = {}
    for name, (radius, K, sort) in dict(r125=(1.25, 32, True), r075=(0.75, 32, True), r025=(0.25, 32, True),
                                        nn05=(0.5, 1, True), unsorted=(0.75, 8, False)).items():
        g = gu.RadiusGraph(runtime_cfg={}, model_cfg=dict(RADIUS=radius, MAX_NUM_NEIGHBORS=K, SORT_BY_DIST=sort,
                                                          RELATIVE_KEY="fxyz"))
        er, eq, _ = g(R.edict(fxyz=f.clone()), R.edict(fxyz=f.clone()))
        cases[name + "_eref"] = er.numpy()
        cases[name + "_equery"] = eq.numpy()
        cases[name + "_cfg"] = np.array([radius, K, int(sort)], np.float64)
    # cross-frame nearest neighbour, the way register_to_next_frame drives the graph (:107-137)
    fr = f[:, 0].round().long()
    a, b = f[fr == 0].clone(), f[fr == 2].clone()
    g = gu.RadiusGraph(runtime_cfg={}, model_cfg=dict(RADIUS=2.5, MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True,
                                                      RELATIVE_KEY="fxyz"))
    g.radius = (2.5 ** 2 + 2 ** 2) ** 0.5
    g.qmin[0] = 2
    g.qmax[0] = 2
    er, eq, _ = g(R.edict(fxyz=b), R.edict(fxyz=a))  # ref = frame 2, query = frame 0
    cases["cross_eref"] = er.numpy()
    cases["cross_equery"] = eq.numpy()
    cases["cross_ref"] = b.numpy()
    cases["cross_query"] = a.numpy()
    cases["points"] = f.numpy()
    np.savez_compressed(os.path.join(OUT, "radius_graph.npz"), **cases)
    print("radius_graph.npz", f.shape, {k: v.shape for k, v in cases.items() if k.endswith("eref")})


def gen_grid_sampling():
    gs = R.load("pcdet.models.model_utils.grid_sampling")
    ct = R.load("pcdet.models.registration.preprocessors.cluster_tracking")
    from torch_scatter import scatter
    _, f, seg = _scene(2, 3, 24, 700)
    out = dict(points=f.numpy())
    for name, size in dict(sub008=[0.08, 0.08, 0.08], lvl0=[0.4, 0.4, 0.6], lvl2=[0.1, 0.1, 0.15]).items():
        s = gs.GridSampling3D(size)
        sampled, inv = s(f.clone(), return_inverse=True)
        out[name + "_sampled"] = sampled.numpy()
        out[name + "_inv"] = inv.numpy()
        out[name + "_size"] = np.array(size)
    # the 0.08 m pick-one subsample of simple_reg.py:119-124
    inv = torch.from_numpy(out["sub008_inv"])
    out["sub008_pick"] = scatter(torch.arange(f.shape[0]), inv, dim_size=int(inv.max()) + 1, dim=0,
                                 reduce="max").numpy()
    # sample_frame (cluster_tracking.py:39-51) on one frame with synthetic components
    fr0 = f[f[:, 0] == 0].clone()
    rng = np.random.default_rng(5)
    comp = torch.from_numpy(rng.integers(0, 40, fr0.shape[0]))
    stat = torch.from_numpy(rng.random(fr0.shape[0]) < 0.3)
    frame = R.edict(fxyz=fr0, stationary=stat, component=comp,
                    frame=torch.zeros(fr0.shape[0], 1, dtype=torch.int32))
    sf = ct.sample_frame(gs.GridSampling3D([0.2, 0.2, 0.3]), frame)
    out.update(sf_in_fxyz=fr0.numpy(), sf_in_comp=comp.numpy(), sf_in_stat=stat.numpy(),
               sf_fxyz=sf.fxyz.numpy(), sf_stat=sf.stationary.numpy(), sf_comp=sf.component.numpy(),
               sf_frame=sf.frame.numpy())
    np.savez_compressed(os.path.join(OUT, "grid_sampling.npz"), **out)
    print("grid_sampling.npz", f.shape)


def gen_proposal():
    """ClusterProposal.propose_cluster (cluster_proposal.py:34-88) with the north-star yaml's GRAPH block."""
    cp = R.load("pcdet.models.registration.preprocessors.cluster_proposal")
    b, f, seg = _scene(4, 13, 24, 600)  # 13 frames -> two 10-frame chunks
    keep = seg < 17
    f = f[keep]
    sweep = b["point_sweep"][keep]
    keys = ["component_rad1x25", "component_rad0x75", "component_rad0x25"]
    with tempfile.TemporaryDirectory() as d:
        cfg = R.edict(GRAPH=dict(TYPE="RadiusGraph", RADIUS=[1.25, 0.75, 0.25], MAX_NUM_NEIGHBORS=32,
                                 SORT_BY_DIST=True, RELATIVE_KEY="fxyz"), COMPONENT_KEYS=keys, DIR=d)
        mod = cp.ClusterProposal(cfg, {})
        seq = R.edict(point_fxyz=f.clone(), point_sweep=sweep.clone(), frame_id=b["frame_id"][0])
        seq = mod.propose_cluster(seq)
    out = dict(points=f.numpy(), sweep=sweep.numpy())
    for k in keys:
        out[k] = seq["point_" + k].numpy()
    np.savez_compressed(os.path.join(OUT, "proposal.npz"), **out)
    print("proposal.npz", f.shape, {k: int(out[k].max()) + 1 for k in keys})


def gen_registration():
    """register_to_next_frame (registration_utils.py:83-206) at the three levels of the yaml."""
    gu = R.load("pcdet.models.model_utils.graph_utils")
    gs = R.load("pcdet.models.model_utils.grid_sampling")
    ru = R.load("pcdet.models.registration.preprocessors.registration_utils")
    ct = R.load("pcdet.models.registration.preprocessors.cluster_tracking")
    from . import cpu_ops as ops
    b, f, seg = _scene(3, 4, 32, 900)
    f = f[seg < 17].numpy()
    comp, _ = ops.propose_clusters(f, 0.75)
    fr = np.rint(f[:, 0]).astype(int)
    out = {}
    for case, (fa, fb) in dict(fwd=(0, 1), bwd=(2, 0)).items():
        A, B = f[fr == fa], f[fr == fb]
        cA = comp[fr == fa]
        cA = cA - cA.min()
        C = int(cA.max()) + 1
        cB = comp[fr == fb]
        # components with a large extent are "stationary" (cluster_tracking.py:860-861)
        fa_pts = R.edict(fxyz=torch.from_numpy(A), component=torch.from_numpy(cA))
        diamA = ct.component_diameter(fa_pts)[torch.from_numpy(cA)].numpy()
        statA = diamA > 12.5
        statB = np.zeros(B.shape[0], bool)
        for lvl, (radius, vsz) in enumerate(zip([2.5, 1.25, 1.0], [[0.4, 0.4, 0.6], [0.2, 0.2, 0.3],
                                                                   [0.1, 0.1, 0.15]])):
            sampler = gs.GridSampling3D(vsz)
            frameA = R.edict(fxyz=torch.from_numpy(A.copy()), stationary=torch.from_numpy(statA),
                             component=torch.from_numpy(cA),
                             frame=torch.full((A.shape[0], 1), fa, dtype=torch.int32))
            frameB = R.edict(fxyz=torch.from_numpy(B.copy()), stationary=torch.from_numpy(statB),
                             component=torch.from_numpy(cB),
                             frame=torch.full((B.shape[0], 1), fb, dtype=torch.int32))
            sA = ct.sample_frame(sampler, frameA)
            sB = ct.sample_frame(sampler, frameB)
            g = gu.RadiusGraph(runtime_cfg={}, model_cfg=dict(RADIUS=radius, MAX_NUM_NEIGHBORS=1,
                                                              SORT_BY_DIST=True, RELATIVE_KEY="fxyz"))
            pre = {k: sA[k].clone().numpy() for k in ["fxyz", "stationary", "component"]}
            preB = {k: sB[k].clone().numpy() for k in ["fxyz", "stationary"]}
            mv, T, l1, ratio = ru.register_to_next_frame(g, sA, sB, C, 10, max_iter=80, stopping_delta=0.05)
            p = f"{case}_l{lvl}_"
            out.update({p + "mov_fxyz": pre["fxyz"], p + "mov_stat": pre["stationary"],
                        p + "mov_comp": pre["component"], p + "ref_fxyz": preB["fxyz"],
                        p + "ref_stat": preB["stationary"], p + "C": np.array(C), p + "radius": np.array(radius),
                        p + "T": T.numpy(), p + "l1": l1.numpy(), p + "ratio": ratio.numpy(),
                        p + "moved": mv.fxyz.numpy()})
    np.savez_compressed(os.path.join(OUT, "registration.npz"), **out)
    print("registration.npz", {k: v.shape for k, v in out.items() if k.endswith("_T")})


def gen_ground():
    """ground_plane_removal (preprocessor_utils.py:352-419) with the north-star yaml's GroundPlaneRemover block."""
    pu = R.load("pcdet.models.registration.preprocessors.preprocessor_utils")
    from . import cpu_ops as ops
    b, f, seg = _scene(6, 6, 32, 900)
    pick = ops.subsample_pick(f.numpy())
    f = f[torch.from_numpy(pick)]
    seg = seg[torch.from_numpy(pick)]
    cfg = R.edict(PILLAR_SIZE=[2, 2], LR=0.01, DECAY_STEPS=[1600], RIGID_WEIGHT=0.5, MAX_NUM_ITERS=10000,
                  TRUNCATE_HEIGHT=[0.5], RANSAC=True, SIGMA2=0.0025, JointOpt=True, K=8)
    torch.manual_seed(0)
    height, horizon, err, pillar_height, pillar_min_z = pu.ground_plane_removal(f.clone(), cfg)
    np.savez_compressed(os.path.join(OUT, "ground.npz"), points=f.numpy(), seg=seg.numpy(),
                        height=height.numpy(), horizon=horizon.numpy(), error=err.numpy(),
                        pillar_height=pillar_height.numpy(), pillar_min_z=pillar_min_z.numpy())
    gm = (height < 0.5).numpy()
    s = seg.numpy()
    print("ground.npz", f.shape, "removed", gm.sum(), "ground-label coverage", (gm & (s >= 17)).sum() / max((s >= 17).sum(), 1),
          "foreground removed", (gm & (s > 0) & (s <= 7)).sum() / max(((s > 0) & (s <= 7)).sum(), 1))


def gen_tracking():
    """ClusterTracking.track_frame (cluster_tracking.py:430-787) for one anchor frame of a small 17-frame scene,
    run by the reference's own code (CPU)."""
    import time
    ct = R.load("pcdet.models.registration.preprocessors.cluster_tracking")
    cu = R.load("pcdet.utils.common_utils")
    from . import cpu_ops as ops
    b, f, seg = _scene(8, 17, 24, 600)
    keep = seg < 17
    f, seg = f[keep], seg[keep]
    sweep = b["point_sweep"][keep]
    comp, _ = ops.propose_clusters(f.numpy(), 0.75)
    comp = torch.from_numpy(comp)
    cfg = R.edict(
        ANGLE_REGULARIZER=10, COMPONENT_KEYS=["component_rad0x75"],
        REGISTRATION=dict(GRAPH=dict(TYPE="RadiusGraph", RADIUS=[2.5, 1.25, 1.0], MAX_NUM_NEIGHBORS=1,
                                     SORT_BY_DIST=True, RELATIVE_KEY="fxyz"),
                          VOXEL_SIZE=[[0.4, 0.4, 0.6], [0.2, 0.2, 0.3], [0.1, 0.1, 0.15]],
                          STOPPING_DELTA=[0.05, 0.05, 0.05]),
        NN_GRAPH=dict(TYPE="RadiusGraph", RADIUS=0.5, MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True, RELATIVE_KEY="fxyz"),
        DIR="/tmp/unused",
        TRACKING_PARAMS=dict(REGISTRATION_ERROR_COEFFICIENT=0.13, TRACK_INTERVAL=8, ANGLE_THRESHOLD=45,
                             MIN_MOVE_FRAME=6))
    mod = ct.ClusterTracking(cfg, {})
    seq_points = R.edict(fxyz=f.clone(), frame=sweep.clone(), gt_box_id=torch.zeros_like(comp) - 1,
                         segmentation_label=seg.clone(), component=comp.clone())
    diam = ct.component_diameter(seq_points)[seq_points.component]
    seq_points.component_diameter = diam
    seq_points.stationary = diam > 12.5
    seq_points.extracted = torch.zeros_like(seq_points.fxyz[:, 0]).bool()
    anchor = 8
    frame_mask = (seq_points.fxyz[:, 0] == anchor).reshape(-1)
    frame_points = R.edict(**cu.filter_dict(seq_points, frame_mask))
    frame_points.component = frame_points.component - frame_points.component.min()
    t0 = time.time()
    torch.manual_seed(0)
    ex = mod.track_frame(seq_points, frame_points, None)
    print(f"reference track_frame took {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(OUT, "tracking.npz"), points=f.numpy(), sweep=sweep.numpy(), seg=seg.numpy(),
                        component=comp.numpy(), anchor=np.array(anchor),
                        ex_fxyz=ex.fxyz.numpy(), ex_component=ex.component.numpy(),
                        ex_original_indices=ex.original_indices.numpy(), ex_frame_indices=ex.frame_indices.numpy(),
                        ex_moving=ex.moving.numpy(), transforms=ex.transforms.numpy())
    print("tracking.npz", f.shape, "extracted", ex.fxyz.shape, "components kept", ex.component.unique().numel(),
          "transforms", tuple(ex.transforms.shape))


def _ref_function(path, name, glb):
    """Compile ONE function of a reference source file (the module itself may not be importable here) and return
    it; the code that runs is the reference's, read from /root/reference at generation time."""
    import ast
    src = open(os.path.join(R.REF, path)).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            code = compile(mod, os.path.join(R.REF, path), "exec")
            ns = dict(glb)
            exec(code, ns)
            return ns[name]
    raise KeyError(name)


def gen_eval_tracking():
    """GT formatting (simple_reg.py:35-101), ClusterProposal.evaluate_proposal (cluster_proposal.py:142-285), one
    ClusterTracking.track_frame with every register_to_next_frame call recorded (teacher forcing), and
    extract_traces_and_update_boxes (cluster_tracking.py:287-428) -- all run by the reference's own code on the CPU."""
    import time
    ct = R.load("pcdet.models.registration.preprocessors.cluster_tracking")
    from easydict import EasyDict  # the shim installed by ref_import
    cp = R.load("pcdet.models.registration.preprocessors.cluster_proposal")
    ru = R.load("pcdet.models.registration.preprocessors.registration_utils")
    cu = R.load("pcdet.utils.common_utils")
    bu = R.load("pcdet.utils.box_utils")
    from . import cpu_ops as ops
    b, f_all, seg_all = _scene(8, 17, 24, 600)
    inst_all = b["instance_label"]
    sweep_all = b["point_sweep"]
    # --- reference GT formatting ------------------------------------------------------------------------------
    fmt = _ref_function("pcdet/models/registration/simple_reg.py", "format_boxes",
                        dict(torch=torch, np=np, EasyDict=EasyDict, common_utils=cu, box_utils=bu))
    seq = EasyDict(dict(point_sweep=sweep_all.clone()))
    for key in ["gt_box_cls_label", "gt_box_attr", "augmented", "num_points_in_gt", "gt_boxes", "obj_ids",
                "gt_box_corners_3d"]:
        seq[key] = b[key][0]
    seq = fmt(None, seq)
    box_out = {f"box_{k}": seq[k].numpy() for k in ["gt_box_attr", "gt_box_cls_label", "gt_box_frame",
                                                    "gt_box_track_label", "gt_box_velo", "moving"]}
    # --- ground removal by label; a third of the ground points stay "above ground" in the full arrays ----------
    rng = np.random.default_rng(5)
    height = torch.where(seg_all < 17, torch.ones(seg_all.shape[0]),
                         torch.from_numpy(np.where(rng.random(seg_all.shape[0]) < 0.3, 0.05, -0.1)).float())
    keep = seg_all < 17
    f, seg, sweep, inst = f_all[keep], seg_all[keep], sweep_all[keep], inst_all[keep]
    comps = {}
    for key, r in (("component_rad1x25", 1.25), ("component_rad0x75", 0.75), ("component_rad0x25", 0.25)):
        comps[key] = torch.from_numpy(ops.propose_clusters(f.numpy(), r)[0])
    # --- reference evaluate_proposal -------------------------------------------------------------------------
    seq.update(dict(point_fxyz=f.clone(), point_sweep=sweep.clone(), segmentation_label=seg.clone(),
                    instance_label=inst.clone(), frame_id=np.array(["seq_golden_000"])))
    for key, c in comps.items():
        seq[f"point_{key}"] = c.clone()
    prop = cp.ClusterProposal.__new__(cp.ClusterProposal)
    torch.nn.Module.__init__(prop)
    prop.component_keys = list(comps.keys())
    prop.model_cfg = EasyDict(dict(DIR=tempfile.mkdtemp()))
    t0 = time.time()
    seq = prop.evaluate_proposal(seq)
    print(f"reference evaluate_proposal took {time.time() - t0:.1f}s")
    eval_out = {f"eval_{k}": seq[k].numpy() for k in ["gt_box_best_iou", "gt_trace_best_iou", "point_gt_box_id",
                                                      "point_gt_trace_id", "point_pred_trace_id", "point_pred_box_id"]}
    # --- reference track_frame with recorded ICP calls -------------------------------------------------------------
    cfg = R.edict(
        ANGLE_REGULARIZER=10, COMPONENT_KEYS=["component_rad0x75"],
        REGISTRATION=dict(GRAPH=dict(TYPE="RadiusGraph", RADIUS=[2.5, 1.25, 1.0], MAX_NUM_NEIGHBORS=1,
                                     SORT_BY_DIST=True, RELATIVE_KEY="fxyz"),
                          VOXEL_SIZE=[[0.4, 0.4, 0.6], [0.2, 0.2, 0.3], [0.1, 0.1, 0.15]],
                          STOPPING_DELTA=[0.05, 0.05, 0.05]),
        NN_GRAPH=dict(TYPE="RadiusGraph", RADIUS=0.5, MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True, RELATIVE_KEY="fxyz"),
        DIR="/tmp/unused",
        TRACKING_PARAMS=dict(REGISTRATION_ERROR_COEFFICIENT=0.13, TRACK_INTERVAL=8, ANGLE_THRESHOLD=45,
                             MIN_MOVE_FRAME=6))
    mod = ct.ClusterTracking(cfg, {})
    comp = comps["component_rad0x75"]
    seq_points = R.edict(fxyz=f.clone(), frame=sweep.clone(), gt_box_id=seq["point_gt_box_id"].clone(),
                         segmentation_label=seg.clone(), instance_label=inst.clone(), component=comp.clone())
    diam = ct.component_diameter(seq_points)[seq_points.component]
    seq_points.component_diameter = diam
    seq_points.stationary = diam > 12.5
    seq_points.extracted = torch.zeros_like(seq_points.fxyz[:, 0]).bool()
    anchor = 8
    frame_mask = (seq_points.fxyz[:, 0] == anchor).reshape(-1)
    frame_points = R.edict(**cu.filter_dict(seq_points, frame_mask))
    frame_points.component = frame_points.component - frame_points.component.min()
    calls = []
    orig_reg = ct.register_to_next_frame

    def recording_reg(graph, moving, ref, num_components, angle_regularizer=10, max_iter=20, stopping_delta=5e-2):
        rec = dict(mov=moving.fxyz.clone().numpy(), mov_comp=moving.component.clone().numpy(),
                   mov_stat=moving.stationary.clone().numpy(), ref=ref.fxyz.clone().numpy(),
                   ref_stat=ref.stationary.clone().numpy(), radius=float(graph.radius), C=int(num_components),
                   max_iter=int(max_iter), delta=float(stopping_delta), reg=float(angle_regularizer))
        out = orig_reg(graph, moving, ref, num_components, angle_regularizer, max_iter, stopping_delta)
        rec.update(T=out[1].clone().numpy(), l1=out[2].clone().numpy(), ratio=out[3].clone().numpy(),
                   moved=out[0].fxyz.clone().numpy())
        calls.append(rec)
        return out

    ct.register_to_next_frame = recording_reg
    t0 = time.time()
    torch.manual_seed(0)
    seq_boxes = mod.format_boxes(seq, 17)
    seq_boxes.best_iou = torch.zeros_like(seq_boxes.attr[:, 0])
    ex = mod.track_frame(seq_points, frame_points, seq_boxes)
    ct.register_to_next_frame = orig_reg
    print(f"reference track_frame took {time.time() - t0:.1f}s, {len(calls)} ICP calls")
    ex_in = {f"ex_{k}": ex[k].numpy() for k in ["fxyz", "component", "segmentation_label", "frame_indices",
                                                "original_indices", "moving", "transforms"]}
    # --- reference extract_traces_and_update_boxes --------------------------------------------------------------------
    all_mask = height > 0
    all_points = R.edict(fxyz=f_all[all_mask].clone(), frame=sweep_all[all_mask].clone(), height=height[all_mask].clone(),
                         full_instance_label=inst_all[all_mask].clone(),
                         full_segmentation_label=seg_all[all_mask].clone())
    mod.visualize = False
    t0 = time.time()
    ex_copy = R.edict(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in ex.items()})
    full, seq_boxes = mod.extract_traces_and_update_boxes(all_points, ex_copy, seq_boxes)
    print(f"reference extract_traces_and_update_boxes took {time.time() - t0:.1f}s")
    full_out = {f"full_{k}": full[k].numpy() for k in ["fxyz", "component", "segmentation_label", "instance_label",
                                                       "original_indices", "frame_indices", "moving", "component_hit",
                                                       "component_size"]}
    np.savez_compressed(
        os.path.join(OUT, "eval_tracking.npz"), points_all=f_all.numpy(), sweep_all=sweep_all.numpy(),
        seg_all=seg_all.numpy(), inst_all=inst_all.numpy(), height_all=height.numpy(),
        raw_gt_box_attr=b["gt_box_attr"][0].numpy(), raw_gt_box_cls_label=b["gt_box_cls_label"][0].numpy(),
        raw_obj_ids=np.asarray(b["obj_ids"][0]).astype(str), raw_augmented=b["augmented"][0].numpy(),
        raw_num_points_in_gt=b["num_points_in_gt"][0].numpy(),
        comp_rad1x25=comps["component_rad1x25"].numpy(), comp_rad0x75=comps["component_rad0x75"].numpy(),
        comp_rad0x25=comps["component_rad0x25"].numpy(), anchor=np.array(anchor),
        best_iou_after_tracking=seq_boxes.best_iou.numpy(), **box_out, **eval_out, **ex_in, **full_out)
    # teacher-forcing records: a spread of the calls (all three levels, near and far target frames)
    pick = sorted(set(list(range(0, 6)) + list(range(6, len(calls), max(1, len(calls) // 14)))))[:20]
    steps = {"n_calls": np.array(len(calls)), "picked": np.array(pick)}
    for i, ci in enumerate(pick):
        for k, v in calls[ci].items():
            steps[f"c{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "tracking_steps.npz"), **steps)
    print("eval_tracking.npz / tracking_steps.npz", f.shape, "extracted", ex.fxyz.shape, "full", full.fxyz.shape,
          "boxes", seq_boxes.attr.shape, "best_iou max", float(seq_boxes.best_iou.max()), "calls", len(calls), "picked", len(pick))



def gen_fullscale():
    """Config-scale fixtures (BASELINE.json configs: 64 beams x 2650 azimuth steps): two frames at full density through
    the reference's ClusterProposal.propose_cluster (3 radii: 1.25 m cells hold thousands of points there and the
    K = 32 truncation is active) and one full-density frame pair through sample_frame + register_to_next_frame."""
    import time
    cp = R.load("pcdet.models.registration.preprocessors.cluster_proposal")
    ct = R.load("pcdet.models.registration.preprocessors.cluster_tracking")
    ru = R.load("pcdet.models.registration.preprocessors.registration_utils")
    gs = R.load("pcdet.models.model_utils.grid_sampling")
    gu = R.load("pcdet.models.model_utils.graph_utils")
    from . import cpu_ops as ops
    b, f, seg = _scene(0, 2, 64, 2650)
    pick = ops.subsample_pick(f.numpy())
    f, seg = f[torch.from_numpy(pick)], seg[torch.from_numpy(pick)]
    keep = seg < 17
    f = f[keep].contiguous()
    sweep = f[:, 0].round().int().reshape(-1, 1)
    cfg = R.edict(GRAPH=dict(TYPE="RadiusGraph", RADIUS=[1.25, 0.75, 0.25], MAX_NUM_NEIGHBORS=32, SORT_BY_DIST=True,
                             RELATIVE_KEY="fxyz"),
                  COMPONENT_KEYS=["component_rad1x25", "component_rad0x75", "component_rad0x25"], DIR=tempfile.mkdtemp())
    mod = cp.ClusterProposal(runtime_cfg={}, model_cfg=cfg)
    t0 = time.time()
    seq = mod.propose_cluster(R.edict(point_fxyz=f.clone(), point_sweep=sweep.clone(), frame_id=np.array(["full_000"])))
    print(f"reference propose_cluster on {f.shape[0]} points took {time.time() - t0:.1f}s")
    out = dict(points=f.numpy())
    for k in cfg.COMPONENT_KEYS:
        out[k] = seq[f"point_{k}"].numpy()
    # one ICP pair at full density, levels 0 and 2
    comp = seq["point_component_rad0x75"]
    fr = f[:, 0].round().long()
    diam = ct.component_diameter(R.edict(fxyz=f, component=comp))[comp]
    stat = diam > 12.5
    a = R.edict(fxyz=f[fr == 0].clone(), component=(comp[fr == 0] - comp[fr == 0].min()), frame=sweep[fr == 0].clone(),
                stationary=stat[fr == 0].clone())
    nb = R.edict(fxyz=f[fr == 1].clone(), component=comp[fr == 1].clone(), frame=sweep[fr == 1].clone(),
                 stationary=stat[fr == 1].clone())
    C = int(a.component.max()) + 1
    for lvl, (radius, vs) in {0: (2.5, [0.4, 0.4, 0.6]), 2: (1.0, [0.1, 0.1, 0.15])}.items():
        sampler = gs.GridSampling3D(vs)
        sa, sb = ct.sample_frame(sampler, a), ct.sample_frame(sampler, nb)
        g = gu.RadiusGraph(runtime_cfg={}, model_cfg=dict(RADIUS=radius, MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True,
                                                          RELATIVE_KEY="fxyz"))
        p = f"icp_l{lvl}_"
        out.update({p + "mov": sa.fxyz.clone().numpy(), p + "mov_comp": sa.component.clone().numpy(),
                    p + "mov_stat": sa.stationary.clone().numpy(), p + "ref": sb.fxyz.clone().numpy(),
                    p + "ref_stat": sb.stationary.clone().numpy(), p + "C": np.array(C), p + "radius": np.array(radius)})
        t0 = time.time()
        mv, T, l1, ratio = ru.register_to_next_frame(g, sa, sb, C, 10, max_iter=80, stopping_delta=0.05)
        print(f"reference register_to_next_frame level {lvl}: {sa.fxyz.shape[0]} x {sb.fxyz.shape[0]} voxels, "
              f"{time.time() - t0:.1f}s")
        out.update({p + "T": T.numpy(), p + "l1": l1.numpy(), p + "ratio": ratio.numpy()})
    np.savez_compressed(os.path.join(OUT, "fullscale.npz"), **out)
    print("fullscale.npz", f.shape, {k: int(out[k].max()) + 1 for k in cfg.COMPONENT_KEYS})

def main():
    """python -m oracle.gen_golden [name ...]   (default: every fixture)"""
    import sys
    os.makedirs(OUT, exist_ok=True)
    gens = dict(radius_graph=gen_radius_graph, grid_sampling=gen_grid_sampling, proposal=gen_proposal,
                registration=gen_registration, ground=gen_ground, tracking=gen_tracking,
                eval_tracking=gen_eval_tracking, fullscale=gen_fullscale)
    for name in (sys.argv[1:] or list(gens)):
        torch.manual_seed(0)
        gens[name]()


if __name__ == "__main__":
    main()
