"""CPU tests: the oracle restatement against the golden vectors recorded from the reference's own Python
(oracle/gen_golden.py), against the reference's single known-answer case, and against independent
cross-checks (brute force, scipy)."""
import os

import numpy as np
import pytest

from oracle import cpu_ops as ops
from oracle import registration_np as reg
from helpers import assert_neighbor_sets_equal, canonical_labels, component_centers, transform_errors


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_reference_known_answer_three_points():
    # pcdet/ops/torch_hash/torch_hash_modules.py:146-151 -- the only KAT the reference ships
    pts = np.array([[0, 0.0, 0.0], [0, 0.1, 0.1], [0, 0.2, 0.2]], np.float32)
    er, eq = ops.radius_graph_build(pts, pts, 0.15, 1, True)
    assert er.tolist() == [0, 1, 2] and eq.tolist() == [0, 1, 2]


@pytest.mark.parametrize("case", ["r125", "r075", "r025", "nn05", "unsorted"])
def test_radius_graph_matches_reference_python(golden_dir, case):
    g = _load(golden_dir, "radius_graph.npz")
    pts = g["points"]
    radius, K, sort = g[case + "_cfg"]
    er, eq = ops.radius_graph_build(pts, pts, float(radius), int(K), bool(sort))
    # same sequential insertion order => identical rows, not just identical sets
    np.testing.assert_array_equal(er, g[case + "_eref"])
    np.testing.assert_array_equal(eq, g[case + "_equery"])


def test_radius_graph_cross_frame(golden_dir):
    g = _load(golden_dir, "radius_graph.npz")
    r = (2.5 ** 2 + 2 ** 2) ** 0.5
    er, eq = ops.radius_graph_build(g["cross_ref"], g["cross_query"], r, 1, True, qmin=[2, -1, -1, -1],
                                    qmax=[2, 1, 1, 1])
    np.testing.assert_array_equal(er, g["cross_eref"])
    np.testing.assert_array_equal(eq, g["cross_equery"])


def test_radius_graph_vs_brute_force():
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, 6, (1500, 4)).astype(np.float32)
    pts[:, 0] = rng.integers(0, 2, 1500)
    er, eq = ops.radius_graph_build(pts, pts, 0.7, 16, True)
    d = ((pts[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    for q in rng.integers(0, 1500, 50):
        nb = np.nonzero(d[q] <= 0.7 ** 2 - 1e-6)[0]
        want = set(nb[np.argsort(d[q][nb])][:16].tolist())
        got = set(er[eq == q].tolist())
        if len(nb) <= 16:
            assert want <= got
        else:
            assert len(got) == 16


@pytest.mark.parametrize("name", ["sub008", "lvl0", "lvl2"])
def test_grid_sampling_matches_reference_python(golden_dir, name):
    g = _load(golden_dir, "grid_sampling.npz")
    sampled, inv = ops.grid_sampling(g["points"], g[name + "_size"].tolist())
    np.testing.assert_array_equal(inv, g[name + "_inv"])
    np.testing.assert_allclose(sampled, g[name + "_sampled"], rtol=0, atol=2e-5)


def test_subsample_pick(golden_dir):
    g = _load(golden_dir, "grid_sampling.npz")
    np.testing.assert_array_equal(ops.subsample_pick(g["points"]), g["sub008_pick"])


def test_sample_frame(golden_dir):
    g = _load(golden_dir, "grid_sampling.npz")
    n = g["sf_in_fxyz"].shape[0]
    sf = reg.sample_frame(g["sf_in_fxyz"], g["sf_in_stat"], g["sf_in_comp"], np.zeros(n, np.int64), [0.2, 0.2, 0.3])
    np.testing.assert_allclose(sf["fxyz"], g["sf_fxyz"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(sf["stationary"], g["sf_stat"])
    np.testing.assert_array_equal(sf["component"], g["sf_comp"])
    np.testing.assert_array_equal(sf["frame"], g["sf_frame"])


def test_connected_components_vs_scipy():
    rng = np.random.default_rng(2)
    for n, e in ((50, 30), (2000, 1500), (2000, 6000)):
        e0 = rng.integers(0, n, e)
        e1 = rng.integers(0, n, e)
        n1, l1 = ops.connected_components(e0, e1, n)
        n2, l2 = ops.connected_components_c(e0, e1, n)
        assert n1 == n2
        np.testing.assert_array_equal(l1, l2)
        np.testing.assert_array_equal(canonical_labels(l1), l1)  # scipy numbering is already canonical


def test_proposal_matches_reference_python(golden_dir):
    g = _load(golden_dir, "proposal.npz")
    for key, r in (("component_rad1x25", 1.25), ("component_rad0x75", 0.75), ("component_rad0x25", 0.25)):
        comp, _ = ops.propose_clusters(g["points"], r)
        np.testing.assert_array_equal(comp, g[key])


@pytest.mark.parametrize("case", ["fwd", "bwd"])
@pytest.mark.parametrize("lvl", [0, 1, 2])
def test_registration_matches_reference_python(golden_dir, case, lvl):
    g = _load(golden_dir, "registration.npz")
    p = f"{case}_l{lvl}_"
    moved, T, l1, ratio, _ = reg.register_to_next_frame(
        g[p + "mov_fxyz"], g[p + "mov_comp"], g[p + "mov_stat"], g[p + "ref_fxyz"], g[p + "ref_stat"],
        int(g[p + "C"]), float(g[p + "radius"]), 10, 80, 0.05)
    C = int(g[p + "C"])
    ang, dt = transform_errors(T, g[p + "T"], component_centers(g[p + "mov_fxyz"], g[p + "mov_comp"], C))
    # north-star tolerance: 1e-4 relative rotation / translation error
    assert ang.max() < 1e-4 and dt.max() < 1e-4
    np.testing.assert_allclose(l1, g[p + "l1"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(ratio, g[p + "ratio"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(moved, g[p + "moved"], rtol=0, atol=5e-4)  # reference centroids are fp32 sums in unspecified order: ~3e-4 m noise on 100 m-wide components


def test_points_in_boxes():
    rng = np.random.default_rng(3)
    boxes = np.array([[0, 0, 0, 4, 2, 1.5, 0.3], [10, 5, 1, 1, 1, 2, -1.0]], np.float32)
    pts = rng.uniform(-3, 12, (500, 3)).astype(np.float32)
    m = ops.points_in_boxes(pts, boxes)
    for b in range(2):
        c, s = np.cos(-boxes[b, 6]), np.sin(-boxes[b, 6])
        d = pts - boxes[b, :3]
        lx, ly = d[:, 0] * c - d[:, 1] * s, d[:, 0] * s + d[:, 1] * c
        inside = (np.abs(d[:, 2]) <= boxes[b, 5] / 2) & (np.abs(lx) < boxes[b, 3] / 2 + 1e-2) & (np.abs(ly) < boxes[b, 4] / 2 + 1e-2)
        assert (m[b].astype(bool) == inside).mean() > 0.995


@pytest.mark.parametrize("case", ["r125", "r075", "r025", "nn05"])
def test_oracle_vs_reference_cuda_op(golden_dir, case):
    """oracle.c against outputs of the reference's OWN CUDA op (torch_hash_cuda compiled from the unmodified
    sources and run on a B200 by oracle/run_ref_op.py).  Slot order in the op is a CAS race, so neighbour
    lists are compared as sets with the tie rule of SURVEY.md A.4; coordinates and dims are exact."""
    g = _load(golden_dir, "radius_graph.npz")
    r = _load(golden_dir, "ref_op_golden.npz")
    pts = g["points"]
    radius, K, sort = g[case + "_cfg"]
    cr, _, dims, _, _ = ops.radius_graph_keys(pts, pts, float(radius))
    np.testing.assert_array_equal(cr, r[case + "_coords"])
    np.testing.assert_array_equal(dims, r[case + "_dims"])
    er, eq = ops.radius_graph_build(pts, pts, float(radius), int(K), bool(sort))
    e = r[case + "_edges"]
    assert_neighbor_sets_equal(pts, pts, (er, eq), (e[:, 0], e[:, 1]))


def test_oracle_vs_reference_cuda_op_cross_frame(golden_dir):
    g = _load(golden_dir, "radius_graph.npz")
    r = _load(golden_dir, "ref_op_golden.npz")
    rad = (2.5 ** 2 + 2 ** 2) ** 0.5
    er, eq = ops.radius_graph_build(g["cross_ref"], g["cross_query"], rad, 1, True, qmin=[2, -1, -1, -1],
                                    qmax=[2, 1, 1, 1])
    e = r["cross_edges"]
    assert_neighbor_sets_equal(g["cross_ref"], g["cross_query"], (er, eq), (e[:, 0], e[:, 1]))


def test_ground_oracle_matches_reference_python(golden_dir):
    """oracle/ground_np.py against the reference's own ground_plane_removal (recorded by oracle/gen_golden.py)."""
    from oracle import ground_np
    g = _load(golden_dir, "ground.npz")
    cfg = dict(PILLAR_SIZE=[2, 2], LR=0.01, DECAY_STEPS=[1600], RIGID_WEIGHT=0.5, MAX_NUM_ITERS=10000,
               TRUNCATE_HEIGHT=[0.5], RANSAC=True, SIGMA2=0.0025, JointOpt=True, K=8)
    height, horizon, err, ph, pmz = ground_np.ground_plane_removal(g["points"], cfg)
    assert ph.shape == g["pillar_height"].shape
    assert np.mean(np.abs(pmz - g["pillar_min_z"]) < 2e-2) > 0.98
    assert np.mean(np.abs(ph - g["pillar_height"]) < 3e-2) > 0.97
    assert np.mean((height < 0.5) == (g["height"] < 0.5)) > 0.995


def test_tracking_oracle_matches_reference_python(golden_dir):
    """oracle/tracking_np.track_frame against the golden recorded from the reference's own track_frame."""
    from oracle import tracking_np as trk
    g = np.load(os.path.join(golden_dir, "tracking.npz"))
    pts, comp = g["points"], g["component"]
    n_comp = int(comp.max()) + 1
    diam = trk.component_diameter(pts[:, 1:], comp, n_comp)[comp]
    trace = []
    ex = trk.track_frame(pts, g["sweep"], comp, diam > 12.5, int(g["anchor"]), trk.tracking_cfg(), trace=trace)
    assert len(trace) == 48
    got_c, want_c = set(np.unique(ex["component"]).tolist()), set(np.unique(g["ex_component"]).tolist())
    assert len(got_c & want_c) / len(got_c | want_c) > 0.97
    pair = lambda o, c: set((np.asarray(o, np.int64) * 100000 + np.asarray(c, np.int64)).tolist())
    got_p, want_p = pair(ex["original_indices"], ex["component"]), pair(g["ex_original_indices"], g["ex_component"])
    assert len(got_p & want_p) / len(got_p | want_p) > 0.98
    both = sorted(got_c & want_c)
    dT = np.abs(ex["transforms"][both] - g["transforms"][both])
    assert np.median(dT[..., :3, :3].max((-1, -2))) < 1e-4


def test_eval_oracle_matches_reference_python(golden_dir):
    """evaluate_proposal / extract_traces restatements against the reference-Python goldens."""
    from oracle import tracking_np as trk
    g = np.load(os.path.join(golden_dir, "eval_tracking.npz"))
    keep = g["seg_all"] < 17
    ev = trk.evaluate_proposal(g["points_all"][keep], [g["comp_rad1x25"], g["comp_rad0x75"], g["comp_rad0x25"]],
                               g["box_gt_box_attr"], g["box_gt_box_frame"], g["box_gt_box_track_label"])
    for k in ["point_gt_box_id", "point_gt_trace_id", "point_pred_trace_id", "point_pred_box_id"]:
        np.testing.assert_array_equal(ev[k], g[f"eval_{k}"], err_msg=k)
    np.testing.assert_allclose(ev["gt_box_best_iou"], g["eval_gt_box_best_iou"], atol=1e-6)
    np.testing.assert_allclose(ev["gt_trace_best_iou"], g["eval_gt_trace_best_iou"], atol=1e-6)
    am = g["height_all"] > 0
    ex = {k: g[f"ex_{k}"] for k in ["fxyz", "component", "moving"]}
    best = np.zeros(g["box_gt_box_attr"].shape[0], np.float32)
    full = trk.extract_traces(g["points_all"][am], g["sweep_all"][am], ex, g["box_gt_box_attr"], g["box_gt_box_frame"],
                              best, 0.5)
    for k in ["fxyz", "component", "original_indices", "frame_indices", "moving", "component_hit", "component_size"]:
        np.testing.assert_array_equal(full[k], g[f"full_{k}"], err_msg=k)
    np.testing.assert_allclose(best, g["best_iou_after_tracking"], atol=1e-6)
