"""GPU: ClusterTracking.track_frame against the golden recorded from the reference's OWN track_frame (CPU run of the
unmodified cluster_tracking.py through oracle/ref_import.py).  The tracker chains ~48 ICP solves, AdamW velocity
smoothing and thresholded stopping tests per anchor, so it is compared on the extracted point sets (Jaccard) and on
transforms within a tolerance, not bit for bit."""
import os

import numpy as np
import pytest
import torch

from helpers import component_centers, rot_angle, transform_errors

pytestmark = pytest.mark.gpu


def test_track_frame_vs_reference_python(golden_dir):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.preprocessors.cluster_tracking import ClusterTracking, component_diameter
    from pcseqlearning_b200.utils import EasyDict, filter_dict
    g = np.load(os.path.join(golden_dir, "tracking.npz"))
    cfg = [p for p in cluster_tracking_cfg().PREPROCESSORS if p.NAME == "ClusterTracking"][0]
    cfg.VERBOSE = False
    mod = ClusterTracking(cfg, {}).cuda()
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    comp = cuda(g["component"])
    seq_points = EasyDict(fxyz=cuda(g["points"]), frame=cuda(g["sweep"]), gt_box_id=torch.zeros_like(comp) - 1,
                          segmentation_label=cuda(g["seg"]), component=comp)
    diam = component_diameter(seq_points)[seq_points.component]
    seq_points.component_diameter = diam
    seq_points.stationary = diam > 12.5
    seq_points.extracted = torch.zeros_like(seq_points.fxyz[:, 0]).bool()
    anchor = int(g["anchor"])
    frame_mask = (seq_points.fxyz[:, 0] == anchor).reshape(-1)
    frame_points = EasyDict(filter_dict(seq_points, frame_mask))
    frame_points.component = frame_points.component - frame_points.component.min()
    ex = mod.track_frame(seq_points, frame_points, None)

    # kept components
    got_c = set(ex.component.unique().tolist())
    want_c = set(np.unique(g["ex_component"]).tolist())
    jac_c = len(got_c & want_c) / max(len(got_c | want_c), 1)
    # extracted (point, component) pairs
    n = g["points"].shape[0]
    got_p = set((ex.original_indices.cpu().numpy().astype(np.int64) * 100000 + ex.component.cpu().numpy()).tolist())
    want_p = set((g["ex_original_indices"].astype(np.int64) * 100000 + g["ex_component"]).tolist())
    jac_p = len(got_p & want_p) / max(len(got_p | want_p), 1)
    # transforms of the components kept by both, every tracked frame
    T, Tw = ex.transforms.cpu().numpy(), g["transforms"]
    assert T.shape == Tw.shape
    both = sorted(got_c & want_c)
    # displacement the two transforms disagree by AT THE COMPONENT (raw t columns carry the 60 m lever arm)
    anchor_pts = g["points"][np.rint(g["points"][:, 0]) == anchor]
    anchor_comp = g["component"][np.rint(g["points"][:, 0]) == anchor]
    anchor_comp = anchor_comp - anchor_comp.min()
    ctr = component_centers(anchor_pts, anchor_comp, T.shape[0])[both]
    ang, dt = transform_errors(T[both], Tw[both], np.repeat(ctr[:, None, :], T.shape[1], axis=1))
    dt = dt * np.maximum(1.0, np.linalg.norm(ctr, axis=-1))[:, None]  # back to metres
    msg = dict(jac_components=jac_c, jac_points=jac_p, ang_med=float(np.median(ang)), ang_p99=float(np.quantile(ang, 0.99)),
               dt_med=float(np.median(dt)), dt_p99=float(np.quantile(dt, 0.99)), n_both=len(both))
    print(msg)
    # Each ICP stops on a loose loss rule (three consecutive improvements < 0.05), so millimetre-level differences in
    # one solve (the reference's fp32 atomics vs fp64 sums here) move the stopping iteration of later solves; the
    # chain of 48 solves per anchor is compared statistically.  Observed: identical kept components, 99.3 % identical
    # extracted points, median rotation difference 1e-7 rad, median displacement 3-5 mm.
    assert jac_c > 0.95 and jac_p > 0.97, msg
    assert np.median(ang) < 1e-3 and np.median(dt) < 2e-2, msg  # metres at the component
    assert np.quantile(ang, 0.9) < 2e-2 and np.quantile(dt, 0.9) < 0.25, msg


def test_smooth_velo_kernel_vs_torch_adamw():
    """Fused velocity-smoothing kernel against the plain PyTorch AdamW loop it replaces."""
    from pcseqlearning_b200.preprocessors.cluster_tracking import smooth_velo
    g = torch.Generator(device="cuda").manual_seed(4)
    C, F = 150, 17
    for a, b in ((9, 12), (2, 8), (8, 9)):
        base = torch.randn(C, 1, 3, generator=g, device="cuda") * 0.5
        diffs = base + 0.05 * torch.randn(C, F, 3, generator=g, device="cuda")
        velos = diffs + 0.1 * torch.randn(C, F, 3, generator=g, device="cuda")
        v1 = smooth_velo(velos.clone(), diffs, a, b, use_kernels=True)
        v2 = smooth_velo(velos.clone(), diffs, a, b, use_kernels=False)
        d = (v1 - v2).abs()
        assert float(d.max()) < 5e-3, (a, b, float(d.max()))
        # untouched entries only see AdamW's weight decay
        assert float((v1[:, :a] - v2[:, :a]).abs().max()) < 1e-5
