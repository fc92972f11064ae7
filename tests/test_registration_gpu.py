"""GPU: the persistent ICP kernel against the golden vectors recorded from the reference's own
register_to_next_frame (oracle/gen_golden.py) and against the numpy oracle on other pairs.

Tolerance (BASELINE.json north_star): registration transforms within 1e-4 relative rotation / translation error."""
import os

import numpy as np
import pytest
import torch

from helpers import component_centers, transform_errors

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", ["fwd", "bwd"])
@pytest.mark.parametrize("lvl", [0, 1, 2])
def test_register_icp_vs_reference_python(golden_dir, case, lvl):
    from pcseqlearning_b200 import ops
    g = np.load(os.path.join(golden_dir, "registration.npz"))
    p = f"{case}_l{lvl}_"
    C = int(g[p + "C"])
    mov, ref = g[p + "mov_fxyz"], g[p + "ref_fxyz"]
    df = int(ref[0, 0]) - int(mov[0, 0])
    moved, T, l1, ratio, info = ops.register_icp(_cuda(mov), _cuda(g[p + "mov_comp"]), _cuda(g[p + "mov_stat"]),
                                                 _cuda(ref), _cuda(g[p + "ref_stat"]), C, float(g[p + "radius"]), df,
                                                 angle_regularizer=10, max_iter=80, stopping_delta=0.05)
    T = T.cpu().numpy()
    ang, dt = transform_errors(T, g[p + "T"], component_centers(mov, g[p + "mov_comp"], C))
    assert ang.max() < 1e-4 and dt.max() < 1e-4, (ang.max(), dt.max(), info.tolist())
    np.testing.assert_allclose(l1.cpu().numpy(), g[p + "l1"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(ratio.cpu().numpy(), g[p + "ratio"], rtol=0, atol=1e-6)
    # reference centroids are fp32 sums in unspecified order: ~3e-4 m noise on 100 m-wide components
    np.testing.assert_allclose(moved.cpu().numpy(), g[p + "moved"], rtol=0, atol=5e-4)
    # every T is a rigid transform
    R = T[:, :3, :3]
    assert np.abs(R @ np.swapaxes(R, 1, 2) - np.eye(3)).max() < 1e-9
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-9


def test_register_icp_vs_oracle_with_stationary_ref():
    from oracle import cpu_ops as oracle, registration_np as reg
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(21)
    # three rigid blobs, each moved by its own small transform in the next frame; one ref blob is stationary
    pts, comp = [], []
    for c, ctr in enumerate([(5, 5, 1), (20, -8, 1.5), (-15, 12, 0.8)]):
        b = rng.normal(0, 0.6, (400, 3)) + np.array(ctr)
        pts.append(b)
        comp.append(np.full(400, c))
    P = np.concatenate(pts).astype(np.float32)
    comp = np.concatenate(comp)
    mov = np.concatenate([np.zeros((P.shape[0], 1), np.float32), P], 1)
    ang = 0.05
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    Q = P.copy()
    for c, ctr in enumerate([(5, 5, 1), (20, -8, 1.5), (-15, 12, 0.8)]):
        m = comp == c
        Q[m] = ((P[m] - ctr) @ Rz.T + ctr + np.array([0.3, -0.2, 0.0]) * (c + 1)).astype(np.float32)
    ref = np.concatenate([np.full((Q.shape[0], 1), 2, np.float32), Q], 1).astype(np.float32)
    ref_stat = np.zeros(ref.shape[0], bool)
    ref_stat[comp == 2] = True
    mov_stat = np.zeros(mov.shape[0], bool)
    want = reg.register_to_next_frame(mov, comp, mov_stat, ref, ref_stat, 3, 2.5, 10, 80, 0.05)
    got = ops.register_icp(_cuda(mov), _cuda(comp), _cuda(mov_stat), _cuda(ref), _cuda(ref_stat), 3, 2.5, 2,
                           angle_regularizer=10, max_iter=80, stopping_delta=0.05)
    a, d = transform_errors(got[1].cpu().numpy(), want[1], component_centers(mov, comp, 3))
    assert a.max() < 1e-4 and d.max() < 1e-4, (a, d, got[4].tolist(), want[4])
    assert int(got[4][1].item()) == want[4]
    np.testing.assert_allclose(got[3].cpu().numpy(), want[3], atol=1e-6)
    np.testing.assert_allclose(got[2].cpu().numpy(), want[2], atol=1e-4)
