"""GPU parity of the voxelization kernels against the golden vectors recorded from the reference's
GridSampling3D / sample_frame (oracle/gen_golden.py) and against the oracle on ragged inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["sub008", "lvl0", "lvl2"])
def test_grid_sampling_vs_reference_python(golden_dir, name):
    from pcseqlearning_b200.grid_sampling import GridSampling3D
    g = _load(golden_dir, "grid_sampling.npz")
    pts = _cuda(g["points"])
    sampler = GridSampling3D(g[name + "_size"].tolist()).cuda()
    sampled, inv = sampler(pts, return_inverse=True)
    np.testing.assert_array_equal(inv.cpu().numpy(), g[name + "_inv"])  # voxel index per point: bit exact
    # fp32 atomic-order noise of the reference's scatter-mean: a few ulp at 75 m
    np.testing.assert_allclose(sampled.cpu().numpy(), g[name + "_sampled"], rtol=0, atol=2e-5)


def test_subsample_pick_and_counts(golden_dir):
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "grid_sampling.npz")
    res = ops.voxelize(_cuda(g["points"]), [0.08, 0.08, 0.08], want_mean=False, want_max=True, want_counts=True)
    np.testing.assert_array_equal(res["maxidx"].cpu().numpy(), g["sub008_pick"])
    np.testing.assert_array_equal(res["counts"].cpu().numpy(), np.bincount(g["sub008_inv"]))
    assert res["num"] == g["sub008_pick"].shape[0]


def test_sample_frame_pieces(golden_dir):
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "grid_sampling.npz")
    fxyz = _cuda(g["sf_in_fxyz"])
    res = ops.voxelize(fxyz, [0.2, 0.2, 0.3], want_mean=True, want_counts=True)
    np.testing.assert_allclose(res["sampled"].cpu().numpy(), g["sf_fxyz"], rtol=0, atol=2e-5)
    med = ops.group_median(_cuda(g["sf_in_comp"]), res["inv"], res["num"], res["counts"])
    np.testing.assert_array_equal(med.cpu().numpy(), g["sf_comp"])


def test_voxelize_vs_oracle_ragged_and_time_ignored():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(11)
    for n in (1, 2, 33, 5000, 200000):
        pts = rng.uniform(-20, 20, (n, 4)).astype(np.float32)
        pts[:, 0] = rng.integers(0, 7, n)
        if n > 100:
            pts[10:20] = pts[0:10]  # exact duplicates
        want_s, want_inv = oracle.grid_sampling(pts, [0.3, 0.3, 0.2])
        res = ops.voxelize(_cuda(pts), [0.3, 0.3, 0.2])
        np.testing.assert_array_equal(res["inv"].cpu().numpy(), want_inv)
        np.testing.assert_allclose(res["sampled"].cpu().numpy(), want_s, rtol=0, atol=2e-5)
    # preprocessor_utils.grid_sample: column 0 zeroed
    z = pts.copy()
    z[:, 0] = 0
    want_s, want_inv = oracle.grid_sampling(z, [0.1, 0.1, 0.03])
    res = ops.voxelize(_cuda(pts), [0.1, 0.1, 0.03], ignore_dim0=True)
    np.testing.assert_array_equal(res["inv"].cpu().numpy(), want_inv)
    np.testing.assert_allclose(res["sampled"].cpu().numpy(), want_s, rtol=0, atol=2e-5)


def test_voxelize_large_properties():
    from pcseqlearning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 3_000_000
    pts = torch.rand(n, 4, generator=g, device="cuda") * torch.tensor([1.0, 150.0, 150.0, 8.0], device="cuda")
    pts[:, 0] = torch.randint(0, 30, (n,), generator=g, device="cuda").float()
    res = ops.voxelize(pts, [0.08, 0.08, 0.08], want_mean=True, want_max=True, want_counts=True)
    inv, V = res["inv"], res["num"]
    assert int(inv.max()) == V - 1 and int(inv.min()) == 0
    assert int(res["counts"].sum()) == n
    keys = res["keys"]
    assert bool((keys[1:] > keys[:-1]).all()), "voxels are not numbered by strictly ascending key"
    # picked representative lies in its own voxel; means lie inside the cell of their members
    pick = res["maxidx"]
    assert torch.equal(inv[pick], torch.arange(V, device="cuda"))
    assert float((res["sampled"][inv] - pts).abs()[:, 1:].max()) <= 0.08 * 1.001
    # idempotence: voxelizing the picked points again gives one point per voxel
    res2 = ops.voxelize(pts[pick], [0.08, 0.08, 0.08], want_mean=False, want_counts=True)
    assert res2["num"] == V and int(res2["counts"].max()) == 1
