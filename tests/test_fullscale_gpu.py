"""GPU: parity at the density of the BASELINE configs (64 beams x 2650 azimuth steps per frame), against fixtures
recorded from the reference's own Python on full-density frames (oracle/gen_golden.py::gen_fullscale): 1.25 m cells hold
thousands of points there and the K = 32 truncation decides which edges exist.  Plus the dense (config 5, ~250 k
rays per frame) sizes through the kernel path, and the reference's clamp-alias quirk of map2key."""
import os

import numpy as np
import pytest
import torch

from helpers import component_centers, transform_errors

pytestmark = pytest.mark.gpu


def test_proposals_at_config_density(golden_dir):
    from pcseqlearning_b200 import ops
    g = np.load(os.path.join(golden_dir, "fullscale.npz"))
    pts = torch.from_numpy(g["points"]).cuda()
    assert pts.shape[0] > 30000
    labels, n_comp = ops.cluster_labels_multi(pts, [1.25, 0.75, 0.25], 32, chunk=10)
    for lab, key in zip(labels, ["component_rad1x25", "component_rad0x75", "component_rad0x25"]):
        np.testing.assert_array_equal(lab.cpu().numpy(), g[key], err_msg=key)
    # and radius by radius through the non-cascaded path
    lab, _ = ops.cluster_labels(pts, 1.25, 32, chunk=10)
    np.testing.assert_array_equal(lab.cpu().numpy(), g["component_rad1x25"])


@pytest.mark.parametrize("lvl", [0, 2])
def test_icp_at_config_density(golden_dir, lvl):
    from pcseqlearning_b200 import ops
    g = np.load(os.path.join(golden_dir, "fullscale.npz"))
    p = f"icp_l{lvl}_"
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    mov, ref, C = g[p + "mov"], g[p + "ref"], int(g[p + "C"])
    moved, T, l1, ratio, info = ops.register_icp(cuda(mov), cuda(g[p + "mov_comp"]), cuda(g[p + "mov_stat"]), cuda(ref),
                                                 cuda(g[p + "ref_stat"]), C, float(g[p + "radius"]), 1,
                                                 angle_regularizer=10, max_iter=80, stopping_delta=0.05)
    comp = g[p + "mov_comp"]
    ctr = component_centers(mov, comp, C)
    ang, dt = transform_errors(T.cpu().numpy(), g[p + "T"], ctr)
    err = np.maximum(ang, dt)
    # 1e-4 against the reference run; at level 2 ONE of the 236 components of that particular reference run sits at
    # 4.6e-4 -- the numpy oracle lands on the same value as the kernel there (checked below), so it is the recorded run
    # (one equal-distance tie resolved the other way), not the kernel, that is off by that much
    assert np.sort(err)[-2 if lvl == 2 else -1] < 1e-4 and err.max() < 1e-3, (np.sort(err)[-3:], info.tolist())
    np.testing.assert_allclose(ratio.cpu().numpy(), g[p + "ratio"], rtol=0, atol=1e-6)
    dl1 = np.sort(np.abs(l1.cpu().numpy() - g[p + "l1"]))
    assert dl1[-2 if lvl == 2 else -1] < 1e-4 and dl1[-1] < 5e-3, dl1[-3:]  # the same outlier component
    from oracle import registration_np as reg
    want = reg.register_to_next_frame(mov, comp, g[p + "mov_stat"], ref, g[p + "ref_stat"], C, float(g[p + "radius"]), 10,
                                      80, 0.05)
    ang, dt = transform_errors(T.cpu().numpy(), want[1], ctr)
    assert ang.max() < 1e-4 and dt.max() < 1e-4, (float(ang.max()), float(dt.max()))
    np.testing.assert_allclose(l1.cpu().numpy(), want[2], rtol=0, atol=1e-4)
    assert int(info[1].item()) == want[4]


def test_dense_config_runs_on_the_kernel_path():
    """Config 5 density (64 x 3900 rays per frame): every limit-guarded solver takes its kernel (an exceeded limit would
    raise PcsError -- there is no eager fallback) and the event log shows the launches."""
    from pcseqlearning_b200 import ops
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    from pcseqlearning_b200.synthetic import generate_sequence
    dev = torch.device("cuda", 0)
    batch = generate_sequence(500, num_frames=4, num_beams=64, num_azimuth=3900, device=dev)
    assert batch["point_bxyz"].shape[0] > 4 * 150000
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_dense_out")
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
        p.SAVE = False
    cfg.SAVE_DIR = None
    model = SimpleReg(cfg, {}, None).to(dev)
    model.train()
    ops.enable_event_log(True)
    try:
        model(batch)
        log = ops.event_log()
        for name in ("ground_ransac", "plane_prune", "l1_heightfield", "hash_build", "radius_search"):
            assert len(log.get(name, [])) >= 1, f"{name} did not run"
    finally:
        ops.enable_event_log(False)
    seq = model.forward_dict["sequences"][0]
    assert len(seq["tracking_results"]) == 3  # one anchor (frame 0) x three component keys
    seq["tracking_batch"].check()


def test_map2key_clamp_alias():
    """Reference quirk (torch_hash_kernel.cu:31-47): a neighbour cell whose coordinate leaves [0, dims_i] is CLAMPED,
    i.e. aliased onto the border cell, so border cells are visited more than once and their points counted again.
    With query offsets reaching beyond `dims` the drop-in op must reproduce exactly that."""
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import torch_hash_cuda as ours
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 3000
    pts = torch.rand(n, 4, generator=g, device="cuda") * torch.tensor([0.0, 3.0, 3.0, 1.0], device="cuda")
    r = 0.5
    vs = torch.tensor([1 - 1e-3, r, r, r], device="cuda")
    lo = pts.min(0)[0]  # NO margin cells: the +-3 offsets below leave the grid on every side
    dims = torch.round((pts.max(0)[0] - lo) / vs).long() + 1
    coord = torch.round((pts - lo) / vs).long()
    H = 2 * n
    keys = torch.full((H,), -1, dtype=torch.int64, device="cuda")
    values = torch.empty(H, 4, device="cuda")
    rev = torch.zeros(H, dtype=torch.int64, device="cuda")
    ours.hash_insert_gpu(keys, values, rev, dims, coord, pts)
    qmin = torch.tensor([0, -3, -3, -3], dtype=torch.int32, device="cuda")
    e = ours.radius_graph_gpu(keys, values, rev, dims, coord, pts, qmin, -qmin, torch.full((n,), r, device="cuda"), -1, True)
    k2, v2, r2 = oracle.new_table(H, 4)
    p = pts.cpu().numpy()
    oracle.hash_insert(k2, v2, r2, dims.cpu().numpy(), coord.cpu().numpy(), p)
    # the oracle's degree pass restates the reference's count kernel, duplicates from aliased cells included
    deg_want = np.zeros(n, np.int32)
    lib = oracle.lib()
    import ctypes
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    qk, qv = c(coord.cpu().numpy(), np.int64), c(p, np.float32)
    lib.oracle_radius_graph_count(k2.ctypes.data_as(ctypes.c_void_p), v2.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int64(H), c(dims.cpu().numpy(), np.int64).ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int(4), qk.ctypes.data_as(ctypes.c_void_p), qv.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int64(n), c([0, -3, -3, -3], np.int32).ctypes.data_as(ctypes.c_void_p),
                                  c([0, 3, 3, 3], np.int32).ctypes.data_as(ctypes.c_void_p),
                                  c(np.full(n, r), np.float32).ctypes.data_as(ctypes.c_void_p), ctypes.c_int(-1),
                                  deg_want.ctypes.data_as(ctypes.c_void_p))
    deg_got = np.bincount(e[:, 1].cpu().numpy(), minlength=n)
    np.testing.assert_array_equal(deg_got, deg_want)
    # aliasing really happened: some (ref, query) rows repeat
    rows = e.cpu().numpy()
    assert np.unique(rows, axis=0).shape[0] < rows.shape[0]
