"""GPU: the whole plugin path (SimpleReg.forward -> GroundPlaneRemover -> ClusterProposal -> ClusterTracking) on a
small synthetic sequence: runs end to end, writes the reference's file layout, and its stages agree with the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipeline_run(tmp_path_factory):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    from pcseqlearning_b200.synthetic import generate_sequence
    out = str(tmp_path_factory.mktemp("out"))
    cfg = cluster_tracking_cfg(out_dir=out)
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
    batch = generate_sequence(11, num_frames=10, num_beams=32, num_azimuth=900, device="cuda")
    model = SimpleReg(cfg, {}, None).cuda()
    model.train()
    ret, tb, disp = model(batch)
    model.first_sequence = model.forward_dict["sequences"][0]
    return cfg, batch, model, ret, out


def test_plugin_contract(pipeline_run):
    cfg, batch, model, ret, out = pipeline_run
    assert ret["loss"].requires_grad and float(ret["loss"]) == 0.0
    ret["loss"].backward()  # the reference's training loop calls backward + optimizer.step on nothing
    leaf_params = [p for m in model.modules() if len(list(m.children())) == 0 for p in m.parameters(recurse=False)]
    assert len(leaf_params) >= 1, "adamW_onecycle needs a leaf module that owns a parameter"
    model.update_global_step()
    model.eval()
    assert model(batch) == ({}, None)


def test_stage_outputs_and_files(pipeline_run):
    cfg, batch, model, ret, out = pipeline_run
    seq = model.first_sequence
    n = seq["point_fxyz"].shape[0]
    for key in ["component_rad1x25", "component_rad0x75", "component_rad0x25"]:
        assert seq[f"point_{key}"].shape[0] == n
    assert seq["full_point_fxyz"].shape[0] > n
    seqname = seq["frame_id"][0][:-4]
    assert os.path.exists(f"{out}/ground_removal/TLS/height/{seqname}/pillar_height.pth")
    tdir = f"{out}/cluster_tracking/TLS_multiradius_every8/{seqname}"
    assert os.path.exists(f"{tdir}/all.pth")
    for frame_id in (0, 8):
        for key in ["component_rad1x25", "component_rad0x75", "component_rad0x25"]:
            ex = torch.load(f"{tdir}/{frame_id:03d}_{key}.pth", weights_only=False)
            for k in ["fxyz", "component", "segmentation_label", "original_indices", "frame_indices", "moving",
                      "transforms"]:
                assert k in ex, (frame_id, key, k)
            T = ex["transforms"]
            assert T.dtype == torch.float64 and T.shape[-2:] == (4, 4)
            R = T[..., :3, :3]
            assert float((R @ R.transpose(-1, -2) - torch.eye(3, dtype=torch.float64, device=R.device)).abs().max()) < 1e-6
    boxes = torch.load(f"{tdir}/all.pth", weights_only=False)
    assert "best_iou" in boxes and float(boxes["best_iou"].max()) > 0.3  # tracked clusters cover GT objects


def test_proposals_match_oracle(pipeline_run):
    from oracle import cpu_ops as oracle
    cfg, batch, model, ret, out = pipeline_run
    seq = model.first_sequence
    pts = seq["point_fxyz"].cpu().numpy()
    for key, r in (("component_rad1x25", 1.25), ("component_rad0x75", 0.75), ("component_rad0x25", 0.25)):
        want, _ = oracle.propose_clusters(pts, r)
        np.testing.assert_array_equal(seq[f"point_{key}"].cpu().numpy(), want)


def test_tracking_follows_moving_vehicle(pipeline_run):
    """Physical sanity: components on moving GT vehicles get transforms whose translation grows with the frame gap."""
    cfg, batch, model, ret, out = pipeline_run
    seq = model.first_sequence
    res = seq["tracking_results"]
    ex = res["000_component_rad0x75"]
    assert ex["fxyz"].shape[0] > 0
    frames = ex["fxyz"][:, 0].round().long()
    assert int(frames.max()) >= 6, "no component was tracked for MIN_MOVE_FRAME frames"


@pytest.mark.gpu
def test_device_prefetcher_roundtrip():
    """DevicePrefetcher / load_data_to_gpu (reference: pcdet/models/__init__.py:44-56): same values, same key rules."""
    import numpy as np
    from pcseqlearning_b200.data_staging import DevicePrefetcher, load_data_to_gpu
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    host = [dict(point_bxyz=torch.from_numpy(rng.standard_normal((1000 + i, 4)).astype(np.float32)).pin_memory(),
                 point_sweep=rng.integers(0, 5, (1000 + i, 1)), frame_id=np.array(["a", "b"]),
                 obj_ids=np.arange(3), batch_size=1) for i in range(4)]
    got = list(DevicePrefetcher(iter(host), dev))
    assert len(got) == 4
    for h, g in zip(host, got):
        assert g["point_bxyz"].is_cuda and torch.equal(g["point_bxyz"].cpu(), h["point_bxyz"])
        assert g["point_sweep"].is_cuda and np.array_equal(g["point_sweep"].cpu().numpy(), h["point_sweep"])
        assert isinstance(g["frame_id"], np.ndarray) and isinstance(g["obj_ids"], np.ndarray)  # skipped keys
        assert g["batch_size"] == 1
        assert not h["point_bxyz"].is_cuda  # the host batch is left untouched
    b = load_data_to_gpu(dict(host[0]), dev)
    assert b["point_bxyz"].is_cuda
