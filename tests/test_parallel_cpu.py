"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: window planning, halos, anchors and the collective that
turns window-local component ids into the reference's sequence-global numbering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcseqlearning_b200 import parallel


def test_frame_windows_cover_and_align():
    for num_frames in (16, 198, 40, 7, 400):
        for world in (1, 2, 4, 8):
            w = parallel.frame_windows(num_frames, world)
            assert len(w) == world
            covered = []
            for s, e in w:
                assert s % 40 == 0 or s == num_frames
                covered += list(range(s, e))
            assert covered == list(range(num_frames))
            anchors = sum((parallel.anchors_of_window(x) for x in w), [])
            assert anchors == list(range(0, num_frames, 8))
            for x in w:
                hs, he = parallel.halo_window(x, num_frames)
                for a in parallel.anchors_of_window(x):
                    assert hs <= max(0, a - 8) and min(num_frames, a + 9) <= he


def test_sequences_round_robin():
    got = sorted(sum((parallel.sequences_of_rank(64, r, 8) for r in range(8)), []))
    assert got == list(range(64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_frames, workdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_ops as oracle
    data = np.load(os.path.join(workdir, "in.npz"))
    pts_all, labels_ref = data["points"], data["labels"]
    frames = np.rint(pts_all[:, 0]).astype(np.int64)
    win = parallel.frame_windows(num_frames, world)[rank]
    mask = (frames >= win[0]) & (frames < win[1])
    pts = pts_all[mask]
    if pts.shape[0] > 0:
        # the rank clusters only its own window (CPU oracle stands in for the GPU kernels in this host-logic test)
        lab, _ = oracle.propose_clusters(pts, 0.75)
        n_chunks = (win[1] - win[0] + 9) // 10
        f = np.rint(pts[:, 0]).astype(np.int64)
        n_comp = np.array([len(np.unique(lab[(f // 10) == (win[0] // 10 + c)])) for c in range(n_chunks)])
    else:
        lab, n_comp, f = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    glob, counts = parallel.globalize_component_ids(torch.from_numpy(lab), torch.from_numpy(f), torch.from_numpy(n_comp),
                                                    win, num_frames)
    ok = np.array_equal(glob.numpy(), labels_ref[mask])
    t = parallel.max_over_ranks(10.0 + rank, "cpu")
    res = torch.tensor([int(ok), int(t == 10.0 + world - 1), int(counts.sum().item() == labels_ref.max() + 1)])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(workdir, "result.npy"), res.numpy())
    dist.destroy_process_group()


def test_globalize_component_ids_two_ranks(golden_dir, tmp_path):
    """Two ranks cluster disjoint frame windows; after the count exchange their ids equal the single-process ids."""
    from oracle import cpu_ops as oracle
    g = np.load(os.path.join(golden_dir, "proposal.npz"))
    # stretch the 13 golden frames over 2 windows of 40 frames so that both ranks own chunks
    pts = g["points"].copy()
    pts[:, 0] = np.rint(pts[:, 0]) * 5
    num_frames = int(pts[:, 0].max()) + 1
    labels_ref, _ = oracle.propose_clusters(pts, 0.75)
    np.savez(os.path.join(tmp_path, "in.npz"), points=pts, labels=labels_ref)
    port = _free_port()
    ctx = mp.get_context("spawn")  # fork after the OpenMP oracle has run in the parent would hang
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_frames, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert np.load(os.path.join(tmp_path, "result.npy")).tolist() == [1, 1, 1]


def test_data_staging_key_rules_cpu():
    """Host-side key rules of load_data_to_gpu / DevicePrefetcher (reference: pcdet/models/__init__.py:44-56):
    string / object arrays and the listed bookkeeping keys stay on the host, numeric arrays and CPU tensors move."""
    import numpy as np
    import torch
    from pcseqlearning_b200.data_staging import _host_tensor
    assert _host_tensor("frame_id", np.array(["a"])) is None
    assert _host_tensor("obj_ids", np.arange(3)) is None
    assert _host_tensor("metadata", np.zeros(2)) is None
    assert _host_tensor("names", np.array(["x", "y"])) is None  # non-numeric arrays are left alone
    assert _host_tensor("batch_size", 4) is None
    t = _host_tensor("point_bxyz", np.zeros((5, 4), np.float32))
    assert isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.shape == (5, 4)
    assert _host_tensor("image_shape", np.array([[3, 4]], np.int64)).dtype == torch.int32
    src = torch.zeros(3)
    assert _host_tensor("point_feat", src) is src


def test_synthetic_dataset_plugin_contract():
    """Dataset plugin boundary (pcdet/datasets/__init__.py:73-98, dataset.py:194-298): attributes the
    driver reads, item layout, and collate_batch key rules -- checked against the collated batch the generator
    builds directly."""
    import numpy as np
    from pcseqlearning_b200.datasets import SyntheticSequenceDataset
    from pcseqlearning_b200.synthetic import generate_sequence
    cfg = dict(NUM_SEQUENCES=2, NUM_SWEEPS=3, NUM_BEAMS=16, NUM_AZIMUTH=180, DEVICE="cpu")
    ds = SyntheticSequenceDataset(cfg, root_path=None, training=True, logger=None)
    assert len(ds) == 2 and ds.num_sweeps == 3 and ds.runtime_cfg["num_sweeps"] == 3
    assert ds.num_point_features == 3 and ds.max_num_points == 16 * 180 * 3
    ds.data_augmentor.set_epoch(1)
    item = ds[1]
    assert set(item) == {"point_wise", "object_wise", "scene_wise"}
    assert all(isinstance(v, np.ndarray) for v in item["point_wise"].values())
    batch = SyntheticSequenceDataset.collate_batch([item])
    ref = generate_sequence(1, num_frames=3, num_beams=16, num_azimuth=180, device="cpu")
    assert batch["batch_size"] == 1
    for k in ["point_bxyz", "point_sweep", "point_feat", "segmentation_label", "instance_label", "is_foreground"]:
        assert np.array_equal(batch[k], ref[k].numpy()), k
    assert batch["gt_box_attr"].shape[0] == 1 and batch["gt_box_attr"].shape[2] == 7
    assert np.array_equal(batch["gt_box_attr"].reshape(-1, 7), ref["gt_box_attr"].numpy().reshape(-1, 7))
    assert batch["gt_box_cls_label"].dtype == np.int32 and batch["gt_box_cls_label"].shape[2] == 1
    assert np.array_equal(batch["gt_box_cls_label"].reshape(-1), ref["gt_box_cls_label"].numpy().reshape(-1))
    assert list(batch["frame_id"][0]) == list(ref["frame_id"][0])
    assert np.array_equal(np.asarray(batch["obj_ids"][0]), np.asarray(ref["obj_ids"][0]))
    # two samples: the sample index lands in column 0 of point_bxyz, boxes are padded to the larger count
    two = SyntheticSequenceDataset.collate_batch([ds[0], item])
    n0 = ds[0]["point_wise"]["point_xyz"].shape[0]
    assert two["batch_size"] == 2 and (two["point_bxyz"][:n0, 0] == 0).all() and (two["point_bxyz"][n0:, 0] == 1).all()
    assert two["gt_box_attr"].shape[0] == 2


def test_chunk_windows_and_anchor_blocks():
    for num_frames in (16, 198, 40, 7, 400):
        for world in (1, 2, 4, 8):
            w = parallel.chunk_windows(num_frames, world)
            assert sum((list(range(s, e)) for s, e in w), []) == list(range(num_frames))
            assert all(s % 10 == 0 or s == num_frames for s, _ in w)
            blocks = parallel.anchor_blocks(num_frames, world)
            assert sum(blocks, []) == list(range(0, num_frames, 8))
            sizes = [len(b) for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _shard_worker(rank, world, port, num_frames, workdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = parallel.FrameSharding(num_frames)
    data = np.load(os.path.join(workdir, "in.npz"))
    frame_all, val_all, flag_all = data["frame"], data["val"], data["flag"]
    own = (frame_all >= sh.window[0]) & (frame_all < sh.window[1])
    frame = torch.from_numpy(frame_all[own])
    out, fr = sh.exchange_frames(dict(val=torch.from_numpy(val_all[own]), flag=torch.from_numpy(flag_all[own])), frame)
    lo, hi = sh.need[rank]
    want = (frame_all >= lo) & (frame_all < hi)
    # a single process selects the needed frames in ascending frame order keeping the row order inside a frame
    order = np.argsort(frame_all[want], kind="stable")
    ok_halo = (np.array_equal(fr.numpy(), frame_all[want][order]) and np.array_equal(out["val"].numpy(), val_all[want][order])
               and np.array_equal(out["flag"].numpy(), flag_all[want][order]) and out["flag"].dtype == torch.bool)
    # bounds: order-preserving uint32 encodings carried in an int32 tensor, min over the first 4 / max over the last 4
    enc = np.array([[5 + rank, 2 ** 31 + 7 - rank, 1, 9, 100 + rank, 2 ** 32 - 1 - rank, 3, 2 ** 31 + rank]], np.uint32)
    b = torch.from_numpy(enc.view(np.int32).copy())
    sh.reduce_bounds(b)
    got = b.numpy().view(np.uint32)[0].tolist()
    ok_bounds = got == [5, 2 ** 31 + 7 - (world - 1), 1, 9, 100 + world - 1, 2 ** 32 - 1, 3, 2 ** 31 + world - 1]
    cat, sizes = sh.all_gather_v(torch.full((rank + 2, 3), float(rank)))
    ok_gather = sizes == [r + 2 for r in range(world)] and cat.shape[0] == sum(sizes) and \
        all(bool((cat[sum(sizes[:r]):sum(sizes[:r + 1])] == r).all()) for r in range(world))
    mx = sh.all_reduce_max(torch.tensor([float(rank), 3.0 - rank]))
    sm = sh.all_reduce_sum(torch.tensor([1, rank]))
    ok_red = mx.tolist() == [world - 1.0, 3.0] and sm.tolist() == [world, world * (world - 1) // 2]
    res = torch.tensor([int(ok_halo), int(ok_bounds), int(ok_gather), int(ok_red)])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(workdir, "shard_result.npy"), res.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [48, 23])
def test_frame_sharding_collectives_two_ranks(tmp_path, num_frames):
    """FrameSharding on gloo: the halo exchange hands every rank exactly the rows a single process would select for
    [first anchor - 8, last anchor + 8]; bounds / max / sum reductions and the variable-length gather."""
    rng = np.random.default_rng(3)
    n = 4000
    frame = np.sort(rng.integers(0, num_frames, n)).astype(np.int64)
    frame = frame[rng.permutation(n)] if num_frames == 23 else frame  # unsorted rows inside a window too
    np.savez(os.path.join(tmp_path, "in.npz"), frame=frame, val=rng.normal(size=(n, 4)).astype(np.float32),
             flag=rng.random(n) < 0.3)
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, num_frames, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert np.load(os.path.join(tmp_path, "shard_result.npy")).tolist() == [1, 1, 1, 1]
