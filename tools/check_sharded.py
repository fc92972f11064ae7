"""Frame-window sharding against the single-GPU pipeline on the same synthetic sequence.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py [frames]

Every rank runs the sharded pipeline collectively; rank 0 then runs the whole sequence alone and compares:
cluster labels (identical), per-box best IoU, and the tracked (anchor, key, frame, row) -> component assignments."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def build(dev):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_shard_out")
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False
        p.LOG_DIR = None
        p.SAVE = False
    cfg.SAVE_DIR = None
    m = SimpleReg(cfg, {}, None).to(dev)
    m.train()
    return m


def canon(fxyz, *cols):
    """rows sorted by (frame, x, y, z) so that differently ordered point sets can be compared"""
    a = fxyz.double().cpu().numpy()
    o = np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
    return [a[o]] + [c.cpu().numpy()[o] for c in cols]


def track_set(seq, keys):
    """{(anchor, key, frame, x, y, z, component)} of everything the tracker extracted"""
    out = set()
    for name, ex in seq["tracking_results"].items():
        if ex["fxyz"].shape[0] == 0:
            continue
        f = ex["fxyz"].cpu().numpy()
        c = ex["component"].cpu().numpy()
        out.update((name,) + tuple(np.round(r, 4)) + (int(k),) for r, k in zip(f, c))
    return out


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pcseqlearning_b200 import parallel
    from pcseqlearning_b200.synthetic import generate_sequence
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import window_batch
    batch = generate_sequence(2, num_frames=frames, num_beams=32, num_azimuth=1200, device=dev)
    shard = parallel.set_sharding(parallel.FrameSharding(frames))
    model = build(dev)
    model(window_batch(batch, *shard.window))
    seq = model.forward_dict["sequences"][0]
    keys = ["component_rad1x25", "component_rad0x75", "component_rad0x25"]
    # gather the sharded result on rank 0
    lab = torch.cat([seq["point_fxyz"].double()] + [seq[f"point_{k}"].double()[:, None] for k in keys], 1)
    lab_all, _ = shard.all_gather_v(lab)
    full_all, _ = shard.all_gather_v(torch.cat([seq["full_point_fxyz"].double(), seq["full_point_height"].double()[:, None],
                                                seq["full_point_horizon"].double()[:, None]], 1))
    tracks = track_set(seq, keys)
    gathered = [None] * world
    dist.all_gather_object(gathered, tracks)
    best_sh = seq["tracking_boxes"]["best_iou"].clone()
    index = seq["tracking_index"].cpu().numpy()
    dist.barrier()
    parallel.set_sharding(None)
    ok = True
    if rank == 0:
        # the ground height field is an iterative float optimisation whose fp64 atomics make it differ by millimetres
        # from run to run on ONE GPU already (tools/dev_ground_determinism.py); to compare everything downstream
        # exactly, the single-GPU run is given the height field of the sharded run (matched point by point)
        import pcseqlearning_b200.preprocessors.ground_plane_remover as gpr
        real = gpr.ground_plane_removal
        stats = {}

        def forced(point_fxyz, cfg, warmup=None, use_kernels=True):
            h, hor, err, ph, pmz = real(point_fxyz, cfg, warmup=warmup, use_kernels=use_kernels)
            a_ = full_all[:, :4].cpu().numpy()
            b_ = point_fxyz.double().cpu().numpy()
            oa = np.lexsort((a_[:, 3], a_[:, 2], a_[:, 1], a_[:, 0]))
            ob = np.lexsort((b_[:, 3], b_[:, 2], b_[:, 1], b_[:, 0]))
            assert a_.shape == b_.shape and np.array_equal(a_[oa], b_[ob]), "subsampled point sets differ"
            hs = np.empty(b_.shape[0], np.float32)
            hs[ob] = full_all[:, 4].cpu().numpy()[oa].astype(np.float32)
            hz = np.empty(b_.shape[0], bool)
            hz[ob] = full_all[:, 5].cpu().numpy()[oa] > 0.5
            d = np.abs(hs - h.cpu().numpy())
            stats.update(max=float(d.max()), mean=float(d.mean()),
                         mask_agree=float(((hs < 0.5) == (h.cpu().numpy() < 0.5)).mean()))
            return torch.from_numpy(hs).to(h), torch.from_numpy(hz).to(hor), err, ph, pmz

        gpr.ground_plane_removal = forced
        model(batch)
        gpr.ground_plane_removal = real
        one = model.forward_dict["sequences"][0]
        print(f"[check_sharded] ground height sharded vs single: max diff {stats['max']:.4f} m, mean {stats['mean']:.5f} m, "
              f"ground-mask agreement {stats['mask_agree']:.6f}", flush=True)
        a = canon(lab_all[:, :4], *[lab_all[:, 4 + i] for i in range(3)])
        b = canon(one["point_fxyz"], *[one[f"point_{k}"] for k in keys])
        same_pts = a[0].shape == b[0].shape and np.array_equal(a[0], b[0])
        same_lab = same_pts and all(np.array_equal(a[1 + i].astype(np.int64), b[1 + i]) for i in range(3))
        sh = set().union(*gathered)
        sg = track_set(one, keys)
        jac = len(sh & sg) / max(len(sh | sg), 1)
        # the same comparison between two single-GPU runs: the run-to-run noise floor of the tracker
        gpr.ground_plane_removal = forced
        model(batch)
        gpr.ground_plane_removal = real
        sg2 = track_set(model.forward_dict["sequences"][0], keys)
        jac_self = len(sg2 & sg) / max(len(sg2 | sg), 1)
        by = lambda st: {n: {x for x in st if x[0] == n} for n in {x[0] for x in st}}
        bs, bg, bg2 = by(sh), by(sg), by(sg2)
        for n in sorted(set(bs) | set(bg)):
            x, y, z = bs.get(n, set()), bg.get(n, set()), bg2.get(n, set())
            print(f"   {n}: sharded {len(x)} single {len(y)} jaccard {len(x & y) / max(len(x | y), 1):.4f} "
                  f"single-vs-single {len(z & y) / max(len(z | y), 1):.4f}", flush=True)
        print(f"[check_sharded] single-GPU run-to-run Jaccard {jac_self:.4f}", flush=True)
        d_iou = float((best_sh - one["tracking_boxes"]["best_iou"]).abs().max())
        n_inst = len(one["tracking_results"])
        print(f"[check_sharded] world={world} frames={frames} points identical={same_pts} labels identical={same_lab} "
              f"tracked-assignment Jaccard={jac:.4f} ({len(sg)} single / {len(sh)} sharded) max |d best_iou|={d_iou:.4f} "
              f"instances single={n_inst} sharded index rows={index.shape[0]}", flush=True)
        # the tracker's thresholded decisions make two single-GPU runs differ (atomic-order ties); the
        # sharded run has to agree with a single-GPU run as well as a second single-GPU run does
        ok = same_lab and jac > jac_self - 0.03 and index.shape[0] == n_inst and stats['mask_agree'] > 0.999
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
