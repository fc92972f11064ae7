"""Developer: run-to-run differences of the batched tracker on identical inputs (atomic-order noise vs. races)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.check_sharded import build
from pcseqlearning_b200.synthetic import generate_sequence
from pcseqlearning_b200.tracker import TrackBatch
dev = torch.device("cuda", 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
batch = generate_sequence(2, num_frames=frames, num_beams=32, num_azimuth=1200, device=dev)
m = build(dev)
m(batch)
seq = m.forward_dict["sequences"][0]
trk = [x for x in m.preprocessors if type(x).__name__ == "ClusterTracking"][0]
keys = ["component_rad1x25", "component_rad0x75", "component_rad0x25"]
comps = [seq[f"point_{k}"] for k in keys]
runs = []
for i in range(3):
    tb = TrackBatch(seq["point_fxyz"], seq["point_sweep"], comps, trk.model_cfg, num_frames=frames)
    for t in tb.steps()[:nsteps]:
        tb.step(t)
    torch.cuda.synchronize()
    snap = {k: tb.t[k].clone() for k in ["transforms", "velos", "centers", "g_stopped", "g_moving", "l1_err", "ratio", "T", "mp"]}
    snap["iters"] = [p.tolist()[8] for p in tb.prof]
    snap["inst_iters"] = tb.sc["iters"].clone()
    n_mv = int(tb.sampler.t["ctr"][1].item())
    mv = tb.mv[:n_mv, 1:].double()
    snap["mv_sorted"] = mv[torch.argsort(mv[:, 0] * 1e6 + mv[:, 1] * 1e3 + mv[:, 2])]
    runs.append(snap)
for i in range(1, len(runs)):
    a, b = runs[0], runs[i]
    msg = [f"run {i} vs 0:"]
    for k in ["transforms", "T", "mp", "velos", "centers", "l1_err", "ratio"]:
        msg.append(f"{k} max|d| {float((a[k].double() - b[k].double()).abs().max()):.3e}")
    for k in ["g_stopped", "g_moving"]:
        msg.append(f"{k} flips {int((a[k] != b[k]).sum())}/{a[k].numel()}")
    msg.append(f"iters {a['iters']} vs {b['iters']}; last-level per-instance iters differ in {int((a['inst_iters'] != b['inst_iters']).sum())} instances")
    same_n = a["mv_sorted"].shape == b["mv_sorted"].shape
    msg.append(f"last-level voxels: {a['mv_sorted'].shape[0]} vs {b['mv_sorted'].shape[0]}" +
               (f" max|d| {float((a['mv_sorted'] - b['mv_sorted']).abs().max()):.3e}" if same_n else ""))
    print(" ".join(msg))
    dT = (a["T"].double() - b["T"].double()).abs().amax(1)
    deg = tb.g_deg
    print("   comps with |dT| > 1e-9: %d, > 1e-6: %d, > 1e-3: %d, > 1e-1: %d of %d; point counts of the > 1e-3 ones: %s" % (
        int((dT > 1e-9).sum()), int((dT > 1e-6).sum()), int((dT > 1e-3).sum()), int((dT > 1e-1).sum()), dT.numel(),
        deg[dT > 1e-3].tolist()[:40]))
