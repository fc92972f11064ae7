import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tools.check_sharded import build
from pcseqlearning_b200.synthetic import generate_sequence
dev = torch.device("cuda", 0)
batch = generate_sequence(2, num_frames=48, num_beams=32, num_azimuth=1200, device=dev)
m = build(dev)
hs = []
for i in range(3):
    m(batch)
    s = m.forward_dict["sequences"][0]
    hs.append((s["full_point_height"].clone(), s["point_fxyz"].shape[0]))
for i in range(1, 3):
    d = (hs[i][0] - hs[0][0]).abs()
    print("run", i, "vs 0: max", float(d.max()), "mean", float(d.mean()), ">1e-3:", int((d > 1e-3).sum()), "N_g", hs[i][1], hs[0][1])
