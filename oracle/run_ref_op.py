"""Run the REFERENCE's own torch_hash CUDA op (oracle/_ref, built from /root/reference's sources by
oracle/build_ref.py) on the B200 box and record its outputs as golden vectors.

TEST INFRASTRUCTURE ONLY.   gpurun -- python -m oracle.run_ref_op
Writes gpurun_out/ref_op_golden.npz; copy it to tests/golden/ref_op_golden.npz and commit it.  The C
oracle (oracle.c) is then pinned against these vectors on the CPU (tests/test_oracle.py) and the CUDA
path is compared with the op live (tests/test_ref_op_gpu.py).
"""
import os
import sys
import time

import numpy as np
import torch

from . import build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_radius_graph(mod, ref, query, radius, K, sort, qmin=(0, -1, -1, -1), qmax=(0, 1, 1, 1)):
    """graph_utils.py:169-205 verbatim in behaviour, driving the reference op."""
    dev = ref.device
    rq = query.new_zeros(query.shape[0]) + radius
    vs = torch.tensor([1 - 1e-3] + [rq.max().item() for _ in range(3)]).to(dev)
    allp = torch.cat([ref, query], 0)
    lo = allp.min(0)[0] - vs * 2
    hi = allp.max(0)[0] + vs * 2
    cr = torch.round((ref - lo) / vs).long() + 1
    cq = torch.round((query - lo) / vs).long() + 1
    dims = torch.round((hi - lo) / vs).long() + 3
    H = int(ref.shape[0] / 0.5)
    keys = cr.new_zeros(H) - 1
    values = ref.new_empty(H, 4)
    rev = cr.new_zeros(H)
    mod.hash_insert_gpu(keys, values, rev, dims, cr, ref)
    qmin_t = torch.tensor(qmin).int().to(dev)
    qmax_t = torch.tensor(qmax).int().to(dev)
    edges = mod.radius_graph_gpu(keys, values, rev, dims, cq, query, qmin_t, qmax_t, rq, K, sort)
    return edges, cr, dims


def main():
    mod = build_ref.load_ref()
    if mod is None:
        print("reference op not built (oracle/_ref missing)")
        return 1
    g = np.load(os.path.join(ROOT, "tests", "golden", "radius_graph.npz"))
    pts = torch.from_numpy(g["points"]).cuda()
    out = {}
    for name in ["r125", "r075", "r025", "nn05", "unsorted"]:
        radius, K, sort = g[name + "_cfg"]
        e, cr, dims = ref_radius_graph(mod, pts, pts, float(radius), int(K), bool(sort))
        out[name + "_edges"] = e.cpu().numpy()
        out[name + "_coords"] = cr.cpu().numpy()
        out[name + "_dims"] = dims.cpu().numpy()
    ref, query = torch.from_numpy(g["cross_ref"]).cuda(), torch.from_numpy(g["cross_query"]).cuda()
    e, _, _ = ref_radius_graph(mod, ref, query, (2.5 ** 2 + 2 ** 2) ** 0.5, 1, True, (2, -1, -1, -1), (2, 1, 1, 1))
    out["cross_edges"] = e.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_op_golden.npz"), **out)
    print("wrote gpurun_out/ref_op_golden.npz", {k: v.shape for k, v in out.items() if k.endswith("edges")})
    return 0


if __name__ == "__main__":
    sys.exit(main())
