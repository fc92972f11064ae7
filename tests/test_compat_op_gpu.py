"""GPU: the drop-in `torch_hash_cuda` module (the reference op's four names / signatures, torch_hash.h:16-32) against
the live reference op (oracle/_ref) driven by the SAME harness, against the C oracle, and the reference's own 3-point
known-answer case (torch_hash_modules.py:146-151)."""
import numpy as np
import pytest
import torch

from helpers import assert_neighbor_sets_equal

pytestmark = pytest.mark.gpu


def _cloud(n, seed, frames=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    pts = torch.rand(n, 4, generator=g, device="cuda") * torch.tensor([1.0, 30.0, 30.0, 4.0], device="cuda")
    pts[:, 0] = torch.randint(0, frames, (n,), generator=g, device="cuda").float()
    return pts


@pytest.mark.parametrize("n,r,K", [(20000, 0.4, 32), (30000, 0.5, 1), (4000, 0.6, -1), (15000, 0.8, 48)])
def test_radius_graph_gpu_signature(n, r, K):
    from oracle import build_ref, cpu_ops as oracle
    from oracle.run_ref_op import ref_radius_graph
    from pcseqlearning_b200 import torch_hash_cuda as ours
    pts = _cloud(n, n)
    e, cr, dims = ref_radius_graph(ours, pts, pts, r, K, True)
    assert e.dtype == torch.int64 and e.shape[1] == 2
    p = pts.cpu().numpy()
    if K == -1:
        # all neighbours within r (the reference's fill kernel writes nothing for -1): brute force
        want = []
        for f in range(3):
            idx = np.nonzero(p[:, 0] == f)[0]
            d = np.linalg.norm(p[idx, None, 1:].astype(np.float64) - p[None, idx, 1:].astype(np.float64), axis=-1)
            a, b = np.nonzero(d <= r * (1 - 1e-6))
            want.append(np.stack([idx[b], idx[a]], 1))
        want = np.concatenate(want)
        got = set(map(tuple, e.cpu().numpy().tolist()))
        assert set(map(tuple, want.tolist())) <= got
        assert len(got) <= want.shape[0] * 1.001 + 8
        return
    ref_mod = build_ref.load_ref()
    if ref_mod is not None and K <= 32:
        ew, _, _ = ref_radius_graph(ref_mod, pts, pts, r, K, True)
        want = (ew[:, 0].cpu().numpy(), ew[:, 1].cpu().numpy())
    else:
        keys, values, rev = oracle.new_table(int(n / 0.5), 4)
        oracle.hash_insert(keys, values, rev, dims.cpu().numpy(), cr.cpu().numpy(), p)
        ew = oracle.radius_graph(keys, values, rev, dims.cpu().numpy(), cr.cpu().numpy(), p, [0, -1, -1, -1],
                                 [0, 1, 1, 1], np.full(n, r, np.float32), K, True)
        want = (ew[:, 0], ew[:, 1])
    assert_neighbor_sets_equal(p, p, (e[:, 0].cpu().numpy(), e[:, 1].cpu().numpy()), want)


def _tables(pts, r):
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import torch_hash_cuda as ours
    dev = pts.device
    vs = torch.tensor([1 - 1e-3, r, r, r], device=dev)
    lo = pts.min(0)[0] - vs * 2
    hi = pts.max(0)[0] + vs * 2
    dims = torch.round((hi - lo) / vs).long() + 3
    coord = lambda x: torch.round((x - lo) / vs).long() + 1
    H = int(pts.shape[0] / 0.5)
    keys = torch.full((H,), -1, dtype=torch.int64, device=dev)
    values = torch.empty(H, 4, device=dev)
    rev = torch.zeros(H, dtype=torch.int64, device=dev)
    ours.hash_insert_gpu(keys, values, rev, dims, coord(pts), pts)
    k2, v2, r2 = oracle.new_table(H, 4)
    oracle.hash_insert(k2, v2, r2, dims.cpu().numpy(), coord(pts).cpu().numpy(), pts.cpu().numpy())
    return (keys, values, rev), (k2, v2, r2), dims, coord


def test_correspondence_no_radius_test():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import torch_hash_cuda as ours
    ref = _cloud(20000, 3)
    qry = _cloud(5000, 4)
    qry[:, 1:] += 0.3  # some queries land next to empty cells
    (keys, values, rev), (k2, v2, r2), dims, coord = _tables(torch.cat([ref, qry]), 0.5)
    # rebuild with the reference points only (the helper above inserted both sets to get common dims)
    ours.hash_insert_gpu(keys.fill_(-1), values, rev, dims, coord(ref), ref)
    k2, v2, r2 = oracle.new_table(k2.shape[0], 4)
    oracle.hash_insert(k2, v2, r2, dims.cpu().numpy(), coord(ref).cpu().numpy(), ref.cpu().numpy())
    qmin = torch.tensor([0, -1, -1, -1], dtype=torch.int32, device="cuda")
    qmax = torch.tensor([0, 1, 1, 1], dtype=torch.int32, device="cuda")
    got = torch.full((qry.shape[0],), -7, dtype=torch.int64, device="cuda")
    ours.correspondence(keys, values, rev, dims, coord(qry), qry, qmin, qmax, got)
    want = oracle.correspondence(k2, v2, r2, dims.cpu().numpy(), coord(qry).cpu().numpy(), qry.cpu().numpy(),
                                 [0, -1, -1, -1], [0, 1, 1, 1])
    g = got.cpu().numpy()
    assert np.array_equal(g >= 0, want >= 0)
    # same nearest distance (ties may pick another index)
    r, q = ref.cpu().numpy(), qry.cpu().numpy()
    m = g >= 0
    dg = np.linalg.norm(r[g[m]] - q[m], axis=-1)
    dw = np.linalg.norm(r[want[m]] - q[m], axis=-1)
    np.testing.assert_array_equal(dg, dw)
    # correspondences beyond the cell size exist: there is no radius test
    assert (dg > 0.5).any()


def test_points_in_radius_strict():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import torch_hash_cuda as ours
    ref = _cloud(20000, 5)
    qry = ref[::7].clone()
    qry[:, 1] += 0.25  # exactly 0.25 away from their source point in x (fp32 exactness aside)
    (keys, values, rev), (k2, v2, r2), dims, coord = _tables(ref, 0.5)
    qmin = torch.tensor([0, -1, -1, -1], dtype=torch.int32, device="cuda")
    qmax = torch.tensor([0, 1, 1, 1], dtype=torch.int32, device="cuda")
    for radius in (0.25, 0.4):
        visited = torch.zeros(ref.shape[0], dtype=torch.int64, device="cuda")
        ours.points_in_radius_gpu(keys, values, rev, dims, coord(qry), qry, qmin, qmax, radius, visited)
        want = oracle.points_in_radius(k2, v2, r2, dims.cpu().numpy(), coord(qry).cpu().numpy(), qry.cpu().numpy(),
                                       [0, -1, -1, -1], [0, 1, 1, 1], radius, ref.shape[0])
        np.testing.assert_array_equal(visited.cpu().numpy(), want)
        assert 0 < int(visited.sum()) < ref.shape[0]


def test_reference_three_point_known_answer():
    """torch_hash_modules.py:146-151: points (0,0,0), (0,.1,.1), (0,.2,.2), ndim = 2, r = 0.15, K = 1 -> each point
    finds itself: edges (ref, query) = (0,0), (1,1), (2,2)."""
    from pcseqlearning_b200 import torch_hash_cuda as ours
    pts = torch.tensor([[0, 0, 0], [0, 0.1, 0.1], [0, 0.2, 0.2]], dtype=torch.float32, device="cuda")
    r = 0.15
    vs = torch.tensor([1 - 1e-3, r, r], device="cuda")
    lo = pts.min(0)[0] - vs * 2
    dims = torch.round((pts.max(0)[0] + vs * 2 - lo) / vs).long() + 3
    coord = torch.round((pts - lo) / vs).long() + 1
    keys = torch.full((8,), -1, dtype=torch.int64, device="cuda")
    values = torch.empty(8, 3, device="cuda")
    rev = torch.zeros(8, dtype=torch.int64, device="cuda")
    ours.hash_insert_gpu(keys, values, rev, dims, coord, pts)
    qmin = torch.tensor([0, -1, -1], dtype=torch.int32, device="cuda")
    qmax = torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda")
    e = ours.radius_graph_gpu(keys, values, rev, dims, coord, pts, qmin, qmax, torch.full((3,), r, device="cuda"), 1, True)
    assert e.cpu().tolist() == [[0, 0], [1, 1], [2, 2]]


def test_errors_raise_instead_of_exit():
    from pcseqlearning_b200 import _lib, torch_hash_cuda as ours
    cpu = torch.zeros(4, 4)
    with pytest.raises(_lib.PcsError):
        ours.hash_insert_gpu(torch.zeros(8, dtype=torch.int64), cpu, torch.zeros(8, dtype=torch.int64),
                             torch.ones(4, dtype=torch.int64), cpu.long(), cpu)
    k = torch.full((8,), -1, dtype=torch.int64, device="cuda")
    with pytest.raises(_lib.PcsError):  # a table that hash_insert_gpu never filled
        ours.correspondence(k, torch.zeros(8, 4, device="cuda"), k.clone(), torch.ones(4, dtype=torch.int64, device="cuda"),
                            torch.zeros(2, 4, dtype=torch.int64, device="cuda"), torch.zeros(2, 4, device="cuda"),
                            torch.zeros(4, dtype=torch.int32, device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda"),
                            torch.zeros(2, dtype=torch.int64, device="cuda"))
