"""GPU parity tests of the voxel hash + neighbour search + connected components, through the C ABI.

Bar (BASELINE.json north_star): voxel keys, neighbour sets and cluster labels bit-exact up to canonical
relabelling (ties at equal distance may be ordered / cut differently)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_neighbor_sets_equal, canonical_labels, d2_f32

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_reference_known_answer_three_points():
    from pcseqlearning_b200 import ops
    pts = torch.tensor([[0, 0.0, 0.0], [0, 0.1, 0.1], [0, 0.2, 0.2]], dtype=torch.float32).cuda()
    er, eq = ops.radius_graph(pts, pts, 0.15, 1, True)  # torch_hash_modules.py:146-151 (ndim=2)
    assert er.tolist() == [0, 1, 2] and eq.tolist() == [0, 1, 2]


@pytest.mark.parametrize("radius", [1.25, 0.75, 0.25, 0.5])
def test_voxel_keys_bit_exact(golden_dir, radius):
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    pts = _load(golden_dir, "radius_graph.npz")["points"]
    cr, cq, dims, _, _ = oracle.radius_graph_keys(pts, pts, radius)
    keys = np.zeros(pts.shape[0], np.int64)
    oracle.lib().oracle_map2key(oracle._p(np.ascontiguousarray(cr)), oracle._p(dims), 4, pts.shape[0], oracle._p(keys))
    t = _cuda(pts)
    grid = ops.CellGrid(t, ops.radius_voxel_size(radius), bounds_sets=[t, t])
    coords, gkeys = grid.voxel_keys(t)
    np.testing.assert_array_equal(grid.seg_dims.cpu().numpy()[0], dims)
    np.testing.assert_array_equal(coords.cpu().numpy(), cr)
    np.testing.assert_array_equal(gkeys.cpu().numpy(), keys)
    ncell = grid.check()
    assert ncell == np.unique(keys).shape[0]
    # the counting sort is a permutation grouped by cell
    sidx = grid.sorted_idx.cpu().numpy()
    assert np.array_equal(np.sort(sidx), np.arange(pts.shape[0]))
    np.testing.assert_array_equal(grid.sorted_pts.cpu().numpy(), pts[sidx])
    k_sorted = keys[sidx]
    change = np.nonzero(k_sorted[1:] != k_sorted[:-1])[0]
    assert change.shape[0] + 1 == ncell, "points of one cell are not contiguous"


@pytest.mark.parametrize("case", ["r125", "r075", "r025", "nn05", "unsorted"])
def test_radius_graph_vs_reference_python(golden_dir, case):
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "radius_graph.npz")
    pts = g["points"]
    radius, K, sort = g[case + "_cfg"]
    t = _cuda(pts)
    er, eq = ops.radius_graph(t, t, float(radius), int(K), bool(sort))
    if case == "unsorted":
        # first-K-discovered in the reference: only degree and validity are defined
        deg = np.bincount(eq.cpu().numpy(), minlength=pts.shape[0])
        np.testing.assert_array_equal(deg, np.bincount(g[case + "_equery"], minlength=pts.shape[0]))
        d2 = d2_f32(pts, pts, er.cpu().numpy(), eq.cpu().numpy())
        assert np.all(d2 <= np.float32(radius) * np.float32(radius))
    else:
        assert_neighbor_sets_equal(pts, pts, (er.cpu().numpy(), eq.cpu().numpy()),
                                   (g[case + "_eref"], g[case + "_equery"]))


def test_radius_graph_cross_frame(golden_dir):
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "radius_graph.npz")
    ref, query = g["cross_ref"], g["cross_query"]
    r = (2.5 ** 2 + 2 ** 2) ** 0.5
    er, eq = ops.radius_graph(_cuda(ref), _cuda(query), r, 1, True, qmin=(2, -1, -1, -1), qmax=(2, 1, 1, 1))
    assert_neighbor_sets_equal(ref, query, (er.cpu().numpy(), eq.cpu().numpy()), (g["cross_eref"], g["cross_equery"]))


def test_radius_graph_vs_oracle_random_and_ragged():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(7)
    for n, m, r, K in ((1, 1, 0.5, 4), (5, 0, 0.5, 4), (3000, 2000, 0.6, 8), (20000, 20000, 0.35, 32), (4000, 100, 3.0, 1)):
        ref = rng.uniform(0, 12, (n, 4)).astype(np.float32)
        ref[:, 0] = rng.integers(0, 3, n)
        query = rng.uniform(0, 12, (m, 4)).astype(np.float32)
        query[:, 0] = rng.integers(0, 3, m)
        # duplicates and exact ties
        if n > 100:
            ref[50:60] = ref[40:50]
        er, eq = ops.radius_graph(_cuda(ref), _cuda(query), r, K, True)
        if m == 0:
            assert er.numel() == 0
            continue
        wr, wq = oracle.radius_graph_build(ref, query, r, K, True)
        assert_neighbor_sets_equal(ref, query, (er.cpu().numpy(), eq.cpu().numpy()), (wr, wq))


def test_per_query_radius():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(8)
    pts = rng.uniform(0, 8, (5000, 4)).astype(np.float32)
    pts[:, 0] = 0
    rad = rng.uniform(0.2, 0.6, 5000).astype(np.float32)
    er, eq = ops.radius_graph(_cuda(pts), _cuda(pts), _cuda(rad), 16, True)
    wr, wq = oracle.radius_graph_build(pts, pts, rad, 16, True)
    assert_neighbor_sets_equal(pts, pts, (er.cpu().numpy(), eq.cpu().numpy()), (wr, wq))


def test_connected_components_vs_scipy():
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(9)
    for n, e in ((10, 0), (50, 30), (5000, 3000), (200000, 150000), (200000, 600000)):
        e0 = rng.integers(0, n, e)
        e1 = rng.integers(0, n, e)
        nc, lab = oracle.connected_components(e0, e1, n)
        gn, glab = ops.connected_components((_cuda(e0), _cuda(e1)), n)
        assert int(gn.sum().item()) == nc
        np.testing.assert_array_equal(glab.cpu().numpy(), lab)  # identical numbering, not just isomorphic


def test_cluster_labels_vs_reference_python(golden_dir):
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "proposal.npz")
    t = _cuda(g["points"])
    for key, r in (("component_rad1x25", 1.25), ("component_rad0x75", 0.75), ("component_rad0x25", 0.25)):
        labels, n_comp = ops.cluster_labels(t, r, 32, chunk=10)
        got = labels.cpu().numpy()
        # chunk-wise numbering with running offset must match the reference exactly
        np.testing.assert_array_equal(got, g[key])
        assert int(n_comp.sum().item()) == int(g[key].max()) + 1
        # and the edge-list route (graph -> union-find over edges) agrees with the fused kernel
        fr = g["points"][:, 0].astype(np.int64)
        m = fr < 10
        tt = _cuda(g["points"][m])
        er, eq = ops.radius_graph(tt, tt, r, 32, True)
        _, lab2 = ops.connected_components((er, eq), int(m.sum()))
        np.testing.assert_array_equal(lab2.cpu().numpy(), g[key][m])


def test_large_scale_properties():
    """Size-independent properties at a size the oracle would not finish quickly."""
    from pcseqlearning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 2_000_000
    pts = torch.rand(n, 4, generator=g, device="cuda") * torch.tensor([1.0, 120.0, 120.0, 6.0], device="cuda")
    pts[:, 0] = torch.randint(0, 20, (n,), generator=g, device="cuda").float()
    r, K = 0.6, 32
    er, eq, d2 = ops.radius_graph(pts, pts, r, K, True, return_dists=True)
    assert bool((eq[1:] >= eq[:-1]).all()), "rows not grouped by ascending query"
    assert bool((d2 <= np.float32(r) * np.float32(r)).all())
    same = eq[1:] == eq[:-1]
    assert bool((d2[1:][same] >= d2[:-1][same]).all()), "lists not sorted by distance"
    dd = (pts[er] - pts[eq]).square().sum(-1)
    assert float((dd - d2).abs().max()) < 1e-5
    assert bool((pts[er, 0] == pts[eq, 0]).all()), "intra-frame graph crossed frames"
    deg = torch.bincount(eq, minlength=n)
    assert int(deg.min()) >= 1 and int(deg.max()) <= K  # self match always present
    # idempotence / determinism
    er2, eq2 = ops.radius_graph(pts, pts, r, K, True)
    assert torch.equal(er, er2) and torch.equal(eq, eq2)
    # labels: fused == edge route, endpoints of every edge share a label
    labels, _ = ops.cluster_labels(pts, r, K, chunk=10, num_frames=20)
    assert bool((labels[er] == labels[eq]).all())


def test_multi_radius_fused_equals_separate_searches(golden_dir):
    """The multi-radius search (one fine pass + one coarse pass over the sparse remainder) gives exactly the labels of
    three independent searches, which equal the reference's."""
    from pcseqlearning_b200 import ops
    g = _load(golden_dir, "proposal.npz")
    t = _cuda(g["points"])
    labels, n_comp = ops.cluster_labels_multi(t, [1.25, 0.75, 0.25], 32, chunk=10)
    for lab, key in zip(labels, ("component_rad1x25", "component_rad0x75", "component_rad0x25")):
        np.testing.assert_array_equal(lab.cpu().numpy(), g[key])
    # dense random cloud: most fine lists are full, so the coarse forests are fed by the fine pass
    gen = torch.Generator(device="cuda").manual_seed(17)
    n = 400_000
    pts = torch.rand(n, 4, generator=gen, device="cuda") * torch.tensor([1.0, 40.0, 40.0, 3.0], device="cuda")
    pts[:, 0] = torch.randint(0, 12, (n,), generator=gen, device="cuda").float()
    labels, _ = ops.cluster_labels_multi(pts, [1.25, 0.75, 0.25], 32, chunk=10, num_frames=12)
    for lab, r in zip(labels, (1.25, 0.75, 0.25)):
        want, _ = ops.cluster_labels(pts, r, 32, chunk=10, num_frames=12)
        assert torch.equal(lab, want), r


@pytest.mark.parametrize("K", [2, 8, 32])
@pytest.mark.parametrize("sorted_cells", [False, True])
def test_thread_search_equals_warp_search(K, sorted_cells):
    """The thread-per-query proposal kernel (pcs_self_search_uf) against the warp-per-query kernel on the same grid:
    identical counts min(#within r, K) and identical components, on a cloud dense enough that most lists overflow
    (replacement path of the shared-memory K-list) and through the skip_full_cnt cascade."""
    from pcseqlearning_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5 + K)
    n = 300_000
    pts = torch.rand(n, 4, generator=gen, device="cuda") * torch.tensor([1.0, 30.0, 30.0, 2.0], device="cuda")
    pts[:, 0] = torch.randint(0, 25, (n,), generator=gen, device="cuda").float()
    pts[: n // 10, 1:] *= 0.1  # a very dense corner: hundreds of candidates per query
    n_seg = 3
    out = {}
    for mode in (1, 0):
        ops.SEARCH_THREADS = mode
        try:
            cnt = None
            parents = []
            for r in (0.2, 0.5):
                grid = ops.CellGrid(pts, ops.radius_voxel_size(r), seg_div=10, n_seg=n_seg, sorted_cells=sorted_cells)
                parent = ops.uf_new(n, pts.device) if not parents else parents[-1].clone()
                _, cnt, _ = grid.search(None, K, r, uf_parent=parent, want_lists=False, skip_full_cnt=cnt, cnt_out=cnt)
                grid.check()
                parents.append(parent)
            seg_of = ops.point_segments(pts, 10, n_seg)
            out[mode] = (cnt.clone(), [ops.uf_labels(p_, seg_of, n_seg)[1] for p_ in parents])
        finally:
            ops.SEARCH_THREADS = 0
    assert torch.equal(out[1][0], out[0][0])
    assert int((out[1][0] >= K).sum()) > n // 20  # the overflow path was exercised
    for a, b in zip(out[1][1], out[0][1]):
        assert torch.equal(a, b)


def test_compact_table_overflow_falls_back():
    """A point set with one point per cell overflows the compact table; the grid must be rebuilt, not corrupted."""
    from oracle import cpu_ops as oracle
    from pcseqlearning_b200 import ops
    rng = np.random.default_rng(31)
    n = 6000
    pts = np.zeros((n, 4), np.float32)
    pts[:, 1:] = rng.uniform(0, 400, (n, 3))  # ~1 point per 0.3 m cell
    t = _cuda(pts)
    grid = ops.compact_grid(t, ops.radius_voxel_size(0.3))
    assert grid.check() > n // 4 and grid.H >= 2 * n
    labels, _ = ops.cluster_labels(t, 0.3, 32, chunk=10, num_frames=1)
    want, _ = oracle.propose_clusters(pts, 0.3)
    np.testing.assert_array_equal(labels.cpu().numpy(), want)


def test_torch_hash_module_api_known_answer_and_chamfer():
    """Module-style API of the op (torch_hash_modules.py): the reference's ndim=2 known-answer case and Chamfer."""
    from pcseqlearning_b200.torch_hash import ChamferDistance, RadiusGraph
    rg = RadiusGraph(ndim=2).cuda()
    pts = torch.tensor([[0, 0.0, 0.0], [0, 0.1, 0.1], [0, 0.2, 0.2]], dtype=torch.float32).cuda()
    eq, er = rg(pts, pts, 0.15, 1, sort_by_dist=True)
    assert eq.tolist() == [0, 1, 2] and er.tolist() == [0, 1, 2]
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.rand(5000, 4, generator=g, device="cuda") * 10
    a[:, 0] = 0
    b = a + 0.01
    b[:, 0] = 0
    cd = ChamferDistance(ndim=3)(a, b, 0.5)
    assert abs(float(cd) - 2 * 3 * 0.01 ** 2) < 1e-5
