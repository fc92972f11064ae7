"""Comparison helpers shared by the CPU (oracle) and GPU (parity) tests.

Canonical orders follow SURVEY.md Appendix A.4: a query's neighbour list is compared as a set, and
where the K-truncation cuts through a group of equal distances any subset of that group is accepted.
"""
import numpy as np


def d2_f32(ref, query, e_ref, e_query):
    """fp32 squared distance with the reference's FMA accumulation order (torch_hash_kernel.cu:364-368)."""
    a = ref[e_ref].astype(np.float32)
    b = query[e_query].astype(np.float32)
    d2 = np.zeros(a.shape[0], np.float32)
    for i in range(a.shape[1]):
        di = (a[:, i] - b[:, i]).astype(np.float32)
        # fused multiply-add: exact product + sum in fp64, one rounding to fp32
        d2 = (di.astype(np.float64) * di.astype(np.float64) + d2.astype(np.float64)).astype(np.float32)
    return d2


def split_by_query(e_ref, e_query, num_queries):
    order = np.argsort(e_query, kind="stable")
    e_ref, e_query = e_ref[order], e_query[order]
    deg = np.bincount(e_query, minlength=num_queries)
    start = np.cumsum(deg) - deg
    return e_ref, deg, start, order


def assert_neighbor_sets_equal(ref, query, got, want, sort_by_dist=True):
    """got / want = (e_ref, e_query).  Rows grouped by ascending query; per-query sets equal up to ties."""
    M = query.shape[0]
    g_ref, g_deg, g_start, g_ord = split_by_query(np.asarray(got[0]), np.asarray(got[1]), M)
    w_ref, w_deg, w_start, w_ord = split_by_query(np.asarray(want[0]), np.asarray(want[1]), M)
    assert np.array_equal(np.asarray(got[1]), np.sort(np.asarray(got[1]))), "rows must be grouped by ascending query"
    np.testing.assert_array_equal(g_deg, w_deg, err_msg="per-query degree differs")
    gq = np.repeat(np.arange(M), g_deg)
    gd = d2_f32(ref, query, g_ref, gq)
    wd = d2_f32(ref, query, w_ref, gq)
    if sort_by_dist:
        # distances of the j-th neighbour must agree exactly (both lists ascending)
        srt = lambda d: d[np.lexsort((d, gq))]
        np.testing.assert_array_equal(srt(gd), srt(wd), err_msg="sorted neighbour distances differ")
        # within a query, got distances are ascending
        same = gq[1:] == gq[:-1]
        assert np.all(gd[1:][same] >= gd[:-1][same]), "neighbour list is not sorted by distance"
    # set equality except where ties at the cut-off distance allow alternatives
    key_g = gq.astype(np.int64) * (ref.shape[0] + 1) + g_ref
    key_w = gq.astype(np.int64) * (ref.shape[0] + 1) + w_ref
    only_g = np.setdiff1d(key_g, key_w)
    only_w = np.setdiff1d(key_w, key_g)
    if only_g.size or only_w.size:
        # every unmatched edge must sit exactly at its query's largest kept distance (a tie at the cut)
        maxd = np.zeros(M, np.float32)
        np.maximum.at(maxd, gq, wd)
        for keys, dist, eref in ((only_g, gd, key_g), (only_w, wd, key_w)):
            idx = np.nonzero(np.isin(eref, keys))[0]
            assert np.all(dist[idx] == maxd[gq[idx]]), "neighbour sets differ beyond cut-off ties"
        assert sort_by_dist, "unsorted lists may only differ through truncation of equal candidates"


def canonical_labels(labels):
    """Relabel so that components are numbered by ascending smallest member index (scipy's numbering)."""
    labels = np.asarray(labels)
    _, first = np.unique(labels, return_index=True)
    order = np.argsort(first)
    remap = np.empty(order.shape[0], np.int64)
    remap[order] = np.arange(order.shape[0])
    _, inv = np.unique(labels, return_inverse=True)
    return remap[inv.reshape(-1)]


def rot_angle(R):
    c = (np.trace(R, axis1=-2, axis2=-1) - 1) / 2
    return np.arccos(np.clip(c, -1, 1))


def transform_errors(T, Tref, centers=None):
    """Relative rotation / translation error of per-component rigid transforms.

    rotation: angle of R Rref^T (rad).  translation: |T c - Tref c| / max(1, |c|) evaluated at the
    component centre c (the origin when centres are not given) -- i.e. the displacement the two
    transforms disagree by, relative to the magnitude of the coordinates they act on.  (The reference
    accumulates its centroids in fp32 in an unspecified order, which alone moves t = mu_r - R mu_m by
    ~1e-4 m at 60 m range; comparing raw t columns would measure that noise.)
    """
    R, Rr = T[..., :3, :3], Tref[..., :3, :3]
    ang = rot_angle(R @ np.swapaxes(Rr, -1, -2))
    if centers is None:
        centers = np.zeros(T.shape[:-2] + (3,))
    a = np.einsum("...ij,...j->...i", R, centers) + T[..., :3, 3]
    b = np.einsum("...ij,...j->...i", Rr, centers) + Tref[..., :3, 3]
    dt = np.linalg.norm(a - b, axis=-1) / np.maximum(1.0, np.linalg.norm(centers, axis=-1))
    return ang, dt


def component_centers(fxyz, comp, C):
    c = np.zeros((C, 3))
    n = np.bincount(comp, minlength=C)
    for k in range(3):
        c[:, k] = np.bincount(comp, weights=fxyz[:, 1 + k].astype(np.float64), minlength=C)
    return c / np.maximum(n, 1)[:, None]
