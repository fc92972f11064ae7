"""numpy restatement of the reference's per-cluster registration (ICP) and frame down-sampling.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows pcdet/models/registration/preprocessors/registration_utils.py:12-206 and
cluster_tracking.py:39-51 of /root/reference.  Pinned by tests/golden/registration_*.npz, which were
produced by running the reference's own Python (imported from /root/reference with CPU shims for its
absent third-party deps) through oracle/gen_golden.py.
"""
import numpy as np

from . import cpu_ops as ops


def segment_mean(data, index, n):
    """efficient_robust_mean (:25-34): argsort + segment_coo(mean); empty groups -> 0.  fp32 in, fp32 out."""
    return ops.scatter(np.asarray(data), index, n, "mean")


def scatter_mean_nonempty(data, index, n):
    """robust_mean (:12-23): sum / count where count > 0.5, sum (=0) elsewhere."""
    return ops.scatter(np.asarray(data), index, n, "mean")


def truncated_mean(data, index, n, trunc=0.3):
    """truncated_robust_mean (:44-58): mean of values clamped to +-trunc around the plain group mean."""
    m = ops.scatter(data, index, n, "mean")
    clamped = np.clip(data, m[index] - trunc, m[index] + trunc)
    return ops.scatter(clamped, index, n, "mean")


def upper_median(values, index, n):
    """robust_median (:60-81): element of rank deg//2 of every group's sorted values; empty groups -> -1e10."""
    values = np.asarray(values, np.int64)
    index = np.asarray(index, np.int64)
    order = np.lexsort((values, index))
    sv, si = values[order], index[order]
    deg = np.bincount(si, minlength=n)
    start = np.cumsum(deg) - deg
    out = np.full(n, -10 ** 10, np.int64)
    has = deg > 0
    out[has] = sv[(start + deg // 2)[has]]
    return out


def sample_frame(fxyz, stationary, component, frame, voxel_size):
    """cluster_tracking.py:39-51: grid mean of fxyz, majority `stationary`, upper-median component / frame."""
    _, inv = ops.grid_sampling(fxyz, voxel_size)
    n = int(inv.max()) + 1
    out = {}
    out["fxyz"] = ops.scatter(np.asarray(fxyz, np.float32), inv, n, "mean")
    out["stationary"] = ops.scatter(np.asarray(stationary, np.float32), inv, n, "mean") > 0.5
    out["component"] = upper_median(component, inv, n)
    out["frame"] = upper_median(np.asarray(frame).reshape(-1), inv, n)
    out["inv"] = inv
    return out


def kabsch_rotation(A):
    """R = V diag(1,1,det(V U^T)) U^T for A = U S V^T (registration_utils.py:167-174), fp64, batched."""
    U, _, Vt = np.linalg.svd(A)
    V = np.swapaxes(Vt, 1, 2)
    Ut = np.swapaxes(U, 1, 2)
    sign = np.ones(A.shape[:1] + (3,))
    sign[:, 2] = np.linalg.det(V @ Ut)
    return (V * sign[:, None, :]) @ Ut


def register_to_next_frame(mov_fxyz, mov_comp, mov_stat, ref_fxyz, ref_stat, num_components, radius,
                           angle_regularizer=10.0, max_iter=20, stopping_delta=5e-2, trace=None):
    """registration_utils.py:83-206.

    mov_fxyz f32[vm,4] (frame,x,y,z) is updated the way the reference updates `moving.fxyz` (fp64 product
    stored back to fp32 every iteration, :179).  Returns (moved_fxyz, T f64[C,4,4], l1_err f64[C],
    edge_ratio f32[C], n_iter).
    """
    mov_fxyz = np.array(mov_fxyz, np.float32, copy=True)
    ref_fxyz = np.asarray(ref_fxyz, np.float32)
    mov_comp = np.asarray(mov_comp, np.int64)
    C = int(num_components)
    df = int(np.int64(ref_fxyz[0, 0]) - np.int64(mov_fxyz[0, 0]))  # :110 (.long() truncation of the frame id)
    r_eff = (radius ** 2 + df ** 2) ** 0.5  # :111, python double
    comp_deg = np.bincount(mov_comp, minlength=C).astype(np.int64)  # :113-114
    T = np.tile(np.eye(4), (C, 1, 1))  # :116
    ns_m = ~np.asarray(mov_stat, bool)
    ns_r = ~np.asarray(ref_stat, bool)
    m = mov_fxyz[ns_m].copy()
    mc = mov_comp[ns_m]
    r = ref_fxyz[ns_r]
    last = 1e10
    countdown = 3
    l1 = np.zeros(C)
    n_iter = 0
    for it in range(max_iter):
        n_iter += 1
        f_ref, f_mov = ops.radius_graph_build(r, m, r_eff, 1, True, qmin=[df, -1, -1, -1], qmax=[df, 1, 1, 1])
        b_mov, b_ref = ops.radius_graph_build(m, r, r_eff, 1, True, qmin=[-df, -1, -1, -1], qmax=[-df, 1, 1, 1])
        e_m = np.concatenate([f_mov, b_mov])  # :144
        e_r = np.concatenate([f_ref, b_ref])  # :145
        e_c = mc[e_m]  # :146
        mu_m = segment_mean(m[e_m, 1:], e_c, C).astype(np.float64)  # :150 (fp32 mean, then double)
        mu_r = segment_mean(r[e_r, 1:], e_c, C).astype(np.float64)  # :151
        P = m[e_m, 1:].astype(np.float64) - mu_m[e_c]  # :152 (fp32 - fp64 -> fp64)
        Q = r[e_r, 1:].astype(np.float64) - mu_r[e_c]  # :153
        dist = np.linalg.norm(P - Q, axis=-1)  # :154
        l1 = truncated_mean(dist, e_c, C)  # :156
        loss = float(np.square(dist).sum())  # :161
        H = P[:, :, None] * Q[:, None, :]  # :163
        cov = scatter_mean_nonempty(H.reshape(-1, 9), e_c, C).reshape(C, 3, 3)  # :164
        R = kabsch_rotation(cov + T[:, :3, :3] * angle_regularizer)  # :165-174
        Ti = np.zeros((C, 4, 4))
        Ti[:, :3, :3] = R
        Ti[:, :3, 3] = mu_r - np.einsum("cj,cij->ci", mu_m, R)  # :177  (mu_m @ R^T)
        Ti[:, 3, 3] = 1.0
        T = Ti @ T  # :178
        moved = np.einsum("nj,nij->ni", m[:, 1:].astype(np.float64), R[mc]) + Ti[mc, :3, 3]
        m[:, 1:] = moved.astype(np.float32)  # :179 stored back into the fp32 array
        if trace is not None:
            trace.append(dict(loss=loss, n_edges=int(e_m.shape[0])))
        if last - loss < stopping_delta:  # :180
            countdown -= 1
        else:
            countdown = 3
        if countdown <= 0:
            break
        last = loss
    # edge ratio: moving (non-stationary) voxels with a ref voxel (stationary included) within r (:189-199)
    f_ref, f_mov = ops.radius_graph_build(ref_fxyz, m, r_eff, 1, True, qmin=[df, -1, -1, -1], qmax=[df, 1, 1, 1])
    cnt = np.bincount(mc[f_mov], minlength=C).astype(np.int64)
    ratio = (cnt / (comp_deg.astype(np.float32) + np.float32(1e-6))).astype(np.float32)
    mov_fxyz[ns_m, 1:] = m[:, 1:]  # :205
    return mov_fxyz, T, l1, ratio, n_iter
