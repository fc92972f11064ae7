"""GPU: the new kernels against the REFERENCE'S OWN CUDA op (oracle/_ref/torch_hash_cuda_ref.so, compiled
from the unmodified sources under /root/reference by oracle/build_ref.py) run live on the same device."""
import numpy as np
import pytest
import torch

from helpers import assert_neighbor_sets_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_mod():
    from oracle import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("reference op not prebuilt (oracle/_ref)")
    return mod


@pytest.mark.parametrize("n,r,K", [(20000, 0.4, 32), (50000, 0.8, 32), (30000, 0.5, 1)])
def test_against_reference_cuda_op(ref_mod, n, r, K):
    from oracle.run_ref_op import ref_radius_graph
    from pcseqlearning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    pts = torch.rand(n, 4, generator=g, device="cuda") * torch.tensor([1.0, 30.0, 30.0, 4.0], device="cuda")
    pts[:, 0] = torch.randint(0, 3, (n,), generator=g, device="cuda").float()
    e, cr, dims = ref_radius_graph(ref_mod, pts, pts, r, K, True)
    er, eq = ops.radius_graph(pts, pts, r, K, True)
    grid = ops.CellGrid(pts, ops.radius_voxel_size(r), bounds_sets=[pts, pts])
    coords, _ = grid.voxel_keys(pts)
    assert torch.equal(coords, cr), "voxel coordinates differ from the reference's torch ops"
    assert torch.equal(grid.seg_dims[0], dims)
    p = pts.cpu().numpy()
    assert_neighbor_sets_equal(p, p, (er.cpu().numpy(), eq.cpu().numpy()), (e[:, 0].cpu().numpy(), e[:, 1].cpu().numpy()))
