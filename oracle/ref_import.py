"""Import the reference's own Python (from /root/reference) on this GPU-less build container.

TEST INFRASTRUCTURE ONLY; used by oracle/gen_golden.py (and by tests that are skipped when
/root/reference is absent, i.e. on the GPU box).

What is real and what is substituted:
  * REAL, imported unmodified from /root/reference: graph_utils.RadiusGraph / connected_components,
    grid_sampling.GridSampling3D, registration_utils.*, cluster_tracking.{sample_frame, smooth_velo,
    component_diameter, ...}, cluster_proposal.ClusterProposal, common_utils.filter_dict.
  * SUBSTITUTED: the CUDA op ``pcdet.ops.torch_hash.torch_hash_cuda`` -> oracle.c through ctypes on CPU
    tensors (itself pinned against the real op on the B200 box, tests/golden/ref_op_*.npz); the absent
    third-party packages -> oracle/shims; ``Tensor.cuda()`` -> identity; package ``__init__`` files
    (which import every detector) -> empty namespace stubs.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REF, "pcdet"))


def _stub_pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    m.__package__ = name
    sys.modules[name] = m
    return m


def _torch_hash_cpu_module():
    from . import cpu_ops as ops
    m = types.ModuleType("pcdet.ops.torch_hash.torch_hash_cuda")

    def hash_insert_gpu(keys, values, reverse_indices, dims, insert_keys, insert_values):
        ops.hash_insert(keys.numpy(), values.numpy(), reverse_indices.numpy(), dims.numpy(),
                        insert_keys.contiguous().numpy(), insert_values.contiguous().numpy())

    def radius_graph_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius,
                         max_num_neighbors, sort_by_dist):
        e = ops.radius_graph(keys.numpy(), values.numpy(), reverse_indices.numpy(), dims.numpy(),
                             query_keys.contiguous().numpy(), query_values.contiguous().numpy(), qmin.numpy(),
                             qmax.numpy(), radius.numpy(), max_num_neighbors, sort_by_dist)
        return torch.from_numpy(e)

    def correspondence(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, corres):
        c = ops.correspondence(keys.numpy(), values.numpy(), reverse_indices.numpy(), dims.numpy(),
                               query_keys.contiguous().numpy(), query_values.contiguous().numpy(), qmin.numpy(),
                               qmax.numpy())
        corres[:c.shape[0]] = torch.from_numpy(c)

    def points_in_radius_gpu(keys, values, reverse_indices, dims, query_keys, query_values, qmin, qmax, radius,
                             visited):
        v = ops.points_in_radius(keys.numpy(), values.numpy(), reverse_indices.numpy(), dims.numpy(),
                                 query_keys.contiguous().numpy(), query_values.contiguous().numpy(), qmin.numpy(),
                                 qmax.numpy(), radius, visited.shape[0])
        visited |= torch.from_numpy(v)

    m.hash_insert_gpu = hash_insert_gpu
    m.radius_graph_gpu = radius_graph_gpu
    m.correspondence = correspondence
    m.points_in_radius_gpu = points_in_radius_gpu
    return m


_installed = False


def install():
    """Populate sys.modules so reference modules can be imported file by file."""
    global _installed
    if _installed:
        return
    assert available(), "/root/reference is not present"
    shims = os.path.join(HERE, "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    # tensors stay on the CPU
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
    p = os.path.join(REF, "pcdet")
    _stub_pkg("pcdet", p)
    _stub_pkg("pcdet.utils", os.path.join(p, "utils"))
    _stub_pkg("pcdet.models", os.path.join(p, "models"))
    _stub_pkg("pcdet.models.model_utils", os.path.join(p, "models/model_utils"))
    _stub_pkg("pcdet.models.registration", os.path.join(p, "models/registration"))
    _stub_pkg("pcdet.models.registration.preprocessors", os.path.join(p, "models/registration/preprocessors"))
    _stub_pkg("pcdet.ops")
    _stub_pkg("pcdet.ops.torch_hash")
    sys.modules["pcdet.ops.torch_hash.torch_hash_cuda"] = _torch_hash_cpu_module()
    _stub_pkg("pcdet.ops.pointops")
    _stub_pkg("pcdet.ops.pointops.functions")
    po = types.ModuleType("pcdet.ops.pointops.functions.pointops")
    po.knnquery = None
    sys.modules[po.__name__] = po
    vox = _stub_pkg("pcdet.ops.voxel")
    vox.VoxelAggregation = None
    vm = types.ModuleType("pcdet.ops.voxel.voxel_modules")
    vm.VoxelAggregation = None
    sys.modules[vm.__name__] = vm
    roi = _stub_pkg("pcdet.ops.roiaware_pool3d")
    ru = types.ModuleType("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")

    def points_in_boxes_cpu(points, boxes):
        from . import cpu_ops as ops
        return torch.from_numpy(ops.points_in_boxes(points.numpy(), boxes.numpy()))

    ru.points_in_boxes_cpu = points_in_boxes_cpu
    sys.modules[ru.__name__] = ru
    roi.roiaware_pool3d_utils = ru
    vis = _stub_pkg("pcdet.models.visualizers")
    vis.GeometryVisualizer = None
    _installed = True


def load(modname):
    """e.g. load('pcdet.models.registration.preprocessors.registration_utils')"""
    install()
    mod = importlib.import_module(modname)
    parent, _, leaf = modname.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod


def edict(**kw):
    install()
    from easydict import EasyDict
    return EasyDict(dict(**kw))
