"""The `pcdet` overlay (pcdet_overlay/): the reference's import paths resolve to this implementation, and
`pcdet.models.build_network` builds SimpleReg from the reference's UNMODIFIED yaml
(tools/cfgs/waymo_models/PCsequence/registration/cluster_tracking_TLS_multiradius_every8.yaml) when /root/reference is
present; a key-for-key rebuilt copy of that yaml's MODEL block (config.cluster_tracking_cfg) is used otherwise."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OVERLAY = os.path.join(ROOT, "pcdet_overlay")
REF_YAML = "/root/reference/tools/cfgs/waymo_models/PCsequence/registration/cluster_tracking_TLS_multiradius_every8.yaml"


@pytest.fixture()
def overlay():
    saved = {k: v for k, v in sys.modules.items() if k == "pcdet" or k.startswith("pcdet.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, OVERLAY)
    try:
        yield importlib.import_module("pcdet")
    finally:
        sys.path.remove(OVERLAY)
        for k in [k for k in sys.modules if k == "pcdet" or k.startswith("pcdet.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_import_paths(overlay):
    from pcdet.models import build_network, load_data_to_gpu, model_fn_decorator  # noqa: F401
    from pcdet.models.model_utils import graph_utils
    from pcdet.models.model_utils.grid_sampling import GridSampling3D  # noqa: F401
    from pcdet.models.registration import __all__ as reg_all
    from pcdet.models.registration.preprocessors import __all__ as pre_all
    from pcdet.ops.torch_hash import ChamferDistance, RadiusGraph, torch_hash_cuda  # noqa: F401
    from pcdet.utils import common_utils
    assert set(reg_all) == {"SimpleReg"}
    assert set(pre_all) == {"GroundPlaneRemover", "ClusterProposal", "ClusterTracking"}
    assert set(graph_utils.GRAPHS) >= {"RadiusGraph"}
    for name in ("hash_insert_gpu", "radius_graph_gpu", "correspondence", "points_in_radius_gpu"):
        assert callable(getattr(torch_hash_cuda, name))  # torch_hash_api.cpp:9-15
    assert callable(common_utils.filter_dict)


def test_build_network_from_reference_yaml(overlay):
    from pcdet.config import cfg, cfg_from_yaml_file
    from pcdet.models import build_network
    if os.path.exists(REF_YAML):
        cfg_from_yaml_file(REF_YAML, cfg)
        model_cfg = cfg.MODEL
    else:  # GPU box: /root/reference is absent
        from pcseqlearning_b200.config import cluster_tracking_cfg
        cfg.MODEL = cluster_tracking_cfg()
        model_cfg = cfg.MODEL
    assert model_cfg.NAME == "SimpleReg"
    model = build_network(model_cfg=model_cfg, cfg=cfg, dataset=None)
    names = [type(m).__name__ for m in model.preprocessors]
    assert names == ["GroundPlaneRemover", "ClusterProposal", "ClusterTracking"]
    trk = model.preprocessors[2]
    assert trk.track_interval == 8 and trk.min_move_frame == 6 and list(trk.radius_list) == [2.5, 1.25, 1.0]
    assert [float(getattr(model.preprocessors[1], f"graph_{k}").radius) for k in model.preprocessors[1].component_keys] == \
        [1.25, 0.75, 0.25]
    # the optimizer of tools/train.py builds its groups from LEAF modules with parameters
    leaf_params = [p for m in model.modules() if not list(m.children()) for p in m.parameters(recurse=False)]
    assert leaf_params and all(p.requires_grad for p in leaf_params)
    assert hasattr(model, "update_global_step") and hasattr(model, "update_ema")
