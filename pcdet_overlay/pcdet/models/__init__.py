"""pcdet/models/__init__.py:16-71 of the reference: build_network dispatches on MODEL.NAME among the capitalised names
of `pcdet.models.registration`; load_data_to_gpu / model_fn_decorator are what tools/train_utils calls every step."""
import re
from collections import namedtuple

import numpy as np
import torch

from .registration import build_registration


def build_network(model_cfg, cfg, dataset):
    import pcdet.models.registration as registration
    builder_dict = {name: build_registration for name in dir(registration) if name[:1].isupper()}
    if model_cfg.NAME not in builder_dict:
        raise KeyError(f"MODEL.NAME={model_cfg.NAME}: only the registration models of the cluster-tracking path are "
                       f"provided by this overlay ({sorted(builder_dict)})")
    model = builder_dict[model_cfg.NAME](model_cfg=model_cfg, runtime_cfg=cfg, dataset=dataset)
    freezed = cfg.get("MODEL", {}).get("FREEZED_MODULES", None) if hasattr(cfg, "get") else None
    if freezed is not None:
        for name, param in model.named_parameters():
            if any(re.match(rx, name) is not None for rx in freezed):
                param.requires_grad = False
    return model


def load_data_to_gpu(batch_dict):
    from pcseqlearning_b200.data_staging import load_data_to_gpu as _load
    return _load(batch_dict)


def model_fn_decorator():
    ModelReturn = namedtuple("ModelReturn", ["loss", "tb_dict", "disp_dict"])

    def model_func(model, batch_dict):
        load_data_to_gpu(batch_dict)
        ret_dict, tb_dict, disp_dict = model(batch_dict)
        loss = ret_dict["loss"].mean()
        (model if hasattr(model, "update_global_step") else model.module).update_global_step()
        return ModelReturn(loss, tb_dict, disp_dict)

    return model_func
