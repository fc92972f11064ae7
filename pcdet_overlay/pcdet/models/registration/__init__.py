"""pcdet/models/registration/__init__.py:4-14 of the reference."""
from pcseqlearning_b200.simple_reg import SimpleReg

__all__ = dict(SimpleReg=SimpleReg)


def build_registration(model_cfg, runtime_cfg, dataset):
    return __all__[model_cfg.NAME](model_cfg=model_cfg, runtime_cfg=runtime_cfg, dataset=dataset)
