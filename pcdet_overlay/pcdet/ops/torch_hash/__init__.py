"""pcdet/ops/torch_hash/__init__.py:1-2 of the reference."""
from pcseqlearning_b200.torch_hash import ChamferDistance, RadiusGraph  # noqa: F401
from . import torch_hash_cuda  # noqa: F401
