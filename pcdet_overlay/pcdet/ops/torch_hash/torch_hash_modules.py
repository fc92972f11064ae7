from pcseqlearning_b200.torch_hash import ChamferDistance, RadiusGraph  # noqa: F401
