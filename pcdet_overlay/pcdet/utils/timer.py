from pcseqlearning_b200.utils import Timer  # noqa: F401
