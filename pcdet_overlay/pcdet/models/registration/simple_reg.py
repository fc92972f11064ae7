from pcseqlearning_b200.simple_reg import SimpleReg  # noqa: F401
