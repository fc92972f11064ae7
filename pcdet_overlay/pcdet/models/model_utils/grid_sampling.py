from pcseqlearning_b200.grid_sampling import GridSampling3D  # noqa: F401
