"""The native module of the reference op (torch_hash_api.cpp:9-15), same four names and signatures."""
from pcseqlearning_b200.torch_hash_cuda import (correspondence, hash_insert_gpu, points_in_radius_gpu,  # noqa: F401
                                                radius_graph_gpu)
