"""pcdet/models/model_utils/graph_utils.py of the reference: build_graph, GRAPHS, RadiusGraph, connected_components."""
from pcseqlearning_b200.graph_utils import *  # noqa: F401,F403
from pcseqlearning_b200.graph_utils import GRAPHS, RadiusGraph, build_graph, connected_components  # noqa: F401
