"""pcdet/models/registration/preprocessors/__init__.py:5-9 of the reference."""
from pcseqlearning_b200.preprocessors import ClusterProposal, ClusterTracking, GroundPlaneRemover

__all__ = dict(GroundPlaneRemover=GroundPlaneRemover, ClusterProposal=ClusterProposal, ClusterTracking=ClusterTracking)
