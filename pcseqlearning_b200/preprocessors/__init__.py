from .cluster_proposal import ClusterProposal
from .ground_plane_remover import GroundPlaneRemover

__all__ = dict(
    GroundPlaneRemover=GroundPlaneRemover,
    ClusterProposal=ClusterProposal,
)
