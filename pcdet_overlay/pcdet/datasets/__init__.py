"""Dataset plugin registry (pcdet/datasets/__init__.py:13-16): the synthetic Waymo-shaped sequence dataset."""
from pcseqlearning_b200.datasets import SyntheticSequenceDataset

__all__ = dict(SyntheticSequenceDataset=SyntheticSequenceDataset)
