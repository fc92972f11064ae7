"""The helpers of pcdet/utils/common_utils.py the path uses (filter_dict :67-78)."""
from pcseqlearning_b200.utils import filter_dict  # noqa: F401
