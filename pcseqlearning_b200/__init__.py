"""pcseqlearning_b200 -- B200-native object-cluster extraction and tracking path of PCSeqLearning.

Only the hot path of ``cluster_tracking_TLS_multiradius_every8.yaml`` lives here:
hand-written sm_100a CUDA kernels behind a C-ABI library (``csrc/`` -> ``libpcseq_b200.so``,
declared in ``include/pcseq_b200.h``) and the Python host mirror of the reference's plugin interface.
"""
__version__ = "0.1.0"
