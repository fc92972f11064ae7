"""`pcdet` overlay: the import paths the reference's driver (tools/train.py) and its yaml resolve, bound to the
B200-native implementation in `pcseqlearning_b200`.

Put this directory FIRST on PYTHONPATH (before, or instead of, the reference's own `pcdet` package) and the unmodified
`tools/train.py --cfg_file tools/cfgs/waymo_models/PCsequence/registration/cluster_tracking_TLS_multiradius_every8.yaml`
builds `SimpleReg` with the GroundPlaneRemover / ClusterProposal / ClusterTracking preprocessors of this repository.
Only the cluster-tracking hot path is provided; detectors, backbones and the other ops of the
reference are out of scope and absent here.  See INTEGRATION.md."""
__version__ = "0.2.0+b200"
