"""Configuration of the cluster-tracking path.

``cluster_tracking_cfg()`` rebuilds, key for key, the MODEL block of the reference's
tools/cfgs/waymo_models/PCsequence/registration/cluster_tracking_TLS_multiradius_every8.yaml (lines 1-78);
``load_yaml`` reads that yaml itself (or any override) into an EasyDict so the unmodified file can be used."""
from .utils import EasyDict


def cluster_tracking_cfg(out_dir="../output/waymo_sequence_registration"):
    keys = ["component_rad1x25", "component_rad0x75", "component_rad0x25"]
    return EasyDict(dict(
        NAME="SimpleReg",
        SCALE=1,
        SAVE_DIR=f"{out_dir}/cluster_tracking/TLS_multiradius_every8/",
        SUBSAMPLE=True,
        PREPROCESSORS=[
            dict(NAME="GroundPlaneRemover", DIR=f"{out_dir}/ground_removal/TLS/height/",
                 LOG_DIR=f"{out_dir}/ground_removal/TLS/log/", PILLAR_SIZE=[2, 2], LR=0.01, DECAY_STEPS=[1600],
                 RIGID_WEIGHT=0.5, MAX_NUM_ITERS=10000, TRUNCATE_HEIGHT=[0.5], RANSAC=True, VISUALIZE=True,
                 SIGMA2=0.0025, JointOpt=True, K=8),
            dict(NAME="ClusterProposal",
                 GRAPH=dict(TYPE="RadiusGraph", RADIUS=[1.25, 0.75, 0.25], MAX_NUM_NEIGHBORS=32, SORT_BY_DIST=True,
                            RELATIVE_KEY="fxyz"),
                 COMPONENT_KEYS=list(keys), VISUALIZE=True, DIR=f"{out_dir}/cluster_proposal/TLS_multiradius/"),
            dict(NAME="ClusterTracking", ANGLE_REGULARIZER=10, COMPONENT_KEYS=list(keys),
                 REGISTRATION=dict(
                     GRAPH=dict(TYPE="RadiusGraph", RADIUS=[2.5, 1.25, 1.0], MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True,
                                RELATIVE_KEY="fxyz"),
                     VOXEL_SIZE=[[0.4, 0.4, 0.6], [0.2, 0.2, 0.3], [0.1, 0.1, 0.15]],
                     STOPPING_DELTA=[0.05, 0.05, 0.05]),
                 NN_GRAPH=dict(TYPE="RadiusGraph", RADIUS=0.5, MAX_NUM_NEIGHBORS=1, SORT_BY_DIST=True,
                               RELATIVE_KEY="fxyz"),
                 DIR=f"{out_dir}/cluster_tracking/TLS_multiradius_every8/",
                 TRACKING_PARAMS=dict(REGISTRATION_ERROR_COEFFICIENT=0.13, TRACK_INTERVAL=8, ANGLE_THRESHOLD=45,
                                      MIN_MOVE_FRAME=6)),
        ],
    ))


def load_yaml(path):
    import yaml
    with open(path) as f:
        return EasyDict(yaml.safe_load(f))
