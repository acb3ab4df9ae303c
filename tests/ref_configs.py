"""The hot-path parameter sets of the reference's shipped configuration files, restated (the reference tree is not
available on the GPU box, and its files are not copied into this repository): every value below is the one
/root/reference/config/<name> holds for that key (SURVEY.md App. C lists them).  tests/test_host_formats.py checks, where
the reference is mounted, that the shim reads the REAL file and this restatement into the same pcdb_params."""
import json


def _config(features, radius, lrf_radius, leaf, bandwidth, distance, normal_radius, normals_method, bbox, avg_rot,
            single_object_max_type):
    return {"ObjectConfig": {
        "Children": {
            "Clustering": {"Type": "None"},
            "Codebook": {
                "Children": {"ActivationStrategy": {"Parameters": {"K": 1, "DistanceRatioThreshold": 0.8,
                                                                   "UseDistanceRatio": False}, "Type": "KNN"}},
                "Parameters": {"UseClassWeight": False, "UseCodewordWeight": False, "UseMatchingWeight": False,
                               "UsePartialShot": False, "UseVoteWeight": False}},
            "FeatureWeighting": {"Parameters": {}, "Type": "Uniform"},
            "Features": {"Parameters": {"Radius": radius, "ReferenceFrameRadius": lrf_radius,
                                        "ReferenceFrameType": "SHOT"}, "Type": features},
            "GlobalFeatures": {"Parameters": {}, "Type": "Dummy"},
            "Keypoints": {"Parameters": {"LeafSize": leaf}, "Type": "VoxelGrid"},
            "Voting": {"Parameters": {
                "AverageRotation": avg_rot, "Bandwidth": bandwidth, "BestK": -1, "BinOrBandwidthFactor": 1.0,
                "BinOrBandwidthType": "Config", "Kernel": "Gaussian", "MaxFilterType": "None", "MaxIter": 1000,
                "MaximaSuppression": "Average", "MinThreshold": 0.0, "MinVotesThreshold": 1,
                "SingleObjectMaxType": single_object_max_type, "SingleObjectMode": True,
                "Threshold": 0.001000000047497451, "UseGlobalFeatures": False}, "Type": "MeanShift"}},
        "Parameters": {"BoundingBoxType": bbox, "ConsistentNormalsMethod": normals_method, "DistanceType": distance,
                       "FLANNExactMatch": False, "FLANNNumKDTrees": 4, "InstanceLabelsPrimary": True,
                       "NormalRadius": normal_radius, "NumThreads": 0, "SingleObjectMode": False,
                       "UseSvmTraining": False, "UseSmoothing": False, "UseStatisticalOutlierRemoval": False,
                       "UseRadiusOutlierRemoval": False, "UseVoxelFiltering": False,
                       "DistanceThresholdDetection": 0.05}}}


# config/qs_input_config.ism: SHOT, chi^2, model units ~ millimetres (BASELINE.json config 1)
QS_INPUT_CONFIG = _config("SHOT", 60, 50, 50, 50, "ChiSquared", 10, 2, "MVBB", True, "None")
# config/default_config_kinect.ism: CSHOT, chi^2, metres (BASELINE.json config 4)
DEFAULT_CONFIG_KINECT = _config("CSHOT", 0.05, 0.05, 0.02, 0.045, "ChiSquared", 0.005, 0, "AABB", True, "Default")


def write(cfg, path):
    with open(path, "w") as f:
        json.dump(cfg, f, indent=3)
    return path
