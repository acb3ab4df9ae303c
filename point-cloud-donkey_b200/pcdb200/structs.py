"""ctypes mirrors of the structs in include/pcdb200.h (shared by the product binding and the test oracle binding)."""
import ctypes as C

import numpy as np

FEATURE_SHOT, FEATURE_CSHOT = 0, 1
DIST_EUCLIDEAN, DIST_CHISQUARED = 0, 1
KERNEL_GAUSSIAN, KERNEL_UNIFORM = 0, 1
SUPPRESS_AVERAGE, SUPPRESS_SUPPRESS = 0, 1
MAXFILTER_NONE, MAXFILTER_SIMPLE, MAXFILTER_MERGE = 0, 1, 2
KNN_AUTO, KNN_SCAN, KNN_GEMM = 0, 1, 2
RADIUS_CONFIG, RADIUS_FIRST_DIM, RADIUS_SECOND_DIM = 0, 1, 2
SOMAX_DEFAULT, SOMAX_BANDWIDTH, SOMAX_VOTING_SPACE, SOMAX_MODEL_RADIUS = 0, 1, 2, 3
RANSAC_FIXED, RANSAC_OBJECT_RADIUS, RANSAC_BBOX_MEDIAN = 0, 1, 2
SHOT_DIM, CSHOT_DIM, MAX_K = 352, 1344, 16

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_CAPACITY, E_STATE, E_UNSUPPORTED, E_COMM = 0, -1, -2, -3, -4, -5, -6, -7


class Params(C.Structure):
    """pcdb_params — field names follow the reference's JSON keys."""

    _fields_ = [
        ("feature_type", C.c_int32),
        ("feature_radius", C.c_double),
        ("lrf_radius", C.c_double),
        ("leaf_size", C.c_float),
        ("distance_type", C.c_int32),
        ("knn_k", C.c_int32),
        ("use_distance_ratio", C.c_int32),
        ("distance_ratio_threshold", C.c_float),
        ("use_class_weight", C.c_int32),
        ("use_vote_weight", C.c_int32),
        ("use_matching_weight", C.c_int32),
        ("use_codeword_weight", C.c_int32),
        ("filter_abs_is_int", C.c_int32),
        ("bandwidth", C.c_float),
        ("ms_threshold", C.c_float),
        ("ms_max_iter", C.c_int32),
        ("ms_kernel", C.c_int32),
        ("maxima_suppression", C.c_int32),
        ("min_threshold", C.c_float),
        ("min_votes_threshold", C.c_int32),
        ("best_k", C.c_int32),
        ("average_rotation", C.c_int32),
        ("single_object_mode", C.c_int32),
        ("normal_radius", C.c_float),
        ("consistent_normals_method", C.c_int32),
        ("max_filter_type", C.c_int32),
        ("radius_type", C.c_int32),
        ("radius_factor", C.c_float),
        ("single_object_max_type", C.c_int32),
        ("ransac_vote_filtering", C.c_int32),
        ("ransac_inlier_threshold", C.c_float),
        ("ransac_threshold_type", C.c_int32),
        ("ransac_refine_model", C.c_int32),
    ]

    @property
    def dim(self):
        return CSHOT_DIM if self.feature_type == FEATURE_CSHOT else SHOT_DIM

    def copy(self):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p


def default_params(**kw):
    """Code defaults of the reference, overridable by keyword."""
    p = Params()
    p.feature_type = FEATURE_SHOT
    p.feature_radius = 0.1
    p.lrf_radius = float(np.float32(0.2))
    p.leaf_size = 0.1
    p.distance_type = DIST_EUCLIDEAN
    p.knn_k = 1
    p.use_distance_ratio = 0
    p.distance_ratio_threshold = 0.95
    p.bandwidth = 0.2
    p.ms_threshold = 1e-3
    p.ms_max_iter = 1000
    p.ms_kernel = KERNEL_GAUSSIAN
    p.maxima_suppression = SUPPRESS_AVERAGE
    p.min_threshold = 0.0
    p.min_votes_threshold = 1
    p.best_k = -1
    p.average_rotation = 0
    p.single_object_mode = 0
    p.normal_radius = 0.05
    p.consistent_normals_method = 2
    p.max_filter_type = MAXFILTER_NONE
    p.radius_type = RADIUS_CONFIG
    p.radius_factor = 1.0
    p.single_object_max_type = SOMAX_DEFAULT
    p.ransac_vote_filtering = 0
    p.ransac_inlier_threshold = 0.1
    p.ransac_threshold_type = RANSAC_FIXED
    p.ransac_refine_model = 0
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        if k == "lrf_radius":  # Features.ReferenceFrameRadius is a float member promoted to double
            v = float(np.float32(v))
        setattr(p, k, v)
    return p


VOTE_DTYPE = np.dtype(
    [
        ("position", np.float32, 3),
        ("weight", np.float32),
        ("keypoint", np.float32, 3),
        ("class_id", np.uint32),
        ("keypoint_training", np.float32, 3),
        ("instance_id", np.uint32),
        ("bbox_quat", np.float32, 4),
        ("bbox_size", np.float32, 3),
        ("codeword_id", np.int32),
    ],
    align=True,
)
assert VOTE_DTYPE.itemsize == 80

MAXIMUM_DTYPE = np.dtype(
    [
        ("position", np.float32, 3),
        ("weight", np.float32),
        ("class_id", np.uint32),
        ("instance_id", np.uint32),
        ("instance_weight", np.float32),
        ("raw_weight", np.float32),
        ("bbox_quat", np.float32, 4),
        ("bbox_size", np.float32, 3),
        ("n_votes", np.int32),
        ("vote_begin", np.int64),
    ],
    align=True,
)
assert MAXIMUM_DTYPE.itemsize == 72


class Stats(C.Structure):
    _fields_ = [
        (n, C.c_int64)
        for n in (
            "n_points n_keypoints n_features n_neighbours_lrf n_neighbours_shot n_votes n_maxima "
            "knn_queries knn_candidates knn_fallback_queries kernel_launches"
        ).split()
    ] + [(n, C.c_double) for n in "features_ms knn_ms knn_gemm_ms votes_ms maxima_ms".split()] + [
        ("comm_bytes", C.c_int64), ("comm_ms", C.c_double),
        ("knn_bound_sweep_ms", C.c_double), ("knn_pool_sweep_ms", C.c_double),
        ("knn_prefilter_dim", C.c_int64), ("knn_prefilter_sample_rows", C.c_int64),
        ("knn_prefilter_resweep_queries", C.c_int64)]


def ptr(a, ctype):
    """numpy array (or None) -> typed ctypes pointer; checks dtype and contiguity."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


def f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def u32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint32)


class Codebook:
    """Flat codebook arrays in the layout pcdb_set_codebook takes (CSR vote table by codeword row)."""

    def __init__(self, words, vote_off, vote_xyz, vote_weight, vote_class, vote_instance, vote_bbox,
                 vote_class_weight, kp_train, codeword_ids, sigma2, codeword_weight=None):
        self.words = f32(words)
        self.vote_off = i64(vote_off)
        self.vote_xyz = f32(vote_xyz).reshape(-1, 3)
        self.vote_weight = f32(vote_weight)
        self.vote_class = u32(vote_class)
        self.vote_instance = u32(vote_instance)
        self.vote_bbox = f32(vote_bbox).reshape(-1, 7)
        self.vote_class_weight = f32(vote_class_weight)
        self.kp_train = f32(kp_train).reshape(-1, 3)
        self.codeword_ids = i32(codeword_ids)
        self.sigma2 = f32(sigma2)
        self.codeword_weight = f32(codeword_weight)

    @property
    def N(self):
        return self.words.shape[0]

    @property
    def D(self):
        return self.words.shape[1]

    @property
    def n_classes(self):
        return self.sigma2.shape[0]

    def rows(self, lo, hi):
        """Row shard [lo, hi) with its CSR vote slice (sharded-codebook mode)."""
        v0, v1 = int(self.vote_off[lo]), int(self.vote_off[hi])
        return Codebook(self.words[lo:hi], self.vote_off[lo:hi + 1] - v0, self.vote_xyz[v0:v1],
                        self.vote_weight[v0:v1], self.vote_class[v0:v1], self.vote_instance[v0:v1],
                        self.vote_bbox[v0:v1], self.vote_class_weight[v0:v1], self.kp_train[lo:hi],
                        self.codeword_ids[lo:hi], self.sigma2,
                        None if self.codeword_weight is None else self.codeword_weight[lo:hi])
