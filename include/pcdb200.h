/*
 * pcdb200.h — C-ABI boundary of the B200-native classification hot path of
 * point-cloud-donkey (SHOT/CSHOT at voxel-grid keypoints -> kNN codebook
 * activation -> Hough vote casting -> mean-shift maxima).
 *
 * The reference (vseib/point-cloud-donkey) has no FFI: its plug-in surface is
 * C++ virtual hooks behind string-keyed factories.  Every entry point below
 * names the reference hook it replaces (file:line under
 * src/implicit_shape_model/ unless stated otherwise).  The host-side C++ shim
 * (point-cloud-donkey_b200/host) wraps these behind the reference's own class
 * names; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature
 *   - all pointers are HOST memory unless the name ends in `_d` (device)
 *   - clouds are concatenated; `*_off` arrays have B+1 int64 entries
 *   - every function returns PCDB_OK (0) or a negative pcdb_status; the text
 *     of the last failure is kept per context (pcdb_last_error)
 *   - there is NO CPU fallback: without a usable sm_100 device pcdb_create
 *     fails with PCDB_E_NO_DEVICE
 */
#ifndef PCDB200_H_
#define PCDB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCDB_ABI_VERSION 5

typedef enum pcdb_status {
  PCDB_OK = 0,
  PCDB_E_INVALID = -1,     /* bad argument / bad parameter (ism3d::BadParamException) */
  PCDB_E_NO_DEVICE = -2,   /* no CUDA device / not sm_100 */
  PCDB_E_CUDA = -3,        /* CUDA runtime failure (text in pcdb_last_error) */
  PCDB_E_CAPACITY = -4,    /* caller-provided output buffer too small */
  PCDB_E_STATE = -5,       /* call order (e.g. classify before set_codebook) */
  PCDB_E_UNSUPPORTED = -6, /* a reference option outside the built hot path */
  PCDB_E_COMM = -7         /* multi-GPU exchange failed */
} pcdb_status;

enum { PCDB_FEATURE_SHOT = 0, PCDB_FEATURE_CSHOT = 1 };           /* features_factory.h:54,62 */
enum { PCDB_DIST_EUCLIDEAN = 0, PCDB_DIST_CHISQUARED = 1 };       /* utils/distance.h:42-75 (squared L2 / chi^2) */
enum { PCDB_KERNEL_GAUSSIAN = 0, PCDB_KERNEL_UNIFORM = 1 };       /* voting_mean_shift.cpp:378-417 */
enum { PCDB_SUPPRESS_AVERAGE = 0, PCDB_SUPPRESS_SUPPRESS = 1 };   /* voting_mean_shift.cpp:98-122 */
enum { PCDB_MAXFILTER_NONE = 0, PCDB_MAXFILTER_SIMPLE = 1, PCDB_MAXFILTER_MERGE = 2 }; /* maxima_handler.cpp:272-383 */
enum { PCDB_KNN_AUTO = 0, PCDB_KNN_SCAN = 1, PCDB_KNN_GEMM = 2 }; /* which exact-kNN kernel family */
/* Voting.BinOrBandwidthType (maxima_handler.cpp:509-521): Config|Fixed, FirstDim|ObjectRadius, SecondDim|BoundingBoxMedian */
enum { PCDB_RADIUS_CONFIG = 0, PCDB_RADIUS_FIRST_DIM = 1, PCDB_RADIUS_SECOND_DIM = 2 };
/* Voting.SingleObjectMaxType (maxima_handler.h:42-58): Default|None, BandwidthVotes, VotingSpaceVotes, ModelRadiusVotes */
enum { PCDB_SOMAX_DEFAULT = 0, PCDB_SOMAX_BANDWIDTH = 1, PCDB_SOMAX_VOTING_SPACE = 2, PCDB_SOMAX_MODEL_RADIUS = 3 };
/* Voting.RansacInlierThresholdType (voting/voting.cpp:50,112-122): the threshold as configured, or times the class'
 * learned object radius / median bounding-box dimension (pcdb_set_class_dimensions) */
enum { PCDB_RANSAC_FIXED = 0, PCDB_RANSAC_OBJECT_RADIUS = 1, PCDB_RANSAC_BBOX_MEDIAN = 2 };

#define PCDB_SHOT_DIM 352
#define PCDB_CSHOT_DIM 1344
#define PCDB_MAX_K 16

/* Hot-path parameters; names and defaults follow the reference's JSON keys
 * (SURVEY.md App. C).  Filled by the host shim from the .ism file. */
typedef struct pcdb_params {
  /* Features (features_shot.cpp:21, features.cpp:32-33) */
  int32_t feature_type;   /* PCDB_FEATURE_*; Features.Type */
  double feature_radius;  /* Features.Radius (double member, features_shot.h:43) */
  double lrf_radius;      /* Features.ReferenceFrameRadius: the float member promoted to double */
  /* Keypoints (keypoints_voxel_grid.cpp:23) */
  float leaf_size;        /* Keypoints.LeafSize */
  /* Activation (activation_strategy_knn.cpp:19, activation_strategy.cpp:19-20) */
  int32_t distance_type;  /* Parameters.DistanceType */
  int32_t knn_k;          /* Codebook.ActivationStrategy.K, 1..PCDB_MAX_K */
  int32_t use_distance_ratio;
  float distance_ratio_threshold;
  /* Vote weights (codebook.cpp:32-35) */
  int32_t use_class_weight, use_vote_weight, use_matching_weight, use_codeword_weight;
  int32_t filter_abs_is_int; /* 1 = treat the unqualified abs() at codeword_distribution.cpp:131 as abs(int) */
  /* Mean-shift voting (voting_mean_shift.cpp:22-26, voting.cpp:28-36) */
  float bandwidth;        /* Voting.Bandwidth */
  float ms_threshold;     /* Voting.Threshold */
  int32_t ms_max_iter;    /* Voting.MaxIter */
  int32_t ms_kernel;      /* PCDB_KERNEL_* */
  int32_t maxima_suppression; /* PCDB_SUPPRESS_* */
  float min_threshold;    /* Voting.MinThreshold */
  int32_t min_votes_threshold; /* Voting.MinVotesThreshold */
  int32_t best_k;         /* Voting.BestK (<=0: keep all) */
  int32_t average_rotation;    /* Voting.AverageRotation */
  int32_t single_object_mode;  /* Voting.SingleObjectMode (only gates cross-class filtering on this path) */
  /* Normals, used when a call is given no normals (implicit_shape_model.cpp:110-112, 940-1037) */
  float normal_radius;               /* Parameters.NormalRadius */
  int32_t consistent_normals_method; /* Parameters.ConsistentNormalsMethod: 0 towards the origin, 1 away from the
                                        centroid, 2 inverted SHOT-LRF z axis (the code default) */
  /* Cross-class maxima filtering when !single_object_mode (voting.cpp:265-268, maxima_handler.cpp:272-383) */
  int32_t max_filter_type;           /* PCDB_MAXFILTER_*; Voting.MaxFilterType */
  /* Per-class search distance (maxima_handler.cpp:509-521, voting_mean_shift.cpp:47-49): Config = Voting.Bandwidth for
   * every class, else the class' learned object radius / median bounding-box dimension (pcdb_set_class_dimensions)
   * times radius_factor */
  int32_t radius_type;               /* PCDB_RADIUS_*; Voting.BinOrBandwidthType */
  float radius_factor;               /* Voting.BinOrBandwidthFactor */
  /* Single-object mode without mean shift (voting_mean_shift.cpp:124-155): one maximum per class at the cloud's
   * centroid, votes collected within the bandwidth / the model radius / the whole voting space.  Only consulted when
   * single_object_mode != 0; needs the cloud, so the fused entries support it and pcdb_find_maxima does not */
  int32_t single_object_max_type;    /* PCDB_SOMAX_*; Voting.SingleObjectMaxType */
  /* RANSAC vote filtering of every maximum before it is reduced (voting/voting.cpp:47-50,110-127,356-433): a rigid
   * transform between the member votes' training keypoints and scene keypoints is fitted by sample consensus (3
   * correspondences per hypothesis, at most 10000 iterations, PCL's adaptive stop at probability 0.99); only the
   * inlier votes stay, and a maximum whose fit fails, has fewer than 3 inliers or is the identity (Eigen isIdentity
   * at 1e-4, as written in the reference) is dropped.  The reference draws its samples from PCL's Boost RNG (seed and
   * shuffle state are PCL internals); here sample `it` of a maximum is a pure function of (class, ordinal of the
   * maximum in its class, it) — same model, same stopping rule, a defined and reproducible sequence.
   * RansacRefineModel = true is rejected (PCDB_E_UNSUPPORTED). */
  int32_t ransac_vote_filtering;     /* Voting.RansacVoteFiltering */
  float ransac_inlier_threshold;     /* Voting.RansacInlierThreshold */
  int32_t ransac_threshold_type;     /* PCDB_RANSAC_*; Voting.RansacInlierThresholdType */
  int32_t ransac_refine_model;       /* Voting.RansacRefineModel: must be 0 */
} pcdb_params;

/* One Hough vote — ism3d::Vote, voting/voting_maximum.h:25-42 (80 bytes). */
typedef struct pcdb_vote {
  float position[3];
  float weight;
  float keypoint[3];
  uint32_t class_id;
  float keypoint_training[3];
  uint32_t instance_id;
  float bbox_quat[4]; /* w,x,y,z  (Utils::BoundingBox::rotQuat) */
  float bbox_size[3];
  int32_t codeword_id;
} pcdb_vote;

/* One voting maximum — ism3d::VotingMaximum, voting/voting_maximum.h:51-88. */
typedef struct pcdb_maximum {
  float position[3];
  float weight;            /* normalised over the cloud's maxima (voting.cpp:441-462) */
  uint32_t class_id;
  uint32_t instance_id;
  float instance_weight;
  float raw_weight;        /* sum of member vote weights before normalisation */
  float bbox_quat[4];      /* w,x,y,z */
  float bbox_size[3];
  int32_t n_votes;
  int64_t vote_begin;      /* first entry of this maximum in the member list (pcdb_get_maximum_votes) */
} pcdb_maximum;

typedef struct pcdb_ctx pcdb_ctx;

/* ---- lifecycle ------------------------------------------------------- */
int pcdb_abi_version(void);
/* Replaces: ImplicitShapeModel ctor + first-call FLANN hook (implicit_shape_model.cpp:131-143,651-660). */
int pcdb_create(pcdb_ctx** out, int device);
void pcdb_destroy(pcdb_ctx* ctx);
const char* pcdb_last_error(const pcdb_ctx* ctx); /* ctx may be NULL: error of a failed pcdb_create */
void pcdb_default_params(pcdb_params* p);         /* code defaults of SURVEY App. C */
int pcdb_set_params(pcdb_ctx* ctx, const pcdb_params* p); /* JSONObject::readObject (utils/json_object.cpp:97-178) */
/* Launch on a caller-owned CUDA stream (cudaStream_t as void*); NULL = the context's own stream. */
int pcdb_set_stream(pcdb_ctx* ctx, void* cuda_stream);

/* ---- model upload ---------------------------------------------------- */
/* Replaces: FlannHelper::createDataset/buildIndex (utils/flann_helper.cpp:21-70) and the in-memory
 * Codebook / CodewordDistribution tables (codebook/codebook.cpp:763-950, codeword_distribution.cpp:395-465).
 * words: N x D row-major in codeword-id order (codebook.cpp:857-859).  Votes are CSR by codeword row:
 * vote_off[N+1]; per vote: LRF-relative offset xyz, learned weight, class, instance, bbox (quat wxyz + size),
 * statistical class weight.  kp_train: N x 3 (Codeword::getFeaturePosition).  codeword_ids: the stored ids
 * (NULL = row index + row_base).  codeword_weight: N (NULL = 1).  class_sigma2[c] for c < n_classes
 * (missing class => pass 1.0f as the reference does, codeword_distribution.cpp:117-121). */
int pcdb_set_codebook(pcdb_ctx* ctx, const float* words, int64_t N, int32_t D,
                      const int64_t* vote_off, const float* vote_xyz, const float* vote_weight,
                      const uint32_t* vote_class, const uint32_t* vote_instance,
                      const float* vote_bbox /* V x 7 */, const float* vote_class_weight /* V or NULL */,
                      const float* kp_train, const int32_t* codeword_ids, const float* codeword_weight,
                      const float* class_sigma2, int32_t n_classes, int64_t row_base);

/* Learned per-class dimensions — Voting::forwardBoxesAndRadii / iLoadData (voting/voting.cpp:497-557,619-650), the
 * table behind BinOrBandwidthType != Config: first[c] = mean object radius, second[c] = mean median bounding-box side of
 * class c's training models.  Classes without an entry: pass 0 (the reference's map::at would throw). */
int pcdb_set_class_dimensions(pcdb_ctx* ctx, const float* first_dim, const float* second_dim, int32_t n_classes);

/* ---- stage-level entry points (one per reference hook) ---------------- */
/* KeypointsVoxelGrid::iComputeKeypoints (keypoints/keypoints_voxel_grid.cpp:30-46 -> pcl::VoxelGrid).
 * rgb: packed 0x00RRGGBB per point or NULL.  kp_capacity in keypoints. */
int pcdb_voxel_keypoints(pcdb_ctx* ctx, const float* xyz, const uint32_t* rgb, const int64_t* cloud_off,
                         int32_t B, float leaf, float* kp_xyz_out, uint32_t* kp_rgb_out,
                         int64_t* kp_off_out, int64_t kp_capacity);

/* Radius neighbourhoods (pcl::search::KdTree::radiusSearch as used at features.cpp:243-249 and
 * features_shot.cpp:37-60): for every keypoint the surface indices (cloud-local) with d^2 < float(r*r),
 * sorted by (d^2, index).  nbr_off has Q+1 entries.  Parity/debug entry; the descriptor kernels fuse this. */
int pcdb_radius_neighbours(pcdb_ctx* ctx, const float* surf_xyz, const int64_t* surf_off,
                           const float* kp_xyz, const int64_t* kp_off, int32_t B, double radius,
                           int64_t* nbr_off_out, int32_t* nbr_idx_out, float* nbr_d2_out, int64_t capacity);

/* Features::computeSHOTReferenceFrames (features/features.cpp:238-252 -> pcl::SHOTLocalReferenceFrameEstimationOMP).
 * lrf9_out: Q x 9 (x_axis, y_axis, z_axis); NaN rows for invalid frames. */
int pcdb_shot_lrf(pcdb_ctx* ctx, const float* surf_xyz, const int64_t* surf_off, const float* kp_xyz,
                  const int64_t* kp_off, int32_t B, double radius, float* lrf9_out);

/* FeaturesSHOT / FeaturesCSHOT::iComputeDescriptors (features/features_shot.cpp:28-81,
 * features/features_cshot.cpp:28-103 -> pcl::SHOT(Color)EstimationOMP) with given reference frames.
 * desc_out: Q x 352 (SHOT) or Q x 1344 (CSHOT); all-NaN rows where the reference yields NaN. */
int pcdb_shot_describe(pcdb_ctx* ctx, int32_t feature_type, const float* surf_xyz, const float* surf_normals,
                       const uint32_t* surf_rgb, const int64_t* surf_off, const float* kp_xyz,
                       const uint32_t* kp_rgb, const float* kp_lrf9, const int64_t* kp_off, int32_t B,
                       double radius, float* desc_out);

/* ImplicitShapeModel::computeNormals for unorganized clouds (implicit_shape_model.cpp:940-1037 ->
 * NormalEstimationOMPWithEigVals, third_party/pcl_normal_3d_omp_with_eigenvalues/; NormalOrientation::processSHOTLRF,
 * utils/normal_orientation.cpp:48-110), with the context's normal_radius / consistent_normals_method.
 * normals_out: P x 3 aligned with xyz (NaN for non-finite points and where the reference yields NaN);
 * curvature_out: P or NULL.  The same estimation runs inside pcdb_compute_features / pcdb_classify_batch(_d) when
 * they are called with normals == NULL (the reference's hasNormals == false path). */
int pcdb_compute_normals(pcdb_ctx* ctx, const float* xyz, const int64_t* cloud_off, int32_t B, float* normals_out,
                         float* curvature_out);
/* The organized branch of ImplicitShapeModel::computeNormals (implicit_shape_model.cpp:948-966), taken by the reference
 * when a cloud isOrganized() (PCD HEIGHT > 1, Kinect-style grids with NaN holes): pcl::IntegralImageNormalEstimation,
 * AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02, NormalSmoothingSize 10, normals flipped towards the sensor origin.
 * xyz and normals_out: height x width x 3, row-major; normals are NaN where PCL leaves them NaN (image border of 10
 * pixels, holes, depth discontinuities). */
int pcdb_compute_normals_organized(pcdb_ctx* ctx, const float* xyz, int32_t width, int32_t height, float* normals_out);

/* Features::operator() + removeNaNFeatures (features/features.cpp:40-116, implicit_shape_model.cpp:1276-1308):
 * keypoints -> LRF -> drop invalid -> descriptor -> drop NaN; uses the context's params.  Outputs are the
 * surviving features (position, LRF, descriptor) per cloud. */
int pcdb_compute_features(pcdb_ctx* ctx, const float* xyz, const float* normals, const uint32_t* rgb,
                          const int64_t* cloud_off, int32_t B, float* feat_xyz_out, float* feat_lrf9_out,
                          float* feat_desc_out, int64_t* feat_off_out, int64_t feat_capacity);

/* ActivationStrategyKNN::activateKNN (activation_strategy/activation_strategy_knn.h:41-126) against the uploaded
 * codebook, exact search (FLANNExactMatch semantics), FLANN functor values.  idx_out/dist_out: Q x k, ascending
 * distance, ties -> lower row; count_out[q] = number of activated rows (0 after a failed ratio test, N if N<=k). */
int pcdb_knn(pcdb_ctx* ctx, const float* queries, int64_t Q, int32_t k, int32_t dist_type, int32_t mode,
             int32_t* idx_out, float* dist_out, int32_t* count_out);

/* Distance functor values for n explicit (a[i], b[i]) row pairs — ism3d::Distance::operator() (utils/distance.cpp:27-52),
 * used by training for the class variances (codebook/codebook.cpp:166-193).  a, b: n x D. */
int pcdb_distance_pairs(pcdb_ctx* ctx, const float* a, const float* b, int64_t n, int32_t D, int32_t dist_type,
                        float* out);

/* Codebook::castVotes + CodewordDistribution::castVotes/castVote (codebook/codebook.cpp:403-555,
 * codebook/codeword_distribution.cpp:73-167).  Votes are emitted per cloud in (feature, activation rank,
 * stored vote) order.  vote_off_out has B+1 entries.  Entries of knn_idx that are negative are skipped (rows owned by
 * another shard of a row-sharded codebook). */
int pcdb_cast_votes(pcdb_ctx* ctx, const float* feat_xyz, const float* feat_lrf9, const int64_t* feat_off,
                    int32_t B, const int32_t* knn_idx, const float* knn_dist, const int32_t* knn_count,
                    int32_t k, pcdb_vote* votes_out, int64_t* vote_off_out, int64_t vote_capacity);

/* Voting::findMaxima + VotingMeanShift::iFindMaxima (voting/voting.cpp:79-328,
 * voting/voting_mean_shift.cpp:39-177).  maxima sorted per cloud by weight (descending). */
int pcdb_find_maxima(pcdb_ctx* ctx, const pcdb_vote* votes, const int64_t* vote_off, int32_t B,
                     pcdb_maximum* maxima_out, int64_t* maxima_off_out, int64_t maxima_capacity);
/* Member votes of the maxima of the last pcdb_find_maxima / pcdb_classify_batch call
 * (VotingMaximum::votes): indices into that call's vote array and the kernel-re-weighted weights. */
int pcdb_get_maximum_votes(pcdb_ctx* ctx, int64_t* vote_index_out, float* vote_weight_out, int64_t capacity,
                           int64_t* n_out);
/* Votes of the last pcdb_classify_batch call (Voting::getVotes, read by training_gui.cpp:1012). */
int pcdb_get_votes(pcdb_ctx* ctx, pcdb_vote* votes_out, int64_t* vote_off_out, int64_t capacity);
/* Sizes of what the last pcdb_classify_batch / pcdb_find_maxima / pcdb_cast_votes call left on the device: votes,
 * maxima kept after the per-cloud threshold / BestK, member entries of pcdb_get_maximum_votes.  Callers size their
 * buffers from these instead of from the point count (std::vector<VotingMaximum> grows on demand in the reference). */
int pcdb_get_last_sizes(pcdb_ctx* ctx, int64_t* n_votes_out, int64_t* n_maxima_out, int64_t* n_members_out);
/* Maxima of the last pcdb_classify_batch / pcdb_find_maxima call again (no recomputation): a caller that passed
 * maxima_out == NULL, or too small a capacity, fetches them once it knows the count. */
int pcdb_get_maxima(pcdb_ctx* ctx, pcdb_maximum* maxima_out, int64_t* maxima_off_out, int64_t maxima_capacity);

/* ---- fused batch path (the throughput entry) -------------------------- */
/* ImplicitShapeModel::detect (implicit_shape_model.cpp:583-712) for B clouds at once, label pick of
 * eval_tool (src/eval_tool/eval_classification.cpp:412-417) included.  Host buffers in, host buffers out.
 * normals == NULL: the reference's hasNormals == false path, normals are estimated (see pcdb_compute_normals).
 * label_out[b] = class of the best maximum or -1.  maxima_out/maxima_off_out may be NULL.
 * times_ms_out[7]: complete, features, keypoints, normals, flann, voting, maxima (implicit_shape_model.cpp:160). */
int pcdb_classify_batch(pcdb_ctx* ctx, const float* xyz, const float* normals, const uint32_t* rgb,
                        const int64_t* cloud_off, int32_t B, int32_t* label_out, pcdb_maximum* maxima_out,
                        int64_t* maxima_off_out, int64_t maxima_capacity, double* times_ms_out);
/* Same, inputs already resident in device memory (xyz P x 3, normals P x 3 or NULL = estimate them, rgb P or NULL,
 * cloud_off on HOST), labels written to device memory. */
int pcdb_classify_batch_d(pcdb_ctx* ctx, const float* xyz_d, const float* normals_d, const uint32_t* rgb_d,
                          const int64_t* cloud_off, int32_t B, int32_t* label_out_d);

/* ---- multi-GPU (one process or thread per GPU, NCCL over NVLink / NVSwitch) ------------------------------- */
/* SURVEY 8(b) `pcdb_comm_init`, 8(e).  The reference has a single FLANN index in one address space
 * (utils/flann_helper.cpp:21-70, searched at codebook/codebook.cpp:483-538) and one keypoint loop
 * (implicit_shape_model.cpp:583-712); these entry points split either over the GPUs of a box.
 * All calls below, and every pcdb_knn / pcdb_classify_batch(_d) on a context whose codebook is sharded or whose
 * keypoints are sharded, are COLLECTIVE: every rank makes them, in the same order.  NCCL is bound at run time
 * (libnccl.so.2); without it pcdb_comm_* return PCDB_E_COMM and everything else works. */
#define PCDB_COMM_ID_BYTES 128
/* ncclGetUniqueId: rank 0 calls it, the host program hands the 128 bytes to the other ranks (MPI, torch, a file). */
int pcdb_comm_unique_id(void* id_out, int32_t bytes);
int pcdb_comm_init(pcdb_ctx* ctx, int32_t rank, int32_t n_ranks, const void* unique_id); /* ncclCommInitRank on ctx's device */
int pcdb_comm_destroy(pcdb_ctx* ctx);
int pcdb_comm_info(pcdb_ctx* ctx, int32_t* rank_out, int32_t* n_ranks_out, int32_t* nccl_version_out);
/* Row-sharded codebook (config C4): this rank holds descriptor rows [row_lo, row_hi) of the N_total x D matrix
 * (`words` points at row row_lo) and the COMPLETE vote tables (all arrays as in pcdb_set_codebook, indexed by global
 * row).  Afterwards pcdb_knn and pcdb_classify_batch(_d) take THIS RANK's queries / clouds and return their results as
 * an unsharded codebook would, bit for bit: query all-gather, local tcgen05 search, one all-to-all of
 * (f32 distance, i32 row) x K per query, device-side merge (ties -> lower row), votes cast by the query's owner.
 * Speed, not results, depends on WHICH rows a shard holds: training appends codewords class by class
 * (implicit_shape_model.cpp:447-475), so deal the rows of the table over the shards before sharding it (INTEGRATION.md D;
 * pcdb200.sharded.interleave_codebook) — a shard without near words for most queries cannot use the pre-filter. */
int pcdb_set_codebook_sharded(pcdb_ctx* ctx, const float* words, int64_t row_lo, int64_t row_hi, int64_t N_total,
                              int32_t D, const int64_t* vote_off, const float* vote_xyz, const float* vote_weight,
                              const uint32_t* vote_class, const uint32_t* vote_instance, const float* vote_bbox,
                              const float* vote_class_weight, const float* kp_train, const int32_t* codeword_ids,
                              const float* codeword_weight, const float* class_sigma2, int32_t n_classes);
/* Keypoint-sharded scene (config C5): when enabled, pcdb_classify_batch(_d) with B == 1 is collective over ranks that
 * all pass the SAME scene and hold the same (replicated) codebook: rank r processes the r-th contiguous slice of the
 * voxel-grid keypoints, the votes are all-gathered in keypoint order and every rank returns the full result. */
int pcdb_comm_shard_keypoints(pcdb_ctx* ctx, int32_t enable);

/* Per-query merge of per-shard top-k lists held in HOST memory (the device-resident exchange above does not
 * need it).  cand_* : S x Q x k (shard-major), global row ids; result: Q x k ascending (distance, row). */
int pcdb_merge_topk(pcdb_ctx* ctx, const int32_t* cand_idx, const float* cand_dist, int32_t S, int64_t Q,
                    int32_t k, int32_t* idx_out, float* dist_out);

/* ---- introspection for bench / profiling ------------------------------ */
typedef struct pcdb_stats {
  int64_t n_points, n_keypoints, n_features, n_neighbours_lrf, n_neighbours_shot, n_votes, n_maxima;
  int64_t knn_queries, knn_candidates, knn_fallback_queries;
  int64_t kernel_launches;  /* launches of this library's own kernels since the last reset */
  /* device time of the last classify batch (CUDA events on the context's stream) */
  double features_ms, knn_ms, knn_gemm_ms /* tcgen05 activation kernel(s) alone */, votes_ms, maxima_ms;
  /* multi-GPU exchange steps: bytes this rank received over NVLink since the last reset; device time of the last
   * batch's exchanges (query all-gather + top-k all-to-all + merge, or the vote all-gather) */
  int64_t comm_bytes;
  double comm_ms;
  /* Euclidean activation with the PCA pre-filter / chi^2 sandwich (last batch): device time of the bound sweep (over
   * the codebook sample / the sqrt rows) and of the pooled sweep (over the projected operands / the sqrt rows again);
   * projected dimension and sample rows in use (0 = plain single sweep); queries handed back to the plain sweep since
   * the last reset */
  double knn_bound_sweep_ms, knn_pool_sweep_ms;
  int64_t knn_prefilter_dim, knn_prefilter_sample_rows, knn_prefilter_resweep_queries;
} pcdb_stats;
int pcdb_get_stats(pcdb_ctx* ctx, pcdb_stats* out);
int pcdb_reset_stats(pcdb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PCDB200_H_ */
