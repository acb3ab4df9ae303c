// stages.h — internal stage drivers of libpcdb200 (device-resident data, asynchronous on ctx->stream unless noted).
#pragma once
#include "common.cuh"

// prep.cu
int stage_compact(pcdb_ctx* ctx, int B, int64_t P, bool has_normals, bool has_rgb);
int stage_cloud_setup(pcdb_ctx* ctx, int B, int64_t n_pts, const float4* extra_kp, const int* extra_kp_cloud,
                      int64_t n_extra, float leaf, double grid_radius);
int stage_voxel_keypoints(pcdb_ctx* ctx, int B, int64_t n_pts, float leaf, int64_t* Q_out);  // syncs
int stage_grid(pcdb_ctx* ctx, int B, int64_t n_surf, int64_t Q, bool color);

// normals.cu: ws.in_xyz -> ws.in_nrm (and optionally the curvature) for clouds without normals; syncs
int stage_normals(pcdb_ctx* ctx, int B, int64_t P, float* curv_out_d);

// normals_organized.cu: ws.in_xyz (H x W x 3) -> ws.in_nrm, the organized branch of computeNormals
int stage_normals_organized(pcdb_ctx* ctx, int W, int H);

// shot.cu
size_t shot_smem_bytes(bool color);
int stage_shot(pcdb_ctx* ctx, int64_t n_surf, int64_t Q, bool color, double r_lrf, double r_shot, bool do_lrf,
               bool do_desc, const float* lrf_in_d, float* lrf_out_d, float* desc_out_d);
int stage_shot_counts(pcdb_ctx* ctx, unsigned long long out[2]);  // syncs

// knn_scan.cu
int stage_knn_scan(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, bool use_ratio,
                   float ratio_thr);
int stage_knn_rerank(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int K, int S, int cap, int dist_type);
int stage_knn_finish(pcdb_ctx* ctx, int64_t Q, int k, int K, bool use_ratio, float ratio_thr);
int stage_knn_chi_pool(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int K, int64_t total, int2* pool_rc,
                       const int* pool_q, const int* q_cnt, int* q_off, int* q_fill, DevBuf* csr_row, DevBuf* csr_d,
                       int dist_type, const int* skip);

int stage_pair_distances(pcdb_ctx* ctx, const float* a_d, const float* b_d, int64_t n, int D, int dist_type,
                         float* out_d);

// knn_gemm.cu
int gemm_prepare_codebook(pcdb_ctx* ctx, int dist_type);  // fp16 operand copy, norms, error bounds, tensor map
int stage_knn_gemm(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, bool use_ratio,
                   float ratio_thr);
bool gemm_supported(const pcdb_ctx* ctx, int dist_type);

// votes.cu
int stage_cast_votes(pcdb_ctx* ctx, const float* feat_xyz_d, const float* feat_lrf_d, const long long* feat_off_d,
                     const int* feat_cloud_d, int B, int64_t F, int k, int64_t* V_out);  // syncs

// meanshift.cu
int stage_votes_unpack(pcdb_ctx* ctx, int B, int64_t V);
// have_cloud: ws.surf4 / ws.surf_off hold the clouds the votes came from (fused path) — the single-object max types need it
int stage_find_maxima(pcdb_ctx* ctx, int B, int64_t V, bool have_cloud, int64_t* M_out, int64_t* members_out);  // syncs

// ransac.cu: Voting.RansacVoteFiltering — replaces ws.mem_idx / mem_w / mem_off by the inlier votes of the surviving
// maxima (M_ptr: device count of maxima, hM its host copy); syncs
int stage_ransac_filter(pcdb_ctx* ctx, int64_t hM, const int* M_ptr, int64_t* hMem_io);

// api.cu: exact kNN of device-resident queries against this context's descriptor rows (GEMM or scan by `mode`)
int pcdb_run_knn(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, int mode, bool use_ratio,
                 float ratio_thr);

// comm.cu (collective calls: every rank of the communicator must make them in the same order)
bool comm_active(const pcdb_ctx* ctx);
bool comm_codebook_sharded(const pcdb_ctx* ctx);   // descriptor rows sharded over the ranks, vote tables replicated
bool comm_keypoints_sharded(const pcdb_ctx* ctx);  // one scene per call, keypoints sharded over the ranks
int stage_knn_sharded(pcdb_ctx* ctx, const float* queries_d, int64_t Q_local, int k, int dist_type, int mode,
                      bool use_ratio, float ratio_thr);
int stage_slice_keypoints(pcdb_ctx* ctx, int B, int64_t Q, int64_t* Q_local_out);
int stage_gather_votes(pcdb_ctx* ctx, int B, int64_t F_local, int k, int64_t V_local, int64_t* V_out);  // syncs
float comm_last_exchange_ms(pcdb_ctx* ctx);
