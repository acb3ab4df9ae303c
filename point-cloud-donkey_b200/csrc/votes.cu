// votes.cu — K8: Hough vote casting.
//
// Replaces Codebook::castVotes (codebook/codebook.cpp:403-555, the bookkeeping under `omp critical`) and
// CodewordDistribution::castVotes / castVote (codebook/codeword_distribution.cpp:73-167) + Voting::vote
// (voting/voting.cpp:58-77).  The reference appends votes from an OpenMP loop in arbitrary order; here the order is
// defined: (feature, activation rank, stored vote), produced by a count pass, an exclusive scan and a write pass —
// a gather from the CSR vote table of the activated codeword rows and one coalesced 80-byte record per vote.
// The rotation goes through the same float quaternion route as Utils::rotateBack (utils/utils.cpp:136-178,342-394,
// 560-566) with boost::math::quaternion's operator*= evaluation order and no FMA contraction.
#include <algorithm>

#include "common.cuh"
#include "stages.h"

namespace {

struct Quat {
  float a, b, c, d;  // w x y z
};
__device__ __forceinline__ Quat qmul(const Quat l, const Quat r) {
  Quat o;
  o.a = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(l.a, r.a), __fmul_rn(l.b, r.b)), __fmul_rn(l.c, r.c)),
                  __fmul_rn(l.d, r.d));
  o.b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(l.a, r.b), __fmul_rn(l.b, r.a)), __fmul_rn(l.c, r.d)),
                  __fmul_rn(l.d, r.c));
  o.c = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(l.a, r.c), __fmul_rn(l.b, r.d)), __fmul_rn(l.c, r.a)),
                  __fmul_rn(l.d, r.b));
  o.d = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(l.a, r.d), __fmul_rn(l.b, r.c)), __fmul_rn(l.c, r.b)),
                  __fmul_rn(l.d, r.a));
  return o;
}
__device__ __forceinline__ Quat qconj(const Quat q) { return Quat{q.a, -q.b, -q.c, -q.d}; }

// Utils::getRotQuaternion + matrix2Quat: rows of the matrix = LRF axes
__device__ Quat lrf_quat(const float* rf) {
  const float m[3][3] = {{rf[0], rf[1], rf[2]}, {rf[3], rf[4], rf[5]}, {rf[6], rf[7], rf[8]}};
  float quat[4] = {0.f, 0.f, 0.f, 0.f};  // x y z w
  float trace = __fadd_rn(__fadd_rn(m[0][0], m[1][1]), m[2][2]);
  if (trace > 0.0f) {
    float root = __fsqrt_rn(__fadd_rn(trace, 1.0f));
    quat[3] = __fmul_rn(0.5f, root);
    root = __fdiv_rn(0.5f, root);
    quat[0] = __fmul_rn(__fsub_rn(m[2][1], m[1][2]), root);
    quat[1] = __fmul_rn(__fsub_rn(m[0][2], m[2][0]), root);
    quat[2] = __fmul_rn(__fsub_rn(m[1][0], m[0][1]), root);
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    float arg = (float)((double)__fsub_rn(__fsub_rn(m[i][i], m[j][j]), m[k][k]) + 1.0);
    float root = __fsqrt_rn(arg);
    quat[i] = __fmul_rn(0.5f, root);
    root = __fdiv_rn(0.5f, root);
    quat[3] = __fmul_rn(__fsub_rn(m[k][j], m[j][k]), root);
    quat[j] = __fmul_rn(__fadd_rn(m[j][i], m[i][j]), root);
    quat[k] = __fmul_rn(__fadd_rn(m[k][i], m[i][k]), root);
  }
  return Quat{quat[3], quat[0], quat[1], quat[2]};
}

struct VoteArgs {
  const float* feat_xyz;   // F x 3
  const float* feat_lrf;   // F x 9
  const int* knn_idx;      // F x k (global row ids)
  const float* knn_dist;   // F x k
  const int* knn_cnt;      // F
  int k;
  long long F;
  long long row_base;
  // codebook
  const long long* vote_off;
  const float* vote_xyz;
  const float* vote_weight;
  const unsigned* vote_class;
  const unsigned* vote_instance;
  const float* vote_bbox;
  const float* vote_class_weight;
  const float* kp_train;
  const int* ids;
  const float* cw_weight;
  const float* sigma2;
  int n_classes;
  int use_class_weight, use_vote_weight, use_matching_weight, use_codeword_weight, abs_is_int;
};

// weight and filter of codeword_distribution.cpp:92-138; returns false when the vote is dropped
__device__ __forceinline__ bool vote_weight_and_filter(const VoteArgs& a, long long v, long long row, float dist,
                                                      float& weight) {
  unsigned cls = a.vote_class[v];
  float classSigma = cls < (unsigned)a.n_classes ? a.sigma2[cls] : 1.0f;
  weight = 1.0f;
  if (a.use_class_weight) weight = __fmul_rn(weight, a.vote_class_weight ? a.vote_class_weight[v] : 1.0f);
  if (a.use_vote_weight) weight = __fmul_rn(weight, a.vote_weight[v]);
  if (a.use_matching_weight) {
    const double kTwoPi = 6.283185307179586476925286766559;
    double s = (double)classSigma;
    float mw = (float)((1 / sqrt(kTwoPi * s)) * exp(-((double)dist * (double)dist) / (2 * s)));
    weight = __fmul_rn(weight, mw);
  }
  if (a.use_codeword_weight) weight = __fmul_rn(weight, a.cw_weight[row]);
  float lim = __fmul_rn(2.f, classSigma);
  if (a.abs_is_int) {
    if ((float)abs((int)dist) > lim) return false;
  } else {
    if (fabsf(dist) > lim) return false;
  }
  if (weight < 1.1920928955078125e-07f) return false;  // std::numeric_limits<float>::epsilon()
  return true;
}

__global__ void k_vote_count(VoteArgs a, int* cnt) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= a.F * a.k) {
    if (t == a.F * a.k) cnt[t] = 0;
    return;
  }
  long long f = t / a.k;
  int j = (int)(t % a.k);
  int c = 0;
  if (j < a.knn_cnt[f] && a.knn_idx[t] >= 0) {  // negative = activated row owned by another codebook shard
    long long row = (long long)a.knn_idx[t] - a.row_base;
    float dist = a.knn_dist[t];
    for (long long v = a.vote_off[row]; v < a.vote_off[row + 1]; ++v) {
      float wgt;
      if (vote_weight_and_filter(a, v, row, dist, wgt)) ++c;
    }
  }
  cnt[t] = c;
}

__global__ void k_vote_write(VoteArgs a, const int* __restrict__ pos, const int* __restrict__ feat_cloud,
                             pcdb_vote* votes, float4* vote_pw, int* vote_cloud, int* vote_feat) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= a.F * a.k) return;
  long long f = t / a.k;
  int j = (int)(t % a.k);
  if (j >= a.knn_cnt[f] || a.knn_idx[t] < 0) return;
  long long row = (long long)a.knn_idx[t] - a.row_base;
  float dist = a.knn_dist[t];
  int o = pos[t];
  float rf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) rf[i] = a.feat_lrf[f * 9 + i];
  const Quat rq = lrf_quat(rf);
  const float kx = a.feat_xyz[f * 3], ky = a.feat_xyz[f * 3 + 1], kz = a.feat_xyz[f * 3 + 2];
  for (long long v = a.vote_off[row]; v < a.vote_off[row + 1]; ++v) {
    float wgt;
    if (!vote_weight_and_filter(a, v, row, dist, wgt)) continue;
    // center = keyPos + q* p q   (codeword_distribution.cpp:156-159)
    Quat p{0.f, a.vote_xyz[3 * v], a.vote_xyz[3 * v + 1], a.vote_xyz[3 * v + 2]};
    Quat r = qmul(qmul(qconj(rq), p), rq);
    Quat bq{a.vote_bbox[7 * v], a.vote_bbox[7 * v + 1], a.vote_bbox[7 * v + 2], a.vote_bbox[7 * v + 3]};
    Quat nb = qmul(bq, rq);  // :162-164
    alignas(16) pcdb_vote out;
    out.position[0] = __fadd_rn(kx, r.b);
    out.position[1] = __fadd_rn(ky, r.c);
    out.position[2] = __fadd_rn(kz, r.d);
    out.weight = wgt;
    out.keypoint[0] = kx;
    out.keypoint[1] = ky;
    out.keypoint[2] = kz;
    out.class_id = a.vote_class[v];
    out.keypoint_training[0] = a.kp_train[3 * row];
    out.keypoint_training[1] = a.kp_train[3 * row + 1];
    out.keypoint_training[2] = a.kp_train[3 * row + 2];
    out.instance_id = a.vote_instance[v];
    out.bbox_quat[0] = nb.a;
    out.bbox_quat[1] = nb.b;
    out.bbox_quat[2] = nb.c;
    out.bbox_quat[3] = nb.d;
    out.bbox_size[0] = a.vote_bbox[7 * v + 4];
    out.bbox_size[1] = a.vote_bbox[7 * v + 5];
    out.bbox_size[2] = a.vote_bbox[7 * v + 6];
    out.codeword_id = a.ids ? a.ids[row] : (int)(row + a.row_base);
    // 80-byte record as five 16-byte stores
    const float4* src = reinterpret_cast<const float4*>(&out);
    float4* dst = reinterpret_cast<float4*>(votes + o);
#pragma unroll
    for (int i = 0; i < 5; ++i) dst[i] = src[i];
    vote_pw[o] = make_float4(out.position[0], out.position[1], out.position[2], wgt);
    vote_cloud[o] = feat_cloud[f];
    if (vote_feat) vote_feat[o] = (int)f;
    ++o;
  }
}

__global__ void k_vote_offsets(const long long* __restrict__ feat_off, int B, int k, const int* __restrict__ pos,
                               long long* vote_off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  vote_off[b] = pos[feat_off[b] * k];
}

}  // namespace

// feat_* / knn_* are device arrays; writes ws.votes / ws.vote_pw / ws.vote_cloud / ws.vote_off and returns V
// (one stream sync: V sizes the voting buffers).
int stage_cast_votes(pcdb_ctx* ctx, const float* feat_xyz_d, const float* feat_lrf_d, const long long* feat_off_d,
                     const int* feat_cloud_d, int B, int64_t F, int k, int64_t* V_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  const pcdb_params& P = ctx->prm;
  PCDB_CUDA(w.vote_off.ensure(sizeof(long long) * (B + 1)));
  *V_out = 0;
  if (F == 0 || cb.N_table == 0) {
    PCDB_CUDA(cudaMemsetAsync(w.vote_off.p, 0, sizeof(long long) * (B + 1), st));
    return PCDB_OK;
  }
  VoteArgs a;
  a.feat_xyz = feat_xyz_d;
  a.feat_lrf = feat_lrf_d;
  a.knn_idx = w.knn_idx.as<int>();
  a.knn_dist = w.knn_dist.as<float>();
  a.knn_cnt = w.knn_cnt.as<int>();
  a.k = k;
  a.F = F;
  a.row_base = cb.row_base;
  a.vote_off = cb.vote_off.as<long long>();
  a.vote_xyz = cb.vote_xyz.as<float>();
  a.vote_weight = cb.vote_weight.as<float>();
  a.vote_class = cb.vote_class.as<unsigned>();
  a.vote_instance = cb.vote_instance.as<unsigned>();
  a.vote_bbox = cb.vote_bbox.as<float>();
  a.vote_class_weight = cb.vote_class_weight.p ? cb.vote_class_weight.as<float>() : nullptr;
  a.kp_train = cb.kp_train.as<float>();
  a.ids = cb.ids.p ? cb.ids.as<int>() : nullptr;
  a.cw_weight = cb.cw_weight.as<float>();
  a.sigma2 = cb.sigma2.as<float>();
  a.n_classes = cb.n_classes;
  a.use_class_weight = P.use_class_weight;
  a.use_vote_weight = P.use_vote_weight;
  a.use_matching_weight = P.use_matching_weight;
  a.use_codeword_weight = P.use_codeword_weight;
  a.abs_is_int = P.filter_abs_is_int;
  const int64_t T = F * k;
  if (T * std::max(1, cb.max_votes_per_word) > 0x7fffff00ll)
    return ctx->fail(PCDB_E_INVALID, "more than 2^31 votes possible in one batch (%lld activations x %d votes per word)",
                     (long long)T, cb.max_votes_per_word);
  PCDB_CUDA(w.vote_cnt.ensure(sizeof(int) * (T + 2)));
  PCDB_CUDA(w.vote_pos.ensure(sizeof(int) * (T + 2)));
  k_vote_count<<<cdiv(T + 1, 128), 128, 0, st>>>(a, w.vote_cnt.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.vote_cnt.as<int>(), w.vote_pos.as<int>(), T + 1));
  int V = 0;
  PCDB_TRY(pcdb_read_small(ctx, &V, w.vote_pos.as<int>() + T, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  PCDB_CUDA(w.votes.ensure(sizeof(pcdb_vote) * (size_t)(V + 1)));
  PCDB_CUDA(w.vote_pw.ensure(sizeof(float4) * (size_t)(V + 1)));
  PCDB_CUDA(w.vote_cloud.ensure(sizeof(int) * (size_t)(V + 1)));
  const bool want_feat = comm_keypoints_sharded(ctx);  // the vote -> feature map restores the global order after the gather
  if (want_feat) PCDB_CUDA(w.vote_feat.ensure(sizeof(int) * (size_t)(V + 1)));
  if (V > 0) {
    k_vote_write<<<cdiv(T, 128), 128, 0, st>>>(a, w.vote_pos.as<int>(), feat_cloud_d, w.votes.as<pcdb_vote>(),
                                               w.vote_pw.as<float4>(), w.vote_cloud.as<int>(),
                                               want_feat ? w.vote_feat.as<int>() : nullptr);
    PCDB_LAUNCH_CHECK();
  }
  k_vote_offsets<<<cdiv(B + 1, 128), 128, 0, st>>>(feat_off_d, B, k, w.vote_pos.as<int>(), w.vote_off.as<long long>());
  PCDB_LAUNCH_CHECK();
  *V_out = V;
  return PCDB_OK;
}
