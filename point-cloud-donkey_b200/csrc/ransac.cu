// ransac.cu — RANSAC vote filtering of the maxima (Voting.RansacVoteFiltering).
//
// Replaces Voting::filterVotesWithRansac (voting/voting.cpp:110-127,356-433): pcl::registration::
// CorrespondenceRejectorSampleConsensus between the member votes' training keypoints (source) and scene keypoints
// (target) of every maximum — rigid model from 3 correspondences, at most 10000 iterations, probability 0.99 with
// PCL's adaptive stop (RandomSampleConsensus::computeModel), sample goodness as SampleConsensusModelRegistration
// (pairwise source distances above ((sum of sqrt eigenvalues of the source covariance) / 3)^2, at most 1000 draws per
// iteration).  A maximum whose fit fails, keeps fewer than 3 inliers or is the identity (Eigen isIdentity(1e-4)) is
// dropped; the others keep their inlier votes only.
//
// One CTA per maximum.  Hypotheses are pure functions of (maximum key, iteration), so the 8 warps evaluate 8 iterations
// at a time and one thread then replays PCL's sequential bookkeeping (best model so far, adaptive iteration bound) over
// the 8 results in order: the outcome is the sequential algorithm's.  Everything that decides an inlier is fp64 and this
// file is compiled with -fmad=false, so the arithmetic is the oracle's operation for operation.
#include <cmath>

#include "common.cuh"
#include "stages.h"

namespace {

constexpr int kRansacThreads = 256;
constexpr int kRansacWarps = kRansacThreads / 32;

__device__ __forceinline__ unsigned ransac_hash(unsigned a, unsigned b, unsigned c, unsigned d) {
  unsigned h = 0x811C9DC5u;
  const unsigned v[4] = {a, b, c, d};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h ^= v[i];
    h *= 0x01000193u;
    h ^= h >> 15;
  }
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

// cyclic Jacobi for a symmetric N x N matrix: eigenvalues end up on the diagonal of A, eigenvectors in the columns of V
template <int N>
__device__ void jacobi_sym(double (&A)[N][N], double (&V)[N][N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
#pragma unroll
    for (int p = 0; p < N; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) off += fabs(A[p][q]);
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (!isfinite(theta)) t = 0.0;
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        A[p][q] = A[q][p] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// Horn's closed form: rotation = eigenvector of the largest eigenvalue of the 4 x 4 matrix of the cross-covariance
__device__ void rigid_from_three(const double (&s)[3][3], const double (&g)[3][3], double (&R)[9], double (&t)[3]) {
  double cs[3], cg[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cs[a] = ((s[0][a] + s[1][a]) + s[2][a]) / 3.0;
    cg[a] = ((g[0][a] + g[1][a]) + g[2][a]) / 3.0;
  }
  double M[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double acc = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += (s[k][a] - cs[a]) * (g[k][b] - cg[b]);
      M[a][b] = acc;
    }
  double Nm[4][4] = {
      {(M[0][0] + M[1][1]) + M[2][2], M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]},
      {0, (M[0][0] - M[1][1]) - M[2][2], M[0][1] + M[1][0], M[2][0] + M[0][2]},
      {0, 0, (M[1][1] - M[0][0]) - M[2][2], M[1][2] + M[2][1]},
      {0, 0, 0, (M[2][2] - M[0][0]) - M[1][1]}};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) Nm[i][j] = Nm[j][i];
  double V[4][4];
  jacobi_sym<4>(Nm, V);
  double bv = Nm[0][0];
  double q[4] = {V[0][0], V[1][0], V[2][0], V[3][0]};
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (Nm[i][i] > bv) {
      bv = Nm[i][i];
#pragma unroll
      for (int r = 0; r < 4; ++r) q[r] = V[r][i];
    }
  const double qn = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] /= qn;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z);
  R[1] = 2.0 * (x * y - w * z);
  R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);
  R[4] = 1.0 - 2.0 * (x * x + z * z);
  R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);
  R[7] = 2.0 * (y * z + w * x);
  R[8] = 1.0 - 2.0 * (x * x + y * y);
#pragma unroll
  for (int a = 0; a < 3; ++a) t[a] = cg[a] - ((R[3 * a] * cs[0] + R[3 * a + 1] * cs[1]) + R[3 * a + 2] * cs[2]);
}

__device__ __forceinline__ double residual2(const double (&R)[9], const double (&t)[3], const float* s, const float* g) {
  double d2 = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double p = ((R[3 * a] * (double)s[0] + R[3 * a + 1] * (double)s[1]) + R[3 * a + 2] * (double)s[2]) + t[a];
    const double d = p - (double)g[a];
    d2 += d * d;
  }
  return d2;
}

struct RansacArgs {
  const int* M_ptr;
  const int* mem_off;
  const long long* mem_idx;
  const pcdb_vote* votes;
  const int* mseg;          // maximum -> (cloud, class) group
  const unsigned* seg_key;  // group -> cloud * n_classes + class
  const int* max_off;       // group -> its first maximum
  const float* cls_mul;     // per class multiplier of the threshold (null: Fixed)
  int n_classes, min_votes;
  float thr;
  int* keep;                // per member entry: 1 = inlier vote of a surviving maximum
};

__global__ void __launch_bounds__(kRansacThreads) k_ransac(RansacArgs a) {
  const int m = blockIdx.x;
  if (m >= *a.M_ptr) return;
  const int o0 = a.mem_off[m], n = a.mem_off[m + 1] - o0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto drop_all = [&]() {
    for (int i = tid; i < n; i += kRansacThreads) a.keep[o0 + i] = 0;
  };
  if (n < a.min_votes || n < 3) {  // voting.cpp:373 (not tried) / fewer than one sample: the maximum goes
    drop_all();
    return;
  }
  const int grp = a.mseg[m];
  const unsigned cls = a.seg_key[grp] % (unsigned)a.n_classes;
  const unsigned key = cls * 65536u + (unsigned)(m - a.max_off[grp]);
  float thr = a.thr;
  if (a.cls_mul) thr *= a.cls_mul[cls];
  const double thr2 = (double)thr * (double)thr;
  auto src = [&](int i) -> const float* { return a.votes[a.mem_idx[o0 + i]].keypoint_training; };
  auto tgt = [&](int i) -> const float* { return a.votes[a.mem_idx[o0 + i]].keypoint; };

  // ---- sample distance threshold: 9 moment sums, 256 strided partial sums combined in thread order (the oracle's tree)
  __shared__ double s_part[kRansacThreads][9];
  __shared__ double s_sdt;
  {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += kRansacThreads) {
      const float* s = src(i);
      const double x = s[0], y = s[1], z = s[2];
      acc[0] += x; acc[1] += y; acc[2] += z;
      acc[3] += x * x; acc[4] += x * y; acc[5] += x * z; acc[6] += y * y; acc[7] += y * z; acc[8] += z * z;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) s_part[tid][j] = acc[j];
  }
  __syncthreads();
  if (tid == 0) {
    double tot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < kRansacThreads; ++t)
#pragma unroll
      for (int j = 0; j < 9; ++j) tot[j] += s_part[t][j];
    const double inv_n = 1.0 / (double)n;
#pragma unroll
    for (int j = 0; j < 9; ++j) tot[j] *= inv_n;
    double C[3][3] = {{tot[3] - tot[0] * tot[0], tot[4] - tot[0] * tot[1], tot[5] - tot[0] * tot[2]},
                      {0, tot[6] - tot[1] * tot[1], tot[7] - tot[1] * tot[2]},
                      {0, 0, tot[8] - tot[2] * tot[2]}};
    C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
    double V3[3][3];
    jacobi_sym<3>(C, V3);
    double sdt = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) sdt += sqrt(C[j][j] > 0 ? C[j][j] : 0.0);
    sdt /= 3.0;
    s_sdt = sdt * sdt;
  }
  __syncthreads();
  const double sdt = s_sdt;

  // ---- hypotheses, 8 iterations at a time
  __shared__ int s_cnt[kRansacWarps], s_smp[kRansacWarps][3];
  __shared__ int s_it, s_best, s_bs[3], s_done;
  __shared__ double s_k;
  if (tid == 0) {
    s_it = 0;
    s_best = -1;
    s_done = 0;
    s_k = 1e300;
  }
  __syncthreads();
  const double inv_n = 1.0 / (double)n;
  auto d2pts = [](const float* p, const float* q) {
    const double dx = (double)p[0] - (double)q[0], dy = (double)p[1] - (double)q[1], dz = (double)p[2] - (double)q[2];
    return (dx * dx + dy * dy) + dz * dz;
  };
  auto load3 = [&](const int (&smp)[3], double (&S)[3][3], double (&G)[3][3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* s = src(smp[j]);
      const float* g = tgt(smp[j]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        S[j][c] = s[c];
        G[j][c] = g[c];
      }
    }
  };
  while (!s_done) {
    const int it = s_it + warp;
    int smp[3] = {-1, -1, -1};
    for (unsigned at = 0; at < 1000u; ++at) {  // max_sample_checks_
      int i0 = (int)(ransac_hash(key, (unsigned)it, at, 0) % (unsigned)n);
      int i1 = (int)(ransac_hash(key, (unsigned)it, at, 1) % (unsigned)(n - 1));
      if (i1 >= i0) ++i1;
      int i2 = (int)(ransac_hash(key, (unsigned)it, at, 2) % (unsigned)(n - 2));
      const int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
      if (i2 >= lo) ++i2;
      if (i2 >= hi) ++i2;
      const float *pa = src(i0), *pb = src(i1), *pc = src(i2);
      if (d2pts(pb, pa) > sdt && d2pts(pc, pa) > sdt && d2pts(pc, pb) > sdt) {
        smp[0] = i0; smp[1] = i1; smp[2] = i2;
        break;
      }
    }
    int cnt = 0;
    if (smp[0] >= 0) {  // warp-uniform
      double S[3][3], G[3][3], R[9], t[3];
      load3(smp, S, G);
      rigid_from_three(S, G, R, t);
      for (int i = lane; i < n; i += 32)
        if (residual2(R, t, src(i), tgt(i)) < thr2) ++cnt;
      cnt = warp_sum(cnt);
    }
    if (lane == 0) {
      s_cnt[warp] = cnt;
      s_smp[warp][0] = smp[0];
      s_smp[warp][1] = smp[1];
      s_smp[warp][2] = smp[2];
    }
    __syncthreads();
    if (tid == 0) {  // RandomSampleConsensus::computeModel, one iteration after the other
      int it_seq = s_it, best = s_best;
      double k = s_k;
      bool done = false;
      for (int j = 0; j < kRansacWarps && !done; ++j) {
        if (!((double)it_seq < k)) { done = true; break; }
        if (s_smp[j][0] < 0) { done = true; break; }  // no good sample
        if (s_cnt[j] > best) {
          best = s_cnt[j];
          s_bs[0] = s_smp[j][0]; s_bs[1] = s_smp[j][1]; s_bs[2] = s_smp[j][2];
          const double w = (double)best * inv_n;
          double p_no = 1.0 - (w * w) * w;
          p_no = p_no > 2.220446049250313e-16 ? p_no : 2.220446049250313e-16;
          p_no = p_no < 1.0 - 2.220446049250313e-16 ? p_no : 1.0 - 2.220446049250313e-16;
          k = -4.605170185988091 / log(p_no);
        }
        ++it_seq;
        if (it_seq > 10000) done = true;
      }
      if (!done && !((double)it_seq < k)) done = true;
      s_it = it_seq;
      s_best = best;
      s_k = k;
      s_done = done ? 1 : 0;
    }
    __syncthreads();
  }
  if (s_best < 3) {  // no model, or fewer than 3 inliers: best_transformation_ stays the identity
    drop_all();
    return;
  }
  double S[3][3], G[3][3], R[9], t[3];
  const int bs[3] = {s_bs[0], s_bs[1], s_bs[2]};
  load3(bs, S, G);
  rigid_from_three(S, G, R, t);
  // Eigen::Matrix4f::isIdentity(1e-4): diagonal isApprox(1), everything else isMuchSmallerThan(1)
  bool identity = true;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (float)R[3 * r + c];
      if (r == c) {
        const float av = fabsf(v);
        if (!(fabsf(v - 1.0f) <= (av < 1.0f ? av : 1.0f) * 1e-4f)) identity = false;
      } else if (!(fabsf(v) <= 1e-4f)) {
        identity = false;
      }
    }
    if (!(fabsf((float)t[r]) <= 1e-4f)) identity = false;
  }
  if (identity) {
    drop_all();
    return;
  }
  for (int i = tid; i < n; i += kRansacThreads) a.keep[o0 + i] = residual2(R, t, src(i), tgt(i)) < thr2 ? 1 : 0;
}

// per maximum: number of kept member entries (one warp per maximum)
__global__ void k_ransac_count(const int* __restrict__ M_ptr, const int* __restrict__ mem_off,
                               const int* __restrict__ keep, int* cnt, int M_cap) {
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (m > M_cap) return;
  if (m >= *M_ptr) {
    if (lane == 0) cnt[m] = 0;
    return;
  }
  int c = 0;
  for (int i = mem_off[m] + lane; i < mem_off[m + 1]; i += 32) c += keep[i];
  c = warp_sum(c);
  if (lane == 0) cnt[m] = c;
}

// stable compaction of the member lists (one warp per maximum, ballot order = member order)
__global__ void k_ransac_pack(const int* __restrict__ M_ptr, const int* __restrict__ mem_off,
                              const int* __restrict__ new_off, const int* __restrict__ keep,
                              const long long* __restrict__ mem_idx, const float* __restrict__ mem_w,
                              long long* mem_idx2, float* mem_w2) {
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (m >= *M_ptr) return;
  const int o0 = mem_off[m], o1 = mem_off[m + 1];
  int d = new_off[m];
  for (int base = o0; base < o1; base += 32) {
    const int i = base + lane;
    const bool k = i < o1 && keep[i] != 0;
    const unsigned mask = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int o = d + __popc(mask & ((1u << lane) - 1u));
      mem_idx2[o] = mem_idx[i];
      mem_w2[o] = mem_w[i];
    }
    d += __popc(mask);
  }
}

}  // namespace

// Filters the member lists in place (ws.mem_idx / ws.mem_w / ws.mem_off are replaced by the compacted ones).
int stage_ransac_filter(pcdb_ctx* ctx, int64_t hM, const int* M_ptr, int64_t* hMem_io) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const pcdb_params& p = ctx->prm;
  if (hM <= 0 || *hMem_io <= 0) return PCDB_OK;
  if (p.ransac_refine_model)
    return ctx->fail(PCDB_E_UNSUPPORTED, "Voting.RansacRefineModel = true is not built (voting.cpp:363)");
  if (!(p.ransac_inlier_threshold > 0.f)) return ctx->fail(PCDB_E_INVALID, "Voting.RansacInlierThreshold must be positive");
  const int C = std::max(1, ctx->cb.n_classes);
  RansacArgs a;
  a.cls_mul = nullptr;
  if (p.ransac_threshold_type != PCDB_RANSAC_FIXED) {
    const std::vector<float>& dims =
        p.ransac_threshold_type == PCDB_RANSAC_OBJECT_RADIUS ? ctx->class_dim_first : ctx->class_dim_second;
    if ((int)dims.size() < C)
      return ctx->fail(PCDB_E_STATE, "Voting.RansacInlierThresholdType needs the learned class dimensions: call "
                                     "pcdb_set_class_dimensions for all %d classes", C);
    PCDB_CUDA(w.rs_mul.ensure(sizeof(float) * (size_t)C));
    PCDB_CUDA(cudaMemcpyAsync(w.rs_mul.p, dims.data(), sizeof(float) * (size_t)C, cudaMemcpyHostToDevice, st));
    PCDB_CUDA(cudaStreamSynchronize(st));  // dims may be reassigned by the caller
    a.cls_mul = w.rs_mul.as<float>();
  }
  const int64_t hMem = *hMem_io;
  PCDB_CUDA(w.rs_keep.ensure(sizeof(int) * (size_t)(hMem + 1)));
  PCDB_CUDA(w.rs_cnt.ensure(sizeof(int) * (size_t)(hM + 2)));
  PCDB_CUDA(w.mem_off2.ensure(sizeof(int) * (size_t)(hM + 2)));
  PCDB_CUDA(w.mem_idx2.ensure(sizeof(long long) * ((size_t)hMem + 1)));
  PCDB_CUDA(w.mem_w2.ensure(sizeof(float) * ((size_t)hMem + 1)));
  a.M_ptr = M_ptr;
  a.mem_off = w.mem_off.as<int>();
  a.mem_idx = w.mem_idx.as<long long>();
  a.votes = w.votes.as<pcdb_vote>();
  a.mseg = w.mseg.as<int>();
  a.seg_key = w.seg2_key.as<unsigned>();
  a.max_off = w.max_off.as<int>();
  a.n_classes = C;
  a.min_votes = p.min_votes_threshold;
  a.thr = p.ransac_inlier_threshold;
  a.keep = w.rs_keep.as<int>();
  k_ransac<<<(unsigned)hM, kRansacThreads, 0, st>>>(a);
  PCDB_LAUNCH_CHECK();
  k_ransac_count<<<cdiv((hM + 1) * 32, 128), 128, 0, st>>>(M_ptr, w.mem_off.as<int>(), w.rs_keep.as<int>(),
                                                           w.rs_cnt.as<int>(), (int)hM);
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.rs_cnt.as<int>(), w.mem_off2.as<int>(), hM + 1));
  k_ransac_pack<<<cdiv(hM * 32, 128), 128, 0, st>>>(M_ptr, w.mem_off.as<int>(), w.mem_off2.as<int>(),
                                                    w.rs_keep.as<int>(), w.mem_idx.as<long long>(),
                                                    w.mem_w.as<float>(), w.mem_idx2.as<long long>(),
                                                    w.mem_w2.as<float>());
  PCDB_LAUNCH_CHECK();
  int total = 0;
  PCDB_TRY(pcdb_read_small(ctx, &total, w.mem_off2.as<int>() + hM, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  std::swap(w.mem_idx, w.mem_idx2);
  std::swap(w.mem_w, w.mem_w2);
  std::swap(w.mem_off, w.mem_off2);
  *hMem_io = total;
  return PCDB_OK;
}
