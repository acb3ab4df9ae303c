// api.cu — the C-ABI of include/pcdb200.h: context lifecycle, model upload, the stage-level entry points (one per
// reference hook) and the fused batch path pcdb_classify_batch (ImplicitShapeModel::detect for B clouds at once).
// Host code only orchestrates: every computation is a kernel of this library (no CPU fallback anywhere).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "stages.h"

#include <chrono>
#include <cstdlib>

void pcdb_trace_point(pcdb_ctx* ctx, const char* name) {
  static const bool on = [] { const char* e = getenv("PCDB_TRACE"); return e && e[0] == '1'; }();
  if (!on) return;
  static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  const auto t_call = std::chrono::steady_clock::now();
  cudaStreamSynchronize(ctx->stream);
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[pcdb trace] %-28s host %8.3f ms, drain %8.3f ms\n", name,
          std::chrono::duration<double, std::milli>(t_call - last).count(),
          std::chrono::duration<double, std::milli>(now - t_call).count());
  last = now;
}

int pcdb_read_small(pcdb_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
  if (bytes == 0) return PCDB_OK;
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (!ctx->pinned || ctx->pinned_used + need > ctx->pinned_cap) {  // no room: plain (blocking) copy
    PCDB_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PCDB_OK;
  }
  PCDB_CUDA(cudaMemcpyAsync(static_cast<char*>(ctx->pinned) + ctx->pinned_used, dev_src, bytes, cudaMemcpyDeviceToHost,
                            ctx->stream));
  ctx->pending_reads.push_back({host_dst, ctx->pinned_used, bytes});
  ctx->pinned_used += need;
  return PCDB_OK;
}

int pcdb_sync_reads(pcdb_ctx* ctx) {
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  for (const auto& r : ctx->pending_reads) std::memcpy(r.dst, static_cast<char*>(ctx->pinned) + r.off, r.bytes);
  ctx->pending_reads.clear();
  ctx->pinned_used = 0;
  if (e != cudaSuccess) return ctx->fail(PCDB_E_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
  return PCDB_OK;
}

namespace {

thread_local std::string g_create_err;

int upload(pcdb_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
  PCDB_CUDA(buf.ensure(bytes + 16));
  if (bytes) PCDB_CUDA(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PCDB_OK;
}
int download(pcdb_ctx* ctx, void* host, const void* dev, size_t bytes) {
  if (bytes) PCDB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return PCDB_OK;
}

int check_offsets(pcdb_ctx* ctx, const int64_t* off, int B, const char* what) {
  if (B < 0 || !off) return ctx->fail(PCDB_E_INVALID, "%s: bad batch", what);
  if (B > 65535) return ctx->fail(PCDB_E_INVALID, "%s: at most 65535 clouds per batch (16-bit cloud field of the sort keys)", what);
  if (off[0] != 0) return ctx->fail(PCDB_E_INVALID, "%s: offsets must start at 0", what);
  for (int b = 0; b < B; ++b)
    if (off[b + 1] < off[b]) return ctx->fail(PCDB_E_INVALID, "%s: offsets must be non-decreasing", what);
  if (off[B] > 0x7fffff00ll) return ctx->fail(PCDB_E_INVALID, "%s: more than 2^31 elements in one batch", what);
  return PCDB_OK;
}

int check_params(pcdb_ctx* ctx, const pcdb_params& p) {
  if (p.feature_type != PCDB_FEATURE_SHOT && p.feature_type != PCDB_FEATURE_CSHOT)
    return ctx->fail(PCDB_E_INVALID, "invalid Features.Type %d", p.feature_type);
  if (!(p.feature_radius > 0) || !(p.lrf_radius > 0)) return ctx->fail(PCDB_E_INVALID, "radii must be positive");
  if (!(p.leaf_size > 0)) return ctx->fail(PCDB_E_INVALID, "Keypoints.LeafSize must be positive");
  if (p.distance_type != PCDB_DIST_EUCLIDEAN && p.distance_type != PCDB_DIST_CHISQUARED)
    return ctx->fail(PCDB_E_INVALID, "invalid DistanceType %d", p.distance_type);
  if (p.knn_k < 1 || p.knn_k > PCDB_MAX_K) return ctx->fail(PCDB_E_INVALID, "K must be in 1..%d", PCDB_MAX_K);
  if (p.ms_kernel != PCDB_KERNEL_GAUSSIAN && p.ms_kernel != PCDB_KERNEL_UNIFORM)
    return ctx->fail(PCDB_E_INVALID, "invalid Voting.Kernel %d", p.ms_kernel);
  if (p.maxima_suppression != PCDB_SUPPRESS_AVERAGE && p.maxima_suppression != PCDB_SUPPRESS_SUPPRESS)
    return ctx->fail(PCDB_E_INVALID, "invalid Voting.MaximaSuppression %d", p.maxima_suppression);
  if (!(p.bandwidth > 0)) return ctx->fail(PCDB_E_INVALID, "Voting.Bandwidth must be positive");
  if (p.radius_type < PCDB_RADIUS_CONFIG || p.radius_type > PCDB_RADIUS_SECOND_DIM)
    return ctx->fail(PCDB_E_INVALID, "invalid Voting.BinOrBandwidthType %d", p.radius_type);
  if (p.radius_type != PCDB_RADIUS_CONFIG && !(p.radius_factor > 0))
    return ctx->fail(PCDB_E_INVALID, "Voting.BinOrBandwidthFactor must be positive");
  if (p.single_object_max_type < PCDB_SOMAX_DEFAULT || p.single_object_max_type > PCDB_SOMAX_MODEL_RADIUS)
    return ctx->fail(PCDB_E_INVALID, "invalid Voting.SingleObjectMaxType %d", p.single_object_max_type);
  if (p.max_filter_type < PCDB_MAXFILTER_NONE || p.max_filter_type > PCDB_MAXFILTER_MERGE)
    return ctx->fail(PCDB_E_INVALID, "invalid Voting.MaxFilterType %d", p.max_filter_type);
  if (p.ransac_vote_filtering) {
    if (p.ransac_threshold_type < PCDB_RANSAC_FIXED || p.ransac_threshold_type > PCDB_RANSAC_BBOX_MEDIAN)
      return ctx->fail(PCDB_E_INVALID, "invalid Voting.RansacInlierThresholdType %d", p.ransac_threshold_type);
    if (!(p.ransac_inlier_threshold > 0.f)) return ctx->fail(PCDB_E_INVALID, "Voting.RansacInlierThreshold must be positive");
    if (p.ransac_refine_model)
      return ctx->fail(PCDB_E_UNSUPPORTED, "Voting.RansacRefineModel = true is not built (voting.cpp:363)");
  }
  return PCDB_OK;
}

__global__ void k_pack_kp(const float* __restrict__ xyz, const unsigned* __restrict__ rgb,
                          const long long* __restrict__ off, int B, long long Q, float4* kp4, int* kp_cloud) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= Q) return;
  kp4[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __uint_as_float(rgb ? rgb[i] : 0u));
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  kp_cloud[i] = lo;
}

__global__ void k_unpack_kp(const float4* __restrict__ kp4, long long Q, float* xyz, unsigned* rgb) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= Q) return;
  float4 k = kp4[i];
  xyz[3 * i] = k.x;
  xyz[3 * i + 1] = k.y;
  xyz[3 * i + 2] = k.z;
  if (rgb) rgb[i] = __float_as_uint(k.w);
}

// Features::operator() drop of non-finite frames (features.cpp:64-76) + removeNaNFeatures
// (implicit_shape_model.cpp:1276-1308): one warp per keypoint
__global__ void k_feat_valid(const float* __restrict__ lrf, const float* __restrict__ desc, long long Q, int D,
                             int* valid) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q > Q) return;
  if (q == Q) {
    if (lane == 0) valid[q] = 0;
    return;
  }
  bool ok = isfinite(lrf[q * 9]) && isfinite(lrf[q * 9 + 3]) && isfinite(lrf[q * 9 + 6]);
  bool nan = false;
  if (ok)
    for (int j = lane; j < D; j += 32) nan |= isnan(desc[q * D + j]);
  nan = __any_sync(0xffffffffu, nan);
  if (lane == 0) valid[q] = (ok && !nan) ? 1 : 0;
}

__global__ void k_feat_compact(const float4* __restrict__ kp4, const int* __restrict__ kp_cloud,
                               const float* __restrict__ lrf, const float* __restrict__ desc, long long Q, int D,
                               const int* __restrict__ valid, const int* __restrict__ pos, float* fxyz, float* flrf,
                               float* fdesc, int* fcloud, int* fkp) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q || !valid[q]) return;
  const long long o = pos[q];
  if (lane == 0) {
    float4 k = kp4[q];
    fxyz[3 * o] = k.x;
    fxyz[3 * o + 1] = k.y;
    fxyz[3 * o + 2] = k.z;
    fcloud[o] = kp_cloud[q];
    fkp[o] = (int)q;  // the keypoint this feature came from (keypoint-sharded scenes restore the global vote order with it)
  }
  if (lane < 9) flrf[9 * o + lane] = lrf[9 * q + lane];
  for (int j = lane * 4; j < D; j += 128)
    *reinterpret_cast<float4*>(fdesc + o * D + j) = *reinterpret_cast<const float4*>(desc + q * D + j);
}

__global__ void k_offsets_from_pos(const long long* __restrict__ in_off, int B, const int* __restrict__ pos,
                                   long long* out_off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  out_off[b] = pos[in_off[b]];
}

__global__ void k_cloud_ids(const long long* __restrict__ off, int B, long long n, int* cloud) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  cloud[i] = lo;
}

// ---- radius-neighbour listing (parity/debug entry) -------------------------------------------------------------
__global__ void k_nbr_count(const float4* __restrict__ surfS, const unsigned long long* __restrict__ skeys,
                            const long long* __restrict__ surf_off, const float4* __restrict__ kp4,
                            const int* __restrict__ kp_cloud, const CloudInfo* __restrict__ ci, float inv_cell,
                            long long Q, float r2, int* cnt) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q > Q) return;
  if (q == Q) {
    if (lane == 0) cnt[q] = 0;
    return;
  }
  float4 k = kp4[q];
  unsigned cloud = (unsigned)kp_cloud[q];
  CloudInfo c = ci[cloud];
  int cx = grid_coord(k.x, c.mn[0], inv_cell), cy = grid_coord(k.y, c.mn[1], inv_cell),
      cz = grid_coord(k.z, c.mn[2], inv_cell);
  long long lo = surf_off[cloud], hi = surf_off[cloud + 1];
  int n = 0;
  for (int r = 0; r < 9; ++r) {
    int y = cy + r % 3 - 1, z = cz + r / 3 - 1;
    if (y < 0 || y > 65535 || z < 0 || z > 65535) continue;
    long long beg = lower_bound_u64(skeys, lo, hi, grid_key(cloud, max(cx - 1, 0), y, z));
    long long end = lower_bound_u64(skeys, beg, hi, grid_key(cloud, min(cx + 1, 65535), y, z) + 1ull);
    for (long long e = beg + lane; e < end; e += 32) {
      float4 p = surfS[e];
      if (sqdist3_rn(k.x, k.y, k.z, p.x, p.y, p.z) < r2) ++n;
    }
  }
  n = warp_sum(n);
  if (lane == 0) cnt[q] = n;
}

__global__ void k_nbr_fill(const float4* __restrict__ surfS, const float4* __restrict__ snrmS,
                           const unsigned long long* __restrict__ skeys, const long long* __restrict__ surf_off,
                           const float4* __restrict__ kp4, const int* __restrict__ kp_cloud,
                           const CloudInfo* __restrict__ ci, float inv_cell, long long Q, float r2,
                           const int* __restrict__ off, unsigned long long* keys) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  float4 k = kp4[q];
  unsigned cloud = (unsigned)kp_cloud[q];
  CloudInfo c = ci[cloud];
  int cx = grid_coord(k.x, c.mn[0], inv_cell), cy = grid_coord(k.y, c.mn[1], inv_cell),
      cz = grid_coord(k.z, c.mn[2], inv_cell);
  long long lo = surf_off[cloud], hi = surf_off[cloud + 1];
  int o = off[q];
  for (int r = 0; r < 9; ++r) {
    int y = cy + r % 3 - 1, z = cz + r / 3 - 1;
    if (y < 0 || y > 65535 || z < 0 || z > 65535) continue;
    long long beg = lower_bound_u64(skeys, lo, hi, grid_key(cloud, max(cx - 1, 0), y, z));
    long long end = lower_bound_u64(skeys, beg, hi, grid_key(cloud, min(cx + 1, 65535), y, z) + 1ull);
    for (long long base = beg; base < end; base += 32) {
      long long e = base + lane;
      bool in = false;
      float d2 = 0.f;
      if (e < end) {
        float4 p = surfS[e];
        d2 = sqdist3_rn(k.x, k.y, k.z, p.x, p.y, p.z);
        in = d2 < r2;
      }
      unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) {
        unsigned idx = (unsigned)__float_as_int(snrmS[e].w);
        keys[o + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)__float_as_uint(d2) << 32) | idx;
      }
      o += __popc(m);
    }
  }
}

__global__ void k_nbr_unpack(const unsigned long long* __restrict__ keys, long long n, int* idx, float* d2) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx[i] = (int)(keys[i] & 0xffffffffull);
  d2[i] = __uint_as_float((unsigned)(keys[i] >> 32));
}

__global__ void k_merge_topk(const int* __restrict__ cand_idx, const float* __restrict__ cand_dist, int S,
                             long long Q, int k, int* idx_out, float* dist_out) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float bd[PCDB_MAX_K];
  int bi[PCDB_MAX_K];
  int cnt = 0;
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < k; ++j) {
      long long o = ((long long)s * Q + q) * k + j;
      int idx = cand_idx[o];
      if (idx < 0) continue;
      float d = cand_dist[o];
      if (cnt == k && !(d < bd[k - 1] || (d == bd[k - 1] && idx < bi[k - 1]))) continue;
      int p = (cnt < k) ? cnt++ : k - 1;
      while (p > 0 && (d < bd[p - 1] || (d == bd[p - 1] && idx < bi[p - 1]))) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
      }
      bd[p] = d;
      bi[p] = idx;
    }
  for (int j = 0; j < k; ++j) {
    idx_out[q * k + j] = j < cnt ? bi[j] : -1;
    dist_out[q * k + j] = j < cnt ? bd[j] : __int_as_float(0x7fc00000);
  }
}

// uploads explicit surface clouds + keypoints and builds the search grid for `radius`
int prepare_explicit(pcdb_ctx* ctx, const float* surf_xyz, const float* surf_normals, const uint32_t* surf_rgb,
                     const int64_t* surf_off, const float* kp_xyz, const uint32_t* kp_rgb, const int64_t* kp_off,
                     int B, double radius, bool color, int64_t* n_surf_out, int64_t* Q_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  PCDB_TRY(check_offsets(ctx, surf_off, B, "surf_off"));
  PCDB_TRY(check_offsets(ctx, kp_off, B, "kp_off"));
  const int64_t P = surf_off[B], Q = kp_off[B];
  PCDB_TRY(upload(ctx, w.in_xyz, surf_xyz, sizeof(float) * 3 * P));
  if (surf_normals) PCDB_TRY(upload(ctx, w.in_nrm, surf_normals, sizeof(float) * 3 * P));
  if (surf_rgb) PCDB_TRY(upload(ctx, w.in_rgb, surf_rgb, sizeof(uint32_t) * P));
  PCDB_TRY(upload(ctx, w.cloud_off, surf_off, sizeof(int64_t) * (B + 1)));
  PCDB_TRY(stage_compact(ctx, B, P, false /* keep NaN-normal points: the caller passes pointsWithoutNaN */,
                         surf_rgb != nullptr));
  // normals (possibly NaN) ride along in snrm4.xyz: re-scatter them when given
  std::vector<int64_t> h_off(B + 1);
  PCDB_TRY(download(ctx, h_off.data(), w.surf_off.p, sizeof(int64_t) * (B + 1)));
  PCDB_CUDA(cudaStreamSynchronize(st));
  if (h_off[B] != P) return ctx->fail(PCDB_E_INVALID, "surface clouds must not contain non-finite points");
  if (surf_normals) {
    // all points kept => compacted index == input index: overwrite the normal part, keep the local index in .w
    PCDB_CUDA(cudaMemcpy2DAsync(w.snrm4.p, sizeof(float4), w.in_nrm.p, sizeof(float) * 3, sizeof(float) * 3, P,
                                cudaMemcpyDeviceToDevice, st));
  }
  PCDB_CUDA(w.kp_in.ensure(sizeof(float) * 3 * Q + 16));
  PCDB_CUDA(w.kp4.ensure(sizeof(float4) * (Q + 1)));
  PCDB_CUDA(w.kp_cloud.ensure(sizeof(int) * (Q + 1)));
  PCDB_TRY(upload(ctx, w.kp_in, kp_xyz, sizeof(float) * 3 * Q));
  PCDB_TRY(upload(ctx, w.kp_off, kp_off, sizeof(int64_t) * (B + 1)));
  DevBuf& rgbbuf = w.feat_valid;  // scratch for the keypoint colours
  if (kp_rgb) PCDB_TRY(upload(ctx, rgbbuf, kp_rgb, sizeof(uint32_t) * Q));
  if (Q > 0) {
    k_pack_kp<<<cdiv(Q, 256), 256, 0, st>>>(w.kp_in.as<float>(), kp_rgb ? rgbbuf.as<unsigned>() : nullptr,
                                            w.kp_off.as<long long>(), B, Q, w.kp4.as<float4>(), w.kp_cloud.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  PCDB_TRY(stage_cloud_setup(ctx, B, P, w.kp4.as<float4>(), w.kp_cloud.as<int>(), Q, 0.f, radius));
  PCDB_TRY(stage_grid(ctx, B, P, Q, color));
  int err[4] = {0, 0, 0, 0};
  PCDB_TRY(download(ctx, err, w.err_flag.p, sizeof(err)));
  PCDB_CUDA(cudaStreamSynchronize(st));
  if (err[0] & 2) return ctx->fail(PCDB_E_INVALID, "search radius too small for the cloud extent (grid > 65535 cells per axis)");
  *n_surf_out = P;
  *Q_out = Q;
  return PCDB_OK;
}

// kNN dispatch: tcgen05 GEMM + exact re-rank on large codebooks (squared L2 directly, chi^2 through the Hellinger
// sandwich), exact scan otherwise
int run_knn(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, int mode, bool use_ratio,
            float ratio_thr) {
  return pcdb_run_knn(ctx, queries_d, Q, k, dist_type, mode, use_ratio, ratio_thr);
}

// this rank's queries against the codebook, whichever way it is laid out over the GPUs
int activate(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, int mode, bool use_ratio,
             float ratio_thr) {
  if (comm_codebook_sharded(ctx)) return stage_knn_sharded(ctx, queries_d, Q, k, dist_type, mode, use_ratio, ratio_thr);
  if (Q == 0) return PCDB_OK;
  return pcdb_run_knn(ctx, queries_d, Q, k, dist_type, mode, use_ratio, ratio_thr);
}

}  // namespace

int pcdb_run_knn(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, int mode, bool use_ratio,
                 float ratio_thr) {
  const int fam = dist_type == PCDB_DIST_CHISQUARED ? 1 : 0;
  const bool big = ctx->cb.N >= 8192 && ctx->cb.N > k + 1;
  if ((mode == PCDB_KNN_GEMM || (mode == PCDB_KNN_AUTO && big)) && !ctx->cb.gemm_tried[fam])
    PCDB_TRY(gemm_prepare_codebook(ctx, dist_type));  // the operand copy of this distance family, built on first use
  bool gemm = false;
  if (mode == PCDB_KNN_GEMM) {
    if (!gemm_supported(ctx, dist_type) || ctx->cb.N <= k + 1)
      return ctx->fail(PCDB_E_UNSUPPORTED, "GEMM activation unavailable for this codebook");
    gemm = true;
  } else if (mode == PCDB_KNN_AUTO) {
    gemm = gemm_supported(ctx, dist_type) && big;
  }
  ctx->stats.knn_queries += Q;
  if (gemm) return stage_knn_gemm(ctx, queries_d, Q, k, dist_type, use_ratio, ratio_thr);
  return stage_knn_scan(ctx, queries_d, Q, k, dist_type, use_ratio, ratio_thr);
}

namespace {

// keypoints -> LRF -> SHOT/CSHOT -> drop invalid, all on the device; leaves feat_* in the workspace
int features_pipeline(pcdb_ctx* ctx, int B, int64_t P, bool has_rgb, int64_t* F_out, int64_t* Q_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const pcdb_params& p = ctx->prm;
  const bool color = p.feature_type == PCDB_FEATURE_CSHOT;
  const int D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  *F_out = 0;
  *Q_out = 0;
  pcdb_trace_point(ctx, "features: begin");
  PCDB_TRY(stage_compact(ctx, B, P, true, has_rgb));
  pcdb_trace_point(ctx, "features: compact");
  int totals[2] = {0, 0};
  PCDB_TRY(pcdb_read_small(ctx, &totals[0], w.pos_pt.as<int>() + P, sizeof(int)));
  PCDB_TRY(pcdb_read_small(ctx, &totals[1], w.pos_sf.as<int>() + P, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  const int64_t n_pts = totals[0], n_surf = totals[1];
  ctx->stats.n_points += n_pts;
  PCDB_TRY(stage_cloud_setup(ctx, B, n_pts, nullptr, nullptr, 0, p.leaf_size,
                             std::max(p.feature_radius, p.lrf_radius)));
  int64_t Q = 0;
  PCDB_TRY(stage_voxel_keypoints(ctx, B, n_pts, p.leaf_size, &Q));
  int err[4] = {0, 0, 0, 0};
  PCDB_TRY(pcdb_read_small(ctx, err, w.err_flag.p, sizeof(err)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  pcdb_trace_point(ctx, "features: voxel keypoints");
  if (err[0] & 1)
    return ctx->fail(PCDB_E_INVALID, "Keypoints.LeafSize too small for the cloud extent (pcl::VoxelGrid index overflow)");
  if (err[0] & 2) return ctx->fail(PCDB_E_INVALID, "Features radius too small for the cloud extent");
  if (comm_keypoints_sharded(ctx)) {  // one scene on several GPUs: this rank keeps its contiguous slice
    if (B != 1) return ctx->fail(PCDB_E_UNSUPPORTED, "keypoint sharding handles one scene per call (B == 1)");
    PCDB_TRY(stage_slice_keypoints(ctx, B, Q, &Q));
  }
  ctx->stats.n_keypoints += Q;
  *Q_out = Q;
  PCDB_CUDA(w.feat_off.ensure(sizeof(long long) * (B + 1)));
  if (Q == 0) {
    PCDB_CUDA(cudaMemsetAsync(w.feat_off.p, 0, sizeof(long long) * (B + 1), st));
    return PCDB_OK;
  }
  PCDB_TRY(stage_grid(ctx, B, n_surf, Q, color));
  pcdb_trace_point(ctx, "features: grid");
  PCDB_CUDA(w.lrf.ensure(sizeof(float) * 9 * Q));
  PCDB_CUDA(w.desc.ensure(sizeof(float) * (size_t)D * Q));
  PCDB_TRY(stage_shot(ctx, n_surf, Q, color, p.lrf_radius, p.feature_radius, true, true, nullptr, w.lrf.as<float>(),
                      w.desc.as<float>()));
  pcdb_trace_point(ctx, "features: shot");
  PCDB_CUDA(w.feat_valid.ensure(sizeof(int) * (Q + 2)));
  PCDB_CUDA(w.feat_pos.ensure(sizeof(int) * (Q + 2)));
  k_feat_valid<<<cdiv((Q + 1) * 32, 256), 256, 0, st>>>(w.lrf.as<float>(), w.desc.as<float>(), Q, D,
                                                        w.feat_valid.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.feat_valid.as<int>(), w.feat_pos.as<int>(), Q + 1));
  int F = 0;
  PCDB_TRY(pcdb_read_small(ctx, &F, w.feat_pos.as<int>() + Q, sizeof(int)));
  unsigned long long nb[2] = {0, 0};
  PCDB_TRY(pcdb_read_small(ctx, nb, w.scalars.as<char>() + 16, sizeof(nb)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  ctx->stats.n_neighbours_lrf += (int64_t)nb[0];
  ctx->stats.n_neighbours_shot += (int64_t)nb[1];
  ctx->stats.n_features += F;
  PCDB_CUDA(w.feat_xyz.ensure(sizeof(float) * 3 * (F + 1)));
  PCDB_CUDA(w.feat_lrf.ensure(sizeof(float) * 9 * (F + 1)));
  PCDB_CUDA(w.feat_desc.ensure(sizeof(float) * (size_t)D * (F + 1)));
  PCDB_CUDA(w.feat_cloud.ensure(sizeof(int) * (F + 1)));
  PCDB_CUDA(w.feat_kp.ensure(sizeof(int) * (F + 1)));
  k_feat_compact<<<cdiv(Q * 32, 256), 256, 0, st>>>(w.kp4.as<float4>(), w.kp_cloud.as<int>(), w.lrf.as<float>(),
                                                    w.desc.as<float>(), Q, D, w.feat_valid.as<int>(),
                                                    w.feat_pos.as<int>(), w.feat_xyz.as<float>(),
                                                    w.feat_lrf.as<float>(), w.feat_desc.as<float>(),
                                                    w.feat_cloud.as<int>(), w.feat_kp.as<int>());
  PCDB_LAUNCH_CHECK();
  k_offsets_from_pos<<<cdiv(B + 1, 128), 128, 0, st>>>(w.kp_off.as<long long>(), B, w.feat_pos.as<int>(),
                                                        w.feat_off.as<long long>());
  PCDB_LAUNCH_CHECK();
  pcdb_trace_point(ctx, "features: compact features");
  *F_out = F;
  return PCDB_OK;
}

// copies the per-cloud sorted maxima out (host compaction of the padded device layout)
int fetch_maxima(pcdb_ctx* ctx, int B, int64_t M, pcdb_maximum* maxima_out, int64_t* maxima_off_out,
                 int64_t maxima_capacity, int32_t* label_out, bool count = true) {
  Workspace& w = ctx->ws;
  std::vector<int> kept(B), first(B), labels(B);
  PCDB_TRY(pcdb_read_small(ctx, kept.data(), w.max_kept.p, sizeof(int) * B));
  PCDB_TRY(pcdb_read_small(ctx, first.data(), w.max_first.p, sizeof(int) * B));
  PCDB_TRY(pcdb_read_small(ctx, labels.data(), w.labels.p, sizeof(int) * B));
  std::vector<pcdb_maximum> all((size_t)std::max<int64_t>(M, 1));
  if (maxima_out && M > 0) PCDB_TRY(pcdb_read_small(ctx, all.data(), w.max_sorted.p, sizeof(pcdb_maximum) * M));
  PCDB_TRY(pcdb_sync_reads(ctx));
  int64_t need = 0;
  for (int b = 0; b < B; ++b) {
    if (label_out) label_out[b] = labels[b];  // labels are complete even when the maxima do not fit
    need += kept[b];
  }
  ctx->last_kept = need;
  if (maxima_out && need > maxima_capacity)
    return ctx->fail(PCDB_E_CAPACITY, "maxima_capacity %lld < %lld maxima (pcdb_get_last_sizes, then pcdb_get_maxima)",
                     (long long)maxima_capacity, (long long)need);
  int64_t total = 0;
  if (maxima_off_out) maxima_off_out[0] = 0;
  for (int b = 0; b < B; ++b) {
    if (maxima_out)
      for (int i = 0; i < kept[b]; ++i) maxima_out[total + i] = all[(size_t)first[b] + i];
    total += kept[b];
    if (maxima_off_out) maxima_off_out[b + 1] = total;
  }
  if (count) ctx->stats.n_maxima += total;
  return PCDB_OK;
}


// Validates everything first, uploads, and only then commits N / D / V: a failed call leaves the context WITHOUT a
// codebook (cb.N == 0), never with stale sizes over unallocated buffers.
// words: rows [word_lo, word_lo + n_words) of the table; the vote tables (CSR over N_table rows) are complete.
int set_codebook_impl(pcdb_ctx* ctx, const float* words, int64_t n_words, int64_t N_table, int64_t word_lo, int32_t D,
                      const int64_t* vote_off, const float* vote_xyz, const float* vote_weight,
                      const uint32_t* vote_class, const uint32_t* vote_instance, const float* vote_bbox,
                      const float* vote_class_weight, const float* kp_train, const int32_t* codeword_ids,
                      const float* codeword_weight, const float* class_sigma2, int32_t n_classes, int64_t row_base) {
  Codebook_d& cb = ctx->cb;
  cb.gemm_ready[0] = cb.gemm_ready[1] = cb.gemm_tried[0] = cb.gemm_tried[1] = false;
  cb.N = 0;  // no codebook until every check and upload below has succeeded
  cb.V = 0;
  cb.N_table = 0;
  if (n_words < 0 || N_table < 0 || D <= 0 || n_classes <= 0) return ctx->fail(PCDB_E_INVALID, "bad codebook sizes");
  if (N_table > 0x7fffff00ll || row_base < 0 || row_base + N_table > 0x7fffff00ll)
    return ctx->fail(PCDB_E_INVALID, "codebook too large for 32-bit row ids");
  if (word_lo < 0 || word_lo + n_words > N_table) return ctx->fail(PCDB_E_INVALID, "word rows outside the vote table");
  if (!class_sigma2) return ctx->fail(PCDB_E_INVALID, "class_sigma2 is required");
  if (N_table > 0 && (!vote_off || !kp_train)) return ctx->fail(PCDB_E_INVALID, "vote_off and kp_train are required");
  if (n_words > 0 && !words) return ctx->fail(PCDB_E_INVALID, "words is required");
  int64_t V = 0;
  int mv = 0;
  if (N_table > 0) {
    if (vote_off[0] != 0) return ctx->fail(PCDB_E_INVALID, "vote_off must start at 0");
    for (int64_t i = 0; i < N_table; ++i) {
      const int64_t n = vote_off[i + 1] - vote_off[i];
      if (n < 0) return ctx->fail(PCDB_E_INVALID, "vote_off must be non-decreasing (row %lld)", (long long)i);
      if (n > 0x7fffffff) return ctx->fail(PCDB_E_INVALID, "too many votes on codeword %lld", (long long)i);
      mv = std::max<int64_t>(mv, n);
    }
    V = vote_off[N_table];
  }
  if (V > 0x7fffff00ll) return ctx->fail(PCDB_E_INVALID, "more than 2^31 stored votes");
  if (V > 0 && (!vote_xyz || !vote_weight || !vote_class || !vote_instance || !vote_bbox))
    return ctx->fail(PCDB_E_INVALID, "the per-vote arrays are required when the table holds votes");
  for (int64_t i = 0; i < V; ++i)
    if (vote_class[i] >= (uint32_t)n_classes)
      return ctx->fail(PCDB_E_INVALID, "vote %lld has class id %u >= n_classes %d", (long long)i, vote_class[i], n_classes);
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(upload(ctx, cb.words, words, sizeof(float) * (size_t)n_words * D));
  PCDB_TRY(upload(ctx, cb.vote_off, vote_off, sizeof(int64_t) * (N_table + 1)));
  PCDB_TRY(upload(ctx, cb.vote_xyz, vote_xyz, sizeof(float) * 3 * V));
  PCDB_TRY(upload(ctx, cb.vote_weight, vote_weight, sizeof(float) * V));
  PCDB_TRY(upload(ctx, cb.vote_class, vote_class, sizeof(uint32_t) * V));
  PCDB_TRY(upload(ctx, cb.vote_instance, vote_instance, sizeof(uint32_t) * V));
  PCDB_TRY(upload(ctx, cb.vote_bbox, vote_bbox, sizeof(float) * 7 * V));
  if (vote_class_weight)
    PCDB_TRY(upload(ctx, cb.vote_class_weight, vote_class_weight, sizeof(float) * V));
  else
    cb.vote_class_weight.release();
  PCDB_TRY(upload(ctx, cb.kp_train, kp_train, sizeof(float) * 3 * N_table));
  if (codeword_ids)
    PCDB_TRY(upload(ctx, cb.ids, codeword_ids, sizeof(int32_t) * N_table));
  else
    cb.ids.release();
  std::vector<float> ones;
  if (!codeword_weight) {
    ones.assign((size_t)std::max<int64_t>(N_table, 1), 1.0f);
    codeword_weight = ones.data();
  }
  PCDB_TRY(upload(ctx, cb.cw_weight, codeword_weight, sizeof(float) * N_table));
  PCDB_TRY(upload(ctx, cb.sigma2, class_sigma2, sizeof(float) * n_classes));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  // commit
  cb.N = n_words;
  cb.N_table = N_table;
  cb.word_lo = word_lo;
  cb.D = D;
  cb.V = V;
  cb.n_classes = n_classes;
  cb.row_base = row_base;
  cb.max_votes_per_word = mv;
  if (n_words > 0) {
    int rc = gemm_prepare_codebook(ctx, ctx->prm.distance_type);
    if (rc != PCDB_OK) {
      cb.N = 0;
      return rc;
    }
  }
  return PCDB_OK;
}

}  // namespace

extern "C" {

int pcdb_abi_version(void) { return PCDB_ABI_VERSION; }

void pcdb_default_params(pcdb_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->feature_type = PCDB_FEATURE_SHOT;
  p->feature_radius = 0.1;
  p->lrf_radius = (double)0.2f;
  p->leaf_size = 0.1f;
  p->distance_type = PCDB_DIST_EUCLIDEAN;
  p->knn_k = 1;
  p->distance_ratio_threshold = 0.95f;
  p->bandwidth = 0.2f;
  p->ms_threshold = 1e-3f;
  p->ms_max_iter = 1000;
  p->ms_kernel = PCDB_KERNEL_GAUSSIAN;
  p->maxima_suppression = PCDB_SUPPRESS_AVERAGE;
  p->min_votes_threshold = 1;
  p->best_k = -1;
  p->normal_radius = 0.05f;
  p->consistent_normals_method = 2;
  p->max_filter_type = PCDB_MAXFILTER_NONE;
  p->radius_type = PCDB_RADIUS_CONFIG;
  p->radius_factor = 1.0f;
  p->single_object_max_type = PCDB_SOMAX_DEFAULT;
  p->ransac_vote_filtering = 0;
  p->ransac_inlier_threshold = 0.1f;
  p->ransac_threshold_type = PCDB_RANSAC_FIXED;
  p->ransac_refine_model = 0;
}

int pcdb_create(pcdb_ctx** out, int device) {
  if (!out) return PCDB_E_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return PCDB_E_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return PCDB_E_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_create_err = "libpcdb200 is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor);
    return PCDB_E_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_err = "cudaSetDevice failed";
    return PCDB_E_CUDA;
  }
  pcdb_ctx* ctx = new pcdb_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  std::memset(&ctx->stats, 0, sizeof(ctx->stats));
  pcdb_default_params(&ctx->prm);
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "cudaStreamCreate failed";
    delete ctx;
    return PCDB_E_CUDA;
  }
  ctx->stream = ctx->own_stream;
  if (cudaMallocHost(&ctx->pinned, 1 << 16) == cudaSuccess) ctx->pinned_cap = 1 << 16;  // else: plain copies
  for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev_knn[i]);
  // CIELab look-up tables, built on the host with powf exactly as the reference does (features_cshot.cpp:52-71)
  std::vector<float> lut(256 + 4000);
  for (int i = 0; i < 256; i++) {
    float f = static_cast<float>(i) / 255.0f;
    lut[i] = (f > 0.04045) ? powf((f + 0.055f) / 1.055f, 2.4f) : f / 12.92f;
  }
  for (int i = 0; i < 4000; i++) {
    float f = static_cast<float>(i) / 4000.0f;
    lut[256 + i] = (f > 0.008856) ? static_cast<float>(powf(f, 0.3333f))
                                  : static_cast<float>((7.787 * f) + (16.0 / 116.0));
  }
  if (cudaMalloc(&ctx->lab_lut_d, sizeof(float) * lut.size()) != cudaSuccess ||
      cudaMemcpy(ctx->lab_lut_d, lut.data(), sizeof(float) * lut.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    g_create_err = "LUT upload failed";
    delete ctx;
    return PCDB_E_CUDA;
  }
  *out = ctx;
  return PCDB_OK;
}

void pcdb_destroy(pcdb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->ws.release();
  ctx->cb.release();
  pcdb_comm_destroy(ctx);
  if (ctx->gemm_state && ctx->gemm_state_free) ctx->gemm_state_free(ctx->gemm_state);
  if (ctx->lab_lut_d) cudaFree(ctx->lab_lut_d);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < 8; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev_knn[i]) cudaEventDestroy(ctx->ev_knn[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* pcdb_last_error(const pcdb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pcdb_set_params(pcdb_ctx* ctx, const pcdb_params* p) {
  if (!ctx || !p) return PCDB_E_INVALID;
  PCDB_TRY(check_params(ctx, *p));
  ctx->prm = *p;
  ctx->prm_set = true;
  return PCDB_OK;
}

int pcdb_set_stream(pcdb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return PCDB_E_INVALID;
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return PCDB_OK;
}

int pcdb_set_codebook(pcdb_ctx* ctx, const float* words, int64_t N, int32_t D, const int64_t* vote_off,
                      const float* vote_xyz, const float* vote_weight, const uint32_t* vote_class,
                      const uint32_t* vote_instance, const float* vote_bbox, const float* vote_class_weight,
                      const float* kp_train, const int32_t* codeword_ids, const float* codeword_weight,
                      const float* class_sigma2, int32_t n_classes, int64_t row_base) {
  if (!ctx) return PCDB_E_INVALID;
  return set_codebook_impl(ctx, words, N, N, 0, D, vote_off, vote_xyz, vote_weight, vote_class, vote_instance, vote_bbox,
                           vote_class_weight, kp_train, codeword_ids, codeword_weight, class_sigma2, n_classes, row_base);
}

int pcdb_set_codebook_sharded(pcdb_ctx* ctx, const float* words, int64_t row_lo, int64_t row_hi, int64_t N_total,
                              int32_t D, const int64_t* vote_off, const float* vote_xyz, const float* vote_weight,
                              const uint32_t* vote_class, const uint32_t* vote_instance, const float* vote_bbox,
                              const float* vote_class_weight, const float* kp_train, const int32_t* codeword_ids,
                              const float* codeword_weight, const float* class_sigma2, int32_t n_classes) {
  if (!ctx) return PCDB_E_INVALID;
  if (!comm_active(ctx)) return ctx->fail(PCDB_E_STATE, "pcdb_set_codebook_sharded before pcdb_comm_init");
  if (row_lo < 0 || row_hi < row_lo || row_hi > N_total) return ctx->fail(PCDB_E_INVALID, "bad row shard");
  return set_codebook_impl(ctx, words, row_hi - row_lo, N_total, row_lo, D, vote_off, vote_xyz, vote_weight, vote_class,
                           vote_instance, vote_bbox, vote_class_weight, kp_train, codeword_ids, codeword_weight,
                           class_sigma2, n_classes, 0);
}

int pcdb_set_class_dimensions(pcdb_ctx* ctx, const float* first_dim, const float* second_dim, int32_t n_classes) {
  if (!ctx) return PCDB_E_INVALID;
  if (n_classes < 0 || (n_classes > 0 && (!first_dim || !second_dim)))
    return ctx->fail(PCDB_E_INVALID, "bad class dimension arguments");
  ctx->class_dim_first.assign(first_dim, first_dim + n_classes);
  ctx->class_dim_second.assign(second_dim, second_dim + n_classes);
  return PCDB_OK;
}

int pcdb_voxel_keypoints(pcdb_ctx* ctx, const float* xyz, const uint32_t* rgb, const int64_t* cloud_off, int32_t B,
                         float leaf, float* kp_xyz_out, uint32_t* kp_rgb_out, int64_t* kp_off_out,
                         int64_t kp_capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, cloud_off, B, "cloud_off"));
  if (!(leaf > 0)) return ctx->fail(PCDB_E_INVALID, "leaf must be positive");
  Workspace& w = ctx->ws;
  const int64_t P = cloud_off[B];
  PCDB_TRY(upload(ctx, w.in_xyz, xyz, sizeof(float) * 3 * P));
  if (rgb) PCDB_TRY(upload(ctx, w.in_rgb, rgb, sizeof(uint32_t) * P));
  PCDB_TRY(upload(ctx, w.cloud_off, cloud_off, sizeof(int64_t) * (B + 1)));
  PCDB_TRY(stage_compact(ctx, B, P, false, rgb != nullptr));
  int n_pts = 0;
  PCDB_TRY(download(ctx, &n_pts, w.pos_pt.as<int>() + P, sizeof(int)));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  PCDB_TRY(stage_cloud_setup(ctx, B, n_pts, nullptr, nullptr, 0, leaf, 0.0));
  int64_t Q = 0;
  PCDB_TRY(stage_voxel_keypoints(ctx, B, n_pts, leaf, &Q));
  int err[4] = {0, 0, 0, 0};
  PCDB_TRY(download(ctx, err, w.err_flag.p, sizeof(err)));
  PCDB_TRY(download(ctx, kp_off_out, w.kp_off.p, sizeof(int64_t) * (B + 1)));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err[0] & 1) return ctx->fail(PCDB_E_INVALID, "leaf size too small for the cloud extent (pcl::VoxelGrid index overflow)");
  if (Q > kp_capacity) return ctx->fail(PCDB_E_CAPACITY, "kp_capacity %lld < %lld keypoints", (long long)kp_capacity, (long long)Q);
  if (Q > 0) {
    PCDB_CUDA(w.kp_in.ensure(sizeof(float) * 3 * Q));
    PCDB_CUDA(w.feat_valid.ensure(sizeof(uint32_t) * Q));
    k_unpack_kp<<<cdiv(Q, 256), 256, 0, ctx->stream>>>(w.kp4.as<float4>(), Q, w.kp_in.as<float>(),
                                                       w.feat_valid.as<unsigned>());
    PCDB_LAUNCH_CHECK();
    PCDB_TRY(download(ctx, kp_xyz_out, w.kp_in.p, sizeof(float) * 3 * Q));
    if (kp_rgb_out) PCDB_TRY(download(ctx, kp_rgb_out, w.feat_valid.p, sizeof(uint32_t) * Q));
    PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return PCDB_OK;
}

int pcdb_radius_neighbours(pcdb_ctx* ctx, const float* surf_xyz, const int64_t* surf_off, const float* kp_xyz,
                           const int64_t* kp_off, int32_t B, double radius, int64_t* nbr_off_out,
                           int32_t* nbr_idx_out, float* nbr_d2_out, int64_t capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (!(radius > 0)) return ctx->fail(PCDB_E_INVALID, "radius must be positive");
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  int64_t P = 0, Q = 0;
  PCDB_TRY(prepare_explicit(ctx, surf_xyz, nullptr, nullptr, surf_off, kp_xyz, nullptr, kp_off, B, radius, false, &P, &Q));
  nbr_off_out[0] = 0;
  if (Q == 0) return PCDB_OK;
  const float r2 = (float)(radius * radius);
  PCDB_CUDA(w.nbr_cnt.ensure(sizeof(int) * (Q + 2)));
  PCDB_CUDA(w.nbr_off.ensure(sizeof(int) * (Q + 2)));
  k_nbr_count<<<cdiv((Q + 1) * 32, 256), 256, 0, st>>>(w.surfS4.as<float4>(), w.gkeys2.as<unsigned long long>(),
                                                       w.surf_off.as<long long>(), w.kp4.as<float4>(),
                                                       w.kp_cloud.as<int>(), w.cinfo.as<CloudInfo>(),
                                                       ctx->grid_inv_cell, Q, r2, w.nbr_cnt.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.nbr_cnt.as<int>(), w.nbr_off.as<int>(), Q + 1));
  std::vector<int> h_off(Q + 1);
  PCDB_TRY(download(ctx, h_off.data(), w.nbr_off.p, sizeof(int) * (Q + 1)));
  PCDB_CUDA(cudaStreamSynchronize(st));
  const int64_t total = h_off[Q];
  for (int64_t q = 0; q <= Q; ++q) nbr_off_out[q] = h_off[q];
  if (total > capacity) return ctx->fail(PCDB_E_CAPACITY, "neighbour capacity %lld < %lld", (long long)capacity, (long long)total);
  if (total == 0) return PCDB_OK;
  PCDB_CUDA(w.nbr_key.ensure(sizeof(unsigned long long) * total));
  PCDB_CUDA(w.nbr_key2.ensure(sizeof(unsigned long long) * total));
  k_nbr_fill<<<cdiv(Q * 32, 256), 256, 0, st>>>(w.surfS4.as<float4>(), w.snrmS4.as<float4>(),
                                                w.gkeys2.as<unsigned long long>(), w.surf_off.as<long long>(),
                                                w.kp4.as<float4>(), w.kp_cloud.as<int>(), w.cinfo.as<CloudInfo>(),
                                                ctx->grid_inv_cell, Q, r2, w.nbr_off.as<int>(),
                                                w.nbr_key.as<unsigned long long>());
  PCDB_LAUNCH_CHECK();
  // (d^2 bits, index) ascending per keypoint == the kd-tree's sorted result order
  PCDB_TRY(pcdb_cub_segmented_sort_u64(ctx, w.nbr_key.as<unsigned long long>(), w.nbr_key2.as<unsigned long long>(),
                                       total, (int)Q, w.nbr_off.as<int>(), w.nbr_off.as<int>() + 1));
  PCDB_CUDA(w.merge_a.ensure(sizeof(int) * total));
  PCDB_CUDA(w.merge_b.ensure(sizeof(float) * total));
  k_nbr_unpack<<<cdiv(total, 256), 256, 0, st>>>(w.nbr_key2.as<unsigned long long>(), total, w.merge_a.as<int>(),
                                                 w.merge_b.as<float>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(download(ctx, nbr_idx_out, w.merge_a.p, sizeof(int) * total));
  PCDB_TRY(download(ctx, nbr_d2_out, w.merge_b.p, sizeof(float) * total));
  PCDB_CUDA(cudaStreamSynchronize(st));
  return PCDB_OK;
}

int pcdb_shot_lrf(pcdb_ctx* ctx, const float* surf_xyz, const int64_t* surf_off, const float* kp_xyz,
                  const int64_t* kp_off, int32_t B, double radius, float* lrf9_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (!(radius > 0)) return ctx->fail(PCDB_E_INVALID, "radius must be positive");
  Workspace& w = ctx->ws;
  int64_t P = 0, Q = 0;
  PCDB_TRY(prepare_explicit(ctx, surf_xyz, nullptr, nullptr, surf_off, kp_xyz, nullptr, kp_off, B, radius, false, &P, &Q));
  if (Q == 0) return PCDB_OK;
  PCDB_CUDA(w.lrf.ensure(sizeof(float) * 9 * Q));
  PCDB_TRY(stage_shot(ctx, P, Q, false, radius, radius, true, false, nullptr, w.lrf.as<float>(), nullptr));
  PCDB_TRY(download(ctx, lrf9_out, w.lrf.p, sizeof(float) * 9 * Q));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_shot_describe(pcdb_ctx* ctx, int32_t feature_type, const float* surf_xyz, const float* surf_normals,
                       const uint32_t* surf_rgb, const int64_t* surf_off, const float* kp_xyz,
                       const uint32_t* kp_rgb, const float* kp_lrf9, const int64_t* kp_off, int32_t B, double radius,
                       float* desc_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (!(radius > 0)) return ctx->fail(PCDB_E_INVALID, "radius must be positive");
  if (feature_type != PCDB_FEATURE_SHOT && feature_type != PCDB_FEATURE_CSHOT)
    return ctx->fail(PCDB_E_INVALID, "invalid feature type %d", feature_type);
  if (!surf_normals || !kp_lrf9) return ctx->fail(PCDB_E_INVALID, "normals and reference frames are required");
  Workspace& w = ctx->ws;
  const bool color = feature_type == PCDB_FEATURE_CSHOT;
  const int D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  int64_t P = 0, Q = 0;
  PCDB_TRY(prepare_explicit(ctx, surf_xyz, surf_normals, surf_rgb, surf_off, kp_xyz, kp_rgb, kp_off, B, radius, color,
                            &P, &Q));
  if (Q == 0) return PCDB_OK;
  PCDB_TRY(upload(ctx, w.lrf, kp_lrf9, sizeof(float) * 9 * Q));
  PCDB_CUDA(w.desc.ensure(sizeof(float) * (size_t)D * Q));
  PCDB_TRY(stage_shot(ctx, P, Q, color, radius, radius, false, true, w.lrf.as<float>(), nullptr, w.desc.as<float>()));
  PCDB_TRY(download(ctx, desc_out, w.desc.p, sizeof(float) * (size_t)D * Q));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_compute_normals(pcdb_ctx* ctx, const float* xyz, const int64_t* cloud_off, int32_t B, float* normals_out,
                         float* curvature_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, cloud_off, B, "cloud_off"));
  if (!normals_out) return ctx->fail(PCDB_E_INVALID, "normals_out is required");
  Workspace& w = ctx->ws;
  const int64_t P = cloud_off[B];
  PCDB_TRY(upload(ctx, w.in_xyz, xyz, sizeof(float) * 3 * P));
  PCDB_TRY(upload(ctx, w.cloud_off, cloud_off, sizeof(int64_t) * (B + 1)));
  PCDB_CUDA(w.nrm_curv.ensure(sizeof(float) * (P + 1)));
  PCDB_TRY(stage_normals(ctx, B, P, w.nrm_curv.as<float>()));
  PCDB_TRY(download(ctx, normals_out, w.in_nrm.p, sizeof(float) * 3 * P));
  if (curvature_out) PCDB_TRY(download(ctx, curvature_out, w.nrm_curv.p, sizeof(float) * P));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_compute_normals_organized(pcdb_ctx* ctx, const float* xyz, int32_t width, int32_t height, float* normals_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (width < 0 || height < 0 || (int64_t)width * height > 0x7fffff00ll)
    return ctx->fail(PCDB_E_INVALID, "bad organized cloud size %d x %d", width, height);
  if (!xyz || !normals_out) return ctx->fail(PCDB_E_INVALID, "xyz and normals_out are required");
  Workspace& w = ctx->ws;
  const int64_t P = (int64_t)width * height;
  PCDB_TRY(upload(ctx, w.in_xyz, xyz, sizeof(float) * 3 * P));
  PCDB_TRY(stage_normals_organized(ctx, width, height));
  PCDB_TRY(download(ctx, normals_out, w.in_nrm.p, sizeof(float) * 3 * P));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_compute_features(pcdb_ctx* ctx, const float* xyz, const float* normals, const uint32_t* rgb,
                          const int64_t* cloud_off, int32_t B, float* feat_xyz_out, float* feat_lrf9_out,
                          float* feat_desc_out, int64_t* feat_off_out, int64_t feat_capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, cloud_off, B, "cloud_off"));
  Workspace& w = ctx->ws;
  const int64_t P = cloud_off[B];
  const int D = ctx->prm.feature_type == PCDB_FEATURE_CSHOT ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  PCDB_TRY(upload(ctx, w.in_xyz, xyz, sizeof(float) * 3 * P));
  if (normals) PCDB_TRY(upload(ctx, w.in_nrm, normals, sizeof(float) * 3 * P));
  if (rgb) PCDB_TRY(upload(ctx, w.in_rgb, rgb, sizeof(uint32_t) * P));
  PCDB_TRY(upload(ctx, w.cloud_off, cloud_off, sizeof(int64_t) * (B + 1)));
  if (!normals) PCDB_TRY(stage_normals(ctx, B, P, nullptr));  // hasNormals == false (implicit_shape_model.cpp:852-858)
  int64_t F = 0, Q = 0;
  PCDB_TRY(features_pipeline(ctx, B, P, rgb != nullptr, &F, &Q));
  PCDB_TRY(download(ctx, feat_off_out, w.feat_off.p, sizeof(int64_t) * (B + 1)));
  if (F > feat_capacity) {
    PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ctx->fail(PCDB_E_CAPACITY, "feat_capacity %lld < %lld features", (long long)feat_capacity, (long long)F);
  }
  PCDB_TRY(download(ctx, feat_xyz_out, w.feat_xyz.p, sizeof(float) * 3 * F));
  PCDB_TRY(download(ctx, feat_lrf9_out, w.feat_lrf.p, sizeof(float) * 9 * F));
  PCDB_TRY(download(ctx, feat_desc_out, w.feat_desc.p, sizeof(float) * (size_t)D * F));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

static void record_sweep_times(pcdb_ctx* ctx) {
  ctx->stats.knn_bound_sweep_ms = ctx->stats.knn_pool_sweep_ms = 0;
  if (!ctx->knn_sweep_events_valid) return;
  float a = 0, b = 0;
  if (cudaEventElapsedTime(&a, ctx->ev_knn[0], ctx->ev_knn[1]) == cudaSuccess) ctx->stats.knn_bound_sweep_ms = a;
  if (cudaEventElapsedTime(&b, ctx->ev_knn[2], ctx->ev_knn[3]) == cudaSuccess) ctx->stats.knn_pool_sweep_ms = b;
}

int pcdb_knn(pcdb_ctx* ctx, const float* queries, int64_t Q, int32_t k, int32_t dist_type, int32_t mode,
             int32_t* idx_out, float* dist_out, int32_t* count_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->cb.N_table == 0) return ctx->fail(PCDB_E_STATE, "pcdb_knn before pcdb_set_codebook");
  if (k < 1 || k > PCDB_MAX_K) return ctx->fail(PCDB_E_INVALID, "k must be in 1..%d", PCDB_MAX_K);
  if (dist_type != PCDB_DIST_EUCLIDEAN && dist_type != PCDB_DIST_CHISQUARED)
    return ctx->fail(PCDB_E_INVALID, "invalid distance type %d", dist_type);
  if (Q < 0) return ctx->fail(PCDB_E_INVALID, "negative query count");
  Workspace& w = ctx->ws;
  PCDB_TRY(upload(ctx, w.feat_desc, queries, sizeof(float) * (size_t)ctx->cb.D * Q));
  ctx->comm_events_valid = false;
  ctx->gemm_events_valid = false;
  ctx->knn_sweep_events_valid = false;
  PCDB_TRY(activate(ctx, w.feat_desc.as<float>(), Q, k, dist_type, mode, ctx->prm.use_distance_ratio != 0,
                    ctx->prm.distance_ratio_threshold));
  ctx->stats.comm_ms = 0;
  PCDB_TRY(download(ctx, idx_out, w.knn_idx.p, sizeof(int) * Q * k));
  PCDB_TRY(download(ctx, dist_out, w.knn_dist.p, sizeof(float) * Q * k));
  PCDB_TRY(download(ctx, count_out, w.knn_cnt.p, sizeof(int) * Q));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.comm_ms = comm_last_exchange_ms(ctx);
  ctx->stats.knn_gemm_ms = 0;
  if (ctx->gemm_events_valid) {
    float g = 0;
    if (cudaEventElapsedTime(&g, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ctx->stats.knn_gemm_ms = g;
  }
  record_sweep_times(ctx);
  return PCDB_OK;
}

int pcdb_distance_pairs(pcdb_ctx* ctx, const float* a, const float* b, int64_t n, int32_t D, int32_t dist_type,
                        float* out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (n < 0 || D <= 0) return ctx->fail(PCDB_E_INVALID, "bad arguments");
  if (dist_type != PCDB_DIST_EUCLIDEAN && dist_type != PCDB_DIST_CHISQUARED)
    return ctx->fail(PCDB_E_INVALID, "invalid distance type %d", dist_type);
  Workspace& w = ctx->ws;
  PCDB_TRY(upload(ctx, w.merge_a, a, sizeof(float) * (size_t)n * D));
  PCDB_TRY(upload(ctx, w.merge_b, b, sizeof(float) * (size_t)n * D));
  PCDB_CUDA(w.knn_dist.ensure(sizeof(float) * (n + 1)));
  PCDB_TRY(stage_pair_distances(ctx, w.merge_a.as<float>(), w.merge_b.as<float>(), n, D, dist_type, w.knn_dist.as<float>()));
  PCDB_TRY(download(ctx, out, w.knn_dist.p, sizeof(float) * n));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_cast_votes(pcdb_ctx* ctx, const float* feat_xyz, const float* feat_lrf9, const int64_t* feat_off, int32_t B,
                    const int32_t* knn_idx, const float* knn_dist, const int32_t* knn_count, int32_t k,
                    pcdb_vote* votes_out, int64_t* vote_off_out, int64_t vote_capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->cb.N_table == 0) return ctx->fail(PCDB_E_STATE, "pcdb_cast_votes before pcdb_set_codebook");
  PCDB_TRY(check_offsets(ctx, feat_off, B, "feat_off"));
  if (k < 1 || k > PCDB_MAX_K) return ctx->fail(PCDB_E_INVALID, "k must be in 1..%d", PCDB_MAX_K);
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const int64_t F = feat_off[B];
  for (int64_t i = 0; i < F; ++i)
    for (int j = 0; j < knn_count[i] && j < k; ++j) {
      if (knn_idx[i * k + j] < 0) continue;  // masked: the row lives in another codebook shard
      int64_t r = (int64_t)knn_idx[i * k + j] - ctx->cb.row_base;
      if (r < 0 || r >= ctx->cb.N_table) return ctx->fail(PCDB_E_INVALID, "activated row %d outside this codebook shard", knn_idx[i * k + j]);
    }
  PCDB_TRY(upload(ctx, w.feat_xyz, feat_xyz, sizeof(float) * 3 * F));
  PCDB_TRY(upload(ctx, w.feat_lrf, feat_lrf9, sizeof(float) * 9 * F));
  PCDB_TRY(upload(ctx, w.feat_off, feat_off, sizeof(int64_t) * (B + 1)));
  PCDB_TRY(upload(ctx, w.knn_idx, knn_idx, sizeof(int) * F * k));
  PCDB_TRY(upload(ctx, w.knn_dist, knn_dist, sizeof(float) * F * k));
  PCDB_TRY(upload(ctx, w.knn_cnt, knn_count, sizeof(int) * F));
  PCDB_CUDA(w.feat_cloud.ensure(sizeof(int) * (F + 1)));
  if (F > 0) {
    k_cloud_ids<<<cdiv(F, 256), 256, 0, st>>>(w.feat_off.as<long long>(), B, F, w.feat_cloud.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  int64_t V = 0;
  PCDB_TRY(stage_cast_votes(ctx, w.feat_xyz.as<float>(), w.feat_lrf.as<float>(), w.feat_off.as<long long>(),
                            w.feat_cloud.as<int>(), B, F, k, &V));
  PCDB_TRY(download(ctx, vote_off_out, w.vote_off.p, sizeof(int64_t) * (B + 1)));
  if (V > vote_capacity) {
    PCDB_CUDA(cudaStreamSynchronize(st));
    return ctx->fail(PCDB_E_CAPACITY, "vote_capacity %lld < %lld votes", (long long)vote_capacity, (long long)V);
  }
  PCDB_TRY(download(ctx, votes_out, w.votes.p, sizeof(pcdb_vote) * V));
  PCDB_CUDA(cudaStreamSynchronize(st));
  ctx->stats.n_votes += V;
  ctx->last_V = V;
  ctx->last_B = B;
  return PCDB_OK;
}

int pcdb_find_maxima(pcdb_ctx* ctx, const pcdb_vote* votes, const int64_t* vote_off, int32_t B,
                     pcdb_maximum* maxima_out, int64_t* maxima_off_out, int64_t maxima_capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, vote_off, B, "vote_off"));
  if (ctx->cb.n_classes <= 0) return ctx->fail(PCDB_E_STATE, "pcdb_find_maxima before pcdb_set_codebook (class count unknown)");
  Workspace& w = ctx->ws;
  const int64_t V = vote_off[B];
  for (int64_t i = 0; i < V; ++i)
    if (votes[i].class_id >= (uint32_t)ctx->cb.n_classes)
      return ctx->fail(PCDB_E_INVALID, "vote %lld has class id %u >= n_classes", (long long)i, votes[i].class_id);
  PCDB_TRY(upload(ctx, w.votes, votes, sizeof(pcdb_vote) * V));
  PCDB_TRY(upload(ctx, w.vote_off, vote_off, sizeof(int64_t) * (B + 1)));
  PCDB_TRY(stage_votes_unpack(ctx, B, V));
  int64_t M = 0, members = 0;
  PCDB_TRY(stage_find_maxima(ctx, B, V, false, &M, &members));
  ctx->last_V = V;
  ctx->last_M = M;
  ctx->last_members = members;
  ctx->last_B = B;
  return fetch_maxima(ctx, B, M, maxima_out, maxima_off_out, maxima_capacity, nullptr);
}

int pcdb_get_maximum_votes(pcdb_ctx* ctx, int64_t* vote_index_out, float* vote_weight_out, int64_t capacity,
                           int64_t* n_out) {
  if (!ctx || !n_out) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  *n_out = ctx->last_members;
  if (capacity < ctx->last_members) return ctx->fail(PCDB_E_CAPACITY, "member capacity too small");
  PCDB_TRY(download(ctx, vote_index_out, ctx->ws.mem_idx.p, sizeof(int64_t) * ctx->last_members));
  PCDB_TRY(download(ctx, vote_weight_out, ctx->ws.mem_w.p, sizeof(float) * ctx->last_members));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_get_votes(pcdb_ctx* ctx, pcdb_vote* votes_out, int64_t* vote_off_out, int64_t capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (capacity < ctx->last_V) return ctx->fail(PCDB_E_CAPACITY, "vote capacity too small");
  PCDB_TRY(download(ctx, votes_out, ctx->ws.votes.p, sizeof(pcdb_vote) * ctx->last_V));
  PCDB_TRY(download(ctx, vote_off_out, ctx->ws.vote_off.p, sizeof(int64_t) * (ctx->last_B + 1)));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_get_last_sizes(pcdb_ctx* ctx, int64_t* n_votes_out, int64_t* n_maxima_out, int64_t* n_members_out) {
  if (!ctx) return PCDB_E_INVALID;
  if (n_votes_out) *n_votes_out = ctx->last_V;
  if (n_maxima_out) *n_maxima_out = ctx->last_kept;
  if (n_members_out) *n_members_out = ctx->last_members;
  return PCDB_OK;
}

int pcdb_get_maxima(pcdb_ctx* ctx, pcdb_maximum* maxima_out, int64_t* maxima_off_out, int64_t maxima_capacity) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (!maxima_out && maxima_capacity > 0) return ctx->fail(PCDB_E_INVALID, "maxima_out is required");
  return fetch_maxima(ctx, ctx->last_B, ctx->last_M, maxima_out, maxima_off_out, maxima_capacity, nullptr, false);
}

static void record_stage_times(pcdb_ctx* ctx, const float t[4]) {
  record_sweep_times(ctx);
  ctx->stats.features_ms = t[0];
  ctx->stats.knn_ms = t[1];
  ctx->stats.votes_ms = t[2];
  ctx->stats.maxima_ms = t[3];
  ctx->stats.knn_gemm_ms = 0;
  if (ctx->gemm_events_valid) {
    float g = 0;
    if (cudaEventElapsedTime(&g, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ctx->stats.knn_gemm_ms = g;
  }
  ctx->stats.comm_ms = comm_last_exchange_ms(ctx);
}

// device-resident core of detect(): inputs already in ws.in_* / ws.cloud_off
static int classify_core(pcdb_ctx* ctx, int B, int64_t P, bool has_rgb, bool has_normals, int64_t* M_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const pcdb_params& p = ctx->prm;
  if (ctx->cb.N_table == 0) return ctx->fail(PCDB_E_STATE, "classify before pcdb_set_codebook");
  if (comm_codebook_sharded(ctx) && comm_keypoints_sharded(ctx))
    return ctx->fail(PCDB_E_UNSUPPORTED, "a sharded codebook and sharded keypoints cannot be combined");
  const int D = p.feature_type == PCDB_FEATURE_CSHOT ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  if (ctx->cb.D != D) return ctx->fail(PCDB_E_INVALID, "codebook dimension %d does not match Features.Type (%d)", ctx->cb.D, D);
  ctx->gemm_events_valid = false;
  ctx->knn_sweep_events_valid = false;
  PCDB_CUDA(cudaEventRecord(ctx->ev[7], st));
  if (B == 0) {  // nothing of this rank's own to classify; a sharded codebook still needs its rows searched
    for (int i = 0; i < 2; ++i) PCDB_CUDA(cudaEventRecord(ctx->ev[i], st));
    ctx->comm_events_valid = false;
    if (comm_codebook_sharded(ctx))
      PCDB_TRY(activate(ctx, nullptr, 0, p.knn_k, p.distance_type, PCDB_KNN_AUTO, p.use_distance_ratio != 0,
                        p.distance_ratio_threshold));
    for (int i = 2; i < 5; ++i) PCDB_CUDA(cudaEventRecord(ctx->ev[i], st));
    ctx->last_V = ctx->last_M = ctx->last_members = 0;
    ctx->last_B = 0;
    *M_out = 0;
    return PCDB_OK;
  }
  if (!has_normals) PCDB_TRY(stage_normals(ctx, B, P, nullptr));  // "normals" bucket of the reference's timing table
  PCDB_CUDA(cudaEventRecord(ctx->ev[0], st));
  int64_t F = 0, Q = 0;
  PCDB_TRY(features_pipeline(ctx, B, P, has_rgb, &F, &Q));
  PCDB_CUDA(cudaEventRecord(ctx->ev[1], st));
  ctx->comm_events_valid = false;
  PCDB_TRY(activate(ctx, w.feat_desc.as<float>(), F, p.knn_k, p.distance_type, PCDB_KNN_AUTO,
                    p.use_distance_ratio != 0, p.distance_ratio_threshold));
  PCDB_CUDA(cudaEventRecord(ctx->ev[2], st));
  pcdb_trace_point(ctx, "activation");
  int64_t V = 0;
  PCDB_TRY(stage_cast_votes(ctx, w.feat_xyz.as<float>(), w.feat_lrf.as<float>(), w.feat_off.as<long long>(),
                            w.feat_cloud.as<int>(), B, F, p.knn_k, &V));
  if (comm_keypoints_sharded(ctx)) PCDB_TRY(stage_gather_votes(ctx, B, F, p.knn_k, V, &V));
  PCDB_CUDA(cudaEventRecord(ctx->ev[3], st));
  pcdb_trace_point(ctx, "votes");
  int64_t M = 0, members = 0;
  PCDB_TRY(stage_find_maxima(ctx, B, V, true, &M, &members));
  PCDB_CUDA(cudaEventRecord(ctx->ev[4], st));
  pcdb_trace_point(ctx, "maxima");
  ctx->stats.n_votes += V;
  ctx->last_V = V;
  ctx->last_M = M;
  ctx->last_members = members;
  ctx->last_B = B;
  *M_out = M;
  return PCDB_OK;
}

int pcdb_classify_batch(pcdb_ctx* ctx, const float* xyz, const float* normals, const uint32_t* rgb,
                        const int64_t* cloud_off, int32_t B, int32_t* label_out, pcdb_maximum* maxima_out,
                        int64_t* maxima_off_out, int64_t maxima_capacity, double* times_ms_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, cloud_off, B, "cloud_off"));
  if (!label_out) return ctx->fail(PCDB_E_INVALID, "label_out is required");
  Workspace& w = ctx->ws;
  const int64_t P = cloud_off[B];
  PCDB_TRY(upload(ctx, w.in_xyz, xyz, sizeof(float) * 3 * P));
  if (normals) PCDB_TRY(upload(ctx, w.in_nrm, normals, sizeof(float) * 3 * P));
  if (rgb) PCDB_TRY(upload(ctx, w.in_rgb, rgb, sizeof(uint32_t) * P));
  PCDB_TRY(upload(ctx, w.cloud_off, cloud_off, sizeof(int64_t) * (B + 1)));
  int64_t M = 0;
  PCDB_TRY(classify_core(ctx, B, P, rgb != nullptr, normals != nullptr, &M));
  PCDB_TRY(fetch_maxima(ctx, B, M, maxima_out, maxima_off_out, maxima_capacity, label_out));
  if (times_ms_out) {
    float t[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], ctx->ev[i], ctx->ev[i + 1]);
    float t_nrm = 0;
    cudaEventElapsedTime(&t_nrm, ctx->ev[7], ctx->ev[0]);
    times_ms_out[0] = t_nrm + t[0] + t[1] + t[2] + t[3];  // complete
    times_ms_out[1] = t[0];                       // features (keypoints included; split below is not tracked)
    times_ms_out[2] = 0;                          // keypoints
    times_ms_out[3] = normals ? 0 : t_nrm;        // normals (0 when given)
    times_ms_out[4] = 0;                          // flann (index build: none, the codebook is resident)
    times_ms_out[5] = t[1] + t[2];                // voting = activation + vote casting
    times_ms_out[6] = t[3];                       // maxima
    record_stage_times(ctx, t);
  }
  return PCDB_OK;
}

int pcdb_classify_batch_d(pcdb_ctx* ctx, const float* xyz_d, const float* normals_d, const uint32_t* rgb_d,
                          const int64_t* cloud_off, int32_t B, int32_t* label_out_d) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  PCDB_TRY(check_offsets(ctx, cloud_off, B, "cloud_off"));
  if (!xyz_d || !label_out_d) return ctx->fail(PCDB_E_INVALID, "xyz_d and label_out_d are required");
  Workspace& w = ctx->ws;
  const int64_t P = cloud_off[B];
  // borrow the caller's device arrays for this call (no copy)
  DevBuf sx = w.in_xyz, sn = w.in_nrm, sr = w.in_rgb;
  w.in_xyz.p = const_cast<float*>(xyz_d);
  if (normals_d) w.in_nrm.p = const_cast<float*>(normals_d);  // else estimated into the workspace's own buffer
  w.in_rgb.p = const_cast<uint32_t*>(rgb_d);
  int rc = upload(ctx, w.cloud_off, cloud_off, sizeof(int64_t) * (B + 1));
  int64_t M = 0;
  if (rc == PCDB_OK) rc = classify_core(ctx, B, P, rgb_d != nullptr, normals_d != nullptr, &M);
  w.in_xyz = sx;
  if (normals_d) w.in_nrm = sn;
  w.in_rgb = sr;
  PCDB_TRY(rc);
  PCDB_CUDA(cudaMemcpyAsync(label_out_d, w.labels.p, sizeof(int) * B, cudaMemcpyDeviceToDevice, ctx->stream));
  float t[4] = {0, 0, 0, 0};
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], ctx->ev[i], ctx->ev[i + 1]);
  record_stage_times(ctx, t);
  return PCDB_OK;
}

int pcdb_merge_topk(pcdb_ctx* ctx, const int32_t* cand_idx, const float* cand_dist, int32_t S, int64_t Q, int32_t k,
                    int32_t* idx_out, float* dist_out) {
  if (!ctx) return PCDB_E_INVALID;
  PCDB_CUDA(cudaSetDevice(ctx->device));
  if (S < 1 || k < 1 || k > PCDB_MAX_K || Q < 0) return ctx->fail(PCDB_E_INVALID, "bad merge arguments");
  Workspace& w = ctx->ws;
  const size_t n = (size_t)S * Q * k;
  PCDB_TRY(upload(ctx, w.merge_a, cand_idx, sizeof(int) * n));
  PCDB_TRY(upload(ctx, w.merge_b, cand_dist, sizeof(float) * n));
  PCDB_CUDA(w.knn_idx.ensure(sizeof(int) * (Q * k + 1)));
  PCDB_CUDA(w.knn_dist.ensure(sizeof(float) * (Q * k + 1)));
  if (Q > 0) {
    k_merge_topk<<<cdiv(Q, 128), 128, 0, ctx->stream>>>(w.merge_a.as<int>(), w.merge_b.as<float>(), S, Q, k,
                                                        w.knn_idx.as<int>(), w.knn_dist.as<float>());
    PCDB_LAUNCH_CHECK();
  }
  PCDB_TRY(download(ctx, idx_out, w.knn_idx.p, sizeof(int) * Q * k));
  PCDB_TRY(download(ctx, dist_out, w.knn_dist.p, sizeof(float) * Q * k));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}

int pcdb_get_stats(pcdb_ctx* ctx, pcdb_stats* out) {
  if (!ctx || !out) return PCDB_E_INVALID;
  *out = ctx->stats;
  return PCDB_OK;
}
int pcdb_reset_stats(pcdb_ctx* ctx) {
  if (!ctx) return PCDB_E_INVALID;
  std::memset(&ctx->stats, 0, sizeof(ctx->stats));
  return PCDB_OK;
}

}  // extern "C"
