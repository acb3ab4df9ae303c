// common.cuh — context, device buffers and small device helpers shared by the sm_100a kernels of libpcdb200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pcdb200.h"

#define PCDB_NUM_SMS 148

// ---- growable device buffer (capacity only ever grows; contents are scratch) -------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) {
      cudaError_t e = cudaFree(p);
      if (e != cudaSuccess) return e;
      p = nullptr;
      cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// Per-cloud geometry derived on the device from the finite points' bounding box.
struct CloudInfo {
  float mn[3];      // min corner of the finite points
  float mx[3];
  int min_b[3];     // pcl::VoxelGrid min_b_ (floor(min * inv_leaf))
  int mul1, mul2;   // divb_mul_ = (1, div_b.x, div_b.x*div_b.y)
  int voxel_overflow;  // PCL "leaf size too small"
  int grid_overflow;   // search-grid coordinate would not fit 16 bits
};

struct Codebook_d {
  // Rows: the vote tables (CSR) cover N_table codewords; the descriptor rows held on this device are table rows
  // [word_lo, word_lo + N).  Unsharded: word_lo = 0, N = N_table.  Row ids leaving the kNN kernels are
  // row_base + word_lo + local row ("global" ids); vote casting maps id - row_base to a table row.
  int64_t N = 0, V = 0, row_base = 0, N_table = 0, word_lo = 0;
  int D = 0, n_classes = 0;
  DevBuf words, vote_off, vote_xyz, vote_weight, vote_class, vote_instance, vote_bbox,
      vote_class_weight, kp_train, ids, cw_weight, sigma2;
  int max_votes_per_word = 0;
  // tensor-core operands (knn_gemm.cu owns the buffers): [0] squared-L2 rows, [1] sqrt rows for the chi^2 sandwich
  bool gemm_ready[2] = {false, false};
  bool gemm_tried[2] = {false, false};
  void release() {
    DevBuf* all[] = {&words, &vote_off, &vote_xyz, &vote_weight, &vote_class, &vote_instance,
                     &vote_bbox, &vote_class_weight, &kp_train, &ids, &cw_weight, &sigma2};
    for (DevBuf* b : all) b->release();
    N = V = N_table = 0;
  }
};

// Workspace of one batch: all device arrays of the fused path, reused (and only ever grown) between calls.
#define PCDB_WS_FIELDS(X) \
  X(in_xyz) X(in_nrm) X(in_rgb) X(cloud_off) X(flag_pt) X(flag_sf) \
  X(pos_pt) X(pos_sf) X(pts4) X(pts_off) X(surf4) X(snrm4) \
  X(surf_off) X(surf_cloud) X(surf_local) X(pts_cloud) X(minmax) X(cinfo) \
  X(err_flag) X(vkeys) X(vvals) X(vkeys2) X(vvals2) X(seg_head) \
  X(seg_id) X(seg_start) X(kp4) X(kp_off) X(kp_cloud) X(kp_in) \
  X(gkeys) X(gvals) X(gkeys2) X(gvals2) X(surfS4) X(snrmS4) \
  X(slabS4) X(kkeys) X(kvals) X(kkeys2) X(kvals2) X(item_head) \
  X(item_id) X(item_start) X(cub_tmp) X(scalars) X(lrf) X(desc) \
  X(feat_valid) X(feat_pos) X(feat_xyz) X(feat_lrf) X(feat_desc) X(feat_off) \
  X(feat_cloud) X(knn_idx) X(knn_dist) X(knn_cnt) X(knn_part_d) X(knn_part_i) \
  X(cand_idx) X(cand_apx) X(cand_cnt) X(cand_thr) X(cand_exact) X(feat_h) \
  X(feat_eps) X(knn_fb_list) X(knn_fb_q) X(vote_cnt) X(vote_pos) X(votes) \
  X(vote_off) X(vote_pw) X(vote_cloud) X(vote_key) X(vote_key2) X(vote_ord) \
  X(vote_ord2) X(vote_pwS) X(vote_w_work) X(vote_w0) X(seg2_head) X(seg2_id) \
  X(seg2_start) X(seg2_key) X(seed_k0) X(seed_k1) X(seed_kS) X(seed_i0) \
  X(seed_i1) X(seed_i2) X(seed_i3) X(seed_s0) X(seed_s1) X(seed_head) \
  X(seed_id) X(seed_key) X(seed_seg) X(seed_first) X(centers) X(ms_cen) \
  X(ms_cen2) X(ms_dens) X(ms_flag) X(max_pos) X(max_n) X(max_off) \
  X(mpos) X(mseg) X(mem_cnt) X(mem_off) X(mem_idx) X(mem_w) \
  X(max_raw) X(max_sorted) X(max_kept) X(max_first) X(labels) X(nbr_cnt) \
  X(nbr_off) X(nbr_key) X(nbr_key2) X(merge_a) X(merge_b) X(nrm_pca) \
  X(nrm_cen) X(nrm_inv) X(nrm_curv) X(max_flag) X(shot_glist) X(item_beg) X(item_len) X(feat_kp) X(vote_feat) \
  X(item_next) X(org_ch) X(org_int) X(org_dist) X(rs_mul) X(rs_keep) X(rs_cnt) X(mem_off2) X(ms_grp) X(ms_close) X(ms_src_dst) X(ms_cls_h) X(ms_cloud_cm) X(mem_idx2) X(mem_w2)
struct Workspace {
#define X(n) DevBuf n;
  PCDB_WS_FIELDS(X)
#undef X
  void release() {
#define X(n) n.release();
    PCDB_WS_FIELDS(X)
#undef X
  }
};

struct pcdb_ctx {
  int device = 0;
  int sm_count = PCDB_NUM_SMS;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  pcdb_params prm;
  bool prm_set = false;
  Codebook_d cb;
  Workspace ws;
  pcdb_stats stats;
  bool gemm_events_valid = false;  // ev[5], ev[6] bracket the last tcgen05 activation kernel
  float grid_inv_cell = 0.f;  // 1 / search-grid cell edge of the current batch
  void* gemm_state = nullptr;            // knn_gemm.cu's per-context buffers / tensor map (opaque here)
  void (*gemm_state_free)(void*) = nullptr;
  void* comm_state = nullptr;            // comm.cu's NCCL communicator and exchange buffers (opaque here)
  void (*comm_state_free)(void*) = nullptr;
  cudaEvent_t ev_comm[4] = {nullptr, nullptr, nullptr, nullptr};  // brackets of the last exchange steps
  bool comm_events_valid = false;
  float* lab_lut_d = nullptr;  // 256 + 4000 floats, built on the host with powf (features_cshot.cpp:52-71)
  // host mirrors of the last batch (for pcdb_get_votes / pcdb_get_maximum_votes)
  std::vector<float> class_dim_first, class_dim_second;  // Voting::m_dimensions_map by class id (pcdb_set_class_dimensions)
  int64_t last_V = 0, last_M = 0, last_members = 0, last_kept = 0;
  int last_B = 0;
  std::vector<int64_t> h_off_a, h_off_b;
  // Small device->host reads (counts, flags) land in a pinned area and are handed to their destinations at the next
  // pcdb_sync_reads: a pageable destination makes every cudaMemcpyAsync a blocking ~20 us round trip of its own.
  void* pinned = nullptr;
  size_t pinned_cap = 0, pinned_used = 0;
  struct PendingRead { void* dst; size_t off, bytes; };
  std::vector<PendingRead> pending_reads;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_knn[4] = {nullptr, nullptr, nullptr, nullptr};  // bound sweep begin/end, pooled sweep begin/end
  bool knn_sweep_events_valid = false;

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define PCDB_CUDA(call)                                                                                      \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return ctx->fail(PCDB_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

#define PCDB_TRY(call)         \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != PCDB_OK) return rc__; \
  } while (0)

#define PCDB_LAUNCH_CHECK()                                                                                 \
  do {                                                                                                      \
    ctx->stats.kernel_launches++;                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                                   \
    if (e__ != cudaSuccess)                                                                                 \
      return ctx->fail(PCDB_E_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

// points the descriptor kernel stages in shared memory for one work item (shot.cu); clouds whose whole surface fits
// are one work item (prep.cu groups their keypoints by cloud instead of by search-grid cell)
constexpr int PCDB_SHOT_CHUNK = 2048;

static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// PCDB_TRACE=1: wall-clock between named points of a call, each after a stream sync (diagnosis only: the syncs
// serialise what normally overlaps).  Prints to stderr.
void pcdb_trace_point(pcdb_ctx* ctx, const char* name);

// ---- device helpers -------------------------------------------------------------------------------------
#ifdef __CUDACC__
// float distance of flann::L2_Simple (x, y, z order; no FMA contraction)
__device__ __forceinline__ float sqdist3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  float d0 = __fsub_rn(ax, bx), d1 = __fsub_rn(ay, by), d2 = __fsub_rn(az, bz);
  float r = __fmul_rn(d0, d0);
  r = __fadd_rn(r, __fmul_rn(d1, d1));
  r = __fadd_rn(r, __fmul_rn(d2, d2));
  return r;
}
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  float r = __fmul_rn(ax, bx);
  r = __fadd_rn(r, __fmul_rn(ay, by));
  r = __fadd_rn(r, __fmul_rn(az, bz));
  return r;
}
__device__ __forceinline__ unsigned enc_float(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// search-grid key: cloud | cz | cy | cx, 16 bits each; consecutive cx are consecutive keys
__device__ __forceinline__ unsigned long long grid_key(unsigned cloud, int cx, int cy, int cz) {
  return ((unsigned long long)cloud << 48) | ((unsigned long long)(unsigned)cz << 32) |
         ((unsigned long long)(unsigned)cy << 16) | (unsigned long long)(unsigned)cx;
}
__device__ __forceinline__ int grid_coord(float v, float mn, float inv) {
  int c = (int)floorf(__fmul_rn(__fsub_rn(v, mn), inv));
  return min(65535, max(0, c));
}
// first index in [lo, hi) whose key is >= k
__device__ __forceinline__ long long lower_bound_u64(const unsigned long long* keys, long long lo, long long hi,
                                                     unsigned long long k) {
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (keys[mid] < k)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
#endif

// api.cu: enqueue a small device->host read / wait for the stream and deliver every pending read
int pcdb_read_small(pcdb_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
int pcdb_sync_reads(pcdb_ctx* ctx);

// ---- stage functions implemented across the .cu files (all asynchronous on ctx->stream unless noted) -----
int pcdb_cub_exclusive_sum_i32(pcdb_ctx* ctx, const int* in, int* out, int64_t n);  // out has n+1 entries (total last)
int pcdb_cub_sort_pairs_u64(pcdb_ctx* ctx, const unsigned long long* kin, unsigned long long* kout, const int* vin,
                            int* vout, int64_t n, int end_bit);
int pcdb_cub_sort_pairs_u32(pcdb_ctx* ctx, const unsigned* kin, unsigned* kout, const int* vin, int* vout, int64_t n,
                            int end_bit);
int pcdb_cub_segmented_sort_u64(pcdb_ctx* ctx, const unsigned long long* kin, unsigned long long* kout, int64_t n,
                                int nseg, const int* seg_begin, const int* seg_end);
