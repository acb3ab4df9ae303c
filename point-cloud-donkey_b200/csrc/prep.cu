// prep.cu — input compaction, per-cloud geometry, K1 voxel-grid keypoints, K2 uniform search grid.
//
// Replaces (reference paths under src/implicit_shape_model/):
//   removeNaNFromPointCloud / filterNormals      implicit_shape_model.cpp:608-625,1040-1068
//   KeypointsVoxelGrid::iComputeKeypoints         keypoints/keypoints_voxel_grid.cpp:30-46 (pcl::VoxelGrid)
//   pcl::search::KdTree (the radius-search index) implicit_shape_model.cpp:823-831
// All of it is integer / byte work bound by HBM: one coalesced pass per array, device-wide radix sorts for the
// two orderings (voxel index; search-grid cell), no per-cloud host round trips.
#include <cub/cub.cuh>

#include "common.cuh"
#include "stages.h"

// ---- CUB plumbing ---------------------------------------------------------------------------------------
int pcdb_cub_exclusive_sum_i32(pcdb_ctx* ctx, const int* in, int* out, int64_t n) {
  size_t tmp = 0;
  PCDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream));
  PCDB_CUDA(ctx->ws.cub_tmp.ensure(tmp));
  PCDB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->ws.cub_tmp.p, tmp, in, out, (int)n, ctx->stream));
  ctx->stats.kernel_launches++;
  return PCDB_OK;
}
// One cloud at a time (the GUI / detect() latency path) sorts ~2000 keys: the device-wide radix sort spends 7 onesweep
// launches of ~9 us each on that.  Up to 4096 pairs are sorted by ONE CTA instead (block radix sort, stable like the
// device-wide one; keys beyond n are padded with all-ones and sort behind every real key of equal low bits).
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_sort_small_u64(const unsigned long long* __restrict__ kin,
                                                            unsigned long long* __restrict__ kout,
                                                            const int* __restrict__ vin, int* __restrict__ vout, int n,
                                                            int end_bit) {
  using Sort = cub::BlockRadixSort<unsigned long long, THREADS, ITEMS, int>;
  __shared__ typename Sort::TempStorage tmp;
  unsigned long long k[ITEMS];
  int v[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = threadIdx.x * ITEMS + i;
    k[i] = idx < n ? kin[idx] : ~0ull;
    v[i] = idx < n ? vin[idx] : -1;
  }
  Sort(tmp).Sort(k, v, 0, end_bit);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int idx = threadIdx.x * ITEMS + i;
    if (idx < n) {
      kout[idx] = k[i];
      vout[idx] = v[i];
    }
  }
}

int pcdb_cub_sort_pairs_u64(pcdb_ctx* ctx, const unsigned long long* kin, unsigned long long* kout, const int* vin,
                            int* vout, int64_t n, int end_bit) {
  if (n > 0 && n <= 1024) {
    k_sort_small_u64<256, 4><<<1, 256, 0, ctx->stream>>>(kin, kout, vin, vout, (int)n, end_bit);
    PCDB_LAUNCH_CHECK();
    return PCDB_OK;
  }
  if (n > 0 && n <= 4096) {
    k_sort_small_u64<512, 8><<<1, 512, 0, ctx->stream>>>(kin, kout, vin, vout, (int)n, end_bit);
    PCDB_LAUNCH_CHECK();
    return PCDB_OK;
  }
  size_t tmp = 0;
  PCDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int)n, 0, end_bit, ctx->stream));
  PCDB_CUDA(ctx->ws.cub_tmp.ensure(tmp));
  PCDB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->ws.cub_tmp.p, tmp, kin, kout, vin, vout, (int)n, 0, end_bit,
                                            ctx->stream));
  ctx->stats.kernel_launches++;
  return PCDB_OK;
}
int pcdb_cub_sort_pairs_u32(pcdb_ctx* ctx, const unsigned* kin, unsigned* kout, const int* vin, int* vout, int64_t n,
                            int end_bit) {
  size_t tmp = 0;
  PCDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int)n, 0, end_bit, ctx->stream));
  PCDB_CUDA(ctx->ws.cub_tmp.ensure(tmp));
  PCDB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->ws.cub_tmp.p, tmp, kin, kout, vin, vout, (int)n, 0, end_bit,
                                            ctx->stream));
  ctx->stats.kernel_launches++;
  return PCDB_OK;
}
int pcdb_cub_segmented_sort_u64(pcdb_ctx* ctx, const unsigned long long* kin, unsigned long long* kout, int64_t n,
                                int nseg, const int* seg_begin, const int* seg_end) {
  size_t tmp = 0;
  PCDB_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp, kin, kout, (int)n, nseg, seg_begin, seg_end, ctx->stream));
  PCDB_CUDA(ctx->ws.cub_tmp.ensure(tmp));
  PCDB_CUDA(cub::DeviceSegmentedSort::SortKeys(ctx->ws.cub_tmp.p, tmp, kin, kout, (int)n, nseg, seg_begin, seg_end,
                                               ctx->stream));
  ctx->stats.kernel_launches++;
  return PCDB_OK;
}

static int ceil_log2(int64_t v) {
  int b = 0;
  while ((1ll << b) < v) ++b;
  return b;
}

// ---- compaction ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cloud_of(const long long* off, int B, long long i) {
  int lo = 0, hi = B;  // largest b with off[b] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__global__ void k_flags(const float* __restrict__ xyz, const float* __restrict__ nrm, long long n, int* f_pt,
                        int* f_sf) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool okp = finite3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  bool okn = nrm ? finite3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]) : true;
  f_pt[i] = okp ? 1 : 0;
  f_sf[i] = (okp && okn) ? 1 : 0;
}

__global__ void k_scatter(const float* __restrict__ xyz, const float* __restrict__ nrm,
                          const unsigned* __restrict__ rgb, const long long* __restrict__ cloud_off, int B, long long n,
                          const int* __restrict__ f_pt, const int* __restrict__ f_sf, const int* __restrict__ pos_pt,
                          const int* __restrict__ pos_sf, float4* pts4, int* pts_cloud, float4* surf4, float4* snrm4,
                          int* surf_cloud, int* surf_local) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!f_pt[i]) return;
  int b = cloud_of(cloud_off, B, i);
  float4 p = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __uint_as_float(rgb ? rgb[i] : 0u));
  pts4[pos_pt[i]] = p;
  pts_cloud[pos_pt[i]] = b;
  if (f_sf[i]) {
    int o = pos_sf[i];
    surf4[o] = p;
    // local index inside the cloud's pointsWithoutNaN (what the kd-tree returns as neighbour index)
    int local = o - pos_sf[cloud_off[b]];
    float4 q = nrm ? make_float4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], __int_as_float(local))
                   : make_float4(0.f, 0.f, 0.f, __int_as_float(local));
    snrm4[o] = q;
    surf_cloud[o] = b;
    surf_local[o] = local;
  }
}

// offsets of the compacted arrays: out[b] = pos[cloud_off[b]], out[B] = total (pos has n+1 entries)
__global__ void k_compact_offsets(const long long* __restrict__ cloud_off, int B, const int* __restrict__ pos_pt,
                                  const int* __restrict__ pos_sf, long long* pts_off, long long* surf_off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  pts_off[b] = pos_pt[cloud_off[b]];
  surf_off[b] = pos_sf[cloud_off[b]];
}

// ---- per-cloud bounding boxes ---------------------------------------------------------------------------------
__global__ void k_minmax_init(unsigned* mm, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 6) return;
  mm[i] = (i % 6 < 3) ? 0xffffffffu : 0u;
}

__global__ void k_minmax(const float4* __restrict__ pts, const int* __restrict__ cloud, long long n, unsigned* mm) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool act = i < n;
  float4 p = act ? pts[i] : make_float4(0, 0, 0, 0);
  int b = act ? cloud[i] : -1;
  int b0 = __shfl_sync(0xffffffffu, b, 0);
  bool uniform = __all_sync(0xffffffffu, b == b0) && b0 >= 0;
  if (uniform) {
    unsigned lo[3] = {enc_float(p.x), enc_float(p.y), enc_float(p.z)};
    unsigned hi[3] = {lo[0], lo[1], lo[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      }
    if ((threadIdx.x & 31) == 0)
      for (int a = 0; a < 3; ++a) {
        atomicMin(&mm[b0 * 6 + a], lo[a]);
        atomicMax(&mm[b0 * 6 + 3 + a], hi[a]);
      }
  } else if (act) {
    atomicMin(&mm[b * 6 + 0], enc_float(p.x));
    atomicMin(&mm[b * 6 + 1], enc_float(p.y));
    atomicMin(&mm[b * 6 + 2], enc_float(p.z));
    atomicMax(&mm[b * 6 + 3], enc_float(p.x));
    atomicMax(&mm[b * 6 + 4], enc_float(p.y));
    atomicMax(&mm[b * 6 + 5], enc_float(p.z));
  }
}

// pcl::VoxelGrid::applyFilter set-up + search-grid origin
__global__ void k_cloud_setup(const unsigned* __restrict__ mm, int B, float leaf, float inv_cell, CloudInfo* ci,
                              int* err_flag) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  CloudInfo c;
  bool empty = mm[b * 6] == 0xffffffffu && mm[b * 6 + 3] == 0u;
  for (int a = 0; a < 3; ++a) {
    c.mn[a] = empty ? 0.f : dec_float(mm[b * 6 + a]);
    c.mx[a] = empty ? 0.f : dec_float(mm[b * 6 + 3 + a]);
  }
  c.voxel_overflow = 0;
  c.grid_overflow = 0;
  if (leaf > 0.f) {
    float inv = __fdiv_rn(1.0f, leaf);
    long long d[3];
    int div_b[3];
    for (int a = 0; a < 3; ++a) {
      d[a] = (long long)(__fmul_rn(__fsub_rn(c.mx[a], c.mn[a]), inv)) + 1;
      c.min_b[a] = (int)floorf(__fmul_rn(c.mn[a], inv));
      int max_b = (int)floorf(__fmul_rn(c.mx[a], inv));
      div_b[a] = max_b - c.min_b[a] + 1;
    }
    if (d[0] * d[1] * d[2] > 2147483647ll) {
      c.voxel_overflow = 1;
      atomicOr(err_flag, 1);
    }
    c.mul1 = div_b[0];
    c.mul2 = div_b[0] * div_b[1];
  } else {
    c.min_b[0] = c.min_b[1] = c.min_b[2] = 0;
    c.mul1 = c.mul2 = 1;
  }
  if (inv_cell > 0.f) {
    for (int a = 0; a < 3; ++a)
      if (floorf(__fmul_rn(__fsub_rn(c.mx[a], c.mn[a]), inv_cell)) > 65534.f) {
        c.grid_overflow = 1;
        atomicOr(err_flag, 2);
      }
  }
  ci[b] = c;
}

// ---- K1: voxel-grid keypoints ------------------------------------------------------------------------------
// key = cloud << 32 | voxel index; value = compacted point index.  The radix sort is stable, so
// points of a voxel stay in ascending original index: the summation order the oracle defines.
__global__ void k_voxel_keys(const float4* __restrict__ pts, const int* __restrict__ cloud, long long n,
                             const CloudInfo* __restrict__ ci, float leaf, unsigned long long* keys, int* vals) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int b = cloud[i];
  const CloudInfo c = ci[b];
  float inv = __fdiv_rn(1.0f, leaf);
  int i0 = (int)(__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)c.min_b[0]));
  int i1 = (int)(__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)c.min_b[1]));
  int i2 = (int)(__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)c.min_b[2]));
  int idx = i0 + i1 * c.mul1 + i2 * c.mul2;
  keys[i] = ((unsigned long long)(unsigned)b << 32) | (unsigned long long)(unsigned)idx;
  vals[i] = (int)i;
}

__global__ void k_heads_u64(const unsigned long long* __restrict__ keys, long long n, int* head) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// Work items of the descriptor kernel: the keypoints of one search-grid cell — or of one whole cloud when the cloud's
// surface fits the kernel's shared-memory stage (small objects: a cell holds 2-3 keypoints, a cloud a few hundred; one
// staging then serves them all and every warp of the CTA has a keypoint in every round).
__global__ void k_item_heads(const unsigned long long* __restrict__ keys, long long n,
                             const long long* __restrict__ surf_off, int item_kp, int* head) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = 1;
  if (i > 0) {
    const unsigned long long a = keys[i], b = keys[i - 1];
    const unsigned cloud = (unsigned)(a >> 48);
    const bool whole = surf_off[cloud + 1] - surf_off[cloud] <= PCDB_SHOT_CHUNK;
    if (!whole) {
      h = a != b;
    } else if ((a >> 48) != (b >> 48)) {
      h = 1;
    } else {
      // a cloud's keypoints are cut into items of item_kp: the warps of a CTA fetch keypoints dynamically and meet at a
      // barrier when the item is exhausted, so an item should hold many keypoints per warp (the wait for the last warp
      // is one keypoint long: 64-keypoint items on 12 warps cost 11 % in barrier stalls) — unless the batch is so
      // small that only finer items keep every SM busy (the host sizes item_kp from the keypoint count)
      const long long first = lower_bound_u64(keys, 0, n, (unsigned long long)cloud << 48);
      const long long last = lower_bound_u64(keys, first, n, (unsigned long long)(cloud + 1) << 48);
      const long long n_items = (last - first + item_kp - 1) / item_kp;       // equal-sized items inside a cloud
      const long long size = (last - first + n_items - 1) / n_items;
      h = ((i - first) % size) == 0;
    }
  }
  head[i] = h;
}

// seg_start[id] = i for every head; seg_start[total] = n
__global__ void k_seg_starts(const int* __restrict__ head, const int* __restrict__ seg_id, long long n,
                             int* seg_start) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    seg_start[seg_id[n]] = (int)n;
    return;
  }
  if (head[i]) seg_start[seg_id[i]] = (int)i;
}

// kp_off[b] = number of voxels of clouds < b  (keys sorted by cloud first; shift = bits below the cloud id)
__global__ void k_offsets_from_keys(const unsigned long long* __restrict__ keys, const int* __restrict__ seg_id,
                                    long long n, int B, int shift, long long* off) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  long long lo = lower_bound_u64(keys, 0, n, (unsigned long long)(unsigned)b << shift);
  off[b] = seg_id[lo];  // seg_id has n+1 entries
}

// one thread per voxel: float sums in sorted (= ascending point index) order, then / n  
__global__ void k_centroids(const float4* __restrict__ pts, const int* __restrict__ order,
                            const int* __restrict__ seg_start, int nseg, const int* __restrict__ cloud, float4* kp,
                            int* kp_cloud) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int a = seg_start[s], e = seg_start[s + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
  for (int t = a; t < e; ++t) {
    float4 p = pts[order[t]];
    sx = __fadd_rn(sx, p.x);
    sy = __fadd_rn(sy, p.y);
    sz = __fadd_rn(sz, p.z);
    unsigned c = __float_as_uint(p.w);
    sr = __fadd_rn(sr, (float)((c >> 16) & 0xFF));
    sg = __fadd_rn(sg, (float)((c >> 8) & 0xFF));
    sb = __fadd_rn(sb, (float)(c & 0xFF));
  }
  float cnt = (float)(e - a);
  unsigned r = (unsigned)__fdiv_rn(sr, cnt), g = (unsigned)__fdiv_rn(sg, cnt), bb = (unsigned)__fdiv_rn(sb, cnt);
  kp[s] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt),
                      __uint_as_float((r << 16) | (g << 8) | bb));
  kp_cloud[s] = cloud[order[a]];
}

// ---- K2: uniform search grid --------------------------------------------------------------------------------
__global__ void k_grid_keys(const float4* __restrict__ pts, const int* __restrict__ cloud, long long n,
                            const CloudInfo* __restrict__ ci, float inv_cell, unsigned long long* keys, int* vals) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int b = cloud[i];
  const CloudInfo c = ci[b];
  keys[i] = grid_key((unsigned)b, grid_coord(p.x, c.mn[0], inv_cell), grid_coord(p.y, c.mn[1], inv_cell),
                     grid_coord(p.z, c.mn[2], inv_cell));
  vals[i] = (int)i;
}

__global__ void k_gather_surface(const int* __restrict__ order, long long n, const float4* __restrict__ surf,
                                 const float4* __restrict__ nrm, float4* surfS, float4* nrmS) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = order[i];
  surfS[i] = surf[o];
  nrmS[i] = nrm[o];
}

// ---- host-side stage drivers ----------------------------------------------------------------------------------
int stage_compact(pcdb_ctx* ctx, int B, int64_t P, bool has_normals, bool has_rgb) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const float* xyz = w.in_xyz.as<float>();
  const float* nrm = has_normals ? w.in_nrm.as<float>() : nullptr;
  const unsigned* rgb = has_rgb ? w.in_rgb.as<unsigned>() : nullptr;
  PCDB_CUDA(w.flag_pt.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.flag_sf.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.pos_pt.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.pos_sf.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.pts4.ensure(sizeof(float4) * (P + 1)));
  PCDB_CUDA(w.surf4.ensure(sizeof(float4) * (P + 1)));
  PCDB_CUDA(w.snrm4.ensure(sizeof(float4) * (P + 1)));
  PCDB_CUDA(w.pts_cloud.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.surf_cloud.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.surf_local.ensure(sizeof(int) * (P + 1)));
  PCDB_CUDA(w.pts_off.ensure(sizeof(long long) * (B + 1)));
  PCDB_CUDA(w.surf_off.ensure(sizeof(long long) * (B + 1)));
  if (P > 0) {
    // the trailing flag is zero so that the exclusive scan's entry P is the total
    PCDB_CUDA(cudaMemsetAsync(w.flag_pt.as<int>() + P, 0, sizeof(int), st));
    PCDB_CUDA(cudaMemsetAsync(w.flag_sf.as<int>() + P, 0, sizeof(int), st));
    k_flags<<<cdiv(P, 256), 256, 0, st>>>(xyz, nrm, P, w.flag_pt.as<int>(), w.flag_sf.as<int>());
    PCDB_LAUNCH_CHECK();
  } else {
    PCDB_CUDA(cudaMemsetAsync(w.flag_pt.p, 0, sizeof(int), st));
    PCDB_CUDA(cudaMemsetAsync(w.flag_sf.p, 0, sizeof(int), st));
  }
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.flag_pt.as<int>(), w.pos_pt.as<int>(), P + 1));
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.flag_sf.as<int>(), w.pos_sf.as<int>(), P + 1));
  if (P > 0) {
    k_scatter<<<cdiv(P, 256), 256, 0, st>>>(xyz, nrm, rgb, w.cloud_off.as<long long>(), B, P, w.flag_pt.as<int>(),
                                             w.flag_sf.as<int>(), w.pos_pt.as<int>(), w.pos_sf.as<int>(),
                                             w.pts4.as<float4>(), w.pts_cloud.as<int>(), w.surf4.as<float4>(),
                                             w.snrm4.as<float4>(), w.surf_cloud.as<int>(), w.surf_local.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  k_compact_offsets<<<cdiv(B + 1, 128), 128, 0, st>>>(w.cloud_off.as<long long>(), B, w.pos_pt.as<int>(),
                                                       w.pos_sf.as<int>(), w.pts_off.as<long long>(),
                                                       w.surf_off.as<long long>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// bounding boxes over pts4[0..n_pts) (+ optional explicit keypoints) and per-cloud voxel / grid set-up
int stage_cloud_setup(pcdb_ctx* ctx, int B, int64_t n_pts, const float4* extra_kp, const int* extra_kp_cloud,
                      int64_t n_extra, float leaf, double grid_radius) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  PCDB_CUDA(w.minmax.ensure(sizeof(unsigned) * 6 * (B + 1)));
  PCDB_CUDA(w.cinfo.ensure(sizeof(CloudInfo) * (B + 1)));
  PCDB_CUDA(w.err_flag.ensure(sizeof(int) * 4));
  PCDB_CUDA(cudaMemsetAsync(w.err_flag.p, 0, sizeof(int) * 4, st));
  k_minmax_init<<<cdiv(B * 6, 128), 128, 0, st>>>(w.minmax.as<unsigned>(), B);
  PCDB_LAUNCH_CHECK();
  if (n_pts > 0) {
    k_minmax<<<cdiv(n_pts, 256), 256, 0, st>>>(w.pts4.as<float4>(), w.pts_cloud.as<int>(), n_pts,
                                                w.minmax.as<unsigned>());
    PCDB_LAUNCH_CHECK();
  }
  if (n_extra > 0) {
    k_minmax<<<cdiv(n_extra, 256), 256, 0, st>>>(extra_kp, extra_kp_cloud, n_extra, w.minmax.as<unsigned>());
    PCDB_LAUNCH_CHECK();
  }
  float inv_cell = 0.f;
  if (grid_radius > 0) inv_cell = 1.0f / (float)(grid_radius * 1.00002);
  ctx->grid_inv_cell = inv_cell;
  k_cloud_setup<<<cdiv(B, 128), 128, 0, st>>>(w.minmax.as<unsigned>(), B, leaf, inv_cell, w.cinfo.as<CloudInfo>(),
                                               w.err_flag.as<int>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// K1.  n_pts: exact number of compacted points (host knows it: the caller read pts_off back, or P when all finite).
// Leaves kp4 / kp_cloud / kp_off on the device and returns Q (one stream sync: sizes downstream buffers).
int stage_voxel_keypoints(pcdb_ctx* ctx, int B, int64_t n_pts, float leaf, int64_t* Q_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  PCDB_CUDA(w.kp_off.ensure(sizeof(long long) * (B + 1)));
  if (n_pts == 0) {
    PCDB_CUDA(cudaMemsetAsync(w.kp_off.p, 0, sizeof(long long) * (B + 1), st));
    *Q_out = 0;
    return PCDB_OK;
  }
  PCDB_CUDA(w.vkeys.ensure(sizeof(unsigned long long) * n_pts));
  PCDB_CUDA(w.vkeys2.ensure(sizeof(unsigned long long) * n_pts));
  PCDB_CUDA(w.vvals.ensure(sizeof(int) * n_pts));
  PCDB_CUDA(w.vvals2.ensure(sizeof(int) * n_pts));
  PCDB_CUDA(w.seg_head.ensure(sizeof(int) * (n_pts + 1)));
  PCDB_CUDA(w.seg_id.ensure(sizeof(int) * (n_pts + 2)));
  PCDB_CUDA(w.seg_start.ensure(sizeof(int) * (n_pts + 2)));
  PCDB_CUDA(w.kp4.ensure(sizeof(float4) * (n_pts + 1)));
  PCDB_CUDA(w.kp_cloud.ensure(sizeof(int) * (n_pts + 1)));
  k_voxel_keys<<<cdiv(n_pts, 256), 256, 0, st>>>(w.pts4.as<float4>(), w.pts_cloud.as<int>(), n_pts,
                                                  w.cinfo.as<CloudInfo>(), leaf, w.vkeys.as<unsigned long long>(),
                                                  w.vvals.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_sort_pairs_u64(ctx, w.vkeys.as<unsigned long long>(), w.vkeys2.as<unsigned long long>(),
                                   w.vvals.as<int>(), w.vvals2.as<int>(), n_pts, 32 + ceil_log2(B + 1)));
  k_heads_u64<<<cdiv(n_pts, 256), 256, 0, st>>>(w.vkeys2.as<unsigned long long>(), n_pts, w.seg_head.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_CUDA(cudaMemsetAsync(w.seg_head.as<int>() + n_pts, 0, sizeof(int), st));
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.seg_head.as<int>(), w.seg_id.as<int>(), n_pts + 1));
  k_seg_starts<<<cdiv(n_pts + 1, 256), 256, 0, st>>>(w.seg_head.as<int>(), w.seg_id.as<int>(), n_pts,
                                                      w.seg_start.as<int>());
  PCDB_LAUNCH_CHECK();
  k_offsets_from_keys<<<cdiv(B + 1, 128), 128, 0, st>>>(w.vkeys2.as<unsigned long long>(), w.seg_id.as<int>(), n_pts,
                                                         B, 32, w.kp_off.as<long long>());
  PCDB_LAUNCH_CHECK();
  int Q = 0;
  PCDB_TRY(pcdb_read_small(ctx, &Q, w.seg_id.as<int>() + n_pts, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  if (Q > 0) {
    k_centroids<<<cdiv(Q, 128), 128, 0, st>>>(w.pts4.as<float4>(), w.vvals2.as<int>(), w.seg_start.as<int>(), Q,
                                               w.pts_cloud.as<int>(), w.kp4.as<float4>(), w.kp_cloud.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  *Q_out = Q;
  return PCDB_OK;
}

// K2.  Sorts the surface by search-grid cell and groups the keypoints by cell (the work items of the SHOT kernel).
int stage_grid(pcdb_ctx* ctx, int B, int64_t n_surf, int64_t Q, bool color) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const float inv_cell = ctx->grid_inv_cell;
  PCDB_CUDA(w.gkeys.ensure(sizeof(unsigned long long) * (n_surf + 1)));
  PCDB_CUDA(w.gkeys2.ensure(sizeof(unsigned long long) * (n_surf + 1)));
  PCDB_CUDA(w.gvals.ensure(sizeof(int) * (n_surf + 1)));
  PCDB_CUDA(w.gvals2.ensure(sizeof(int) * (n_surf + 1)));
  PCDB_CUDA(w.surfS4.ensure(sizeof(float4) * (n_surf + 1)));
  PCDB_CUDA(w.snrmS4.ensure(sizeof(float4) * (n_surf + 1)));
  if (n_surf > 0) {
    k_grid_keys<<<cdiv(n_surf, 256), 256, 0, st>>>(w.surf4.as<float4>(), w.surf_cloud.as<int>(), n_surf,
                                                    w.cinfo.as<CloudInfo>(), inv_cell,
                                                    w.gkeys.as<unsigned long long>(), w.gvals.as<int>());
    PCDB_LAUNCH_CHECK();
    PCDB_TRY(pcdb_cub_sort_pairs_u64(ctx, w.gkeys.as<unsigned long long>(), w.gkeys2.as<unsigned long long>(),
                                     w.gvals.as<int>(), w.gvals2.as<int>(), n_surf, 48 + ceil_log2(B + 1)));
    k_gather_surface<<<cdiv(n_surf, 256), 256, 0, st>>>(w.gvals2.as<int>(), n_surf, w.surf4.as<float4>(),
                                                         w.snrm4.as<float4>(), w.surfS4.as<float4>(),
                                                         w.snrmS4.as<float4>());
    PCDB_LAUNCH_CHECK();
  }
  (void)color;
  if (Q == 0) return PCDB_OK;
  PCDB_CUDA(w.kkeys.ensure(sizeof(unsigned long long) * (Q + 1)));
  PCDB_CUDA(w.kkeys2.ensure(sizeof(unsigned long long) * (Q + 1)));
  PCDB_CUDA(w.kvals.ensure(sizeof(int) * (Q + 1)));
  PCDB_CUDA(w.kvals2.ensure(sizeof(int) * (Q + 1)));
  PCDB_CUDA(w.item_head.ensure(sizeof(int) * (Q + 1)));
  PCDB_CUDA(w.item_id.ensure(sizeof(int) * (Q + 2)));
  PCDB_CUDA(w.item_start.ensure(sizeof(int) * (Q + 2)));
  k_grid_keys<<<cdiv(Q, 256), 256, 0, st>>>(w.kp4.as<float4>(), w.kp_cloud.as<int>(), Q, w.cinfo.as<CloudInfo>(),
                                             inv_cell, w.kkeys.as<unsigned long long>(), w.kvals.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_sort_pairs_u64(ctx, w.kkeys.as<unsigned long long>(), w.kkeys2.as<unsigned long long>(),
                                   w.kvals.as<int>(), w.kvals2.as<int>(), Q, 48 + ceil_log2(B + 1)));
  // whole-cloud items are not cut any more: CTAs that run out of items help with the ones still in progress (k_shot
  // draws keypoints from a global per-item counter), so one staging serves a whole cloud and the launch has no item tail
  const int item_kp = 1 << 30;
  k_item_heads<<<cdiv(Q, 256), 256, 0, st>>>(w.kkeys2.as<unsigned long long>(), Q, w.surf_off.as<long long>(), item_kp,
                                             w.item_head.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_CUDA(cudaMemsetAsync(w.item_head.as<int>() + Q, 0, sizeof(int), st));
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.item_head.as<int>(), w.item_id.as<int>(), Q + 1));
  k_seg_starts<<<cdiv(Q + 1, 256), 256, 0, st>>>(w.item_head.as<int>(), w.item_id.as<int>(), Q,
                                                  w.item_start.as<int>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;  // the item count stays on the device (item_id[Q]); the SHOT kernel reads it there
}
