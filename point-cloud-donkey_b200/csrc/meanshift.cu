// meanshift.cu — K9 mean-shift maxima search and K10 per-maximum reduction / label pick.
//
// Replaces (reference paths under src/implicit_shape_model/voting/):
//   VotingMeanShift::iFindMaxima, createSeeds, iDoMeanShift, computeMeanShift, estimateDensity(AndReweightVotes)
//                                   voting_mean_shift.cpp:39-177,201-244,247-328,331-376,431-481
//   MaximaHandler::averageNeighborMaxima / suppressNeighborMaxima   maxima_handler.cpp:94-157,51-92
//   Voting::findMaxima (per-maximum reduce, sort, normalise, threshold, best-K)   voting.cpp:79-328
//   label pick of eval_tool         src/eval_tool/eval_classification.cpp:412-417
// The reference runs this per class, sequentially, with a kd-tree per class; here all (cloud, class) vote groups of a
// batch are processed at once: votes are ordered by (cloud, class) with one radix sort, seeds come from a second
// sort of the bin keys, every seed iterates in its own warp over the group's votes (16 B per vote, L1/L2 resident),
// and the small order-dependent tails (average / suppress / cumulative re-weighting) run one warp or block per
// group exactly in the reference's order.  Membership tests are the kd-tree's float test d^2 < float(h*h).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "stages.h"

namespace {

struct MsP {
  float h_cfg, thr;        // Voting.Bandwidth, Voting.Threshold
  int max_iter, kernel, suppression, min_votes, best_k, average_rotation, n_classes;
  float min_threshold;
  int cross_class_filter;  // PCDB_MAXFILTER_* applied per cloud when !single_object_mode
  int radius_type;         // PCDB_RADIUS_*
  int single_max;          // PCDB_SOMAX_* when single_object_mode, else 0 (mean shift)
  // (h, r2 = float(double h * h), hh = float h * h, bin = 2 h / sqrt 2) of every (cloud, class) vote group: the search
  // distance is per class (BinOrBandwidthType) or even per group (single-object max types), k_group_params fills it
  const float4* grp;
  const float* cls_h;      // [n_classes] MaximaHandler::getSearchDistForClass for the cross-class filters
};

__device__ __forceinline__ float4 make_hp(float h) {
  return make_float4(h, (float)((double)h * (double)h), __fmul_rn(h, h), __fdiv_rn(__fmul_rn(h, 2.0f), sqrtf(2.f)));
}

__device__ __forceinline__ float ms_profile(int kernel, float u) {
  // kernelGaussian: float profile = exp(-0.5 * x) evaluated in double (voting_mean_shift.cpp:396-400)
  return kernel == PCDB_KERNEL_GAUSSIAN ? (float)exp(-0.5 * (double)u) : 1.0f;
}
__device__ __forceinline__ float ms_g(int kernel, float u, float w) {
  // g = -kernelDerivative(u) * w ; derivative = -0.5f * profile (Gaussian) or 1 (Uniform)  (:359-362,:402-417)
  float der = kernel == PCDB_KERNEL_GAUSSIAN ? __fmul_rn(-0.5f, ms_profile(kernel, u)) : 1.0f;
  return __fmul_rn(-der, w);
}
__device__ __forceinline__ float norm3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  float d0 = __fsub_rn(ax, bx), d1 = __fsub_rn(ay, by), d2 = __fsub_rn(az, bz);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
}

__global__ void k_votes_unpack(const pcdb_vote* __restrict__ votes, long long V, const long long* __restrict__ off,
                               int B, float4* pw, int* cloud) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  pw[v] = make_float4(votes[v].position[0], votes[v].position[1], votes[v].position[2], votes[v].weight);
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  cloud[v] = lo;
}

__global__ void k_vote_keys(const pcdb_vote* __restrict__ votes, const int* __restrict__ cloud, long long V,
                            int n_classes, unsigned* keys, int* vals) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  keys[v] = (unsigned)cloud[v] * (unsigned)n_classes + votes[v].class_id;
  vals[v] = (int)v;
}

__global__ void k_gather_pw(const int* __restrict__ ord, long long V, const float4* __restrict__ pw, MsP P,
                            const int* __restrict__ seg_id, const int* __restrict__ seg_head, float4* pwS,
                            float* wwork, unsigned long long* seedkey, int* ident) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= V) return;
  float4 p = pw[ord[t]];
  pwS[t] = p;
  wwork[t] = p.w;
  const float bin = P.grp[seg_id[t] + seg_head[t] - 1].w;
  // createSeeds: key = (int)floor(pos / binSize + 0.5)  (float division, double add/floor)  (:431-449)
  long long kx = (long long)floor((double)__fdiv_rn(p.x, bin) + 0.5);
  long long ky = (long long)floor((double)__fdiv_rn(p.y, bin) + 0.5);
  long long kz = (long long)floor((double)__fdiv_rn(p.z, bin) + 0.5);
  const long long lim = (1ll << 20) - 1;
  kx = max(-lim, min(lim, kx)) + (1ll << 20);
  ky = max(-lim, min(lim, ky)) + (1ll << 20);
  kz = max(-lim, min(lim, kz)) + (1ll << 20);
  seedkey[t] = ((unsigned long long)kz << 42) | ((unsigned long long)ky << 21) | (unsigned long long)kx;
  ident[t] = (int)t;
}

__global__ void k_heads_u32(const unsigned* __restrict__ keys, long long n, int* head) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i > n) return;
  head[i] = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}

// seg_start[id] = i, seg_key[id] = key for every head; seg_start[nseg] = n
__global__ void k_seg_info(const int* __restrict__ head, const int* __restrict__ id, const unsigned* __restrict__ keys,
                           long long n, int* seg_start, unsigned* seg_key) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    seg_start[id[n]] = (int)n;
    return;
  }
  if (head[i]) {
    seg_start[id[i]] = (int)i;
    seg_key[id[i]] = keys[i];
  }
}

// group of sorted vote t = (number of heads up to and including t) - 1; seg_id is the EXCLUSIVE scan of the heads
__global__ void k_gather_segkey(const int* __restrict__ t1, const int* __restrict__ seg_id,
                                const int* __restrict__ head, long long V, unsigned* segkey, int* ident) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= V) return;
  const int t = t1[i];
  segkey[i] = (unsigned)(seg_id[t] + head[t] - 1);
  ident[i] = (int)i;
}

__global__ void k_seed_heads(const unsigned* __restrict__ segS, const int* __restrict__ val2,
                             const unsigned long long* __restrict__ sk1, long long V, unsigned long long* seedS,
                             int* head) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i > V) return;
  if (i == V) {
    head[i] = 0;
    return;
  }
  unsigned long long k = sk1[val2[i]];
  seedS[i] = k;
  bool h = true;
  if (i > 0) h = segS[i] != segS[i - 1] || k != sk1[val2[i - 1]];
  head[i] = h ? 1 : 0;
}

// seed list (key, segment) and per-segment first seed id; seed_first[nseg] = n_seeds
__global__ void k_seed_list(const int* __restrict__ head, const int* __restrict__ id,
                            const unsigned long long* __restrict__ seedS, const unsigned* __restrict__ segS,
                            long long V, const int* __restrict__ nseg_ptr, unsigned long long* seed_key,
                            int* seed_seg, int* seed_first) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i > V) return;
  if (i == V) {
    seed_first[*nseg_ptr] = id[V];
    return;
  }
  if (head[i]) {
    seed_key[id[i]] = seedS[i];
    seed_seg[id[i]] = (int)segS[i];
    if (i == 0 || segS[i] != segS[i - 1]) seed_first[segS[i]] = id[i];
  }
}

// ---- per-group search distance (voting_mean_shift.cpp:47-49,124-155; maxima_handler.cpp:509-521) -------------------
// One warp per (cloud, class) group.  Mean-shift mode: h = the class' search distance.  Single-object max types (no
// mean shift): BandwidthVotes = the same, ModelRadiusVotes = the cloud's model radius, VotingSpaceVotes = the distance
// of the farthest vote of the group from the cloud's centroid; the group's single maximum sits at that centroid.
__global__ void k_group_params(const int* __restrict__ nseg_ptr, const unsigned* __restrict__ seg_key,
                               const int* __restrict__ seg_start, const float4* __restrict__ pw,
                               const int* __restrict__ ord, MsP P, const float4* __restrict__ cloud_cm, float4* grp,
                               int* seed_first, float4* max_pos, int* n_max) {
  const int seg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nseg = *nseg_ptr;
  if (seg > nseg) return;
  if (seg == nseg) {
    if (lane == 0 && P.single_max) seed_first[seg] = seg;
    return;
  }
  const unsigned cls = seg_key[seg] % (unsigned)P.n_classes, cloud = seg_key[seg] / (unsigned)P.n_classes;
  float h = P.cls_h[cls];
  if (P.single_max) {
    const float4 cm = cloud_cm[cloud];  // centroid xyz, model radius
    if (P.single_max == PCDB_SOMAX_MODEL_RADIUS) h = cm.w;
    if (P.single_max == PCDB_SOMAX_VOTING_SPACE) {  // SingleObjectHelper::getVotingSpaceSize: sqrt(max squaredNorm)
      float mx = 0.f;
      for (int t = seg_start[seg] + lane; t < seg_start[seg + 1]; t += 32) {
        const float4 p = pw[ord[t]];
        mx = fmaxf(mx, sqdist3_rn(p.x, p.y, p.z, cm.x, cm.y, cm.z));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      h = __fsqrt_rn(mx);
    }
    if (lane == 0) {
      seed_first[seg] = seg;  // slot of the group's single maximum
      max_pos[seg] = make_float4(cm.x, cm.y, cm.z, 1.f);
      n_max[seg] = 1;
    }
  }
  if (lane == 0) grp[seg] = make_hp(h);
}

// pcl::compute3DCentroid of the cloud Voting::findMaxima is handed (the surface points: finite points with finite
// normals) — sequential float sums and one division, as the Eigen::Vector4f accumulation does — and
// SingleObjectHelper::getModelRadius (farthest point from it).  One warp per cloud; lane 0 owns the ordered sums.
__global__ void k_cloud_centroid(const float4* __restrict__ surf, const long long* __restrict__ surf_off, int B,
                                 float4* cloud_cm) {
  const int b = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long s0 = surf_off[b], s1 = surf_off[b + 1];
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (lane == 0) {
    for (long long i = s0; i < s1; ++i) {
      const float4 p = surf[i];
      cx = __fadd_rn(cx, p.x);
      cy = __fadd_rn(cy, p.y);
      cz = __fadd_rn(cz, p.z);
    }
    const float n = (float)(s1 - s0);
    cx = __fdiv_rn(cx, n);
    cy = __fdiv_rn(cy, n);
    cz = __fdiv_rn(cz, n);
  }
  cx = __shfl_sync(0xffffffffu, cx, 0);
  cy = __shfl_sync(0xffffffffu, cy, 0);
  cz = __shfl_sync(0xffffffffu, cz, 0);
  float r = 0.f;
  for (long long i = s0 + lane; i < s1; i += 32) {
    const float4 p = surf[i];
    r = fmaxf(r, norm3_rn(p.x, p.y, p.z, cx, cy, cz));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  if (lane == 0) cloud_cm[b] = make_float4(cx, cy, cz, r);
}

// ---- iDoMeanShift / computeMeanShift: one warp per seed -------------------------------------------------------
__global__ void __launch_bounds__(128) k_meanshift(const int* __restrict__ n_seeds_ptr,
                                                   const unsigned long long* __restrict__ seed_key,
                                                   const int* __restrict__ seed_seg,
                                                   const int* __restrict__ seg_start, const float4* __restrict__ pwS,
                                                   MsP P, float4* centers) {
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= *n_seeds_ptr) return;
  const unsigned long long key = seed_key[warp];
  const int seg = seed_seg[warp];
  const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
  const float4 hp = P.grp[seg];
  // seed position = key * binSize (int -> float, float multiply)  (:466-470)
  float cx = __fmul_rn((float)((long long)(key & 0x1fffff) - (1ll << 20)), hp.w);
  float cy = __fmul_rn((float)((long long)((key >> 21) & 0x1fffff) - (1ll << 20)), hp.w);
  float cz = __fmul_rn((float)((long long)((key >> 42) & 0x1fffff) - (1ll << 20)), hp.w);
  int iter = 0;
  float diff = 0.f;
  bool skip = false;
  do {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    double tw = 0.0;
    int cnt = 0;
    for (int t = s0 + lane; t < s1; t += 32) {
      float4 p = pwS[t];
      float d2 = sqdist3_rn(cx, cy, cz, p.x, p.y, p.z);
      if (d2 < hp.y) {
        float u = __fdiv_rn(d2, hp.z);
        float g = ms_g(P.kernel, u, p.w);
        sx = __fadd_rn(sx, __fmul_rn(g, p.x));
        sy = __fadd_rn(sy, __fmul_rn(g, p.y));
        sz = __fadd_rn(sz, __fmul_rn(g, p.z));
        tw += (double)g;
        ++cnt;
      }
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    tw = warp_sum(tw);
    cnt = warp_sum(cnt);
    if (cnt == 0) {
      skip = true;
      break;
    }
    if (tw != 0.0) {
      float twf = (float)tw;
      sx = __fdiv_rn(sx, twf);
      sy = __fdiv_rn(sy, twf);
      sz = __fdiv_rn(sz, twf);
    }
    diff = norm3_rn(cx, cy, cz, sx, sy, sz);
    cx = sx;
    cy = sy;
    cz = sz;
    ++iter;
  } while (diff > P.thr && iter <= P.max_iter);
  if (lane == 0) centers[warp] = make_float4(cx, cy, cz, skip ? 0.f : 1.f);
}

// density of one position over a vote group, warp-cooperative (estimateDensity :247-285)
__device__ float group_density(const float4* __restrict__ pwS, const float* __restrict__ w, int s0, int s1, float x,
                               float y, float z, const MsP& P, const float4 hp, int lane) {
  float dens = 0.f;
  for (int t = s0 + lane; t < s1; t += 32) {
    float4 p = pwS[t];
    float d2 = sqdist3_rn(x, y, z, p.x, p.y, p.z);
    if (d2 < hp.y) dens = __fadd_rn(dens, __fmul_rn(ms_profile(P.kernel, __fdiv_rn(d2, hp.z)), w[t]));
  }
  return warp_sum(dens);
}

// ---- per-group tail: densities, average, suppress (one warp per (cloud, class) group) -------------------------
__global__ void __launch_bounds__(32) k_ms_tail(const int* __restrict__ nseg_ptr, const int* __restrict__ seg_start,
                                                const int* __restrict__ seed_first,
                                                const float4* __restrict__ centers, const float4* __restrict__ pwS,
                                                const float* __restrict__ w0, MsP P, float4* cen, float4* cen2,
                                                float* dens, int* flag, float4* max_pos, int* n_max) {
  const int seg = blockIdx.x, lane = threadIdx.x;
  if (seg >= *nseg_ptr) return;
  const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
  const int f0 = seed_first[seg], f1 = seed_first[seg + 1];
  const float4 hp = P.grp[seg];
  float4* C = cen + f0;
  float4* C2 = cen2 + f0;
  float* Dn = dens + f0;
  int* Fl = flag + f0;
  // cluster centres = seeds that found votes, in seed order
  int M = 0;
  for (int base = f0; base < f1; base += 32) {
    int i = base + lane;
    float4 c = i < f1 ? centers[i] : make_float4(0, 0, 0, 0);
    bool ok = i < f1 && c.w != 0.f;
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) C[M + __popc(m & ((1u << lane) - 1))] = c;
    M += __popc(m);
  }
  __syncwarp();
  for (int i = 0; i < M; ++i) {
    float4 c = C[i];
    float d = group_density(pwS, w0, s0, s1, c.x, c.y, c.z, P, hp, lane);
    if (lane == 0) Dn[i] = d;
  }
  __syncwarp();
  if (P.suppression == PCDB_SUPPRESS_AVERAGE) {
    // averageNeighborMaxima (maxima_handler.cpp:94-157): Fl = duplicate flag
    for (int i = lane; i < M; i += 32) Fl[i] = 0;
    __syncwarp();
    for (int k = 0; k < M; ++k) {
      float4 a = C[k];
      if (Fl[k]) {  // already merged into an earlier maximum: its own group has size 1 -> re-emitted as is
        if (lane == 0) C2[k] = a;
        __syncwarp();
        continue;
      }
      // mark later, not-yet-duplicate centres closer than h; accumulate in index order
      float ax = 0.f, ay = 0.f, az = 0.f, sd = 0.f;
      int members = 1;
      if (lane == 0) {
        float dk = Dn[k];
        ax = __fmul_rn(a.x, dk);
        ay = __fmul_rn(a.y, dk);
        az = __fmul_rn(a.z, dk);
        sd = dk;
      }
      for (int base = k + 1; base < M; base += 32) {
        int j = base + lane;
        bool hit = false;
        float4 b = make_float4(0, 0, 0, 0);
        if (j < M && !Fl[j]) {
          b = C[j];
          hit = norm3_rn(a.x, a.y, a.z, b.x, b.y, b.z) < hp.x;
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) Fl[j] = 1;
        members += __popc(m);
        // sequential accumulation in ascending j (the reference's order) by lane 0
        while (m) {
          int l = __ffs(m) - 1;
          m &= m - 1;
          float bx = __shfl_sync(0xffffffffu, b.x, l), by = __shfl_sync(0xffffffffu, b.y, l),
                bz = __shfl_sync(0xffffffffu, b.z, l);
          if (lane == 0) {
            float dj = Dn[base + l];
            ax = __fadd_rn(ax, __fmul_rn(bx, dj));
            ay = __fadd_rn(ay, __fmul_rn(by, dj));
            az = __fadd_rn(az, __fmul_rn(bz, dj));
            sd = __fadd_rn(sd, dj);
          }
        }
      }
      if (lane == 0) {
        if (members == 1)
          C2[k] = a;
        else
          C2[k] = make_float4(__fdiv_rn(ax, sd), __fdiv_rn(ay, sd), __fdiv_rn(az, sd), 1.f);
      }
      __syncwarp();
    }
    for (int i = lane; i < M; i += 32) C[i] = C2[i];
    __syncwarp();
    for (int i = 0; i < M; ++i) {
      float4 c = C[i];
      float d = group_density(pwS, w0, s0, s1, c.x, c.y, c.z, P, hp, lane);
      if (lane == 0) Dn[i] = d;
    }
    __syncwarp();
  }
  // suppressNeighborMaxima (maxima_handler.cpp:51-92): greedy, max density first (first index on ties)
  int nm = 0;
  while (true) {
    float bd = -1.f;
    int bi = 0x7fffffff;
    for (int i = lane; i < M; i += 32) {
      float d = Dn[i];
      if (d > bd) {  // strictly greater keeps the first index within a lane's ascending scan
        bd = d;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, bd, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od > bd || (od == bd && oi < bi)) {
        bd = od;
        bi = oi;
      }
    }
    if (!(bd != -1.f) || bi == 0x7fffffff) break;
    float4 c = C[bi];
    if (lane == 0) max_pos[f0 + nm] = c;
    ++nm;
    for (int i = lane; i < M; i += 32) {
      float4 b = C[i];
      if (i == bi || norm3_rn(c.x, c.y, c.z, b.x, b.y, b.z) < hp.x) Dn[i] = -1.f;
    }
    __syncwarp();
  }
  if (lane == 0) n_max[seg] = nm;
}

// maximum list: for every group with maxima, (group, position); one thread per group
__global__ void k_max_list(const int* __restrict__ nseg_ptr, const int* __restrict__ n_max,
                           const int* __restrict__ max_off, const int* __restrict__ seed_first,
                           const float4* __restrict__ max_pos, float4* mpos, int* mseg) {
  int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= *nseg_ptr) return;
  for (int m = 0; m < n_max[seg]; ++m) {
    mpos[max_off[seg] + m] = max_pos[seed_first[seg] + m];
    mseg[max_off[seg] + m] = seg;
  }
}

// members per maximum (geometric test only), one warp per maximum
__global__ void k_member_count(const int* __restrict__ M_ptr, const float4* __restrict__ mpos,
                               const int* __restrict__ mseg, const int* __restrict__ seg_start,
                               const float4* __restrict__ pwS, MsP P, int* mem_cnt) {
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int M = *M_ptr;
  if (m > M) return;
  if (m == M) {
    if (lane == 0) mem_cnt[m] = 0;
    return;
  }
  float4 c = mpos[m];
  int seg = mseg[m];
  const float r2 = P.grp[seg].y;
  int cnt = 0;
  for (int t = seg_start[seg] + lane; t < seg_start[seg + 1]; t += 32) {
    float4 p = pwS[t];
    if (sqdist3_rn(c.x, c.y, c.z, p.x, p.y, p.z) < r2) ++cnt;
  }
  cnt = warp_sum(cnt);
  if (lane == 0) mem_cnt[m] = cnt;
}

// estimateDensityAndReweightVotes (:289-328), cumulative over the group's maxima in order; one warp per group
__global__ void __launch_bounds__(32) k_ms_reweight(const int* __restrict__ nseg_ptr, const int* __restrict__ n_max,
                                                    const int* __restrict__ max_off,
                                                    const int* __restrict__ seg_start,
                                                    const float4* __restrict__ mpos, const float4* __restrict__ pwS,
                                                    const int* __restrict__ ord, const int* __restrict__ mem_off,
                                                    MsP P, float* wwork, long long* mem_idx, float* mem_w) {
  const int seg = blockIdx.x, lane = threadIdx.x;
  if (seg >= *nseg_ptr) return;
  const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
  const float4 hp = P.grp[seg];
  for (int m = max_off[seg]; m < max_off[seg] + n_max[seg]; ++m) {
    float4 c = mpos[m];
    int o = mem_off[m];
    for (int base = s0; base < s1; base += 32) {
      int t = base + lane;
      bool in = false;
      float nw = 0.f;
      if (t < s1) {
        float4 p = pwS[t];
        float d2 = sqdist3_rn(c.x, c.y, c.z, p.x, p.y, p.z);
        if (d2 < hp.y) {
          in = true;
          nw = __fmul_rn(ms_profile(P.kernel, __fdiv_rn(d2, hp.z)), wwork[t]);
          wwork[t] = nw;
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, in);
      if (in) {
        int pos = o + __popc(mask & ((1u << lane) - 1));
        mem_idx[pos] = ord[t];
        mem_w[pos] = nw;
      }
      o += __popc(mask);
    }
    __syncwarp();
  }
}

// symmetric 4x4 Jacobi (double) for Utils::quatWeightedAverage (utils/utils.cpp:617-665)
__device__ void quat_avg_eig(const float S[4][4], float q[4]) {
  double A[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      A[i][j] = (double)S[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p)
      for (int r = p + 1; r < 4; ++r) off += fabs(A[p][r]);
    if (off == 0) break;
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        double apq = A[p][r];
        if (apq == 0.0) continue;
        double theta = (A[r][r] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (!isfinite(theta)) t = 0.0;
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          double akp = A[k][p], akq = A[k][r];
          A[k][p] = c * akp - s * akq;
          A[k][r] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          double apk = A[p][k], aqk = A[r][k];
          A[p][k] = c * apk - s * aqk;
          A[r][k] = s * apk + c * aqk;
        }
        A[p][r] = A[r][p] = 0.0;
        for (int k = 0; k < 4; ++k) {
          double vkp = V[k][p], vkq = V[k][r];
          V[k][p] = c * vkp - s * vkq;
          V[k][r] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  float maxEv = 0.f;
  for (int i = 0; i < 4; ++i)
    if ((float)A[i][i] > maxEv) {
      maxEv = (float)A[i][i];
      best = i;
    }
  double sgn = V[0][best] < 0 ? -1.0 : 1.0;
  for (int i = 0; i < 4; ++i) q[i] = (float)(sgn * V[i][best]);
}

// ---- K10: per-maximum reduction (voting.cpp:130-236), one warp per maximum -------------------------------------
constexpr int kInstSlots = 256;  // per-warp hash set of the distinct instance ids of one maximum (k_max_reduce)
__global__ void k_max_reduce(const int* __restrict__ M_ptr, const float4* __restrict__ mpos,
                             const int* __restrict__ mseg, const unsigned* __restrict__ seg_key,
                             const int* __restrict__ mem_off, const long long* __restrict__ mem_idx,
                             const float* __restrict__ mem_w, const pcdb_vote* __restrict__ votes, MsP P,
                             pcdb_maximum* out) {
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (m >= *M_ptr) return;
  const int o0 = mem_off[m], o1 = mem_off[m + 1], n = o1 - o0;
  float wsum = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = o0 + lane; i < o1; i += 32) {
    const pcdb_vote& v = votes[mem_idx[i]];
    float w = mem_w[i];
    wsum = __fadd_rn(wsum, w);
    sx = __fadd_rn(sx, __fmul_rn(w, v.bbox_size[0]));
    sy = __fadd_rn(sy, __fmul_rn(w, v.bbox_size[1]));
    sz = __fadd_rn(sz, __fmul_rn(w, v.bbox_size[2]));
  }
  wsum = warp_sum(wsum);
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  sz = warp_sum(sz);
  // instance with the largest summed weight (std::map order: ascending id, strict >)  (:140-167).  Every instance's
  // weights are summed in member order (the order the reference adds them to its map entry).  The distinct ids of the
  // maximum are collected in a small per-warp hash set first; then every lane owns one id and walks the members once:
  // n * ceil(I / 32) steps instead of the n^2 of "every member rescans all members" (a dense scene puts 10^4..10^5
  // votes into one maximum but only a few dozen training instances).
  __shared__ unsigned s_set[4][kInstSlots];
  __shared__ int s_nset[4];
  unsigned* set = s_set[threadIdx.x >> 5];
  for (int i = lane; i < kInstSlots; i += 32) set[i] = 0xffffffffu;
  if (lane == 0) s_nset[threadIdx.x >> 5] = 0;
  __syncwarp();
  bool overflow = false;
  for (int i = o0 + lane; i < o1; i += 32) {
    const unsigned inst = votes[mem_idx[i]].instance_id;
    if (inst == 0xffffffffu) { overflow = true; continue; }  // the empty marker itself: slow path
    unsigned h = (inst * 2654435761u) % kInstSlots;
    for (int probe = 0; probe < kInstSlots; ++probe) {
      const unsigned prev = atomicCAS(&set[h], 0xffffffffu, inst);
      if (prev == 0xffffffffu) { atomicAdd(&s_nset[threadIdx.x >> 5], 1); break; }
      if (prev == inst) break;
      h = (h + 1) % kInstSlots;
    }
    if (s_nset[threadIdx.x >> 5] > kInstSlots / 2) { overflow = true; break; }
  }
  overflow = __any_sync(0xffffffffu, overflow);
  __syncwarp();
  float bw = 0.f;
  unsigned bid = 0xffffffffu;
  float w_inst0 = 0.f;
  if (!overflow) {
    for (int base = 0; base < kInstSlots; base += 32) {
      const unsigned inst = set[base + lane];
      if (!__any_sync(0xffffffffu, inst != 0xffffffffu)) continue;
      float tot = 0.f;
      for (int j = o0; j < o1; ++j)
        if (votes[mem_idx[j]].instance_id == inst) tot = __fadd_rn(tot, mem_w[j]);
      if (inst == 0xffffffffu) continue;
      if (inst == 0) w_inst0 = tot;
      if (tot > bw || (tot == bw && tot > 0.f && inst < bid)) {
        bw = tot;
        bid = inst;
      }
    }
  } else {
    for (int i = o0 + lane; i < o1; i += 32) {
      unsigned inst = votes[mem_idx[i]].instance_id;
      float tot = 0.f;
      for (int j = o0; j < o1; ++j)
        if (votes[mem_idx[j]].instance_id == inst) tot = __fadd_rn(tot, mem_w[j]);
      if (inst == 0) w_inst0 = tot;
      if (tot > bw || (tot == bw && tot > 0.f && inst < bid)) {
        bw = tot;
        bid = inst;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ow = __shfl_xor_sync(0xffffffffu, bw, o);
    unsigned oi = __shfl_xor_sync(0xffffffffu, bid, o);
    float o0w = __shfl_xor_sync(0xffffffffu, w_inst0, o);
    if (o0w != 0.f) w_inst0 = o0w;
    if (ow > bw || (ow == bw && ow > 0.f && oi < bid)) {
      bw = ow;
      bid = oi;
    }
  }
  if (bid == 0xffffffffu) {  // no instance weight > 0: the reference leaves the id uninitialised; defined as 0
    bid = 0;
    bw = w_inst0;
  }
  // quaternion scatter matrix with weights / maxWeight  (:186-215)
  float S[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) S[i][j] = 0.f;
  if (P.average_rotation) {
    for (int i = o0 + lane; i < o1; i += 32) {
      const pcdb_vote& v = votes[mem_idx[i]];
      float w = __fdiv_rn(mem_w[i], wsum);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          S[a][b] = __fadd_rn(S[a][b], __fmul_rn(__fmul_rn(w, v.bbox_quat[a]), v.bbox_quat[b]));
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) S[a][b] = warp_sum(S[a][b]);
  }
  if (lane == 0) {
    pcdb_maximum r;
    float4 c = mpos[m];
    r.position[0] = c.x;
    r.position[1] = c.y;
    r.position[2] = c.z;
    r.weight = wsum;
    r.raw_weight = wsum;
    r.class_id = seg_key[mseg[m]] % (unsigned)P.n_classes;
    r.instance_id = bid;
    r.instance_weight = bw;
    r.bbox_quat[0] = 1.f;
    r.bbox_quat[1] = r.bbox_quat[2] = r.bbox_quat[3] = 0.f;
    if (P.average_rotation && n > 0) quat_avg_eig(S, r.bbox_quat);
    r.bbox_size[0] = __fdiv_rn(sx, wsum);
    r.bbox_size[1] = __fdiv_rn(sy, wsum);
    r.bbox_size[2] = __fdiv_rn(sz, wsum);
    r.n_votes = n;
    r.vote_begin = o0;
    out[m] = r;
  }
}

// ---- per-cloud sort / normalise / threshold / best-K / label (voting.cpp:272-323), one warp per cloud ----------
__global__ void k_cloud_finalize(int B, const int* __restrict__ nseg_ptr, const unsigned* __restrict__ seg_key,
                                 const int* __restrict__ max_off, const int* __restrict__ mem_off, pcdb_maximum* raw,
                                 MsP P, int* flag, int* close_list, int* src_dst, pcdb_maximum* sorted, int* kept,
                                 int* first, int* label) {
  const int b = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int nseg = *nseg_ptr;
  // groups of this cloud: keys in [b*C, (b+1)*C)
  int lo = 0, hi = nseg;
  unsigned klo = (unsigned)b * (unsigned)P.n_classes, khi = klo + (unsigned)P.n_classes;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (seg_key[mid] < klo) lo = mid + 1; else hi = mid;
  }
  int g0 = lo;
  hi = nseg;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (seg_key[mid] < khi) lo = mid + 1; else hi = mid;
  }
  int g1 = lo;
  const int m0 = max_off[g0], m1 = max_off[g1];
  if (P.cross_class_filter == PCDB_MAXFILTER_SIMPLE) {
    // MaximaHandler::suppressNeighborMaxima2 (maxima_handler.cpp:227-268): keep the heaviest pending maximum (the first
    // of equal ones), drop every maximum of any class closer than the radius, repeat.  The radius is
    // MaximaHandler::m_radius as iFindMaxima left it (voting_mean_shift.cpp:47-49): Voting.Bandwidth for
    // BinOrBandwidthType "Config", else the search distance of the class processed BEFORE the last one.  The list is
    // short and the loop order dependent: one lane.  Dropped maxima get n_votes = 0 and fall out below.
    const float radius = (P.radius_type != PCDB_RADIUS_CONFIG && g1 - g0 >= 2) ? P.grp[g1 - 2].x : P.h_cfg;
    if (lane == 0) {
      for (int i = m0; i < m1; ++i) flag[i] = (raw[i].n_votes >= P.min_votes && raw[i].n_votes > 0) ? 0 : 2;
      for (;;) {
        int best = -1;
        for (int i = m0; i < m1; ++i)
          if (flag[i] == 0 && (best < 0 || raw[best].weight < raw[i].weight)) best = i;
        if (best < 0) break;
        flag[best] = 1;
        const float cx = raw[best].position[0], cy = raw[best].position[1], cz = raw[best].position[2];
        for (int i = m0; i < m1; ++i)
          if (flag[i] == 0 &&
              norm3_rn(cx, cy, cz, raw[i].position[0], raw[i].position[1], raw[i].position[2]) < radius)
            flag[i] = 2;
      }
      for (int i = m0; i < m1; ++i)
        if (flag[i] == 2) raw[i].n_votes = 0;
    }
    __syncwarp();
  }
  if (P.cross_class_filter == PCDB_MAXFILTER_MERGE) {
    // MaximaHandler::mergeAndFilterMaxima(maxima, true) + mergeMaxima (maxima_handler.cpp:296-383,386-443), order
    // dependent: one lane.  flag: 0 pending, 1 subsumed (dirty), 2 below MinVotesThreshold.  `sorted` serves as the
    // output list until the ranking below re-reads `raw`; src_dst[s] = where source maximum s's member votes go in the
    // re-packed member arrays (-1: dropped), the cloud's members keep their range [mem_off[m0], mem_off[m1]).
    if (lane == 0) {
      int n_out = 0, new_off = mem_off[m0];
      for (int i = m0; i < m1; ++i) {
        flag[i] = (raw[i].n_votes >= P.min_votes && raw[i].n_votes > 0) ? 0 : 2;
        src_dst[i] = -1;
      }
      for (int i = m0; i < m1; ++i) {
        if (flag[i] != 0) continue;
        const float sd = P.cls_h[raw[i].class_id];
        int n_close = 0;
        for (int j = i + 1; j < m1; ++j) {
          if (flag[j] != 0) continue;
          const float dist = norm3_rn(raw[j].position[0], raw[j].position[1], raw[j].position[2], raw[i].position[0],
                                      raw[i].position[1], raw[i].position[2]);
          if (dist < sd && P.cls_h[raw[j].class_id] <= sd) {
            close_list[m0 + n_close++] = j;
            flag[j] = 1;
          }
        }
        if (n_close == 0) {
          pcdb_maximum r = raw[i];
          src_dst[i] = new_off;
          r.vote_begin = new_off;
          new_off += r.n_votes;
          sorted[m0 + n_out++] = r;
          continue;
        }
        close_list[m0 + n_close++] = i;  // the maximum itself comes last
        // classes of the group in ascending order (std::map), each merged in list order; the heaviest merged one stays
        pcdb_maximum best;
        best.weight = 0.f;
        best.n_votes = 0;
        unsigned best_cls = 0xffffffffu;
        long long prev_cls = -1;
        for (;;) {
          long long cls = -1;
          for (int c = 0; c < n_close; ++c) {
            const long long k = (long long)raw[close_list[m0 + c]].class_id;
            if (k > prev_cls && (cls < 0 || k < cls)) cls = k;
          }
          if (cls < 0) break;
          prev_cls = cls;
          pcdb_maximum r;
          r.position[0] = r.position[1] = r.position[2] = 0.f;
          r.bbox_size[0] = r.bbox_size[1] = r.bbox_size[2] = 0.f;
          r.bbox_quat[0] = 1.f;
          r.bbox_quat[1] = r.bbox_quat[2] = r.bbox_quat[3] = 0.f;
          r.weight = 0.f;
          r.n_votes = 0;
          r.class_id = (unsigned)cls;
          r.instance_id = 0;
          r.instance_weight = 0.f;
          for (int c = 0; c < n_close; ++c) {
            const pcdb_maximum& mx = raw[close_list[m0 + c]];
            if ((long long)mx.class_id != cls) continue;
            const float rw = r.weight, mw = mx.weight, den = __fadd_rn(rw, mw);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              r.position[a] = __fdiv_rn(__fadd_rn(__fmul_rn(r.position[a], rw), __fmul_rn(mx.position[a], mw)), den);
              r.bbox_size[a] = __fdiv_rn(__fadd_rn(__fmul_rn(r.bbox_size[a], rw), __fmul_rn(mx.bbox_size[a], mw)), den);
            }
            float S[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b2 = 0; b2 < 4; ++b2)
                S[a][b2] = __fadd_rn(__fadd_rn(0.f, __fmul_rn(__fmul_rn(rw, r.bbox_quat[a]), r.bbox_quat[b2])),
                                     __fmul_rn(__fmul_rn(mw, mx.bbox_quat[a]), mx.bbox_quat[b2]));
            quat_avg_eig(S, r.bbox_quat);
            r.weight = __fadd_rn(r.weight, mx.weight);
            r.n_votes += mx.n_votes;
            // instance weights summed per id over the merged maxima so far; largest wins, ascending id on ties (std::map)
            float bw = 0.f;
            unsigned bid = 0;
            float w_id0 = 0.f;
            bool any = false;
            for (int d = 0; d <= c; ++d) {
              const pcdb_maximum& md = raw[close_list[m0 + d]];
              if ((long long)md.class_id != cls) continue;
              bool first_of_id = true;
              for (int e = 0; e < d; ++e) {
                const pcdb_maximum& me = raw[close_list[m0 + e]];
                if ((long long)me.class_id == cls && me.instance_id == md.instance_id) first_of_id = false;
              }
              if (!first_of_id) continue;
              float tot = 0.f;
              for (int e = d; e <= c; ++e) {
                const pcdb_maximum& me = raw[close_list[m0 + e]];
                if ((long long)me.class_id == cls && me.instance_id == md.instance_id)
                  tot = __fadd_rn(tot, me.instance_weight);
              }
              if (md.instance_id == 0) w_id0 = tot;
              if (tot > bw || (tot == bw && tot > 0.f && any && md.instance_id < bid)) {
                bw = tot;
                bid = md.instance_id;
                any = true;
              }
            }
            if (!any) {  // no instance weight > 0: the reference leaves the id uninitialised; defined as 0
              bid = 0;
              bw = w_id0;
            }
            r.instance_id = bid;
            r.instance_weight = bw;
          }
          r.raw_weight = r.weight;
          if (r.weight > best.weight) {
            best = r;
            best_cls = (unsigned)cls;
          }
        }
        if (best_cls != 0xffffffffu) {
          best.vote_begin = new_off;
          for (int c = 0; c < n_close; ++c) {
            const int s = close_list[m0 + c];
            if (raw[s].class_id != best_cls) continue;
            src_dst[s] = new_off;
            new_off += raw[s].n_votes;
          }
          sorted[m0 + n_out++] = best;
        }
      }
      for (int i = 0; i < n_out; ++i) raw[m0 + i] = sorted[m0 + i];
      for (int i = m0 + n_out; i < m1; ++i) raw[i].n_votes = 0;
    }
    __syncwarp();
  }
  // stable rank by weight (descending) among maxima that pass MinVotesThreshold
  int n_ok = 0;
  for (int i = m0 + lane; i < m1; i += 32) {
    const pcdb_maximum& a = raw[i];
    bool ok = a.n_votes >= P.min_votes && a.n_votes > 0;
    if (!ok) continue;
    int rank = 0;
    for (int j = m0; j < m1; ++j) {
      const pcdb_maximum& c = raw[j];
      if (!(c.n_votes >= P.min_votes && c.n_votes > 0)) continue;
      if (c.weight > a.weight || (c.weight == a.weight && j < i)) ++rank;
    }
    sorted[m0 + rank] = a;
    ++n_ok;
  }
  n_ok = warp_sum(n_ok);
  __syncwarp();
  if (lane == 0) {
    float sum = 0.f, sum_inst = 0.f;
    for (int i = 0; i < n_ok; ++i) {
      sum = __fadd_rn(sum, sorted[m0 + i].weight);
      sum_inst = __fadd_rn(sum_inst, sorted[m0 + i].instance_weight);
    }
    for (int i = 0; i < n_ok; ++i) {
      sorted[m0 + i].weight = sum != 0.f ? __fdiv_rn(sorted[m0 + i].weight, sum) : 0.f;
      sorted[m0 + i].instance_weight = sum_inst != 0.f ? __fdiv_rn(sorted[m0 + i].instance_weight, sum_inst) : 0.f;
    }
    float thr = P.min_threshold;
    if (thr < 0.f) thr = __fmul_rn(-thr, n_ok > 0 ? sorted[m0].weight : 0.f);
    int keep = 0;
    for (int i = 0; i < n_ok; ++i)
      if (sorted[m0 + i].weight >= thr) {
        if (keep != i) sorted[m0 + keep] = sorted[m0 + i];
        ++keep;
      }
    if (P.best_k > 0 && keep >= P.best_k) keep = P.best_k;
    kept[b] = keep;
    first[b] = m0;
    label[b] = keep > 0 ? (int)sorted[m0].class_id : -1;  // eval_classification.cpp:412-417
  }
}

// "Merge" filter: member votes of the surviving maxima, re-packed in merge order; one warp per source maximum
__global__ void k_repack_members(const int* __restrict__ M_ptr, const int* __restrict__ mem_off,
                                 const int* __restrict__ src_dst, const long long* __restrict__ mem_idx,
                                 const float* __restrict__ mem_w, long long* mem_idx2, float* mem_w2) {
  const int m = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (m >= *M_ptr || src_dst[m] < 0) return;
  const int o0 = mem_off[m], n = mem_off[m + 1] - o0, d0 = src_dst[m];
  for (int i = lane; i < n; i += 32) {
    mem_idx2[d0 + i] = mem_idx[o0 + i];
    mem_w2[d0 + i] = mem_w[o0 + i];
  }
}

__global__ void k_fill_labels(int B, int* kept, int* first, int* label) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  kept[b] = 0;
  first[b] = 0;
  label[b] = -1;
}

}  // namespace

// votes (device, V records), vote_pw, vote_cloud already in the workspace.  Leaves per-cloud sorted maxima in
// ws.max_sorted (+ ws.max_kept / ws.max_first), labels in ws.labels and member lists in ws.mem_idx / ws.mem_w.
int stage_find_maxima(pcdb_ctx* ctx, int B, int64_t V, bool have_cloud, int64_t* M_out, int64_t* members_out) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const pcdb_params& p = ctx->prm;
  const int C = std::max(1, ctx->cb.n_classes);
  *M_out = 0;
  *members_out = 0;
  PCDB_CUDA(w.labels.ensure(sizeof(int) * (B + 1)));
  PCDB_CUDA(w.max_kept.ensure(sizeof(int) * (B + 1)));
  PCDB_CUDA(w.max_first.ensure(sizeof(int) * (B + 1)));
  if (V == 0) {
    k_fill_labels<<<cdiv(B, 128), 128, 0, st>>>(B, w.max_kept.as<int>(), w.max_first.as<int>(), w.labels.as<int>());
    PCDB_LAUNCH_CHECK();
    return PCDB_OK;
  }
  if ((int64_t)B * C > 0xffffffffll) return ctx->fail(PCDB_E_INVALID, "B * n_classes overflows the 32-bit group key");
  MsP P;
  P.h_cfg = p.bandwidth;
  P.thr = p.ms_threshold;
  P.max_iter = p.ms_max_iter;
  P.kernel = p.ms_kernel;
  P.suppression = p.maxima_suppression;
  P.min_votes = p.min_votes_threshold;
  P.best_k = p.best_k;
  P.average_rotation = p.average_rotation;
  P.n_classes = C;
  P.min_threshold = p.min_threshold;
  P.cross_class_filter = p.single_object_mode ? PCDB_MAXFILTER_NONE : p.max_filter_type;
  P.radius_type = p.radius_type;
  P.single_max = p.single_object_mode ? p.single_object_max_type : PCDB_SOMAX_DEFAULT;
  if (!(p.bandwidth > 0.f)) return ctx->fail(PCDB_E_INVALID, "Voting.Bandwidth must be positive");
  if (P.single_max != PCDB_SOMAX_DEFAULT && !have_cloud)
    return ctx->fail(PCDB_E_UNSUPPORTED,
                     "Voting.SingleObjectMaxType other than Default needs the cloud (centroid, model radius): use "
                     "pcdb_classify_batch");
  // MaximaHandler::getSearchDistForClass per class (maxima_handler.cpp:509-521)
  std::vector<float> cls_h((size_t)C, p.bandwidth);
  if (p.radius_type != PCDB_RADIUS_CONFIG) {
    const std::vector<float>& dims = p.radius_type == PCDB_RADIUS_FIRST_DIM ? ctx->class_dim_first : ctx->class_dim_second;
    if ((int)dims.size() < C)
      return ctx->fail(PCDB_E_STATE, "Voting.BinOrBandwidthType needs the learned class dimensions: call "
                                     "pcdb_set_class_dimensions for all %d classes", C);
    for (int c = 0; c < C; ++c) cls_h[c] = dims[c] * p.radius_factor;
  }
  PCDB_CUDA(w.ms_cls_h.ensure(sizeof(float) * (size_t)C));
  PCDB_CUDA(cudaMemcpyAsync(w.ms_cls_h.p, cls_h.data(), sizeof(float) * (size_t)C, cudaMemcpyHostToDevice, st));
  PCDB_CUDA(cudaStreamSynchronize(st));  // cls_h is a local
  P.cls_h = w.ms_cls_h.as<float>();

  const size_t n = (size_t)V + 2;
  PCDB_CUDA(w.vote_key.ensure(sizeof(unsigned) * n));
  PCDB_CUDA(w.vote_key2.ensure(sizeof(unsigned) * n));
  PCDB_CUDA(w.vote_ord.ensure(sizeof(int) * n));
  PCDB_CUDA(w.vote_ord2.ensure(sizeof(int) * n));
  PCDB_CUDA(w.vote_pwS.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.vote_w_work.ensure(sizeof(float) * n));
  PCDB_CUDA(w.vote_w0.ensure(sizeof(float) * n));
  PCDB_CUDA(w.seg2_head.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seg2_id.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seg2_start.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seg2_key.ensure(sizeof(unsigned) * n));
  PCDB_CUDA(w.seed_k0.ensure(sizeof(unsigned long long) * n));
  PCDB_CUDA(w.seed_k1.ensure(sizeof(unsigned long long) * n));
  PCDB_CUDA(w.seed_kS.ensure(sizeof(unsigned long long) * n));
  PCDB_CUDA(w.seed_i0.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_i1.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_i2.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_i3.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_s0.ensure(sizeof(unsigned) * n));
  PCDB_CUDA(w.seed_s1.ensure(sizeof(unsigned) * n));
  PCDB_CUDA(w.seed_head.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_id.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_key.ensure(sizeof(unsigned long long) * n));
  PCDB_CUDA(w.seed_seg.ensure(sizeof(int) * n));
  PCDB_CUDA(w.seed_first.ensure(sizeof(int) * n));
  PCDB_CUDA(w.centers.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.ms_cen.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.ms_cen2.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.ms_dens.ensure(sizeof(float) * n));
  PCDB_CUDA(w.ms_flag.ensure(sizeof(int) * n));
  PCDB_CUDA(w.max_pos.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.max_n.ensure(sizeof(int) * n));
  PCDB_CUDA(w.max_off.ensure(sizeof(int) * n));
  PCDB_CUDA(w.mpos.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.mseg.ensure(sizeof(int) * n));
  PCDB_CUDA(w.mem_cnt.ensure(sizeof(int) * n));
  PCDB_CUDA(w.mem_off.ensure(sizeof(int) * n));
  PCDB_CUDA(w.max_raw.ensure(sizeof(pcdb_maximum) * n));
  PCDB_CUDA(w.max_sorted.ensure(sizeof(pcdb_maximum) * n));
  PCDB_CUDA(w.max_flag.ensure(sizeof(int) * n));
  PCDB_CUDA(w.ms_grp.ensure(sizeof(float4) * n));
  PCDB_CUDA(w.ms_close.ensure(sizeof(int) * n));
  PCDB_CUDA(w.ms_src_dst.ensure(sizeof(int) * n));
  P.grp = w.ms_grp.as<float4>();

  const unsigned gV = cdiv(V, 256), gV1 = cdiv(V + 1, 256);
  const pcdb_vote* votes = w.votes.as<pcdb_vote>();
  // (cloud, class) grouping
  k_vote_keys<<<gV, 256, 0, st>>>(votes, w.vote_cloud.as<int>(), V, C, w.vote_key.as<unsigned>(), w.vote_ord.as<int>());
  PCDB_LAUNCH_CHECK();
  int bits = 1;
  while ((1ll << bits) < (int64_t)B * C && bits < 32) ++bits;
  PCDB_TRY(pcdb_cub_sort_pairs_u32(ctx, w.vote_key.as<unsigned>(), w.vote_key2.as<unsigned>(), w.vote_ord.as<int>(),
                                   w.vote_ord2.as<int>(), V, bits));
  k_heads_u32<<<gV1, 256, 0, st>>>(w.vote_key2.as<unsigned>(), V, w.seg2_head.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.seg2_head.as<int>(), w.seg2_id.as<int>(), V + 1));
  k_seg_info<<<gV1, 256, 0, st>>>(w.seg2_head.as<int>(), w.seg2_id.as<int>(), w.vote_key2.as<unsigned>(), V,
                                  w.seg2_start.as<int>(), w.seg2_key.as<unsigned>());
  PCDB_LAUNCH_CHECK();
  const int* nseg_ptr = w.seg2_id.as<int>() + V;
  const unsigned gseg = (unsigned)std::min<int64_t>(V, (int64_t)B * C);
  PCDB_CUDA(cudaMemsetAsync(w.max_n.p, 0, sizeof(int) * n, st));
  if (P.single_max != PCDB_SOMAX_DEFAULT) {  // centroid + model radius of the cloud handed to Voting::findMaxima
    PCDB_CUDA(w.ms_cloud_cm.ensure(sizeof(float4) * (size_t)(B + 1)));
    k_cloud_centroid<<<cdiv((int64_t)B * 32, 128), 128, 0, st>>>(w.surf4.as<float4>(), w.surf_off.as<long long>(), B,
                                                                 w.ms_cloud_cm.as<float4>());
    PCDB_LAUNCH_CHECK();
  }
  // search distance of every (cloud, class) group; in the single-object max types also the group's one maximum
  k_group_params<<<cdiv(((int64_t)gseg + 1) * 32, 128), 128, 0, st>>>(
      nseg_ptr, w.seg2_key.as<unsigned>(), w.seg2_start.as<int>(), w.vote_pw.as<float4>(), w.vote_ord2.as<int>(), P,
      w.ms_cloud_cm.as<float4>(), w.ms_grp.as<float4>(), w.seed_first.as<int>(), w.max_pos.as<float4>(),
      w.max_n.as<int>());
  PCDB_LAUNCH_CHECK();
  k_gather_pw<<<gV, 256, 0, st>>>(w.vote_ord2.as<int>(), V, w.vote_pw.as<float4>(), P, w.seg2_id.as<int>(),
                                  w.seg2_head.as<int>(), w.vote_pwS.as<float4>(), w.vote_w_work.as<float>(),
                                  w.seed_k0.as<unsigned long long>(), w.seed_i0.as<int>());
  PCDB_LAUNCH_CHECK();
  PCDB_CUDA(cudaMemcpyAsync(w.vote_w0.p, w.vote_w_work.p, sizeof(float) * V, cudaMemcpyDeviceToDevice, st));
  if (P.single_max == PCDB_SOMAX_DEFAULT) {
    // seeds: sort bin keys, then stable by group
    PCDB_TRY(pcdb_cub_sort_pairs_u64(ctx, w.seed_k0.as<unsigned long long>(), w.seed_k1.as<unsigned long long>(),
                                     w.seed_i0.as<int>(), w.seed_i1.as<int>(), V, 63));
    k_gather_segkey<<<gV, 256, 0, st>>>(w.seed_i1.as<int>(), w.seg2_id.as<int>(), w.seg2_head.as<int>(), V,
                                        w.seed_s0.as<unsigned>(), w.seed_i2.as<int>());
    PCDB_LAUNCH_CHECK();
    int sbits = 1;
    while ((1ll << sbits) < V + 1 && sbits < 32) ++sbits;
    PCDB_TRY(pcdb_cub_sort_pairs_u32(ctx, w.seed_s0.as<unsigned>(), w.seed_s1.as<unsigned>(), w.seed_i2.as<int>(),
                                     w.seed_i3.as<int>(), V, sbits));
    k_seed_heads<<<gV1, 256, 0, st>>>(w.seed_s1.as<unsigned>(), w.seed_i3.as<int>(),
                                      w.seed_k1.as<unsigned long long>(), V, w.seed_kS.as<unsigned long long>(),
                                      w.seed_head.as<int>());
    PCDB_LAUNCH_CHECK();
    PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.seed_head.as<int>(), w.seed_id.as<int>(), V + 1));
    k_seed_list<<<gV1, 256, 0, st>>>(w.seed_head.as<int>(), w.seed_id.as<int>(), w.seed_kS.as<unsigned long long>(),
                                     w.seed_s1.as<unsigned>(), V, nseg_ptr, w.seed_key.as<unsigned long long>(),
                                     w.seed_seg.as<int>(), w.seed_first.as<int>());
    PCDB_LAUNCH_CHECK();
    const int* nseeds_ptr = w.seed_id.as<int>() + V;
    // mean shift: one warp per seed (at most V seeds)
    k_meanshift<<<cdiv(V * 32, 128), 128, 0, st>>>(nseeds_ptr, w.seed_key.as<unsigned long long>(),
                                                   w.seed_seg.as<int>(), w.seg2_start.as<int>(),
                                                   w.vote_pwS.as<float4>(), P, w.centers.as<float4>());
    PCDB_LAUNCH_CHECK();
    k_ms_tail<<<gseg, 32, 0, st>>>(nseg_ptr, w.seg2_start.as<int>(), w.seed_first.as<int>(), w.centers.as<float4>(),
                                   w.vote_pwS.as<float4>(), w.vote_w0.as<float>(), P, w.ms_cen.as<float4>(),
                                   w.ms_cen2.as<float4>(), w.ms_dens.as<float>(), w.ms_flag.as<int>(),
                                   w.max_pos.as<float4>(), w.max_n.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  // maxima offsets per group (entries past nseg are zero), total M
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.max_n.as<int>(), w.max_off.as<int>(), (int64_t)gseg + 1));
  const int* M_ptr = w.max_off.as<int>() + gseg;
  k_max_list<<<cdiv(gseg, 128), 128, 0, st>>>(nseg_ptr, w.max_n.as<int>(), w.max_off.as<int>(), w.seed_first.as<int>(),
                                              w.max_pos.as<float4>(), w.mpos.as<float4>(), w.mseg.as<int>());
  PCDB_LAUNCH_CHECK();
  k_member_count<<<cdiv((V + 1) * 32, 128), 128, 0, st>>>(M_ptr, w.mpos.as<float4>(), w.mseg.as<int>(),
                                                          w.seg2_start.as<int>(), w.vote_pwS.as<float4>(), P,
                                                          w.mem_cnt.as<int>());
  PCDB_LAUNCH_CHECK();
  // mem_cnt has M+1 valid entries; scan over V+1 upper bound needs zeros beyond: cleared by the count kernel only up
  // to M, so read M first.
  int hM = 0;
  PCDB_TRY(pcdb_read_small(ctx, &hM, M_ptr, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, w.mem_cnt.as<int>(), w.mem_off.as<int>(), (int64_t)hM + 1));
  int hMem = 0;
  PCDB_TRY(pcdb_read_small(ctx, &hMem, w.mem_off.as<int>() + hM, sizeof(int)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  PCDB_CUDA(w.mem_idx.ensure(sizeof(long long) * ((size_t)hMem + 1)));
  PCDB_CUDA(w.mem_w.ensure(sizeof(float) * ((size_t)hMem + 1)));
  k_ms_reweight<<<gseg, 32, 0, st>>>(nseg_ptr, w.max_n.as<int>(), w.max_off.as<int>(), w.seg2_start.as<int>(),
                                     w.mpos.as<float4>(), w.vote_pwS.as<float4>(), w.vote_ord2.as<int>(),
                                     w.mem_off.as<int>(), P, w.vote_w_work.as<float>(), w.mem_idx.as<long long>(),
                                     w.mem_w.as<float>());
  PCDB_LAUNCH_CHECK();
  if (p.ransac_vote_filtering && hM > 0) {  // voting.cpp:110-127: only the inlier votes of every maximum go on
    int64_t mem_io = hMem;
    PCDB_TRY(stage_ransac_filter(ctx, hM, M_ptr, &mem_io));
    hMem = (int)mem_io;
  }
  if (hM > 0) {
    k_max_reduce<<<cdiv((int64_t)hM * 32, 128), 128, 0, st>>>(M_ptr, w.mpos.as<float4>(), w.mseg.as<int>(),
                                                              w.seg2_key.as<unsigned>(), w.mem_off.as<int>(),
                                                              w.mem_idx.as<long long>(), w.mem_w.as<float>(), votes, P,
                                                              w.max_raw.as<pcdb_maximum>());
    PCDB_LAUNCH_CHECK();
  }
  k_cloud_finalize<<<cdiv((int64_t)B * 32, 128), 128, 0, st>>>(
      B, nseg_ptr, w.seg2_key.as<unsigned>(), w.max_off.as<int>(), w.mem_off.as<int>(), w.max_raw.as<pcdb_maximum>(), P,
      w.max_flag.as<int>(), w.ms_close.as<int>(), w.ms_src_dst.as<int>(), w.max_sorted.as<pcdb_maximum>(),
      w.max_kept.as<int>(), w.max_first.as<int>(), w.labels.as<int>());
  PCDB_LAUNCH_CHECK();
  if (P.cross_class_filter == PCDB_MAXFILTER_MERGE && hM > 0) {  // member votes of merged maxima, in merge order
    PCDB_CUDA(w.mem_idx2.ensure(sizeof(long long) * ((size_t)hMem + 1)));
    PCDB_CUDA(w.mem_w2.ensure(sizeof(float) * ((size_t)hMem + 1)));
    k_repack_members<<<cdiv((int64_t)hM * 32, 128), 128, 0, st>>>(M_ptr, w.mem_off.as<int>(), w.ms_src_dst.as<int>(),
                                                                  w.mem_idx.as<long long>(), w.mem_w.as<float>(),
                                                                  w.mem_idx2.as<long long>(), w.mem_w2.as<float>());
    PCDB_LAUNCH_CHECK();
    std::swap(w.mem_idx, w.mem_idx2);
    std::swap(w.mem_w, w.mem_w2);
  }
  *M_out = hM;
  *members_out = hMem;
  return PCDB_OK;
}

int stage_votes_unpack(pcdb_ctx* ctx, int B, int64_t V) {
  Workspace& w = ctx->ws;
  PCDB_CUDA(w.vote_pw.ensure(sizeof(float4) * (size_t)(V + 1)));
  PCDB_CUDA(w.vote_cloud.ensure(sizeof(int) * (size_t)(V + 1)));
  if (V > 0) {
    k_votes_unpack<<<cdiv(V, 256), 256, 0, ctx->stream>>>(w.votes.as<pcdb_vote>(), V, w.vote_off.as<long long>(), B,
                                                         w.vote_pw.as<float4>(), w.vote_cloud.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  return PCDB_OK;
}
