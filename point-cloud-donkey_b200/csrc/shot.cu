// shot.cu — K3 SHOT local reference frame, K4 SHOT-352, K5 CSHOT-1344, fused with the radius search.
//
// Replaces (reference paths under src/implicit_shape_model/):
//   Features::computeSHOTReferenceFrames    features/features.cpp:238-252  (pcl::SHOTLocalReferenceFrameEstimationOMP;
//                                           in-repo near-copy third_party/pcl_shot_na_lrf/shot_na_lrf.hpp:48-178)
//   FeaturesSHOT::iComputeDescriptors       features/features_shot.cpp:28-81   (pcl::SHOTEstimationOMP)
//   FeaturesCSHOT::iComputeDescriptors      features/features_cshot.cpp:28-103 (pcl::SHOTColorEstimationOMP)
//   the per-keypoint kd-tree radius searches behind both
//
// Work item = one occupied search-grid cell of keypoints.  The 27 surrounding cells (9 contiguous runs of the
// cell-sorted surface) are staged once into shared memory as float4 (xyz|rgb), float4 (normal|index) [+ float4 Lab]
// (dense scenes whose 27 cells hold more than kChunk points skip the staging: each warp filters the runs once from
// L2 into a global list of in-radius points and the passes read those points by index)
// and every keypoint of the cell is processed by one warp from that staged copy: pass A weighted covariance (fp64,
// warp-shuffle reduction), 3x3 symmetric eigen-solve, pass B sign disambiguation, pass C quadrilinear soft histogram
// accumulated in 32-bit fixed point in shared memory (native integer atomics => bit-reproducible across runs).
// Geometry follows the reference's float/double choices; membership d^2 < r^2 is bit-exact.
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "stages.h"

namespace {

// Warps per CTA: all of them share ONE staged neighbourhood, so more warps per CTA = more resident warps per byte of
// shared memory.  SHOT: 16 warps, 2 CTAs per SM = 32 resident warps at 64 registers (round 1 ran 8 x 2 = 16 and was
// latency-bound: warps_active 25 %, stall_wait 2.4 per issue; 12 x 2 measured 36 %).  The stage is 28 B per point (the
// normal as three float arrays, the point's index in the unused w of its position) so that two such CTAs fit the SM.
// CSHOT (44 B per staged point, 5.4 KB histograms): 16 warps, 1 CTA.
constexpr int kWarpsShot = 16, kWarpsCshot = 16, kWarpsMax = 16;
constexpr int kChunk = PCDB_SHOT_CHUNK;     // staged points per work item
// Per-warp list of in-radius staged points, filled kList entries at a time: a neighbourhood larger than one fill is
// processed in several fills (the passes only accumulate), so the list costs 2 KB per warp instead of 2 bytes per
// staged point (4 KB) — that difference is what lets the extra warps fit.
constexpr int kList = 1024;

constexpr double PST_RAD_45 = 0.78539816339744830961566084581988;
constexpr double PST_RAD_90 = 1.5707963267948966192313216916398;
constexpr double PST_RAD_135 = 2.3561944901923449288469825374596;
constexpr double PST_RAD_PI_7_8 = 2.7488935718910690836548129603691;

struct ShotArgs {
  const float4* surfS;
  const float4* snrmS;
  const float4* slabS;
  const unsigned long long* skeys;
  const long long* surf_off;
  const float4* kp4;
  const int* kp_order;
  const unsigned long long* kp_keys;
  const int* item_start;
  const int* n_items_ptr;
  double r_lrf, r_shot;
  float r2_lrf, r2_shot;
  float t2_gt_r12, t2_gt_r34, t2_ge_r14;  // float thresholds on d^2 equivalent to the fp64 shell tests (see stage_shot)
  const float* lrf_in;
  float* lrf_out;
  float* desc_out;
  int do_lrf, do_desc;
  int* work_counter;
  unsigned long long* nbr_counts;
  const float* lab_lut;
  const long long* item_beg;  // [items][9] first sorted-surface index of each of the nine cell runs of an item
  const int* item_len;        // [items][9] run lengths (precomputed by k_item_ranges: no serial searches in k_shot)
  int dense;         // 0: staged launch (CTA per item, skips dense items); 1: dense launch (warp per keypoint)
  int stage_cap;     // points the launch stages per item (kChunk), 0 for the dense launch (no staging arrays)
  long long n_kp;    // keypoints of the batch (sorted positions the dense launch walks)
  const int* item_id;    // exclusive scan of the item heads over the sorted keypoints
  const int* item_head;
  int* work_counter_dense;
  int* item_next;    // staged launch: keypoints drawn so far from every item (global, so that helper CTAs share an item)
  int blocked;       // pass C: lane-blocked (1) or lane-strided (0) walk of the in-radius list
  float4* glist;     // dense neighbourhoods (more than kChunk points in the 27 cells): per-warp lists of the in-radius
                     // points themselves, (x, y, z, sorted-surface index bits): the passes read them back coalesced
  long long gcap;    // entries per warp (>= the largest 27-cell population of the batch)
};

// cyclic Jacobi, ascending eigenvalues; V columns are eigenvectors (the reference: Eigen::SelfAdjointEigenSolver)
__device__ void eig3_sym(const double Ain[6], double w[3], double V[3][3]) {
  double A[3][3] = {{Ain[0], Ain[1], Ain[2]}, {Ain[1], Ain[3], Ain[4]}, {Ain[2], Ain[4], Ain[5]}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (!isfinite(theta)) t = 0.0;
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        A[p][q] = A[q][p] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  double d[3] = {A[0][0], A[1][1], A[2][2]};
  int o0 = 0, o1 = 1, o2 = 2;
  // stable ascending sort of three
  if (d[o1] < d[o0]) { int t = o0; o0 = o1; o1 = t; }
  if (d[o2] < d[o1]) { int t = o1; o1 = o2; o2 = t; }
  if (d[o1] < d[o0]) { int t = o0; o0 = o1; o1 = t; }
  double Vs[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Vs[i][0] = V[i][o0];
    Vs[i][1] = V[i][o1];
    Vs[i][2] = V[i][o2];
  }
  w[0] = d[o0];
  w[1] = d[o1];
  w[2] = d[o2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = Vs[i][j];
}

// PCL RGB2CIELAB with the host-built LUTs, normalised L/100, a/120, b/120
__device__ __forceinline__ float3 rgb_to_lab_norm(unsigned rgb, const float* __restrict__ lut) {
  float fr = lut[(rgb >> 16) & 0xFF], fg = lut[(rgb >> 8) & 0xFF], fb = lut[rgb & 0xFF];
  const float* xyzl = lut + 256;
  float x = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.412453f), __fmul_rn(fg, 0.357580f)), __fmul_rn(fb, 0.180423f));
  float y = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.212671f), __fmul_rn(fg, 0.715160f)), __fmul_rn(fb, 0.072169f));
  float z = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.019334f), __fmul_rn(fg, 0.119193f)), __fmul_rn(fb, 0.950227f));
  float vx = __fdiv_rn(x, 0.95047f), vy = y, vz = __fdiv_rn(z, 1.08883f);
  vx = xyzl[min(3999, max(0, (int)__fmul_rn(vx, 4000.f)))];
  vy = xyzl[min(3999, max(0, (int)__fmul_rn(vy, 4000.f)))];
  vz = xyzl[min(3999, max(0, (int)__fmul_rn(vz, 4000.f)))];
  float L = __fsub_rn(__fmul_rn(116.0f, vy), 16.0f);
  if (L > 100.f) L = 100.0f;
  float A = __fmul_rn(500.0f, __fsub_rn(vx, vy));
  if (A > 120.f) A = 120.0f; else if (A < -120.f) A = -120.0f;
  float B = __fmul_rn(200.0f, __fsub_rn(vy, vz));
  if (B > 120.f) B = 120.0f; else if (B < -120.f) B = -120.0f;
  return make_float3(__fdiv_rn(L, 100.0f), __fdiv_rn(A, 120.0f), __fdiv_rn(B, 120.0f));
}

__global__ void k_lab(const float4* __restrict__ pts, long long n, const float* __restrict__ lut, float4* lab) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float3 l = rgb_to_lab_norm(__float_as_uint(pts[i].w), lut);
  lab[i] = make_float4(l.x, l.y, l.z, 0.f);
}

// Histogram bins are 32-bit unsigned fixed point in shared memory: native ATOMS.ADD (a 64-bit shared atomic is a CAS
// spin loop on sm_100 and serialises badly when many neighbours hit one bin), and integer adds commute, so the
// descriptor is bit-reproducible from run to run.  Every contribution is >= 0 and <= 4, a bin receives at most one
// such contribution per staged point, hence scale = 2^floor(log2(2^32 / (4 T + 4))) cannot overflow.
__device__ __forceinline__ void hist_add(unsigned* hist, int bin, float v, float scale) {
  // the reference narrows every contribution to float before adding it (shot[..] += static_cast<float>(..))
  atomicAdd(hist + bin, __float2uint_rn(__fmul_rn(v, scale)));
}

// ---- fast fp32 helpers for the CONTINUOUS interpolation weights (never for a discrete decision) -------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// acos on [-1, 1]: sqrt(1 - |x|) * P7(|x|) (Abramowitz & Stegun 4.4.46), |error| <= 5e-7 rad in fp32
__device__ __forceinline__ float acos_fast(float x) {
  const float ax = fabsf(x);
  float p = -0.0012624911f;
  p = fmaf(p, ax, 0.0066700901f);
  p = fmaf(p, ax, -0.0170881256f);
  p = fmaf(p, ax, 0.0308918810f);
  p = fmaf(p, ax, -0.0501743046f);
  p = fmaf(p, ax, 0.0889789874f);
  p = fmaf(p, ax, -0.2145988016f);
  p = fmaf(p, ax, 1.5707963050f);
  const float r = sqrt_approx(1.0f - ax) * p;
  return x < 0.f ? 3.14159265358979f - r : r;
}
// atan2 for (y, x) != (0, 0): odd minimax polynomial of degree 15 on the reduced argument, |error| <= 3e-7 rad in fp32
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mn * rcp_approx(mx);
  const float s = t * t;
  float p = -0.004054560326039791f;
  p = fmaf(p, s, 0.021862933412194252f);
  p = fmaf(p, s, -0.05591229349374771f);
  p = fmaf(p, s, 0.09642194956541061f);
  p = fmaf(p, s, -0.1390862911939621f);
  p = fmaf(p, s, 0.19946566224098206f);
  p = fmaf(p, s, -0.33329859375953674f);
  p = fmaf(p, s, 0.9999993443489075f);
  float r = p * t;
  if (ay > ax) r = 1.57079632679490f - r;
  if (x < 0.f) r = 3.14159265358979f - r;
  return copysignf(r, y);
}

// Work distribution of the staged launch.  Pass 0 hands every item (a whole small cloud, or one search-grid cell of a
// large one) to one CTA.  A CTA that finds no pass-0 item left becomes a HELPER: it walks the last items in reverse
// order (those are the ones still being worked on), stages the same neighbourhood and draws keypoints from the item's
// global counter.  The launch therefore ends about one keypoint after the last keypoint was drawn, instead of one
// item after the last item was drawn (round 2 measured sm__issue_active min/max over the SMs = 51 / 68 %: the item tail).
template <bool COLOR, bool DENSE>
__global__ void __launch_bounds__((COLOR ? kWarpsCshot : kWarpsShot) * 32, COLOR ? 1 : 2) k_shot(ShotArgs a) {
  constexpr int D = COLOR ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  constexpr int kWarps = COLOR ? kWarpsCshot : kWarpsShot;
  constexpr int kThreads = kWarps * 32;
  constexpr int kStage = DENSE ? 0 : kChunk;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // staged neighbourhood: float4 (x, y, z, surface index bits), the normal as three float arrays, float4 Lab (CSHOT)
  float4* s_pts = reinterpret_cast<float4*>(smem_raw);
  float* s_nx = reinterpret_cast<float*>(s_pts + kStage);
  float* s_ny = s_nx + kStage;
  float* s_nz = s_ny + kStage;
  float4* s_lab = reinterpret_cast<float4*>(s_nz + kStage);  // only touched when COLOR
  unsigned* s_hist = reinterpret_cast<unsigned*>(smem_raw + (size_t)kStage * (COLOR ? 44 : 28));
  // per-warp compacted list of the staged points inside the current search radius (positions in the stage)
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_hist + (size_t)kWarps * D);
  __shared__ long long s_rbeg_all[DENSE ? kWarps : 1][9];  // dense launch: one row per warp
  __shared__ int s_rlen_all[DENSE ? kWarps : 1][9];
  __shared__ int s_pref[10];
  __shared__ int s_item;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = *a.n_items_ptr;
  unsigned* hist = s_hist + (size_t)warp * D;
  unsigned short* list = s_list + (size_t)warp * (DENSE ? 0 : kList);
  long long* s_rbeg = s_rbeg_all[DENSE ? warp : 0];
  int* s_rlen = s_rlen_all[DENSE ? warp : 0];
  float4* glist = DENSE ? a.glist + ((size_t)blockIdx.x * kWarps + warp) * a.gcap : nullptr;
  // helper passes over the last n_tail items (see above); few items (a single cloud) => many CTAs share each of them
  const int n_tail = min(n_items, (int)gridDim.x);
  const int n_pass = DENSE ? 0 : (n_items >= (int)gridDim.x ? 4 : min(64, max(4, 2 * (int)gridDim.x / max(1, n_items))));
  const long long n_units = (long long)n_items + (long long)(n_pass - 1) * n_tail;

  while (true) {
    int k0, k1, T, item = 0;
    if (!DENSE) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const long long u = (long long)atomicAdd(a.work_counter, 1);
        int it = -1;
        if (u < n_units) {
          if (u < n_items) {
            it = (int)u;
          } else {  // helper unit: skip an item whose keypoints have all been drawn
            it = n_items - 1 - (int)((u - n_items) % n_tail);
            const int drawn = *reinterpret_cast<volatile int*>(a.item_next + it);
            if (a.item_start[it] + drawn >= a.item_start[it + 1]) it = -2;
          }
        }
        s_item = it;
      }
      __syncthreads();
      item = s_item;
      if (item == -1) break;
      if (item == -2) continue;
      k0 = a.item_start[item];
      k1 = a.item_start[item + 1];
      if (threadIdx.x < 9) {
        s_rbeg[threadIdx.x] = a.item_beg[(size_t)item * 9 + threadIdx.x];
        s_rlen[threadIdx.x] = a.item_len[(size_t)item * 9 + threadIdx.x];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int acc = 0;
        for (int r = 0; r < 9; ++r) {
          s_pref[r] = acc;
          acc += s_rlen[r];
        }
        s_pref[9] = acc;
      }
      __syncthreads();
      T = s_pref[9];
      if (T > kChunk) continue;  // dense neighbourhood: left to the dense launch (block-uniform)
      for (int e = threadIdx.x; e < T; e += kThreads) {
        int r = 0;
        while (e >= s_pref[r + 1]) ++r;
        const long long src = s_rbeg[r] + (e - s_pref[r]);
        const float4 p = a.surfS[src], nq = a.snrmS[src];
        s_pts[e] = make_float4(p.x, p.y, p.z, nq.w);
        s_nx[e] = nq.x;
        s_ny[e] = nq.y;
        s_nz[e] = nq.z;
        if (COLOR) s_lab[e] = a.slabS[src];
      }
      __syncthreads();
    } else {
      // Dense launch: nothing is shared between the warps of a CTA, so every warp walks the sorted keypoints on its own
      // (a cell of a dense cloud holds 2-3 keypoints: tying the warps of a CTA to it left most of them waiting at a barrier)
      int p = 0;
      if (lane == 0) p = atomicAdd(a.work_counter_dense, 1);
      p = __shfl_sync(0xffffffffu, p, 0);
      if (p >= a.n_kp) break;
      const int it = a.item_id[p] + a.item_head[p] - 1;
      __syncwarp();
      if (lane < 9) {
        s_rbeg[lane] = a.item_beg[(size_t)it * 9 + lane];
        s_rlen[lane] = a.item_len[(size_t)it * 9 + lane];
      }
      __syncwarp();
      T = 0;
      for (int r = 0; r < 9; ++r) T += s_rlen[r];
      if (T <= kChunk) continue;  // staged launch's item (warp-uniform)
      k0 = p;
      k1 = p + 1;
    }
    const float fix_scale = exp2f(floorf(log2f(4294967296.0f / (4.0f * (float)T + 4.0f))));
    const float fix_inv = 1.0f / fix_scale;
    const float r2_max = fmaxf(a.do_lrf ? a.r2_lrf : 0.f, a.do_desc ? a.r2_shot : 0.f);

    // Staged launch: every warp draws the next keypoint of the item when it is free (the counter is global so that
    // helper CTAs draw from the same item).  Dense launch: the warp's single keypoint.
    bool first = true;
    while (true) {
      int kq = k0;
      if (!DENSE) {
        if (lane == 0) kq = k0 + atomicAdd(a.item_next + item, 1);
        kq = __shfl_sync(0xffffffffu, kq, 0);
      } else if (!first) {
        break;
      }
      first = false;
      if (kq >= k1) break;
      const int kidx = a.kp_order[kq];
      float kx, ky, kz;
      unsigned krgb;
      {
        const float4 k4 = a.kp4[kidx];
        kx = k4.x; ky = k4.y; kz = k4.z;
        krgb = __float_as_uint(k4.w);
      }
      // Radius filter first, heavy math second: the in-radius points of the stage are compacted into this warp's list
      // (ballot + popc, order = staging order), so the passes below run with all 32 lanes busy instead of diverging on
      // the 30-50 % of the staged neighbourhood that lies inside the sphere.  One fill holds at most kList entries:
      // `from` is where the scan of the staged points resumes, the return value the number of entries.
      auto fill = [&](int& from, int cnt, float r2) -> int {
        int n = 0;
        int e0 = from;
        __syncwarp();  // every lane is done reading the previous fill
        for (; e0 < cnt && n <= kList - 64; e0 += 64) {  // two independent loads / distance chains per lane
          const int ea = e0 + lane, eb = ea + 32;
          bool ina = false, inb = false;
          if (ea < cnt) {
            const float4 p = s_pts[ea];
            ina = sqdist3_rn(kx, ky, kz, p.x, p.y, p.z) < r2;
          }
          if (eb < cnt) {
            const float4 p = s_pts[eb];
            inb = sqdist3_rn(kx, ky, kz, p.x, p.y, p.z) < r2;
          }
          const unsigned ma = __ballot_sync(0xffffffffu, ina), mb = __ballot_sync(0xffffffffu, inb);
          const unsigned below = (1u << lane) - 1u;
          if (ina) list[n + __popc(ma & below)] = (unsigned short)ea;
          n += __popc(ma);
          if (inb) list[n + __popc(mb & below)] = (unsigned short)eb;
          n += __popc(mb);
        }
        from = e0;
        __syncwarp();
        return n;
      };
      // Dense neighbourhood: one sweep over the nine runs (coalesced float4 reads, served by L2, two independent loads
      // per lane and iteration) copies the points inside the larger of the two radii into this warp's global list; the
      // passes below then read only those (typically < 10 % of the 27-cell population), coalesced and without a
      // dependent index load, and apply their own radius inline.  No block-wide barrier in this mode.
      int n_g = 0;
      if (DENSE) {
        for (int r = 0; r < 9; ++r) {
          const long long rb = s_rbeg[r];
          const int rl = s_rlen[r];
          for (int e0 = 0; e0 < rl; e0 += 64) {
            const int ea = e0 + lane, eb = ea + 32;
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
            if (ea < rl) pa = a.surfS[rb + ea];
            if (eb < rl) pb = a.surfS[rb + eb];
            const bool ina = ea < rl && sqdist3_rn(kx, ky, kz, pa.x, pa.y, pa.z) < r2_max;
            const bool inb = eb < rl && sqdist3_rn(kx, ky, kz, pb.x, pb.y, pb.z) < r2_max;
            const unsigned ma = __ballot_sync(0xffffffffu, ina), mb = __ballot_sync(0xffffffffu, inb);
            const unsigned below = (1u << lane) - 1u;
            if (ina) glist[n_g + __popc(ma & below)] = make_float4(pa.x, pa.y, pa.z, __int_as_float((int)(rb + ea)));
            n_g += __popc(ma);
            if (inb) glist[n_g + __popc(mb & below)] = make_float4(pb.x, pb.y, pb.z, __int_as_float((int)(rb + eb)));
            n_g += __popc(mb);
          }
        }
        __syncwarp();
      }
      auto pt_at = [&](int i) -> float4 {  // .w: sorted-surface index (dense) / index inside the cloud (staged), as bits
        if constexpr (DENSE) return glist[i]; else return s_pts[list[i]];
      };
      auto idx_at = [&](int i) -> int {  // index of the point inside its cloud (tie-break key of the sorted kd-tree result)
        if constexpr (DENSE) return __float_as_int(a.snrmS[__float_as_int(glist[i].w)].w);
        else return __float_as_int(s_pts[list[i]].w);
      };
      float rf[9];
      bool lrf_ok = true;
      // ---------------------------------------------------------------- LRF
      if (a.do_lrf) {
        double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0, sw = 0;
        int valid = 0, nall = 0, n_lrf = 0;
        int from = 0, done = 0;
        do {
          n_lrf = DENSE ? n_g : fill(from, T, a.r2_lrf);
          // lane <-> entry assignment of ONE long list, whatever the number of fills (keeps the fp64 sums' order)
          const int start = DENSE ? lane : ((lane + 32 - (done & 31)) & 31);
          for (int i = start; i < n_lrf; i += 32) {
            const float4 p = pt_at(i);
            const float d2 = sqdist3_rn(kx, ky, kz, p.x, p.y, p.z);
            if (!DENSE || d2 < a.r2_lrf) {
              ++nall;
              if (!(p.x == kx && p.y == ky && p.z == kz)) {
                const double vx = (double)__fsub_rn(p.x, kx), vy = (double)__fsub_rn(p.y, ky),
                             vz = (double)__fsub_rn(p.z, kz);
                const double wgt = a.r_lrf - sqrt((double)d2);
                c00 += wgt * (vx * vx);
                c01 += wgt * (vx * vy);
                c02 += wgt * (vx * vz);
                c11 += wgt * (vy * vy);
                c12 += wgt * (vy * vz);
                c22 += wgt * (vz * vz);
                sw += wgt;
                ++valid;
              }
            }
          }
          done += n_lrf;
        } while (!DENSE && from < T);
        const bool whole = DENSE || done == n_lrf;  // the list still holds the complete LRF neighbourhood
        c00 = warp_sum(c00); c01 = warp_sum(c01); c02 = warp_sum(c02);
        c11 = warp_sum(c11); c12 = warp_sum(c12); c22 = warp_sum(c22);
        sw = warp_sum(sw);
        valid = warp_sum(valid);
        nall = warp_sum(nall);
        if (lane == 0 && a.nbr_counts) atomicAdd(&a.nbr_counts[0], (unsigned long long)nall);
        double x[3] = {0, 0, 0}, z[3] = {0, 0, 0};
        lrf_ok = valid >= 5;
        if (lrf_ok) {
          double cov[6] = {c00 / sw, c01 / sw, c02 / sw, c11 / sw, c12 / sw, c22 / sw};
          double ev[3], V[3][3];
          eig3_sym(cov, ev, V);
          if (!isfinite(ev[0]) || !isfinite(ev[1]) || !isfinite(ev[2])) lrf_ok = false;
          x[0] = V[0][2]; x[1] = V[1][2]; x[2] = V[2][2];
          z[0] = V[0][0]; z[1] = V[1][0]; z[2] = V[2][0];
        }
        // pass B: sign disambiguation
        int plusX = 0, plusZ = 0;
        if (lrf_ok) {
          int fromB = 0;
          do {
            const int nb = whole ? n_lrf : fill(fromB, T, a.r2_lrf);  // the list of pass A when it is complete
            for (int i = lane; i < nb; i += 32) {
              const float4 p = pt_at(i);
              if (DENSE && !(sqdist3_rn(kx, ky, kz, p.x, p.y, p.z) < a.r2_lrf)) continue;
              if (!(p.x == kx && p.y == ky && p.z == kz)) {
                const double vx = (double)__fsub_rn(p.x, kx), vy = (double)__fsub_rn(p.y, ky),
                             vz = (double)__fsub_rn(p.z, kz);
                if (vx * x[0] + vy * x[1] + vz * x[2] >= 0) ++plusX;
                if (vx * z[0] + vy * z[1] + vz * z[2] >= 0) ++plusZ;
              }
            }
          } while (!whole && fromB < T);
        }
        plusX = warp_sum(plusX);
        plusZ = warp_sum(plusZ);
        if (lrf_ok) {
          int sx = 2 * plusX - valid, sz = 2 * plusZ - valid;
          if (sx == 0 || sz == 0) {
            // Tie (about 2% of real keypoints): the reference looks at the 5 neighbours around the median of the
            // (d^2, index)-sorted valid list (shot_na_lrf.hpp:141-153).  Rank selection by bisection on the 63-bit
            // key (d^2 bits << 32 | index): 63 counting passes over the staged points + 5 successive-minimum passes.
            auto scan = [&](auto&& f) {  // key_of applies the LRF radius itself
              if (whole) {           // the list of pass A
                for (int i = lane; i < n_lrf; i += 32) f(pt_at(i), idx_at(i));
              } else {               // a neighbourhood of several fills: walk the staged points
                for (int e = lane; e < T; e += 32) f(s_pts[e], __float_as_int(s_pts[e].w));
              }
            };
            const unsigned long long kInvalid = ~0ull;
            auto key_of = [&](const float4& p, int idx) -> unsigned long long {
              float d2 = sqdist3_rn(kx, ky, kz, p.x, p.y, p.z);
              if (!(d2 < a.r2_lrf) || (p.x == kx && p.y == ky && p.z == kz)) return kInvalid;
              return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)idx;
            };
            const int m = valid / 2;
            unsigned long long lo = 0, hi = 0x7f80000000000000ull;  // every valid key is below hi
            while (lo < hi) {  // smallest K with #{key <= K} >= m - 1  ==  key of 0-based rank m - 2
              const unsigned long long mid = lo + ((hi - lo) >> 1);
              int c = 0;
              scan([&](const float4& p, int idx) { c += key_of(p, idx) <= mid ? 1 : 0; });
              c = warp_sum(c);
              if (c >= m - 1) hi = mid; else lo = mid + 1;
            }
            int cntX = 0, cntZ = 0;
            unsigned long long prev = lo;  // first of the five
            for (int j = 0; j < 5; ++j) {
              unsigned long long bk = kInvalid;
              float bx = 0.f, by = 0.f, bz = 0.f;
              scan([&](const float4& p, int idx) {
                unsigned long long k = key_of(p, idx);
                bool take = (j == 0) ? (k >= prev) : (k > prev);
                if (take && k < bk) { bk = k; bx = p.x; by = p.y; bz = p.z; }
              });
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
                float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o),
                      oz = __shfl_xor_sync(0xffffffffu, bz, o);
                if (ok < bk) { bk = ok; bx = ox; by = oy; bz = oz; }
              }
              if (bk == kInvalid) break;
              prev = bk;
              double vx = (double)__fsub_rn(bx, kx), vy = (double)__fsub_rn(by, ky), vz = (double)__fsub_rn(bz, kz);
              if (vx * x[0] + vy * x[1] + vz * x[2] > 0) ++cntX;
              if (vx * z[0] + vy * z[1] + vz * z[2] > 0) ++cntZ;
            }
            if (sx == 0) sx = (cntX < 3) ? -1 : 1;
            if (sz == 0) sz = (cntZ < 3) ? -1 : 1;
          }
          if (sx < 0) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
          if (sz < 0) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
          rf[0] = (float)x[0]; rf[1] = (float)x[1]; rf[2] = (float)x[2];
          rf[6] = (float)z[0]; rf[7] = (float)z[1]; rf[8] = (float)z[2];
          rf[3] = __fsub_rn(__fmul_rn(rf[7], rf[2]), __fmul_rn(rf[8], rf[1]));
          rf[4] = __fsub_rn(__fmul_rn(rf[8], rf[0]), __fmul_rn(rf[6], rf[2]));
          rf[5] = __fsub_rn(__fmul_rn(rf[6], rf[1]), __fmul_rn(rf[7], rf[0]));
        } else {
#pragma unroll
          for (int i = 0; i < 9; ++i) rf[i] = __int_as_float(0x7fc00000);
        }
        if (a.lrf_out && lane < 9) {
          float v = rf[0];
#pragma unroll
          for (int i = 1; i < 9; ++i)
            if (lane == i) v = rf[i];
          a.lrf_out[(size_t)kidx * 9 + lane] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) rf[i] = a.lrf_in[(size_t)kidx * 9 + i];
      }
      if (!a.do_desc) continue;  // next keypoint
      // Features::operator() drops keypoints whose frame is not finite (features.cpp:64-76)
      const bool frame_ok = isfinite(rf[0]) && isfinite(rf[3]) && isfinite(rf[6]);
      // ---------------------------------------------------------------- SHOT / CSHOT
      for (int j = lane; j < D; j += 32) hist[j] = 0;
      __syncwarp();
      float LRef = 0.f, aRef = 0.f, bRef = 0.f;
      if (COLOR && frame_ok) {
        float3 l = rgb_to_lab_norm(krgb, a.lab_lut);
        LRef = l.x; aRef = l.y; bRef = l.z;
      }
      // Discrete decisions (volume index, histogram bin, which neighbour bin) are taken exactly as the reference takes
      // them: on float-exact quantities, on the fp64 bin position (through an fp32 evaluation that falls back to the
      // fp64 one whenever it lands within 1e-4 of a rounding boundary: the two differ by < 5e-6), and on the radial
      // shells through float thresholds on d^2 that are equivalent to the reference's fp64 "sqrt(d^2) > r/2" tests
      // (computed on the host, stage_shot).  The CONTINUOUS interpolation weights (cosine / colour bin fraction, radial,
      // inclination, azimuth) are evaluated in fp32 with hardware approximations and short polynomials: their error
      // (< 1e-6) is far inside the 1e-4 descriptor bar and removes the software fp64 / libm code that dominated this loop.
      const float r34f = (float)((a.r_shot * 3) / 4), r14f = (float)(a.r_shot / 4);
      const float inv_r12f = (float)(1.0 / (a.r_shot / 2));
      const float inv_90f = (float)(1.0 / PST_RAD_90), inv_45f = (float)(1.0 / PST_RAD_45);
      int nshot = 0;
      if (frame_ok) {  // warp-uniform
        int fromC = 0;
        do {
        const int n_in = DENSE ? n_g : fill(fromC, T, a.r2_shot);
        // lane L takes the contiguous entries [L m, (L + 1) m): the list is in staging (= cell) order, so the 32 points
        // a warp accumulates at the same time are far apart and rarely hit the same histogram bin (same-address shared
        // atomics serialise); m is odd, which keeps the 16-bit list reads of the 32 lanes on different banks
        const int m_blk = a.blocked ? (((n_in + 31) >> 5) | 1) : ((n_in + 31) >> 5);
        const int i_base = a.blocked ? lane * m_blk : lane, i_step = a.blocked ? 1 : 32;
        for (int j = 0; j < m_blk; ++j) {
          const int i = i_base + j * i_step;
          if (i >= n_in) continue;
          int gi = 0, li = 0;
          float4 p;
          if (DENSE) {
            p = glist[i];
            gi = __float_as_int(p.w);
          } else {
            li = list[i];
            p = s_pts[li];
          }
          const float d2 = sqdist3_rn(kx, ky, kz, p.x, p.y, p.z);
          if (DENSE && !(d2 < a.r2_shot)) continue;
          ++nshot;
          float nrx, nry, nrz;
          if (DENSE) {
            const float4 nr = a.snrmS[gi];
            nrx = nr.x; nry = nr.y; nrz = nr.z;
          } else {
            nrx = s_nx[li]; nry = s_ny[li]; nrz = s_nz[li];
          }
          if (!finite3(nrx, nry, nrz)) continue;
          const float cosf_ = fminf(1.0f, fmaxf(-1.0f, dot3_rn(nrx, nry, nrz, rf[6], rf[7], rf[8])));
          if (d2 < 1E-30f) continue;  // reference: fabs(sqrt(d2)) < 1e-15
          // bin position bdS = ((1 + cos) * 10) / 2, stepS = floor(bdS + 0.5), fraction bdS - stepS
          int stepS;
          float fracS;
          {
            const float bd = fmaf(cosf_, 5.0f, 5.0f);
            const float t = bd + 0.5f, st = floorf(t), fr = t - st;
            if (fr < 1e-4f || fr > 0.9999f) {  // near a rounding boundary: the reference's fp64 evaluation decides
              double bdS = ((1.0 + (double)cosf_) * 10) / 2;
              const int s = (int)floor(bdS + 0.5);
              bdS -= s;
              stepS = s;
              fracS = (float)bdS;
            } else {
              stepS = (int)st;
              fracS = bd - st;
            }
          }
          const float inv_d = rsqrt_approx(d2);
          const float distance = d2 * inv_d;
          const float dx = __fsub_rn(p.x, kx), dy = __fsub_rn(p.y, ky), dz = __fsub_rn(p.z, kz);
          float xIn = dot3_rn(dx, dy, dz, rf[0], rf[1], rf[2]);
          float yIn = dot3_rn(dx, dy, dz, rf[3], rf[4], rf[5]);
          float zIn = dot3_rn(dx, dy, dz, rf[6], rf[7], rf[8]);
          if (fabsf(yIn) < 1E-30f) yIn = 0.f;
          if (fabsf(xIn) < 1E-30f) xIn = 0.f;
          if (fabsf(zIn) < 1E-30f) zIn = 0.f;
          const bool same_sign = (xIn > 0.f && yIn > 0.f) || (xIn < 0.f && yIn < 0.f);  // xIn * yIn > 0 in fp64
          int bit4 = ((yIn > 0.f) || ((yIn == 0.f) && (xIn < 0.f))) ? 1 : 0;
          int bit3 = ((xIn > 0.f) || ((xIn == 0.f) && (yIn > 0.f))) ? !bit4 : bit4;
          int di = ((bit4 << 3) + (bit3 << 2)) << 1;
          if (same_sign || (xIn == 0.f))
            di += (fabsf(xIn) >= fabsf(yIn)) ? 0 : 4;
          else
            di += (fabsf(xIn) > fabsf(yIn)) ? 4 : 0;
          di += zIn > 0.f ? 1 : 0;
          const bool outer = d2 >= a.t2_gt_r12;  // sqrt((double)d2) > r/2
          di += outer ? 2 : 0;

          const int volS = di * 11;
          float wS = 1.0f - fabsf(fracS);
          if (fracS > 0.f)
            hist_add(hist, volS + ((stepS + 1) % 10), fracS, fix_scale);
          else
            hist_add(hist, volS + ((stepS - 1 + 10) % 10), -fracS, fix_scale);
          int stepC = 0, volC = 0;
          float wC = 0.f;
          if (COLOR) {
            float4 lb;
            if (DENSE) lb = a.slabS[gi]; else lb = s_lab[li];
            const float cdist = __fdiv_rn(
                __fadd_rn(fabsf(__fsub_rn(LRef, lb.x)),
                          __fdiv_rn(__fadd_rn(fabsf(__fsub_rn(aRef, lb.y)), fabsf(__fsub_rn(bRef, lb.z))), 2.0f)),
                3.0f);
            const float cd = fminf(1.0f, fmaxf(0.0f, cdist));
            float fracC;
            {
              const float bd = cd * 30.0f;
              const float t = bd + 0.5f, st = floorf(t), fr = t - st;
              if (fr < 1e-4f || fr > 0.9999f) {
                double bdC = (double)cd * 30;
                const int s = (int)floor(bdC + 0.5);
                bdC -= s;
                stepC = s;
                fracC = (float)bdC;
              } else {
                stepC = (int)st;
                fracC = bd - st;
              }
            }
            volC = 352 + di * 31;
            wC = 1.0f - fabsf(fracC);
            if (fracC > 0.f)
              hist_add(hist, volC + ((stepC + 1) % 30), fracC, fix_scale);
            else
              hist_add(hist, volC + ((stepC - 1 + 30) % 30), -fracC, fix_scale);
          }
#define SHOT_NEIGHBOUR(DI, VAL)                                           \
  do {                                                                    \
    hist_add(hist, (DI) * 11 + stepS, (VAL), fix_scale);                  \
    if (COLOR) hist_add(hist, 352 + (DI) * 31 + stepC, (VAL), fix_scale); \
  } while (0)
          float wAdd = 0.f;
          if (outer) {
            const float rd = (distance - r34f) * inv_r12f;
            if (d2 >= a.t2_gt_r34)  // sqrt((double)d2) > 3r/4
              wAdd += 1.f - rd;
            else {
              wAdd += 1.f + rd;
              SHOT_NEIGHBOUR(di - 2, -rd);
            }
          } else {
            const float rd = (distance - r14f) * inv_r12f;
            if (d2 < a.t2_ge_r14)  // sqrt((double)d2) < r/4
              wAdd += 1.f + rd;
            else {
              wAdd += 1.f - rd;
              SHOT_NEIGHBOUR(di + 2, rd);
            }
          }
          // the reference branches on acos(z/d) > 90 deg (ties: z <= 0), which is z <= 0 for every representable input
          const float incl = acos_fast(fminf(1.f, fmaxf(-1.f, zIn * inv_d)));
          if (zIn <= 0.f) {
            const float id = (incl - (float)PST_RAD_135) * inv_90f;
            if (id > 0.f) {
              wAdd += 1.f - id;
            } else {
              wAdd += 1.f + id;
              SHOT_NEIGHBOUR(di + 1, -id);
            }
          } else {
            const float id = (incl - (float)PST_RAD_45) * inv_90f;
            if (id < 0.f) {
              wAdd += 1.f + id;
            } else {
              wAdd += 1.f - id;
              SHOT_NEIGHBOUR(di - 1, id);
            }
          }
          if (yIn != 0.f || xIn != 0.f) {
            const float az = atan2_fast(yIn, xIn);
            const int sel = di >> 2;
            float ad = (az - (float)(-PST_RAD_PI_7_8 + PST_RAD_45 * sel)) * inv_45f;
            ad = fmaxf(-0.5f, fminf(ad, 0.5f));
            if (ad > 0.f) {
              wAdd += 1.f - ad;
              SHOT_NEIGHBOUR((di + 4) % 32, ad);
            } else {
              wAdd += 1.f + ad;
              SHOT_NEIGHBOUR((di - 4 + 32) % 32, -ad);
            }
          }
#undef SHOT_NEIGHBOUR
          hist_add(hist, volS + stepS, wS + wAdd, fix_scale);
          if (COLOR) hist_add(hist, volC + stepC, wC + wAdd, fix_scale);
        }
        } while (!DENSE && fromC < T);
      }
      nshot = warp_sum(nshot);
      if (lane == 0 && a.nbr_counts && frame_ok) atomicAdd(&a.nbr_counts[1], (unsigned long long)nshot);
      __syncwarp();
      {
        float* out = a.desc_out + (size_t)kidx * D;
        if (!frame_ok || nshot < 5) {
          for (int j = lane; j < D; j += 32) out[j] = __int_as_float(0x7fc00000);
        } else {
          double acc = 0.0;
          for (int j = lane; j < D; j += 32) {
            float s = __fmul_rn((float)hist[j], fix_inv);
            acc += (double)__fmul_rn(s, s);
          }
          acc = warp_sum(acc);
          float nrm = (float)sqrt(acc);
          for (int j = lane; j < D; j += 32) out[j] = __fdiv_rn(__fmul_rn((float)hist[j], fix_inv), nrm);
        }
      }
      __syncwarp();
    }
  }
}

// The nine cell runs (three x-neighbours are contiguous in the cell-sorted surface) of every work item, one thread
// per (item, run): 18 binary searches that used to sit, serially, at the head of every item inside k_shot.  The
// largest 27-cell population sizes the per-warp lists of the dense mode.
__global__ void k_item_ranges(const unsigned long long* __restrict__ kp_keys, const int* __restrict__ item_start,
                              const int* __restrict__ n_items_ptr, const unsigned long long* __restrict__ skeys,
                              const long long* __restrict__ surf_off, long long* item_beg, int* item_len) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int item = (int)(t / 9), j = (int)(t % 9);
  const bool live = item < *n_items_ptr;
  long long beg = 0, end = 0;
  if (live) {
    const unsigned long long key = kp_keys[item_start[item]];
    const unsigned cloud = (unsigned)(key >> 48);
    const int cz = (int)((key >> 32) & 0xffff), cy = (int)((key >> 16) & 0xffff), cx = (int)(key & 0xffff);
    const int y = cy + j % 3 - 1, z = cz + j / 3 - 1;
    const long long lo = surf_off[cloud], hi = surf_off[cloud + 1];
    if (hi - lo <= kChunk) {  // whole-cloud item (prep.cu:k_item_heads): one run, the cloud itself
      if (j == 0) { beg = lo; end = hi; }
    } else if (y >= 0 && y <= 65535 && z >= 0 && z <= 65535) {
      beg = lower_bound_u64(skeys, lo, hi, grid_key(cloud, max(cx - 1, 0), y, z));
      end = lower_bound_u64(skeys, beg, hi, grid_key(cloud, min(cx + 1, 65535), y, z) + 1ull);
    }
    item_beg[t] = beg;
    item_len[t] = (int)(end - beg);
  }
}
__global__ void k_item_population(const int* __restrict__ item_len, const int* __restrict__ n_items_ptr,
                                  unsigned long long* max_pop) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= *n_items_ptr) return;
  long long T = 0;
  for (int j = 0; j < 9; ++j) T += item_len[(size_t)item * 9 + j];
  atomicMax(max_pop, (unsigned long long)T);
}

}  // namespace

static int shot_warps(bool color) { return color ? kWarpsCshot : kWarpsShot; }
static size_t shot_smem_for(bool color, int stage_cap) {
  const int D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  const size_t W = (size_t)shot_warps(color);
  return (size_t)stage_cap * (color ? 44 : 28) + sizeof(unsigned) * W * D +
         (stage_cap ? sizeof(unsigned short) * W * kList : 0);
}
size_t shot_smem_bytes(bool color) { return shot_smem_for(color, kChunk); }

// Runs the fused LRF + descriptor kernel over the keypoint items prepared by stage_grid.
// smallest float f with sqrt((double)f) > r (strict) or >= r: "sqrt((double)d2) > r"  <=>  d2 >= f for float d2
static float shell_threshold(double r, bool strict) {
  auto pass = [&](float f) { double s = std::sqrt((double)f); return strict ? s > r : s >= r; };
  float f = (float)(r * r);
  while (f > 0.f && pass(f)) f = std::nextafterf(f, 0.f);
  while (!pass(f)) f = std::nextafterf(f, INFINITY);
  return f;
}

int stage_shot(pcdb_ctx* ctx, int64_t n_surf, int64_t Q, bool color, double r_lrf, double r_shot, bool do_lrf,
               bool do_desc, const float* lrf_in_d, float* lrf_out_d, float* desc_out_d) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  if (Q == 0) return PCDB_OK;
  PCDB_CUDA(w.scalars.ensure(256));
  PCDB_CUDA(cudaMemsetAsync(w.scalars.p, 0, 256, st));
  if (color && do_desc) {
    PCDB_CUDA(w.slabS4.ensure(sizeof(float4) * (n_surf + 1)));
    if (n_surf > 0) {
      k_lab<<<cdiv(n_surf, 256), 256, 0, st>>>(w.surfS4.as<float4>(), n_surf, ctx->lab_lut_d, w.slabS4.as<float4>());
      PCDB_LAUNCH_CHECK();
    }
  }
  ShotArgs a;
  a.surfS = w.surfS4.as<float4>();
  a.snrmS = w.snrmS4.as<float4>();
  a.slabS = w.slabS4.as<float4>();
  a.skeys = w.gkeys2.as<unsigned long long>();
  a.surf_off = w.surf_off.as<long long>();
  a.kp4 = w.kp4.as<float4>();
  a.kp_order = w.kvals2.as<int>();
  a.kp_keys = w.kkeys2.as<unsigned long long>();
  a.item_start = w.item_start.as<int>();
  a.n_items_ptr = w.item_id.as<int>() + Q;
  a.r_lrf = r_lrf;
  a.r_shot = r_shot;
  a.r2_lrf = (float)(r_lrf * r_lrf);    // pcl::KdTreeFLANN::radiusSearch: float(radius*radius)
  a.r2_shot = (float)(r_shot * r_shot);
  a.t2_gt_r12 = shell_threshold(r_shot / 2, true);
  a.t2_gt_r34 = shell_threshold((r_shot * 3) / 4, true);
  a.t2_ge_r14 = shell_threshold(r_shot / 4, false);
  a.lrf_in = lrf_in_d;
  a.lrf_out = lrf_out_d;
  a.desc_out = desc_out_d;
  a.do_lrf = do_lrf ? 1 : 0;
  a.do_desc = do_desc ? 1 : 0;
  a.work_counter = w.scalars.as<int>();
  a.nbr_counts = reinterpret_cast<unsigned long long*>(w.scalars.as<char>() + 16);
  a.lab_lut = ctx->lab_lut_d;
  const size_t smem = shot_smem_bytes(color);
  // persistent grid: a multiple of the SM count, bounded by the number of keypoints
  int per_sm = color ? 1 : 2;
  int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * per_sm, Q);
  // dense scenes: size the per-warp in-radius lists from the largest 27-cell population of this batch (one sync)
  unsigned long long* max_pop = reinterpret_cast<unsigned long long*>(w.scalars.as<char>() + 48);
  PCDB_CUDA(w.item_beg.ensure(sizeof(long long) * 9 * (size_t)(Q + 1)));
  PCDB_CUDA(w.item_len.ensure(sizeof(int) * 9 * (size_t)(Q + 1)));
  k_item_ranges<<<cdiv(Q * 9, 256), 256, 0, st>>>(a.kp_keys, a.item_start, a.n_items_ptr, a.skeys, a.surf_off,
                                                  w.item_beg.as<long long>(), w.item_len.as<int>());
  PCDB_LAUNCH_CHECK();
  k_item_population<<<cdiv(Q, 256), 256, 0, st>>>(w.item_len.as<int>(), a.n_items_ptr, max_pop);
  PCDB_LAUNCH_CHECK();
  a.item_beg = w.item_beg.as<long long>();
  a.item_len = w.item_len.as<int>();
  unsigned long long h_pop = 0;
  PCDB_TRY(pcdb_read_small(ctx, &h_pop, max_pop, sizeof(h_pop)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  a.glist = nullptr;
  a.gcap = 0;
  a.dense = 0;
  a.stage_cap = kChunk;
  a.n_kp = Q;
  a.item_id = w.item_id.as<int>();
  a.item_head = w.item_head.as<int>();
  a.work_counter_dense = w.scalars.as<int>() + 2;
  // dense launch: no staging arrays; every warp takes keypoints on its own
  const int kWarps = shot_warps(color), kThreads = kWarps * 32;
  const int dense_grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * (color ? 1 : 2), cdiv(Q, kWarps));
  if (h_pop > (unsigned long long)kChunk) {
    a.gcap = (long long)((h_pop + 31) & ~31ull);
    const size_t bytes = sizeof(float4) * (size_t)a.gcap * (size_t)dense_grid * kWarps;
    if (bytes > (32ull << 30))
      return ctx->fail(PCDB_E_INVALID, "a 27-cell neighbourhood holds %llu points: radius too large for this cloud density",
                       h_pop);
    PCDB_CUDA(w.shot_glist.ensure(bytes));
    a.glist = w.shot_glist.as<float4>();
  }
  // keypoints drawn per item (the helper scheme of the staged launch)
  PCDB_CUDA(w.item_next.ensure(sizeof(int) * (size_t)(Q + 1)));
  PCDB_CUDA(cudaMemsetAsync(w.item_next.p, 0, sizeof(int) * (size_t)(Q + 1), st));
  a.item_next = w.item_next.as<int>();
  static const int blocked = [] { const char* e = getenv("PCDB_SHOT_BLOCKED"); return e ? atoi(e) : 1; }();
  a.blocked = blocked;
  // Staged launch.  With both stages requested it runs as two launches of the same kernel — frames first, descriptors
  // second (reading the frames back): the 32 resident warps of an SM then loop over ~5 KB of code instead of ~25 KB
  // spread over five phases, which is what the 6 KB L0 instruction caches can hold (one fused launch measured 2.4
  // no-instruction stalls per issue).  The cloud is staged twice; that is 57 KB per item against ~4 M instructions.
  static const int split_env = [] { const char* e = getenv("PCDB_SHOT_SPLIT"); return e ? atoi(e) : 1; }();
  const bool split = split_env && do_lrf && do_desc && lrf_out_d != nullptr;
  for (int phase = 0; phase < (split ? 2 : 1); ++phase) {
    ShotArgs b = a;
    if (split) {
      b.do_lrf = phase == 0;
      b.do_desc = phase == 1;
      b.lrf_in = lrf_out_d;
      if (phase == 1) {
        PCDB_CUDA(cudaMemsetAsync(w.scalars.p, 0, 4, st));  // the staged launch's work counter
        PCDB_CUDA(cudaMemsetAsync(w.item_next.p, 0, sizeof(int) * (size_t)(Q + 1), st));
      }
    }
    if (color) {
      PCDB_CUDA(cudaFuncSetAttribute(k_shot<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shot<true, false><<<grid, kThreads, smem, st>>>(b);
    } else {
      PCDB_CUDA(cudaFuncSetAttribute(k_shot<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shot<false, false><<<grid, kThreads, smem, st>>>(b);
    }
    PCDB_LAUNCH_CHECK();
  }
  if (a.glist) {  // the keypoints whose 27 cells do not fit the stage: second launch, warp per keypoint
    a.dense = 1;
    a.stage_cap = 0;
    const size_t dsmem = shot_smem_for(color, 0);
    if (color) {
      PCDB_CUDA(cudaFuncSetAttribute(k_shot<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
      k_shot<true, true><<<dense_grid, kThreads, dsmem, st>>>(a);
    } else {
      k_shot<false, true><<<dense_grid, kThreads, dsmem, st>>>(a);
    }
    PCDB_LAUNCH_CHECK();
  }
  return PCDB_OK;
}

int stage_shot_counts(pcdb_ctx* ctx, unsigned long long out[2]) {
  PCDB_CUDA(cudaMemcpyAsync(out, ctx->ws.scalars.as<char>() + 16, sizeof(unsigned long long) * 2,
                            cudaMemcpyDeviceToHost, ctx->stream));
  PCDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return PCDB_OK;
}
