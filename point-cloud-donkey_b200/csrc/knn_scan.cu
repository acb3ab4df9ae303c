// knn_scan.cu — K7 / K7chi: exact kNN over the device-resident codebook with the FLANN functor arithmetic.
//
// Replaces ActivationStrategyKNN::activateKNN (activation_strategy/activation_strategy_knn.h:41-126) in its
// FLANNExactMatch form and the functors flann::L2<float> / flann::ChiSquareDistance<float> behind
// utils/distance.h:42-75.  Distances are accumulated per (query, codeword) pair in exactly FLANN's order
// (4-wide groups ((d0^2+d1^2)+d2^2)+d3^2 added to the running sum; chi^2 element by element with an
// IEEE division), so the values — and with them the neighbour order and the 2*sigma^2 vote filter — are bit-exact
// with the CPU functor.  Ties resolve to the lower row.
//
// Two uses: (1) the complete exact scan (small codebooks, chi^2, and the fallback of the tensor-core path),
// register-tiled 64 queries x 64 codewords per CTA with both operands staged through shared memory;
// (2) k_rerank: exact distances for the candidate lists produced by the tcgen05 GEMM (knn_gemm.cu).
#include "common.cuh"
#include "stages.h"

namespace {

constexpr int TQ = 64, TC = 64, TD = 32;  // tile: queries x codewords x dims per stage
constexpr int PADD = TD + 4;              // row pitch in floats (keeps 16-byte alignment, spreads banks)
constexpr int KMAX = PCDB_MAX_K + 1;

struct TopK {  // per-query running list in shared memory, ascending (dist, idx)
  float d[KMAX];
  int i[KMAX];
};

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) {
  return da < db || (da == db && ia < ib);
}

template <int DIST>
__device__ __forceinline__ float accum4(float acc, const float4 q, const float4 c) {
  if (DIST == PCDB_DIST_EUCLIDEAN) {
    float d0 = __fsub_rn(q.x, c.x), d1 = __fsub_rn(q.y, c.y), d2 = __fsub_rn(q.z, c.z), d3 = __fsub_rn(q.w, c.w);
    float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)),
                        __fmul_rn(d3, d3));
    return __fadd_rn(acc, t);
  } else {
    // functor(codeword, query): a = codeword, b = query (codeword_distribution.cpp:87); symmetric bit for bit
    const float qa[4] = {q.x, q.y, q.z, q.w}, ca[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float s = __fadd_rn(ca[e], qa[e]);
      if (s > 0.f) {
        float diff = __fsub_rn(ca[e], qa[e]);
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(diff, diff), s));
      }
    }
    return acc;
  }
}

// Chi-squared tile pass: the same terms in the same order as the exact functor, but with the fast reciprocal-multiply
// division (__fdividef, <= 2 ulp) instead of the IEEE one (which costs ~10 instructions per dimension).  The result is
// only used to decide which pairs deserve the exact evaluation, see k_knn_scan.
__device__ __forceinline__ float accum4_chi2_fast(float acc, const float4 q, const float4 c) {
  const float qa[4] = {q.x, q.y, q.z, q.w}, ca[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float s = __fadd_rn(ca[e], qa[e]);
    const float diff = __fsub_rn(ca[e], qa[e]);
    const float t = __fdividef(__fmul_rn(diff, diff), s);
    acc = __fadd_rn(acc, s > 0.f ? t : 0.f);
  }
  return acc;
}

// the exact functor over one pair of rows in FLANN order (one thread; used for the few pairs that pass the filter)
template <int DIST>
__device__ float exact_pair(const float* __restrict__ q, const float* __restrict__ c, int D) {
  float acc = 0.f;
  for (int j = 0; j < D; j += 4)
    acc = accum4<DIST>(acc, *reinterpret_cast<const float4*>(q + j), __ldg(reinterpret_cast<const float4*>(c + j)));
  return acc;
}

// grid = (q tiles, splits).  Each CTA scans codeword rows [split*rows_per_split, ...) for its 64 queries and writes
// its K best per query to part_* [split][q][K].
template <int DIST>
__global__ void __launch_bounds__(256) k_knn_scan(const float* __restrict__ queries, long long Q,
                                                  const float* __restrict__ words, long long N, int D, int K,
                                                  long long rows_per_split, float* part_d, int* part_i) {
  __shared__ __align__(16) float sQ[TQ][PADD];
  __shared__ __align__(16) float sC[TC][PADD];
  __shared__ float sDist[TQ][TC + 1];
  __shared__ TopK sTop[TQ];
  __shared__ int sCnt[TQ];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long q0 = (long long)blockIdx.x * TQ;
  const long long n_begin = (long long)blockIdx.y * rows_per_split;
  const long long n_end = min(N, n_begin + rows_per_split);
  if (threadIdx.x < TQ) sCnt[threadIdx.x] = 0;
  // |approx - exact| <= rho * exact with rho = (D + 16) 2^-23: 2.5 ulp per term (fast vs IEEE division) plus the
  // propagation of those differences through D sequential fp32 additions of non-negative terms
  const float chi2_slack = 1.0f + 2.0f * (float)(D + 16) * 1.1920929e-07f;
  __syncthreads();
  for (long long n0 = n_begin; n0 < n_end; n0 += TC) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < D; d0 += TD) {
      // stage 64 x 32 floats of each operand: 512 float4 per operand, 2 per thread
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        int f = threadIdx.x + r * 256;  // float4 index
        int row = f >> 3, col = (f & 7) * 4;
        long long qr = min(q0 + row, Q - 1);
        long long cr = min(n0 + row, N - 1);
        float4 qv = *reinterpret_cast<const float4*>(queries + qr * D + d0 + col);
        float4 cv = __ldg(reinterpret_cast<const float4*>(words + cr * D + d0 + col));
        *reinterpret_cast<float4*>(&sQ[row][col]) = qv;
        *reinterpret_cast<float4*>(&sC[row][col]) = cv;
      }
      __syncthreads();
#pragma unroll
      for (int g = 0; g < TD; g += 4) {
        float4 qv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&sQ[ty + 16 * i][g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const float4*>(&sC[tx + 16 * j][g]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[i][j] = DIST == PCDB_DIST_CHISQUARED ? accum4_chi2_fast(acc[i][j], qv[i], cv[j])
                                                     : accum4<DIST>(acc[i][j], qv[i], cv[j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sDist[ty + 16 * i][tx + 16 * j] = acc[i][j];
    __syncthreads();
    if (threadIdx.x < TQ) {
      const int q = threadIdx.x;
      TopK& t = sTop[q];
      int cnt = sCnt[q];
      const int lim = (int)min((long long)TC, n_end - n0);
      for (int c = 0; c < lim; ++c) {
        float d = sDist[q][c];
        int idx = (int)(n0 + c);
        if (DIST == PCDB_DIST_CHISQUARED) {
          // d is the fast approximation (relative error <= rho, see above): a pair that cannot reach the running
          // k-th best even at the favourable end of that interval is dropped, the others get the exact functor, so
          // the list holds exact FLANN-order values and the result equals the all-exact scan bit for bit
          if (cnt == K && d > t.d[K - 1] * chi2_slack) continue;
          d = exact_pair<DIST>(queries + min(q0 + q, Q - 1) * D, words + (n0 + c) * D, D);
        }
        if (cnt == K && !cand_less(d, idx, t.d[K - 1], t.i[K - 1])) continue;
        int p = (cnt < K) ? cnt++ : K - 1;
        while (p > 0 && cand_less(d, idx, t.d[p - 1], t.i[p - 1])) {
          t.d[p] = t.d[p - 1];
          t.i[p] = t.i[p - 1];
          --p;
        }
        t.d[p] = d;
        t.i[p] = idx;
      }
      sCnt[q] = cnt;
    }
    __syncthreads();
  }
  if (threadIdx.x < TQ && q0 + threadIdx.x < Q) {
    const int q = threadIdx.x;
    long long o = ((long long)blockIdx.y * Q + q0 + q) * K;
    for (int j = 0; j < K; ++j) {
      bool ok = j < sCnt[q];
      part_d[o + j] = ok ? sTop[q].d[j] : __int_as_float(0x7f800000);
      part_i[o + j] = ok ? sTop[q].i[j] : -1;
    }
  }
}

// Merge S partial lists per query, apply the N<=k shortcut semantics and the detection-time distance-ratio test
// (activation_strategy_knn.h:50-54,75-85).  One thread per query.
__global__ void k_knn_merge(const float* __restrict__ part_d, const int* __restrict__ part_i, int S, long long Q,
                            int K /* searched */, int k /* returned */, int use_ratio, float ratio_thr,
                            long long row_base, int* idx_out, float* dist_out, int* cnt_out) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float bd[KMAX];
  int bi[KMAX];
  int cnt = 0;
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < K; ++j) {
      long long o = ((long long)s * Q + q) * K + j;
      int idx = part_i[o];
      if (idx < 0) break;
      float d = part_d[o];
      if (cnt == K && !cand_less(d, idx, bd[K - 1], bi[K - 1])) break;  // lists are ascending
      int p = (cnt < K) ? cnt++ : K - 1;
      while (p > 0 && cand_less(d, idx, bd[p - 1], bi[p - 1])) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
      }
      bd[p] = d;
      bi[p] = idx;
    }
  int use = min(cnt, k);
  if (use_ratio && k == 1 && cnt >= 2) {
    if (__fdiv_rn(bd[0], bd[1]) > ratio_thr) use = 0;
  }
  for (int j = 0; j < k; ++j) {
    idx_out[q * k + j] = j < use ? (int)(bi[j] + row_base) : -1;
    dist_out[q * k + j] = j < use ? bd[j] : __int_as_float(0x7fc00000);
  }
  cnt_out[q] = use;
}

// ---- exact re-rank of GEMM candidates ---------------------------------------------------------------------------
// One warp per query.  cand_* hold, per (query, split), up to cap (row, approximate distance) pairs; only candidates
// whose approximate distance can still contain the true k nearest (<= kth-best-approx + margin) are evaluated, one
// lane per candidate, each running the full FLANN-order chain.  The K best are then picked by successive minima.
template <int DIST>
__global__ void __launch_bounds__(256) k_rerank(const float* __restrict__ queries, long long Q,
                                                const float* __restrict__ words, int D, int S, int cap,
                                                const int* __restrict__ cand_idx, const float* __restrict__ cand_apx,
                                                const int* __restrict__ cand_cnt, const float* __restrict__ cand_thr,
                                                int K, float* exact_scratch, float* part_d, int* part_i,
                                                unsigned long long* n_evaluated) {
  extern __shared__ __align__(16) float s_q[];  // 8 warps x D
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= Q) return;
  float* qv = s_q + (size_t)warp * D;
  for (int j = lane * 4; j < D; j += 128)
    *reinterpret_cast<float4*>(qv + j) = *reinterpret_cast<const float4*>(queries + q * D + j);
  __syncwarp();
  // final pruning threshold = min over splits of (kth best approx + margin)
  float thr = __int_as_float(0x7f800000);
  for (int s = lane; s < S; s += 32) thr = fminf(thr, cand_thr[(long long)s * Q + q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) thr = fminf(thr, __shfl_xor_sync(0xffffffffu, thr, o));
  const long long base = q * (long long)S * cap;
  int evaluated = 0;
  for (int s = 0; s < S; ++s) {
    const int cnt = min(cap, cand_cnt[(long long)s * Q + q]);
    for (int c = lane; c < cnt; c += 32) {
      const long long o = base + (long long)s * cap + c;
      float ex = __int_as_float(0x7f800000);
      if (cand_apx[o] <= thr) {
        const float* w = words + (long long)cand_idx[o] * D;
        float acc = 0.f;
        for (int j = 0; j < D; j += 4)
          acc = accum4<DIST>(acc, *reinterpret_cast<const float4*>(qv + j),
                             __ldg(reinterpret_cast<const float4*>(w + j)));
        ex = acc;
        ++evaluated;
      }
      exact_scratch[o] = ex;
    }
  }
  __syncwarp();
  evaluated = warp_sum(evaluated);
  if (lane == 0 && n_evaluated) atomicAdd(n_evaluated, (unsigned long long)evaluated);
  // successive minima over (exact, row)
  float last_d = -1.f;
  int last_i = -1;
  for (int j = 0; j < K; ++j) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int s = 0; s < S; ++s) {
      const int cnt = min(cap, cand_cnt[(long long)s * Q + q]);
      for (int c = lane; c < cnt; c += 32) {
        const long long o = base + (long long)s * cap + c;
        float d = exact_scratch[o];
        int idx = cand_idx[o];
        if (!(d < __int_as_float(0x7f800000))) continue;
        if (!cand_less(last_d, last_i, d, idx)) continue;  // already emitted
        if (cand_less(d, idx, bd, bi)) {
          bd = d;
          bi = idx;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, bd, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (cand_less(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
      }
    }
    if (lane == 0) {
      bool ok = bi != 0x7fffffff;
      part_d[q * K + j] = ok ? bd : __int_as_float(0x7f800000);
      part_i[q * K + j] = ok ? bi : -1;
    }
    last_d = bd;
    last_i = bi;
    if (bi == 0x7fffffff) {
      for (int jj = j + 1; jj < K; ++jj)
        if (lane == 0) {
          part_d[q * K + jj] = __int_as_float(0x7f800000);
          part_i[q * K + jj] = -1;
        }
      break;
    }
  }
}

// all rows, in row order (N <= k shortcut): distances only
template <int DIST>
__global__ void k_all_rows(const float* __restrict__ queries, long long Q, const float* __restrict__ words, int N,
                           int D, int k, long long row_base, int* idx_out, float* dist_out, int* cnt_out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Q * k) return;
  long long q = t / k;
  int j = (int)(t % k);
  if (j == 0) cnt_out[q] = N;
  if (j >= N) {
    idx_out[t] = -1;
    dist_out[t] = __int_as_float(0x7fc00000);
    return;
  }
  float acc = 0.f;
  for (int e = 0; e < D; e += 4)
    acc = accum4<DIST>(acc, *reinterpret_cast<const float4*>(queries + q * D + e),
                       *reinterpret_cast<const float4*>(words + (long long)j * D + e));
  idx_out[t] = (int)(j + row_base);
  dist_out[t] = acc;
}

template <int DIST>
__global__ void k_pair_dist(const float* __restrict__ a, const float* __restrict__ b, long long n, int D, float* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  int e = 0;
  for (; e + 3 < D; e += 4)
    acc = accum4<DIST>(acc, *reinterpret_cast<const float4*>(b + i * D + e),
                       *reinterpret_cast<const float4*>(a + i * D + e));
  for (; e < D; ++e) {  // FLANN's tail loop
    float x = a[i * D + e], y = b[i * D + e];
    if (DIST == PCDB_DIST_EUCLIDEAN) {
      float d = __fsub_rn(x, y);
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    } else {
      float s = __fadd_rn(x, y);
      if (s > 0.f) {
        float d = __fsub_rn(x, y);
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(d, d), s));
      }
    }
  }
  out[i] = acc;
}

// ---- chi^2 sandwich, pooled second pass -------------------------------------------------------------------------
// exact functor value of every pooled pair, evaluated IN POOL ORDER: the sweep appends slice by slice, so consecutive
// entries hit the same L2-sized part of the codebook (grouped by query first, the 27 M pairs of a C3 step re-read the
// 1.5 GB of fp32 rows from HBM ~25 times over).  The value goes into the entry's second word.
template <int DIST>
__global__ void __launch_bounds__(128) k_pool_eval(const float* __restrict__ queries, const float* __restrict__ words,
                                                   int D, long long n, int2* pool_rc, const int* __restrict__ pool_q,
                                                   const int* __restrict__ skip) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int q = pool_q[e];
  if (skip && skip[q]) return;
  pool_rc[e].y = __float_as_int(exact_pair<DIST>(queries + (long long)q * D, words + (long long)pool_rc[e].x * D, D));
}
__global__ void k_pool_scatter(const int2* __restrict__ pool_rc, const int* __restrict__ pool_q, long long n,
                               const int* __restrict__ q_off, const int* __restrict__ skip, int* q_fill, int* csr_row,
                               float* csr_d) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int q = pool_q[e];
  if (skip && skip[q]) return;  // a query past its pool cap: its entries are dropped, the fallback path searches it
  const int p = q_off[q] + atomicAdd(q_fill + q, 1);
  const int2 rc = pool_rc[e];
  csr_row[p] = rc.x;
  csr_d[p] = __int_as_float(rc.y);
}
// one warp per query: K successive minima over (distance, row); the order inside a CSR segment is arbitrary
__global__ void k_chi_select(long long Q, int K, const int* __restrict__ q_off, const int* __restrict__ csr_row,
                             const float* __restrict__ csr_d, float* part_d, int* part_i) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int beg = q_off[q], end = q_off[q + 1];
  float last_d = -1.f;
  int last_i = -1;
  for (int j = 0; j < K; ++j) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int c = beg + lane; c < end; c += 32) {
      const float d = csr_d[c];
      const int idx = csr_row[c];
      if (!(d < __int_as_float(0x7f800000))) continue;
      if (!cand_less(last_d, last_i, d, idx)) continue;  // already emitted
      if (cand_less(d, idx, bd, bi)) {
        bd = d;
        bi = idx;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, bd, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (cand_less(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
      }
    }
    const bool ok = bi != 0x7fffffff;
    if (lane == 0) {
      part_d[q * K + j] = ok ? bd : __int_as_float(0x7f800000);
      part_i[q * K + j] = ok ? bi : -1;
    }
    last_d = bd;
    last_i = bi;
    if (!ok) {
      for (int jj = j + 1; jj < K; ++jj)
        if (lane == 0) {
          part_d[q * K + jj] = __int_as_float(0x7f800000);
          part_i[q * K + jj] = -1;
        }
      break;
    }
  }
}

}  // namespace

int stage_pair_distances(pcdb_ctx* ctx, const float* a_d, const float* b_d, int64_t n, int D, int dist_type,
                         float* out_d) {
  if (n == 0) return PCDB_OK;
  if (D % 4 != 0) return ctx->fail(PCDB_E_UNSUPPORTED, "row length must be a multiple of 4");
  if (dist_type == PCDB_DIST_CHISQUARED)
    k_pair_dist<PCDB_DIST_CHISQUARED><<<cdiv(n, 128), 128, 0, ctx->stream>>>(a_d, b_d, n, D, out_d);
  else
    k_pair_dist<PCDB_DIST_EUCLIDEAN><<<cdiv(n, 128), 128, 0, ctx->stream>>>(a_d, b_d, n, D, out_d);
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// Exact scan of queries_d (Q x D, device) against the uploaded codebook; results to ws.knn_* (device).
int stage_knn_scan(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, bool use_ratio,
                   float ratio_thr) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  PCDB_CUDA(w.knn_idx.ensure(sizeof(int) * (Q * k + 1)));
  PCDB_CUDA(w.knn_dist.ensure(sizeof(float) * (Q * k + 1)));
  PCDB_CUDA(w.knn_cnt.ensure(sizeof(int) * (Q + 1)));
  if (Q == 0) return PCDB_OK;
  if (cb.D % TD != 0) return ctx->fail(PCDB_E_UNSUPPORTED, "descriptor length %d is not a multiple of %d", cb.D, TD);
  if (cb.N <= k) {
    if (dist_type == PCDB_DIST_CHISQUARED)
      k_all_rows<PCDB_DIST_CHISQUARED><<<cdiv(Q * k, 128), 128, 0, st>>>(queries_d, Q, cb.words.as<float>(),
                                                                         (int)cb.N, cb.D, k, cb.row_base + cb.word_lo,
                                                                         w.knn_idx.as<int>(), w.knn_dist.as<float>(),
                                                                         w.knn_cnt.as<int>());
    else
      k_all_rows<PCDB_DIST_EUCLIDEAN><<<cdiv(Q * k, 128), 128, 0, st>>>(queries_d, Q, cb.words.as<float>(),
                                                                        (int)cb.N, cb.D, k, cb.row_base + cb.word_lo,
                                                                        w.knn_idx.as<int>(), w.knn_dist.as<float>(),
                                                                        w.knn_cnt.as<int>());
    PCDB_LAUNCH_CHECK();
    return PCDB_OK;
  }
  const int K = use_ratio ? k + 1 : k;
  const int qtiles = (int)cdiv(Q, TQ);
  // enough CTAs for ~4 waves, split along the codebook in multiples of the 64-row tile
  int64_t want = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 8) / qtiles);
  int64_t tiles = cdiv(cb.N, TC);
  int S = (int)std::min<int64_t>(want, tiles);
  S = std::min(S, 65535);
  int64_t rows_per_split = cdiv(tiles, S) * TC;
  S = (int)cdiv(cb.N, rows_per_split);
  PCDB_CUDA(w.knn_part_d.ensure(sizeof(float) * ((int64_t)S * Q * K + 1)));
  PCDB_CUDA(w.knn_part_i.ensure(sizeof(int) * ((int64_t)S * Q * K + 1)));
  dim3 grid(qtiles, S);
  if (dist_type == PCDB_DIST_CHISQUARED)
    k_knn_scan<PCDB_DIST_CHISQUARED><<<grid, 256, 0, st>>>(queries_d, Q, cb.words.as<float>(), cb.N, cb.D, K,
                                                           rows_per_split, w.knn_part_d.as<float>(),
                                                           w.knn_part_i.as<int>());
  else
    k_knn_scan<PCDB_DIST_EUCLIDEAN><<<grid, 256, 0, st>>>(queries_d, Q, cb.words.as<float>(), cb.N, cb.D, K,
                                                          rows_per_split, w.knn_part_d.as<float>(),
                                                          w.knn_part_i.as<int>());
  PCDB_LAUNCH_CHECK();
  k_knn_merge<<<cdiv(Q, 128), 128, 0, st>>>(w.knn_part_d.as<float>(), w.knn_part_i.as<int>(), S, Q, K, k,
                                            use_ratio ? 1 : 0, ratio_thr, cb.row_base + cb.word_lo, w.knn_idx.as<int>(),
                                            w.knn_dist.as<float>(), w.knn_cnt.as<int>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// Exact re-rank of the GEMM candidate lists (see knn_gemm.cu): the K best per query by the exact functor -> ws.knn_part_*.
int stage_knn_rerank(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int K, int S, int cap, int dist_type) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  PCDB_CUDA(w.knn_part_d.ensure(sizeof(float) * (Q * K + 1)));
  PCDB_CUDA(w.knn_part_i.ensure(sizeof(int) * (Q * K + 1)));
  PCDB_CUDA(w.cand_exact.ensure(sizeof(float) * ((int64_t)Q * S * cap + 1)));
  const size_t smem = sizeof(float) * 8 * cb.D;
  unsigned long long* n_eval = reinterpret_cast<unsigned long long*>(w.scalars.as<char>() + 64);
  if (dist_type == PCDB_DIST_CHISQUARED) {
    PCDB_CUDA(cudaFuncSetAttribute(k_rerank<PCDB_DIST_CHISQUARED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    k_rerank<PCDB_DIST_CHISQUARED><<<cdiv(Q, 8), 256, smem, st>>>(
        queries_d, Q, cb.words.as<float>(), cb.D, S, cap, w.cand_idx.as<int>(), w.cand_apx.as<float>(),
        w.cand_cnt.as<int>(), w.cand_thr.as<float>(), K, w.cand_exact.as<float>(), w.knn_part_d.as<float>(),
        w.knn_part_i.as<int>(), n_eval);
  } else {
    PCDB_CUDA(cudaFuncSetAttribute(k_rerank<PCDB_DIST_EUCLIDEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    k_rerank<PCDB_DIST_EUCLIDEAN><<<cdiv(Q, 8), 256, smem, st>>>(
        queries_d, Q, cb.words.as<float>(), cb.D, S, cap, w.cand_idx.as<int>(), w.cand_apx.as<float>(),
        w.cand_cnt.as<int>(), w.cand_thr.as<float>(), K, w.cand_exact.as<float>(), w.knn_part_d.as<float>(),
        w.knn_part_i.as<int>(), n_eval);
  }
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// ws.knn_part_* (one ascending list of K per query) -> ws.knn_idx / knn_dist / knn_cnt with the N<=k and ratio-test
// semantics of activateKNN and global row ids
int stage_knn_finish(pcdb_ctx* ctx, int64_t Q, int k, int K, bool use_ratio, float ratio_thr) {
  Workspace& w = ctx->ws;
  const Codebook_d& cb = ctx->cb;
  k_knn_merge<<<cdiv(Q, 128), 128, 0, ctx->stream>>>(w.knn_part_d.as<float>(), w.knn_part_i.as<int>(), 1, Q, K, k,
                                                     use_ratio ? 1 : 0, ratio_thr, cb.row_base + cb.word_lo,
                                                     w.knn_idx.as<int>(), w.knn_dist.as<float>(), w.knn_cnt.as<int>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

// Second half of the chi^2 sandwich (knn_gemm.cu): the pooled (query, row) pairs of the second sweep are grouped by
// query (count -> scan -> scatter), every pair gets the exact FLANN-order chi^2 (one thread per pair, a flat and
// therefore balanced launch: a query with 10^4 survivors costs what 10^4 queries with one survivor cost), and one warp
// per query picks its K best by (distance, row).
int stage_knn_chi_pool(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int K, int64_t total, int2* pool_rc,
                       const int* pool_q, const int* q_cnt, int* q_off, int* q_fill, DevBuf* csr_row, DevBuf* csr_d,
                       int dist_type, const int* skip) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  PCDB_CUDA(w.knn_part_d.ensure(sizeof(float) * (Q * K + 1)));
  PCDB_CUDA(w.knn_part_i.ensure(sizeof(int) * (Q * K + 1)));
  PCDB_CUDA(csr_row->ensure(sizeof(int) * (size_t)(total + 1)));
  PCDB_CUDA(csr_d->ensure(sizeof(float) * (size_t)(total + 1)));
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, q_cnt, q_off, Q + 1));
  PCDB_CUDA(cudaMemsetAsync(q_fill, 0, sizeof(int) * (Q + 1), st));
  if (total > 0) {
    if (dist_type == PCDB_DIST_CHISQUARED)
      k_pool_eval<PCDB_DIST_CHISQUARED><<<cdiv(total, 128), 128, 0, st>>>(queries_d, cb.words.as<float>(), cb.D, total,
                                                                         pool_rc, pool_q, skip);
    else
      k_pool_eval<PCDB_DIST_EUCLIDEAN><<<cdiv(total, 128), 128, 0, st>>>(queries_d, cb.words.as<float>(), cb.D, total,
                                                                        pool_rc, pool_q, skip);
    PCDB_LAUNCH_CHECK();
    k_pool_scatter<<<cdiv(total, 256), 256, 0, st>>>(pool_rc, pool_q, total, q_off, skip, q_fill, csr_row->as<int>(),
                                                     csr_d->as<float>());
    PCDB_LAUNCH_CHECK();
  }
  k_chi_select<<<cdiv(Q * 32, 256), 256, 0, st>>>(Q, K, q_off, csr_row->as<int>(), csr_d->as<float>(),
                                                  w.knn_part_d.as<float>(), w.knn_part_i.as<int>());
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}
