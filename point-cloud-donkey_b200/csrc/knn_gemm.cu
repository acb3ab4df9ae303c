// knn_gemm.cu — K6: codebook activation as a dense contraction on the 5th-generation tensor cores.
//
// Replaces the FLANN index search behind ActivationStrategyKNN::activateKNN
// (activation_strategy/activation_strategy_knn.h:57-92; index built at utils/flann_helper.cpp:21-70) for
// DistanceType "Euclidean":  d(q,c) = |q|^2 + |c|^2 - 2 q.c.  The q.c term is an fp16 x fp16 -> fp32 GEMM issued with
// tcgen05.mma (cta_group::2: a CTA pair on one TPC computes a 256 x 256 tile, M=256, N=256, K=16) from TMA-staged,
// 128B-swizzled shared-memory tiles with the accumulator double-buffered in TMEM.  The M x N score matrix is never
// written: eight epilogue warps per CTA read each accumulator tile back with tcgen05.ld and keep, per query row, every
// column whose approximate distance is within a RIGOROUS error margin of the running k-th best.  Those
// candidates are re-ranked with the exact FLANN fp32 arithmetic (knn_scan.cu:k_rerank), so the neighbour set equals
// the exact scan's.
//
// Tiling: each CTA of a pair owns 128 queries whose fp16 descriptors stay RESIDENT in its shared memory for the whole
// sweep (D=352 + 16 augmented columns: 6 K-blocks of 64, the last one partly used — the TMA zero-fills the columns
// beyond the row, and only 3 of its 4 MMAs are issued) while 256-codeword tiles stream through a 7-stage mbarrier ring, each
// CTA fetching HALF of every codebook tile (128 rows).  Per SM that is 256 flop per streamed byte: a single-CTA
// 128-row tile needs 64 B/clk/SM from L2 at full tensor rate, i.e. 9.5 KB/clk chip-wide against a measured L2 cap of
// about 6.3 KB/clk (the first version of this kernel sat exactly on that cap, profiles/r01_gemm_c3_1cta.txt); the pair
// needs half.  For D=1344 the query tile does not fit and both operands stream.
//
// Sweep order: the codebook is cut into slices of <= 20 MB of fp16 rows and the work units (slice, query-tile pair) are
// handed out SLICE-MAJOR, so all 74 pairs stream the same L2-resident slice at any time and the codebook is read from
// HBM about once per launch (a query-major sweep of the 0.75 GB C3 codebook let the pairs drift apart: L2 hit rate
// 53 %, 480 GB of DRAM reads per launch, and on a power-capped part those reads cost clock).  The candidate filter
// survives the slicing because the running k-th-best bound of every query is carried from slice to slice in global
// memory (atomic min; a stale value is still an upper bound), and the per-query candidate lists are shared by all
// slices (a work unit stages its candidates privately and appends them with one atomic reservation).
//
// Pair protocol (leader = even CTA of the cluster): both CTAs issue their TMA loads with .cta_group::2 so the bytes
// are counted on the LEADER's "full" barrier; the leader's elected thread issues every MMA; tcgen05.commit
// ...multicast::cluster arrives on the "empty"/"tmem_full" barriers of BOTH CTAs; the epilogue warps of both CTAs
// release an accumulator by arriving on the leader's "tmem_empty" barrier.
//
// The |c|^2 term rides in the GEMM: operand rows are AUGMENTED by one K step of 16 columns, queries with
// [.., 1, 1, 0 x 14], codewords with [-2 ch, hi, lo, 0 x 14] where hi + lo is an fp16 split of |c|^2 (relative error
// 2^-22).  The accumulator is then |c|^2 - 2 qh.ch directly and the epilogue is tcgen05.ld + a min tree: no FFMA, no
// |c|^2 staging, no barrier between epilogue warps.  (On a power-capped B200 the epilogue's instructions cost
// throughput, not just issue slots.)  The streaming-query variant (D=1344) is not augmented: a 22nd, mostly empty
// K block would cost a full 64 KB stage per tile pair on a kernel that sits on the L2->SM limit; it adds |c|^2 in the
// epilogue from a shared-memory staged tile and keeps the query-major sweep.
//
// Error margin (DESIGN.md "activation"): with qh = fp16(q), ch = fp16(c),
//   |q.c - fl(qh.ch)| <= |q-qh| |c| + |qh| |c-ch| + |q-qh| |c-ch| + D 2^-22 |qh| |ch|
// per-query norms are computed in the query-prep kernel, codebook maxima at upload time.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "stages.h"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int BN_HALF = BN / 2;         // codebook rows each CTA of a pair stages per tile
constexpr int THREADS = 384;            // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int EPI_THREADS = 256;        // two epilogue warps per TMEM lane quarter, one per 128-column half
constexpr int EPI_WARPS = EPI_THREADS / 32;
constexpr int A_BOX_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_BOX_BYTES = BN_HALF * BK * 2;  // 16 KB per CTA
constexpr int MAX_STAGES = 7;
constexpr unsigned kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA
constexpr int KB_RES_MAX = 6;             // resident K-blocks (D + 16 <= 384)
constexpr int K_AUG = 16;                 // extra K columns carrying the |c|^2 term (one UMMA_K step)
constexpr int CAND_CAP_MAX = 256;         // private staging entries per epilogue thread (64 / 128 / 256 by K)
constexpr unsigned TMEM_COLS = 512;
constexpr int STASH_SLOTS = 2;            // parked chunks per epilogue thread and tile (POOL): 8 warps x 2 x 4 KB = 4 operand boxes

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_ns() - t0 > 4000000000ull) {  // 4 s
      printf("pcdb200 knn_gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// pair load: lands in the issuing CTA's shared memory, transaction bytes are counted on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned n_clusters_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                           unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// issue only; the registers are valid after tc_ld_wait()
__device__ __forceinline__ void tc_ld_32x32b_x32(unsigned taddr, unsigned* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14),
// LBO in [16,30) (unused for swizzled K-major, set to 1), SBO = 1024 B (8 rows x 128 B) in [32,46), version 1 in
// [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)2 << 61;
  return d;
}
__device__ __forceinline__ unsigned long long desc64(unsigned lo, unsigned hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// atomic min on a float that may be negative (+inf initialised): non-negative values order like signed ints,
// negative ones inversely like unsigned ints
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ bool elect_one() {  // the same lane every time for a full warp
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=F32 (bits 4-5 = 1), A=B=F16 (0), both K-major,
// N>>3 in [17,23), M>>4 in [24,29); M is the pair's 256 rows.
constexpr unsigned kIdesc = (1u << 4) | ((unsigned)(BN >> 3) << 17) | ((unsigned)((2 * BM) >> 4) << 24);

struct GemmArgs {
  long long Q, N;
  int D;                 // augmented operand row length (descriptor dimension + K_AUG)
  int n_mpairs, n_ntiles, tiles_per_split, n_splits;  // n_mpairs: 256-query tile pairs
  const float* cnorm;    // streaming-query variant only: |c|^2, padded to a multiple of BN with +inf
  const float* margin;   // per query: 2 * (bound on |approx - exact|)
  int cand_cap;          // capacity of one shared list: stage_cap x (slices of one query tile that can be in flight at once)
  int* cand_idx;         // [Q][2][cand_cap]: one list per (query, column half), shared by all codebook slices
  float* cand_apx;
  int* cand_cnt;         // [2][Q], appended with atomics (zeroed before the launch)
  float* bound;          // [Q] running upper bound on the k-th best approximate distance (+inf before the launch)
  int2* stage;           // [CTAs][EPI_THREADS][stage_cap] private staging lists of the epilogue threads
  int stage_cap;         // entries of one private list: 64 for K <= 4, 128 for K <= 8, 256 beyond (longer staircases)
  // POOL variant (second pass of the chi^2 sandwich): `margin` holds a FIXED per-query threshold on the accumulator
  // (-inf = skip the query) and every column at or below it is appended to one global pool of (query, row) pairs
  int2* pool_rc;         // (row, unused)
  int* pool_q;
  unsigned long long* pool_count;  // entries wanted so far (may run past pool_cap: the host grows the pool and re-runs)
  long long pool_cap;
  int* q_cnt;            // [Q] pool entries per query (for the CSR the re-rank works on)
  int q_cap;             // > 0: a query whose pool entries exceed this stops appending (the host re-searches it exactly)
  int stash;             // POOL, <= 2 resident K blocks: chunks with a passing column are parked in shared memory and
                         // picked apart AFTER the accumulator has been handed back (see the epilogue)
};

// moves a thread's staged candidates into the global pool: one returning atomic per flush
__device__ __forceinline__ void pool_flush(const GemmArgs& g, long long row, const int2* stage, int cnt) {
  const int before = atomicAdd(g.q_cnt + row, cnt);
  if (g.q_cap > 0 && before > g.q_cap) return;  // hopeless query (PCA pre-filter): flagged from q_cnt afterwards
  const unsigned long long pos = atomicAdd(g.pool_count, (unsigned long long)cnt);
  for (int i = 0; i < cnt; ++i)
    if ((long long)(pos + i) < g.pool_cap) {
      g.pool_rc[pos + i] = stage[i];
      g.pool_q[pos + i] = (int)row;
    }
}

struct __align__(16) Barriers {
  float cn[2][BN];  // streaming-query variant only: |c|^2 of the tile in each accumulator buffer
  unsigned long long full[MAX_STAGES], empty[MAX_STAGES], a_full, a_empty, tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

template <bool A_RES, int KT, bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_knn_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // layout (identical in both CTAs of the pair — the MMA addresses both through one descriptor):
  //   [A resident: KB_RES_MAX boxes]? [STAGES x (B half box + A box if !A_RES)] [barriers]
  constexpr int STAGES = A_RES ? 7 : 6;
  constexpr int STAGE_BYTES = B_BOX_BYTES + (A_RES ? 0 : A_BOX_BYTES);
  unsigned char* a_res = smem;
  unsigned char* stage0 = smem + (A_RES ? KB_RES_MAX * A_BOX_BYTES : 0);
  Barriers* bars = reinterpret_cast<Barriers*>(stage0 + STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const bool leader = rank == 0;
  const int pair_id = (int)cluster_id_x(), n_pairs = (int)n_clusters_x();
  const int KB = (g.D + BK - 1) / BK;
  const int k_steps_last = (g.D - (KB - 1) * BK) / UMMA_K;
  const int n_units = g.n_mpairs * g.n_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);   // leader's producer (arrive.expect_tx); bytes come from both CTAs
      mbar_init(&bars->empty[s], 1);  // multicast tcgen05.commit
    }
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 2 * EPI_WARPS);  // one arrival per epilogue warp of both CTAs (leader's copy)
    }
    fence_barrier_init();
  }
  if (warp == 2) {  // the same warp of both CTAs: the pair allocation is collective
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers initialised and TMEM allocated in both CTAs before anyone signals across
  tc_fence_after();
  const unsigned tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane in each CTA)
    if (lane == 0) {
      int stage = 0;
      unsigned phase = 0, a_phase = 0;
      for (int unit = pair_id; unit < n_units; unit += n_pairs) {
        const int mp = unit % g.n_mpairs, split = unit / g.n_mpairs;
        const int t0 = split * g.tiles_per_split, t1 = min(g.n_ntiles, t0 + g.tiles_per_split);
        const int q_row0 = (mp * 2 + (int)rank) * BM;
        if (A_RES) {
          mbar_wait(&bars->a_empty, a_phase ^ 1);  // previous unit's MMAs are done with the resident tile
          if (leader) mbar_expect_tx(&bars->a_full, (unsigned)(2 * KB * A_BOX_BYTES));
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d_pair(a_res + kb * A_BOX_BYTES, &map_a, &bars->a_full, kb * BK, q_row0);
          a_phase ^= 1;
        }
        for (int t = t0; t < t1; ++t)
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            unsigned char* sb = stage0 + stage * STAGE_BYTES;
            if (leader) mbar_expect_tx(&bars->full[stage], (unsigned)(2 * STAGE_BYTES));
            tma_load_2d_pair(sb, &map_b, &bars->full[stage], 0, ((t * 2 + (int)rank) * KB + kb) * BN_HALF);  // box-major operand
            if (!A_RES) tma_load_2d_pair(sb + B_BOX_BYTES, &map_a, &bars->full[stage], kb * BK, q_row0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA; the warp stays converged and one
    // elected lane issues, so every loop variable is warp-uniform and the descriptors live in uniform registers: the
    // issue loop must stay well under the 128 clk an M=256 x N=256 x K=16 MMA takes, or it becomes the bottleneck)
    if (leader) {
      int stage = 0, acc = 0;
      unsigned phase = 0, acc_phase = 0, a_phase = 0;
      // descriptor halves: hi = SBO | version | SWIZZLE_128B (constant), lo = (address >> 4) | LBO; a K step of 16
      // halves (32 bytes) inside the 128-byte swizzle atom adds 2 to lo
      constexpr unsigned kDescHi = (unsigned)(1024 >> 4) | (1u << 14) | (2u << 29);
      const unsigned a_lo0 = ((smem_u32(a_res) & 0x3FFFFu) >> 4) | (1u << 16);
      const unsigned s_lo0 = ((smem_u32(stage0) & 0x3FFFFu) >> 4) | (1u << 16);
      for (int unit = pair_id; unit < n_units; unit += n_pairs) {
        const int split = unit / g.n_mpairs;
        const int t0 = split * g.tiles_per_split, t1 = min(g.n_ntiles, t0 + g.tiles_per_split);
        if (A_RES) {
          mbar_wait(&bars->a_full, a_phase);
          a_phase ^= 1;
        }
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);  // both epilogues have drained this accumulator
          tc_fence_after();
          const unsigned tmem_d = tmem_base + (unsigned)(acc * BN);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const unsigned b_lo = s_lo0 + (unsigned)(stage * (STAGE_BYTES >> 4));
            const unsigned a_lo = A_RES ? a_lo0 + (unsigned)(kb * (A_BOX_BYTES >> 4)) : b_lo + (B_BOX_BYTES >> 4);
            if (elect_one()) {
              if (kb < KB - 1 || k_steps_last == BK / UMMA_K) {
                tc_mma_f16(tmem_d, desc64(a_lo, kDescHi), desc64(b_lo, kDescHi), kIdesc, kb ? 1u : 0u);
                tc_mma_f16(tmem_d, desc64(a_lo + 2, kDescHi), desc64(b_lo + 2, kDescHi), kIdesc, 1u);
                tc_mma_f16(tmem_d, desc64(a_lo + 4, kDescHi), desc64(b_lo + 4, kDescHi), kIdesc, 1u);
                tc_mma_f16(tmem_d, desc64(a_lo + 6, kDescHi), desc64(b_lo + 6, kDescHi), kIdesc, 1u);
              } else {
                for (int k = 0; k < k_steps_last; ++k)
                  tc_mma_f16(tmem_d, desc64(a_lo + 2 * k, kDescHi), desc64(b_lo + 2 * k, kDescHi), kIdesc,
                             (kb | k) ? 1u : 0u);
              }
              tc_commit_pair(&bars->empty[stage]);  // frees the slot in both CTAs once these MMAs have read it
              if (kb == KB - 1) tc_commit_pair(&bars->tmem_full[acc]);  // accumulator complete -> both epilogues
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        if (A_RES) {
          if (elect_one()) tc_commit_pair(&bars->a_empty);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: TMEM -> registers -> running candidate filter
    const int grp = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which 128 columns of the 256-wide accumulator this warp filters
    int acc = 0;
    unsigned acc_phase = 0;
    for (int unit = pair_id; unit < n_units; unit += n_pairs) {
      const int mp = unit % g.n_mpairs, split = unit / g.n_mpairs;
      const int t0 = split * g.tiles_per_split, t1 = min(g.n_ntiles, t0 + g.tiles_per_split);
      const long long row = (long long)(mp * 2 + (int)rank) * BM + grp * 32 + lane;
      const bool active = row < g.Q;
      const float margin = active ? g.margin[row] : (POOL ? __int_as_float(0xff800000) : 0.f);
      float best[KT];
#pragma unroll
      for (int i = 0; i < KT; ++i) best[i] = __int_as_float(0x7f800000);
      // bound carried over from the slices already swept for this query (any stale value is still an upper bound)
      const float gb = (active && !POOL) ? __ldcg(g.bound + row) : __int_as_float(0x7f800000);
      float thr = POOL ? margin : gb + margin;
      // pre-filter pool: a query already past its cap is searched by the plain sweep afterwards — stop collecting for it
      bool hopeless = false;
      if (POOL && g.q_cap > 0 && active && __ldcg(g.q_cnt + row) > g.q_cap) thr = __int_as_float(0xff800000);
      // Candidates of this unit are staged in a list private to this thread (plain stores, local counter) and moved
      // to the query's shared list at the end of the unit with ONE atomic reservation: a returning global atomic per
      // append put a ~1 us round trip into the filter loop (C4: 1724 -> 1412 TFLOP/s).
      int cnt = 0;
      int2* stage = g.stage + ((size_t)blockIdx.x * EPI_THREADS + (size_t)((warp - 4) * 32 + lane)) * g.stage_cap;
      const int stage_cap = g.stage_cap;
      // Streaming-query variant (D = 1344): the operands are not augmented (a 17th partial K block would cost a full
      // 64 KB stage per tile pair on a kernel that already sits on the L2->SM limit), so |c|^2 is added here from a
      // shared-memory staged tile, one column per epilogue thread, software-prefetched.
      const int epi_tid = (warp - 4) * 32 + lane;
      // POOL with a short K (the projected sweep: 8 MMAs per tile, 1024 clk): the accumulator has to go back to the MMA
      // warp within one tile time, but picking a passing chunk apart (32 compares, mask, appends) costs ONE lane ~350 clk
      // while its 31 siblings wait, and with ~80 pooled rows per query some warp of the pair meets such a chunk in nearly
      // every tile — the hand-over then waits for the slowest of 16 warps.  So a lane that sees a passing chunk only
      // PARKS its 32 values (8 x 16-byte stores into the unused resident-operand boxes 2..5, lane-interleaved: no bank
      // conflicts whichever lanes hit) and takes them apart after its warp has arrived on tmem_empty (pooled sweep at
      // 115 rows per query: 67.7 -> 60.9 ms; with a threshold nothing passes the kernel takes 51 ms = 1300 clk per tile).
      // Tried against that floor and rejected (profiles/r02_pool_variants_c3_*_rejected.json): sixteen epilogue warps
      // with both tcgen05.ld of a warp in flight (floor unchanged: the epilogue's length is not what the hand-over waits
      // for), and four N = 128 accumulators with three epilogue warp sets (2250 clk per tile).
      const bool stash_on = POOL && A_RES && (g.stash & 1) != 0 && KB <= 2;
      float4* const stash = reinterpret_cast<float4*>(a_res + 2 * A_BOX_BYTES) + (warp - 4) * (STASH_SLOTS * 8 * 32) + lane;
      int n_stash = 0, stash_nb[STASH_SLOTS];
#pragma unroll
      for (int i = 0; i < STASH_SLOTS; ++i) stash_nb[i] = 0;
      // one passing chunk -> private list: the passing columns as a bit mask (32 independent compares), then one append
      // per set bit (a column-by-column scan is ~300 dependent instructions of ONE warp while its 7 siblings and the MMA
      // pipe wait for the accumulator: measured 0.5 ms per pooled row per query at Q = 88 k)
      // passing columns of one chunk (bit e = column n_base + e) -> private list, one append per set bit
      auto pool_append = [&](unsigned pm, int n_base) {
        if (!active) pm = 0;
        if (n_base + 32 > g.N) pm &= n_base < g.N ? (0xffffffffu >> (32 - (int)(g.N - n_base))) : 0u;
        while (pm) {
          const int e = __ffs(pm) - 1;
          pm &= pm - 1;
          if (cnt < stage_cap) {
            stage[cnt] = make_int2(n_base + e, 0);
            ++cnt;
          } else if (g.q_cap > 0) {  // a full private list inside ONE unit: the query is hopeless
            hopeless = true;
            thr = __int_as_float(0xff800000);
            pm = 0;
          } else {
            pool_flush(g, row, stage, cnt);
            stage[0] = make_int2(n_base + e, 0);
            cnt = 1;
          }
        }
      };
      // in place (no parking slot left, or long rows): the passing columns as a bit mask (32 independent compares — a
      // column-by-column scan is ~300 dependent instructions of ONE warp while its 7 siblings and the MMA pipe wait for
      // the accumulator: measured 0.5 ms per pooled row per query at Q = 88 k)
      auto pool_take = [&](const unsigned (&v)[32], int n_base) {
        unsigned pm = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) pm |= (__uint_as_float(v[e]) <= thr ? 1u : 0u) << e;
        pool_append(pm, n_base);
      };
      float cn_next = 0.f;
      if (!A_RES) cn_next = __ldg(g.cnorm + (size_t)t0 * BN + epi_tid);
      for (int t = t0; t < t1; ++t) {
        if (!A_RES) {
          // publish this tile's |c|^2; the barrier also orders it after every thread's reads of tile t-2
          bars->cn[acc][epi_tid] = cn_next;
          if (t + 1 < t1) cn_next = __ldg(g.cnorm + (size_t)(t + 1) * BN + epi_tid);
          epi_bar_sync();
        }
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const int col0 = half * (BN / 2);
        const unsigned taddr = tmem_base + ((unsigned)(grp * 32) << 16) + (unsigned)(acc * BN + col0);
        const float* cn_s = bars->cn[acc] + col0;
        unsigned ra[32], rb[32];
        tc_ld_32x32b_x32(taddr, ra);
        // Branch-free common path: 32 distances, one min-tree, ONE compare per 32 columns; the per-column scan only runs
        // when some column of the chunk can still be among the k best (rare after the first tiles of a sweep).
#define PCDB_FILTER_CHUNK(REG, C)                                                                  \
  {                                                                                                \
    if (!A_RES) {                                                                                  \
      _Pragma("unroll") for (int j4 = 0; j4 < 8; ++j4) {                                           \
        const float4 cn = *reinterpret_cast<const float4*>(cn_s + (C) * 32 + j4 * 4);              \
        REG[j4 * 4 + 0] = __float_as_uint(fmaf(-2.f, __uint_as_float(REG[j4 * 4 + 0]), cn.x));     \
        REG[j4 * 4 + 1] = __float_as_uint(fmaf(-2.f, __uint_as_float(REG[j4 * 4 + 1]), cn.y));     \
        REG[j4 * 4 + 2] = __float_as_uint(fmaf(-2.f, __uint_as_float(REG[j4 * 4 + 2]), cn.z));     \
        REG[j4 * 4 + 3] = __float_as_uint(fmaf(-2.f, __uint_as_float(REG[j4 * 4 + 3]), cn.w));     \
      }                                                                                            \
    }                                                                                              \
    float m16[16];                                                                                 \
    _Pragma("unroll") for (int i = 0; i < 16; ++i)                                                 \
      m16[i] = fminf(__uint_as_float(REG[i]), __uint_as_float(REG[i + 16]));                       \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) m16[i] = fminf(m16[i], m16[i + 8]);              \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) m16[i] = fminf(m16[i], m16[i + 4]);              \
    const float mn = fminf(fminf(m16[0], m16[1]), fminf(m16[2], m16[3]));                          \
    if (mn <= thr) {                                                                               \
      const int n_base = t * BN + col0 + (C) * 32;                                                 \
      if (POOL) {                                                                                  \
        if (stash_on && n_stash < STASH_SLOTS) {                                                   \
          float4* sp = stash + n_stash * (8 * 32);                                                 \
          _Pragma("unroll") for (int j4 = 0; j4 < 8; ++j4)                                         \
            sp[j4 * 32] = make_float4(__uint_as_float(REG[j4 * 4 + 0]), __uint_as_float(REG[j4 * 4 + 1]), \
                                      __uint_as_float(REG[j4 * 4 + 2]), __uint_as_float(REG[j4 * 4 + 3])); \
          _Pragma("unroll") for (int i = 0; i < STASH_SLOTS; ++i)                                  \
            if (i == n_stash) stash_nb[i] = n_base;                                                \
          ++n_stash;                                                                               \
        } else {                                                                                   \
          pool_take(REG, n_base);                                                                  \
        }                                                                                          \
      } else {                                                                                     \
        _Pragma("unroll") for (int e = 0; e < 32; ++e) {                                           \
          const float d = __uint_as_float(REG[e]);  /* |c|^2 - 2 q.c */                             \
          if (d <= thr && active && n_base + e < g.N) {                                            \
            if (cnt < stage_cap) stage[cnt] = make_int2(n_base + e, __float_as_int(d));            \
            ++cnt;                                                                                 \
            float x = d;                                                                           \
            _Pragma("unroll") for (int i = 0; i < KT; ++i) {                                       \
              float lo = fminf(best[i], x);                                                        \
              x = fmaxf(best[i], x);                                                               \
              best[i] = lo;                                                                        \
            }                                                                                      \
            thr = fminf(best[KT - 1], gb) + margin;                                                \
          }                                                                                        \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
  }
        auto hand_back = [&]() {  // the accumulator returns to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[acc]);
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        };
        if (POOL && A_RES && (g.stash & 2)) {
          // short rows: the accumulator is handed back as soon as its last chunk sits in registers — the hand-over, not
          // the length of the epilogue, paces this sweep (see the parking note above) — and that chunk is filtered after
          tc_ld_wait();
          tc_ld_32x32b_x32(taddr + 32, rb);
          PCDB_FILTER_CHUNK(ra, 0)
          tc_ld_wait();
          tc_ld_32x32b_x32(taddr + 64, ra);
          PCDB_FILTER_CHUNK(rb, 1)
          tc_ld_wait();
          tc_ld_32x32b_x32(taddr + 96, rb);
          PCDB_FILTER_CHUNK(ra, 2)
          tc_ld_wait();
          hand_back();
          PCDB_FILTER_CHUNK(rb, 3)
        } else if (POOL && (g.stash & 4)) {  // experiment: no read-back at all (what the hand-over alone costs)
          tc_ld_wait();
          hand_back();
        } else {
#pragma unroll 1
          for (int c = 0; c < BN / 64; c += 2) {
            tc_ld_wait();                                  // chunk c is in ra
            tc_ld_32x32b_x32(taddr + (c + 1) * 32, rb);    // chunk c+1 in flight while c is filtered
            PCDB_FILTER_CHUNK(ra, c)
            tc_ld_wait();
            if (c + 2 < BN / 64) tc_ld_32x32b_x32(taddr + (c + 2) * 32, ra);
            PCDB_FILTER_CHUNK(rb, c + 1)
          }
          hand_back();
        }
#undef PCDB_FILTER_CHUNK
        if (POOL && n_stash > 0) {
          // parked chunks of this tile, off the accumulator's critical path.  (Taking them apart with the whole warp —
          // lane e compares column e, one ballot is the mask — measured SLOWER than the owner lane doing it alone, 68.3
          // against 59.9 ms per sweep: the owner's work runs in the shadow of its siblings' wait for the next
          // accumulator, a whole-warp drain serialises in front of that wait.)
#pragma unroll
          for (int i = 0; i < STASH_SLOTS; ++i)
            if (i < n_stash) {
              unsigned v[32];
              const float4* sp = stash + i * (8 * 32);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 x = sp[j4 * 32];
                v[j4 * 4 + 0] = __float_as_uint(x.x);
                v[j4 * 4 + 1] = __float_as_uint(x.y);
                v[j4 * 4 + 2] = __float_as_uint(x.z);
                v[j4 * 4 + 3] = __float_as_uint(x.w);
              }
              pool_take(v, stash_nb[i]);
            }
          n_stash = 0;
        }
      }
      if (POOL) {
        // end of the unit: ONE reservation in the global pool per warp (a returning atomic per thread and unit on the
        // single pool counter serialised the whole launch: 12 M same-address atomics per C3 step)
        int take = active ? cnt : 0;
        if (take > 0 || hopeless) {
          const int before = atomicAdd(g.q_cnt + row, hopeless ? g.q_cap + 1 + take : take);
          if (hopeless || (g.q_cap > 0 && before > g.q_cap)) take = 0;  // flagged from q_cnt afterwards
        }
        int incl = take;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot > 0) {  // warp-uniform
          unsigned long long base = 0;
          if (lane == 31) base = atomicAdd(g.pool_count, (unsigned long long)tot);
          base = __shfl_sync(0xffffffffu, base, 31);
          const unsigned long long pos = base + (unsigned long long)(incl - take);
          // the 32 private lists are copied one after the other by the whole warp: coalesced, independent loads (a
          // per-lane copy loop was a chain of dependent L2 round trips that held the accumulator back)
          for (int src = 0; src < 32; ++src) {
            const int n_src = __shfl_sync(0xffffffffu, take, src);
            if (n_src == 0) continue;
            const unsigned long long p_src = __shfl_sync(0xffffffffu, pos, src);
            const long long r_src = __shfl_sync(0xffffffffu, row, src);
            const int2* st_src = stage + ((long long)src - lane) * stage_cap;  // lane `src`'s private list
            for (int i = lane; i < n_src; i += 32)
              if ((long long)(p_src + i) < g.pool_cap) {
                g.pool_rc[p_src + i] = __ldcg(st_src + i);
                g.pool_q[p_src + i] = (int)r_src;
              }
          }
        }
      } else if (active && cnt > 0) {
        if (best[KT - 1] < gb) atomic_min_float(g.bound + row, best[KT - 1]);
        // a staging list that overflowed, or a shared list above its capacity, marks the query for the exact-scan
        // fallback (the count is pushed past any capacity)
        const int base = atomicAdd(g.cand_cnt + (size_t)half * g.Q + row, cnt > stage_cap ? (1 << 24) : cnt);
        const size_t cbase = ((size_t)row * 2 + half) * g.cand_cap;
        for (int i = 0; i < min(cnt, stage_cap) && base + i < g.cand_cap; ++i) {
          const int2 c = stage[i];
          g.cand_idx[cbase + base + i] = c.x;
          g.cand_apx[cbase + base + i] = __int_as_float(c.y);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // every MMA has completed and been read back in both CTAs; no remote arrival is still in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- codebook / query preparation ------------------------------------------------------------------------------
// One warp per row: the fp16 operand row of the augmented GEMM (pitch D + K_AUG), |row|^2 (double accumulate),
// |fp16(row)|, |row - fp16(row)|.  CODEBOOK rows are stored as [-2 fp16(c), hi, lo, 0 x 14] with hi + lo = |c|^2 split
// into two halves; query rows as [fp16(q), 1, 1, 0 x 14]: the accumulator of the GEMM is |c|^2 - 2 qh.ch.
// SQRT: the row is sqrt-transformed first (chi^2 sandwich, see stage_knn_gemm): everything above then refers to the
// fp32 vector s = fl(sqrt(x)); rows with a negative or non-finite entry are flagged (the sandwich needs x >= 0).
template <bool CODEBOOK, bool SQRT>
__global__ void k_prep_rows(const float* __restrict__ x, long long n, int D, int aug, __half* xh, float* norm2,
                            float* norm, float* err, int* row_bad, int* any_bad) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int Dh = D + (aug ? K_AUG : 0);  // aug == 0: plain fp16 copy (streaming-query variant)
  // Codebook operand: stored box by box — [128-row block][K block][128 rows][64 columns] — so that every TMA box of the
  // sweep is ONE contiguous 16 KB run of HBM instead of 128 segments of 128 B at a row stride (a lone cloud's activation
  // streams the codebook straight from HBM: 2.9 TB/s row-major).  Queries stay row-major (loaded once per work unit).
  const int KBt = (Dh + BK - 1) / BK;
  auto at = [&](int j) -> size_t {
    if (CODEBOOK) return ((((size_t)(r >> 7) * KBt + (j >> 6)) << 7 | (size_t)(r & 127)) << 6) | (size_t)(j & 63);
    return (size_t)r * Dh + j;
  };
  double s2 = 0, e2 = 0, h2 = 0;
  bool bad = false;
  for (int j = lane; j < D; j += 32) {
    float v = x[r * D + j];
    if (SQRT) {
      bad |= !(v >= 0.f) || !(v < __int_as_float(0x7f800000));
      v = __fsqrt_rn(fmaxf(v, 0.f));
    }
    __half h = __float2half_rn(v);
    float hv = __half2float(h);
    xh[at(j)] = (CODEBOOK && aug) ? __float2half_rn(-2.0f * hv) : h;  // exact: a power-of-two multiple
    s2 += (double)v * v;
    h2 += (double)hv * hv;
    double dd = (double)v - (double)hv;
    e2 += dd * dd;
  }
  s2 = warp_sum(s2);
  e2 = warp_sum(e2);
  h2 = warp_sum(h2);
  if (SQRT) {
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      if (row_bad) row_bad[r] = bad ? 1 : 0;
      if (bad && any_bad) atomicOr(any_bad, 1);
    }
  }
  if (aug && lane < K_AUG) {
    __half a = __float2half_rn(0.f);
    if (CODEBOOK) {
      const float cn = (float)s2;
      const __half hi = __float2half_rn(cn);
      if (lane == 0) a = hi;
      if (lane == 1) a = __float2half_rn(cn - __half2float(hi));
    } else if (lane < 2) {
      a = __float2half_rn(1.0f);
    }
    xh[at(D + lane)] = a;
  }
  if (lane == 0) {
    if (norm2) norm2[r] = (float)s2;
    norm[r] = (float)sqrt(h2) * 1.0000002f;   // |fp16(row)|, rounded up
    err[r] = (float)sqrt(e2) * 1.0000002f;    // |row - fp16(row)|, rounded up
  }
}

__global__ void k_pad_inf(float* a, long long from, long long to) {
  long long i = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < to) a[i] = __int_as_float(0x7f800000);
}

__global__ void k_max2(const float* __restrict__ a, const float* __restrict__ b, long long n, unsigned* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float va = i < n ? a[i] : 0.f, vb = i < n ? b[i] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, o));
    vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], __float_as_uint(va));  // non-negative floats order like their bit patterns
    atomicMax(&out[1], __float_as_uint(vb));
  }
}

// margin[q] = 2 * eps_d(q); eps_d bounds |accumulator - (|c|^2 - 2 q.c)|.  eps_out (optional) receives eps_d itself,
// in the chi^2 sandwich enlarged by the rounding of the fp32 square roots (the bound then holds against the
// real-arithmetic Hellinger form sum (sqrt(q_i) - sqrt(c_i))^2 - sum q_i).
__global__ void k_margin(const float* __restrict__ qnorm_h, const float* __restrict__ qerr, long long Q, int D, int aug,
                         float cmax_h, float cerr_max, float cmax2, int sqrt_rows, float* margin, float* eps_out) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  // |q.c - fl(qh.ch)| <= |q-qh||c| + |qh||c-ch| + accumulation error; |c| <= |ch| + |c-ch|
  const double p22 = 2.384185791015625e-07 /* 2^-22 */;
  double cmax = (double)cmax_h + (double)cerr_max;
  double eps_dot = (double)qerr[q] * cmax + (double)qnorm_h[q] * (double)cerr_max;  // operand rounding to fp16
  // fp32 accumulation of D + 2 products whose magnitudes sum to at most 2 |qh||ch| + |c|^2; the fp16 hi/lo split of
  // |c|^2 (relative 2^-22, subnormal floor 6e-8) and its own fp32 rounding
  double eps_acc = (double)(D + 2) * p22 * (2.0 * (double)qnorm_h[q] * (double)cmax_h + (double)cmax2);
  double eps_cn = p22 * (double)cmax2 + 1.2e-7;
  double eps_d = 2.0 * eps_dot + eps_acc + eps_cn;
  if (!aug)  // epilogue-side |c|^2: fp32 dot accumulation + one fmaf with a correctly rounded |c|^2
    eps_d = 2.0 * (eps_dot + (double)D * p22 * (double)qnorm_h[q] * (double)cmax_h) +
            2.0 * 5.9604644775390625e-08 /* 2^-24 */ * ((double)cmax2 + 2.0);
  if (sqrt_rows) {
    // s = fl(sqrt(x)) is within 2^-24 relative of the real root: | |s - t|^2 - H^2 | <= 2 H eta + eta^2 with
    // eta = 2^-24 (|s| + |t|) and H <= |s| + |t|
    double st = (double)qnorm_h[q] + (double)qerr[q] + cmax;
    eps_d += 1.25 * p22 * 0.5 * st * st;
  }
  margin[q] = (float)(2.0 * eps_d * 1.0001);
  if (eps_out) eps_out[q] = (float)(eps_d * 1.0001);
}

__global__ void k_fill_f32(float* a, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
// final pruning threshold of the re-rank: carried bound + margin, written for both column halves
__global__ void k_final_thr(const float* __restrict__ bound, const float* __restrict__ margin, long long Q, float* thr) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float t = bound[q] + margin[q];
  thr[q] = t;
  thr[Q + q] = t;
}

// queries whose candidate list overflowed in some split, or (chi^2) that hold a negative entry -> exact-scan fallback
__global__ void k_overflow_flags(const int* __restrict__ cand_cnt, int S, long long Q, int cap,
                                 const int* __restrict__ row_bad, int* flag) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q > Q) return;
  int f = 0;
  if (q < Q) {
    for (int s = 0; s < S; ++s) f |= cand_cnt[(size_t)s * Q + q] > cap;
    if (row_bad) f |= row_bad[q];
  }
  flag[q] = f;
}
__global__ void k_overflow_gather(const int* __restrict__ flag, const int* __restrict__ pos, long long Q, int D,
                                  const float* __restrict__ queries, int* list, float* out) {
  const long long q = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q || !flag[q]) return;
  const int o = pos[q];
  if (lane == 0) list[o] = (int)q;
  for (int j = lane; j < D; j += 32) out[(size_t)o * D + j] = queries[q * D + j];
}
__global__ void k_overflow_scatter(const int* __restrict__ list, int n, int k, const int* __restrict__ idx,
                                   const float* __restrict__ dist, const int* __restrict__ cnt, int* idx_out,
                                   float* dist_out, int* cnt_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int q = list[i];
  for (int j = 0; j < k; ++j) {
    idx_out[(size_t)q * k + j] = idx[(size_t)i * k + j];
    dist_out[(size_t)q * k + j] = dist[(size_t)i * k + j];
  }
  cnt_out[q] = cnt[i];
}

// Chi^2 sandwich, threshold of the second pass.  For non-negative rows  H^2 <= chi^2 <= 2 H^2  with
// H^2 = sum (sqrt(a_i) - sqrt(b_i))^2  (term by term: (a-b)^2/(a+b) = (sqrt a - sqrt b)^2 (1 + 2 sqrt(ab)/(a+b))).
// U = the K-th smallest FLANN-order chi^2 over ANY K codewords bounds the K-th smallest overall; a row of the true top
// K therefore has  H^2 <= chi^2_real <= chi^2_fl / (1 - g) <= U (1 + 1.01 g)  with g = (D + 8) 2^-24 bounding the
// rounding of the functor's D sequential non-negative fp32 terms, and its accumulator value (H^2 - sum q_i, known to
// within eps) is at most  U (1 + 1.01 g) - qn + eps.
__global__ void k_chi_thr(const float* __restrict__ part_d, int K, const float* __restrict__ qn2,
                          const float* __restrict__ eps, const int* __restrict__ skip, long long Q, int D, float* thr) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (skip[q]) {
    thr[q] = __int_as_float(0xff800000);
    return;
  }
  const double g = 1.01 * (double)(D + 8) * 5.9604644775390625e-08;
  const double t = (double)part_d[q * K + K - 1] * (1.0 + g) - (double)qn2[q] * (1.0 - 2.4e-7) + (double)eps[q] + 1e-12;
  float f = (float)t;
  if ((double)f < t) f = __uint_as_float(__float_as_uint(f) + (f >= 0.f ? 1u : -1u));  // round up
  thr[q] = f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// blocked: the box-major codebook operand (k_prep_rows), a [rows][BK] array of which every box is one contiguous run
int make_map(pcdb_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int D, int box_rows, bool blocked = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return ctx->fail(PCDB_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (blocked) D = BK;
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(PCDB_E_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return PCDB_OK;
}

// codebook-side operand of one distance family: [0] the rows themselves (squared L2), [1] their square roots (chi^2)
struct GemmOperand {
  CUtensorMap map_b;
  DevBuf words_h, cnorm, cnorm_h, cerr;
  float cmax_h = 0, cerr_max = 0, cmax2 = 0;
  int64_t n = 0;   // rows
  int D = 0;       // row length before augmentation
  bool ready = false;
  void release() {
    DevBuf* all[] = {&words_h, &cnorm, &cnorm_h, &cerr};
    for (DevBuf* b : all) b->release();
    ready = false;
    n = 0;
  }
};

// PCA pre-filter of the Euclidean activation (large codebooks, large batches):
//   * `sample`: every f-th codeword (one random row of each block of f), full-length rows.  A sweep over it gives U(q),
//     the exact K-th nearest distance INSIDE the sample, an upper bound of the K-th nearest distance overall.
//   * `proj`: every codeword projected on the d leading principal axes P of the codebook, y = P^T (c - mu).  With P
//     orthonormal |y_q - y_c| <= |q - c|, so ONE sweep over the d-dimensional operands (d + 16 instead of D + 16 columns
//     of MMA work) pools every codeword whose projected distance can still be <= U; the exact functor decides.
// Sound whatever the basis is (orthonormality is checked numerically and its defect enters the threshold); the basis
// only decides how many rows survive.  Measured on C3 (profiles/r02_prefilter_experiment.json): d = 128, f = 16 keeps a
// median of 30 codewords per query out of 1.07 M; 1 % of the queries keep more than 2048 and are searched the old way.
struct PcaFilter {
  bool ready = false;
  int d = 0, f = 0;
  GemmOperand sample, proj;
  DevBuf s2g;      // sample row -> codebook row
  DevBuf basis;    // [D][d] fp32, orthonormal columns
  DevBuf mean;     // [D]
  double orth_defect = 0;  // bound on ||P^T P - I||_2
  float eta_c = 0;         // bound on |fl(y_c) - y_c| over the codebook (fp32 projection rounding)
  float mean_norm = 0;
  // Back-off: a batch whose queries mostly run past their pool cap (scene clutter far from every codeword: C5 sent 73 %
  // of its queries to the plain sweep after paying for both pre-filter sweeps) switches the filter off for the next
  // `backoff` eligible batches, doubling up to 256 while retries keep failing; results are exact either way.
  int backoff = 0, backoff_left = 0;
  // per batch
  DevBuf yq, yq_h, rnorm_q, qn2p, qnormp, qerrp, epsp, marginp;
  void release() {
    sample.release();
    proj.release();
    DevBuf* all[] = {&s2g, &basis, &mean, &yq, &yq_h, &rnorm_q, &qn2p, &qnormp, &qerrp, &epsp, &marginp};
    for (DevBuf* b : all) b->release();
    ready = false;
    backoff = backoff_left = 0;
  }
};

struct FallbackBufs {
  DevBuf q, idx, dist, cnt, list;
  void release() {
    DevBuf* all[] = {&q, &idx, &dist, &cnt, &list};
    for (DevBuf* b : all) b->release();
  }
};

struct GemmState {
  GemmOperand op[2];
  PcaFilter pca;
  FallbackBufs fb[2];  // [0]: queries the pre-filtered search hands to the plain sweep, [1]: that sweep's own fallback
  DevBuf qnorm_h, qerr, qn2, qbad, margin, eps, bound, stage, fb_flag, fb_pos;
  // chi^2 second pass: pooled candidates and their CSR by query
  DevBuf thr2, pool_rc, pool_q, q_cnt, q_off, q_fill, csr_row, csr_q, csr_d;
  int64_t pool_cap = 0;
  int max_clusters[2] = {0, 0};  // co-resident CTA pairs of the streaming / resident-query kernel on this device
  ~GemmState() {
    DevBuf* all[] = {&qnorm_h, &qerr, &qn2, &qbad, &margin, &eps, &bound, &stage, &fb_flag, &fb_pos, &thr2, &pool_rc,
                     &pool_q, &q_cnt, &q_off, &q_fill, &csr_row, &csr_q, &csr_d};
    for (DevBuf* b : all) b->release();
    op[0].release();
    op[1].release();
    pca.release();
    fb[0].release();
    fb[1].release();
  }
};
// owned by the context (freed by pcdb_destroy through gemm_state_free)
GemmState* state_of(pcdb_ctx* ctx) {
  if (!ctx->gemm_state) {
    ctx->gemm_state = new GemmState();
    ctx->gemm_state_free = [](void* p) { delete static_cast<GemmState*>(p); };
  }
  return static_cast<GemmState*>(ctx->gemm_state);
}

template <bool A_RES, int KT, bool POOL>
int launch_gemm(pcdb_ctx* ctx, const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmArgs& g, int grid) {
  const size_t smem = (A_RES ? KB_RES_MAX * A_BOX_BYTES : 0) +
                      (size_t)(A_RES ? 7 : 6) * (B_BOX_BYTES + (A_RES ? 0 : A_BOX_BYTES)) + sizeof(Barriers) + 64;
  // co-resident CTA pairs: GPCs with an odd number of usable SMs leave one SM without a partner, and a pair that
  // cannot be resident from the start would run after the others and double the sweep time
  PCDB_CUDA(cudaFuncSetAttribute(k_knn_gemm<A_RES, KT, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int& max_clusters = state_of(ctx)->max_clusters[A_RES ? 1 : 0];  // same footprint for every KT / POOL
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (ctx->sm_count / 2));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    PCDB_CUDA(cudaOccupancyMaxActiveClusters(&n, k_knn_gemm<A_RES, KT, POOL>, &cfg));
    if (n < 1) return ctx->fail(PCDB_E_CUDA, "no CTA pair of k_knn_gemm can be resident on this device");
    max_clusters = n;
  }
  grid = 2 * std::min(grid / 2, max_clusters);
  k_knn_gemm<A_RES, KT, POOL><<<grid, THREADS, smem, ctx->stream>>>(map_a, map_b, g);
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}

template <bool A_RES>
int launch_gemm_kt(pcdb_ctx* ctx, int K, const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmArgs& g,
                   int grid) {
  if (K <= 1) return launch_gemm<A_RES, 1, false>(ctx, map_a, map_b, g, grid);
  if (K <= 2) return launch_gemm<A_RES, 2, false>(ctx, map_a, map_b, g, grid);
  if (K <= 4) return launch_gemm<A_RES, 4, false>(ctx, map_a, map_b, g, grid);
  if (K <= 8) return launch_gemm<A_RES, 8, false>(ctx, map_a, map_b, g, grid);
  return launch_gemm<A_RES, 17, false>(ctx, map_a, map_b, g, grid);
}

// fp16 operand copy of n fp32 rows of length D (the rows, or their square roots for chi^2): box-major rows, |row|^2,
// error maxima and the tensor map.  op.ready stays false when the rows cannot be used (non-finite / negative entries).
int prepare_operand(pcdb_ctx* ctx, const float* rows_d, int64_t n, int D, bool sqrt_rows, GemmOperand& op) {
  cudaStream_t st = ctx->stream;
  op.ready = false;
  op.n = n;
  op.D = D;
  const int aug = (D + K_AUG + BK - 1) / BK <= KB_RES_MAX ? 1 : 0;  // resident-query variant <=> augmented operands
  const int Dh = D + (aug ? K_AUG : 0);
  const int64_t n_pad = (int64_t)cdiv(n, BN) * BN;
  const int KBt = (Dh + BK - 1) / BK;
  const int64_t blk_rows = (int64_t)cdiv(n, BN_HALF) * KBt * BN_HALF;  // rows of the box-major [rows][BK] operand
  PCDB_CUDA(op.words_h.ensure(sizeof(__half) * (size_t)blk_rows * BK + 256));
  PCDB_CUDA(cudaMemsetAsync(op.words_h.p, 0, sizeof(__half) * (size_t)blk_rows * BK, st));  // padding rows / columns
  PCDB_CUDA(op.cnorm.ensure(sizeof(float) * (n_pad + 4)));
  PCDB_CUDA(op.cnorm_h.ensure(sizeof(float) * (n + 1)));
  PCDB_CUDA(op.cerr.ensure(sizeof(float) * (n + 1)));
  PCDB_CUDA(ctx->ws.scalars.ensure(256));
  unsigned* mx = reinterpret_cast<unsigned*>(ctx->ws.scalars.as<char>() + 128);
  PCDB_CUDA(cudaMemsetAsync(mx, 0, 32, st));
  if (sqrt_rows)
    k_prep_rows<true, true><<<cdiv(n * 32, 256), 256, 0, st>>>(rows_d, n, D, aug, op.words_h.as<__half>(),
                                                              op.cnorm.as<float>(), op.cnorm_h.as<float>(),
                                                              op.cerr.as<float>(), nullptr, reinterpret_cast<int*>(mx + 4));
  else
    k_prep_rows<true, false><<<cdiv(n * 32, 256), 256, 0, st>>>(rows_d, n, D, aug, op.words_h.as<__half>(),
                                                               op.cnorm.as<float>(), op.cnorm_h.as<float>(),
                                                               op.cerr.as<float>(), nullptr, nullptr);
  PCDB_LAUNCH_CHECK();
  if (n_pad > n) {
    k_pad_inf<<<cdiv(n_pad - n, 256), 256, 0, st>>>(op.cnorm.as<float>(), n, n_pad);
    PCDB_LAUNCH_CHECK();
  }
  k_max2<<<cdiv(n, 256), 256, 0, st>>>(op.cnorm_h.as<float>(), op.cerr.as<float>(), n, mx);
  PCDB_LAUNCH_CHECK();
  k_max2<<<cdiv(n, 256), 256, 0, st>>>(op.cnorm.as<float>(), op.cnorm.as<float>(), n, mx + 2);
  PCDB_LAUNCH_CHECK();
  float h[8];
  PCDB_CUDA(cudaMemcpyAsync(h, mx, 32, cudaMemcpyDeviceToHost, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  op.cmax_h = h[0];
  op.cerr_max = h[1];
  op.cmax2 = h[2];
  int any_bad = 0;
  std::memcpy(&any_bad, &h[4], sizeof(int));
  if (!std::isfinite(h[0]) || !std::isfinite(h[2]) || any_bad) {
    op.release();  // non-finite rows, or negative entries under chi^2: scan path only
    return PCDB_OK;
  }
  PCDB_TRY(make_map(ctx, &op.map_b, op.words_h.p, blk_rows, Dh, BN_HALF, true));
  op.ready = true;
  return PCDB_OK;
}

// ---- PCA pre-filter set-up (one-off, at codebook upload) ---------------------------------------------------------
// column sums of n rows taken with a stride (fp64 atomics on D accumulators)
__global__ void k_col_sums(const float* __restrict__ x, long long n, long long stride, int D, double* sums) {
  const int j = blockIdx.y * blockDim.x + threadIdx.x;  // grid.y covers the columns in blocks of blockDim.x
  double acc = 0;
  for (long long r = blockIdx.x; r < n; r += gridDim.x)
    if (j < D) acc += (double)x[r * stride * D + j];
  if (j < D) atomicAdd(&sums[j], acc);
}
// covariance tile: C[16 a .. 16 a + 15][16 b .. 16 b + 15] over n strided rows, fp64; one CTA of 256 threads per tile
__global__ void __launch_bounds__(256) k_cov_tile(const float* __restrict__ x, long long n, long long stride, int D,
                                                  const double* __restrict__ mean, double* cov) {
  __shared__ float sa[64][17], sb[64][17];
  const int ta = blockIdx.x * 16, tb = blockIdx.y * 16;
  const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
  double acc = 0;
  for (long long r0 = 0; r0 < n; r0 += 64) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int rr = e >> 4, cc = e & 15;
      const long long r = r0 + rr;
      sa[rr][cc] = r < n ? (float)((double)x[r * stride * D + ta + cc] - mean[ta + cc]) : 0.f;
      sb[rr][cc] = r < n ? (float)((double)x[r * stride * D + tb + cc] - mean[tb + cc]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) acc += (double)sa[rr][i] * (double)sb[rr][j];
    __syncthreads();
  }
  cov[(size_t)(ta + i) * D + tb + j] = acc / (double)n;
}
// y = P^T (x - mu): rows x [n][D] -> y [n][d]; CTA = 64 rows x d columns, K staged 16 at a time; DC = ceil(d / 32)
template <int DC>
__global__ void __launch_bounds__(256) k_project(const float* __restrict__ x, long long n, int D, int d,
                                                 const float* __restrict__ mean, const float* __restrict__ basis,
                                                 float* __restrict__ y) {
  __shared__ float sx[16][65];
  __shared__ float sp[16][DC * 32];
  const long long r0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty: 8 row groups of 8, tx: columns tx + 32 c
  float acc[8][DC];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < DC; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int rr = e >> 4, kk = e & 15;
      const long long r = r0 + rr;
      sx[kk][rr] = (r < n && k0 + kk < D) ? __fsub_rn(x[r * D + k0 + kk], mean[k0 + kk]) : 0.f;
    }
    for (int e = threadIdx.x; e < 16 * DC * 32; e += 256) {
      const int kk = e / (DC * 32), c = e % (DC * 32);
      sp[kk][c] = (k0 + kk < D && c < d) ? basis[(size_t)(k0 + kk) * d + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float pv[DC];
#pragma unroll
      for (int c = 0; c < DC; ++c) pv[c] = sp[kk][tx + 32 * c];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float xv = sx[kk][ty * 8 + a];
#pragma unroll
        for (int c = 0; c < DC; ++c) acc[a][c] = fmaf(xv, pv[c], acc[a][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const long long r = r0 + ty * 8 + a;
    if (r < n)
#pragma unroll
      for (int c = 0; c < DC; ++c)
        if (tx + 32 * c < d) y[r * d + tx + 32 * c] = acc[a][c];
  }
}
// |x - mu| per row, rounded up (one warp per row); out_max (optional): maximum over the rows
__global__ void k_centered_norm(const float* __restrict__ x, long long n, int D, const float* __restrict__ mean,
                                float* out, unsigned* out_max) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  double s = 0;
  for (int j = lane; j < D; j += 32) {
    const double v = (double)x[r * D + j] - (double)mean[j];
    s += v * v;
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float v = (float)sqrt(s) * 1.000001f;
    if (out) out[r] = v;
    if (out_max) atomicMax(out_max, __float_as_uint(v));
  }
}
__global__ void k_sample_rows(const float* __restrict__ words, long long N, int D, int f, long long n_s, int* s2g,
                              float* out) {
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_s) return;
  unsigned h = (unsigned)i * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  long long g = i * f + (long long)(h % (unsigned)f);  // one pseudo-random row of every block of f
  if (g >= N) g = N - 1;
  if (lane == 0) s2g[i] = (int)g;
  for (int j = lane; j < D; j += 32) out[i * D + j] = words[g * D + j];
}
// candidate lists of the sample sweep: sample rows -> codebook rows (the re-rank reads the codebook)
__global__ void k_remap_cands(int* cand_idx, const int* __restrict__ cand_cnt, long long Q, int S, int cap,
                              const int* __restrict__ s2g) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Q * S * cap) return;
  const long long list = t / cap;  // [s][q]
  const int i = (int)(t % cap);
  if (i < min(cand_cnt[list], cap)) {
    const long long s = list / Q, q = list % Q;
    int* p = cand_idx + ((size_t)q * S + s) * cap + i;
    *p = s2g[*p];
  }
}
// Threshold of the pooled sweep over the projected operands.  A codeword c with |q - c|^2 <= U' has
//   |y_q - y_c| <= sqrt(1 + e) |q - c| + eta_q + eta_c =: T      (e: orthonormality defect of the basis, eta: fp32
//   rounding of the two projections), hence accumulator = |y_c|^2 - 2 y_q.y_c (known to within eps) <= T^2 - |y_q|^2 + eps.
// U' = U (1 + 1.01 g): U is an fp32 FLANN-order functor value, g bounds its rounding (see k_chi_thr).
__global__ void k_pca_thr(const float* __restrict__ part_d, int K, const float* __restrict__ qn2p,
                          const float* __restrict__ epsp, const float* __restrict__ rnorm_q, int* skip,
                          long long Q, int D, int d, double orth_defect, float eta_c, float* thr) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float U = part_d[q * K + K - 1];
  if (skip[q] || !(U < __int_as_float(0x7f800000))) {
    skip[q] = 1;  // no bound (fewer than K usable sample candidates): searched by the plain sweep
    thr[q] = __int_as_float(0xff800000);
    return;
  }
  const double g = 1.01 * (double)(D + 8) * 5.9604644775390625e-08;
  const double eta_q = 1.01 * (double)(D + 3) * 5.9604644775390625e-08 * sqrt((double)d) * (double)rnorm_q[q];
  const double T = sqrt((1.0 + orth_defect) * (double)U * (1.0 + g)) + eta_q + (double)eta_c;
  const double t = T * T * (1.0 + 1e-9) - (double)qn2p[q] * (1.0 - 2.4e-7) + (double)epsp[q] + 1e-12;
  float f = (float)t;
  if ((double)f < t) f = __uint_as_float(__float_as_uint(f) + (f >= 0.f ? 1u : -1u));  // round up
  thr[q] = f;
}
// queries whose pool went past the cap: flag them for the plain sweep and drop their pool entries
__global__ void k_pool_cap_flags(int* q_cnt, long long Q, int cap, int* flag) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q > Q) return;
  if (q == Q) {
    q_cnt[q] = 0;
    return;
  }
  if (q_cnt[q] > cap) {
    flag[q] = 1;
    q_cnt[q] = 0;
  } else if (flag[q]) {
    q_cnt[q] = 0;
  }
}

// symmetric eigen-decomposition on the host (cyclic Jacobi, fp64): only the ORTHONORMALITY of V matters for soundness
// and plane rotations keep it to rounding whatever the sweep count; the sweeps just have to find the leading subspace
void host_jacobi(std::vector<double>& A, int n, std::vector<double>& V, int sweeps) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sw = 0; sw < sweeps; ++sw) {
    double off = 0, diag = 0;
    for (int p = 0; p < n; ++p) {
      diag += A[(size_t)p * n + p] * A[(size_t)p * n + p];
      for (int q = p + 1; q < n; ++q) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
    }
    if (off <= 1e-22 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        double* Ap = &A[(size_t)p * n];
        double* Aq = &A[(size_t)q * n];
        for (int k = 0; k < n; ++k) {  // rows p, q
          const double a = Ap[k], b = Aq[k];
          Ap[k] = c * a - sn * b;
          Aq[k] = sn * a + c * b;
        }
        for (int k = 0; k < n; ++k) {  // columns p, q
          const double a = A[(size_t)k * n + p], b = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * a - sn * b;
          A[(size_t)k * n + q] = sn * a + c * b;
        }
        double* Vp = &V[(size_t)p * n];  // V stored by ROWS = eigenvectors (transposed accumulation)
        double* Vq = &V[(size_t)q * n];
        for (int k = 0; k < n; ++k) {
          const double a = Vp[k], b = Vq[k];
          Vp[k] = c * a - sn * b;
          Vq[k] = sn * a + c * b;
        }
      }
  }
}

// Leading d-dimensional invariant subspace of a symmetric n x n matrix (fp64) for LONG rows (CSHOT: n = 1344, where
// the cyclic Jacobi above would cost minutes): block power iteration on p = d + 16 vectors with modified Gram-Schmidt,
// then Rayleigh-Ritz on the p x p projection.  Only the SPAN matters for how many rows the filter keeps, and only the
// orthonormality of the result for its soundness (re-checked in fp64 on the fp32 values by the caller).
// out: d basis vectors of length n, stored by rows.
void host_subspace(const std::vector<double>& A, int n, int d, std::vector<double>& out) {
  const int p = std::min(n, d + 16);
  std::vector<double> X((size_t)n * p), Y((size_t)n * p);
  unsigned long long rng = 0x9e3779b97f4a7c15ull;
  for (double& v : X) {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    v = (double)(int)(rng >> 40) / 8388608.0 - 1.0;
  }
  auto orthonormalise = [&](std::vector<double>& M) {  // columns of M [n][p], two Gram-Schmidt passes
    for (int pass = 0; pass < 2; ++pass)
      for (int c = 0; c < p; ++c) {
        for (int b = 0; b < c; ++b) {
          double dot = 0;
          for (int i = 0; i < n; ++i) dot += M[(size_t)i * p + c] * M[(size_t)i * p + b];
          for (int i = 0; i < n; ++i) M[(size_t)i * p + c] -= dot * M[(size_t)i * p + b];
        }
        double nn = 0;
        for (int i = 0; i < n; ++i) nn += M[(size_t)i * p + c] * M[(size_t)i * p + c];
        nn = std::sqrt(nn);
        if (!(nn > 1e-300)) {  // degenerate direction: replace by a unit vector (orthogonalised on the second pass)
          for (int i = 0; i < n; ++i) M[(size_t)i * p + c] = (i == c % n) ? 1.0 : 0.0;
        } else {
          for (int i = 0; i < n; ++i) M[(size_t)i * p + c] /= nn;
        }
      }
  };
  auto multiply = [&](const std::vector<double>& Xin, std::vector<double>& Yout) {  // Y = A X
    std::fill(Yout.begin(), Yout.end(), 0.0);
    for (int i = 0; i < n; ++i) {
      double* y = &Yout[(size_t)i * p];
      for (int k = 0; k < n; ++k) {
        const double a = A[(size_t)i * n + k];
        const double* x = &Xin[(size_t)k * p];
        for (int c = 0; c < p; ++c) y[c] += a * x[c];
      }
    }
  };
  orthonormalise(X);
  for (int it = 0; it < 12; ++it) {
    multiply(X, Y);
    X.swap(Y);
    orthonormalise(X);
  }
  // Rayleigh-Ritz: T = X^T A X (p x p), its eigenvectors rotate X onto the best approximations inside the span
  multiply(X, Y);
  std::vector<double> T((size_t)p * p, 0.0), W;
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < p; ++a) {
      const double xa = X[(size_t)i * p + a];
      for (int b = 0; b < p; ++b) T[(size_t)a * p + b] += xa * Y[(size_t)i * p + b];
    }
  for (int a = 0; a < p; ++a)
    for (int b = a + 1; b < p; ++b) T[(size_t)a * p + b] = T[(size_t)b * p + a] = 0.5 * (T[(size_t)a * p + b] + T[(size_t)b * p + a]);
  host_jacobi(T, p, W, 12);  // W by rows = eigenvectors of T, eigenvalues on the diagonal of T
  std::vector<int> order(p);
  for (int i = 0; i < p; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return T[(size_t)a * p + a] > T[(size_t)b * p + b]; });
  out.assign((size_t)d * n, 0.0);
  for (int c = 0; c < d; ++c) {
    const double* w = &W[(size_t)order[c] * p];
    for (int i = 0; i < n; ++i) {
      double v = 0;
      for (int a = 0; a < p; ++a) v += X[(size_t)i * p + a] * w[a];
      out[(size_t)c * n + i] = v;
    }
  }
  // one more Gram-Schmidt pass over the d result vectors (rows of out)
  for (int pass = 0; pass < 2; ++pass)
    for (int c = 0; c < d; ++c) {
      double* vc = &out[(size_t)c * n];
      for (int b = 0; b < c; ++b) {
        const double* vb = &out[(size_t)b * n];
        double dot = 0;
        for (int i = 0; i < n; ++i) dot += vc[i] * vb[i];
        for (int i = 0; i < n; ++i) vc[i] -= dot * vb[i];
      }
      double nn = 0;
      for (int i = 0; i < n; ++i) nn += vc[i] * vc[i];
      nn = std::sqrt(nn);
      if (nn > 1e-300)
        for (int i = 0; i < n; ++i) vc[i] /= nn;
    }
}

template <int DC>
int launch_project(pcdb_ctx* ctx, const float* x, int64_t n, int D, int d, const float* mean, const float* basis,
                   float* y) {
  k_project<DC><<<cdiv(n, 64), 256, 0, ctx->stream>>>(x, n, D, d, mean, basis, y);
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}
int project_rows(pcdb_ctx* ctx, const float* x, int64_t n, int D, int d, const float* mean, const float* basis, float* y) {
  switch ((d + 31) / 32) {
    case 2: return launch_project<2>(ctx, x, n, D, d, mean, basis, y);
    case 3: return launch_project<3>(ctx, x, n, D, d, mean, basis, y);
    case 4: return launch_project<4>(ctx, x, n, D, d, mean, basis, y);
    case 5: return launch_project<5>(ctx, x, n, D, d, mean, basis, y);
    case 6: return launch_project<6>(ctx, x, n, D, d, mean, basis, y);
  }
  return ctx->fail(PCDB_E_INVALID, "PCA pre-filter: unsupported dimension %d", d);
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// builds the sample and projected operands of the PCA pre-filter for the resident codebook (Euclidean family)
int pca_prepare(pcdb_ctx* ctx) {
  Codebook_d& cb = ctx->cb;
  cudaStream_t st = ctx->stream;
  GemmState* gs = state_of(ctx);
  PcaFilter& pf = gs->pca;
  pf.release();
  const int d = env_int("PCDB_GEMM_PCA_D", 112), f = env_int("PCDB_GEMM_SAMPLE", 16);
  const int64_t min_rows = env_int("PCDB_GEMM_PCA_MIN_ROWS", 131072);
  if (!env_int("PCDB_GEMM_PCA", 1) || cb.N < min_rows || cb.D % 16 != 0 || d % 16 != 0 || d < 48 || d > 192 ||
      d + K_AUG >= cb.D || f < 2)
    return PCDB_OK;
  // long rows (CSHOT-1344: streaming-query bound sweep over the sample, basis by block power iteration); measured at
  // 1.07 M x 1344, 104 k queries: sweeps 33 ms against 172 ms for the plain streaming sweep (PCDB_GEMM_PCA_WIDE=0)
  if ((cb.D + K_AUG + BK - 1) / BK > KB_RES_MAX && !env_int("PCDB_GEMM_PCA_WIDE", 1)) return PCDB_OK;
  const int D = cb.D;
  const float* words = cb.words.as<float>();
  // ---- mean and covariance of up to 65536 strided rows
  const int64_t stride = std::max<int64_t>(1, cb.N / 65536), n_cov = cb.N / stride;
  DevBuf dsum, dcov;
  PCDB_CUDA(dsum.ensure(sizeof(double) * D));
  PCDB_CUDA(dcov.ensure(sizeof(double) * (size_t)D * D));
  auto free_tmp = [&]() { dsum.release(); dcov.release(); };
  PCDB_CUDA(cudaMemsetAsync(dsum.p, 0, sizeof(double) * D, st));
  k_col_sums<<<dim3(256, (D + 255) / 256), 256, 0, st>>>(words, n_cov, stride, D, dsum.as<double>());
  PCDB_LAUNCH_CHECK();
  std::vector<double> hmean(D);
  PCDB_CUDA(cudaMemcpyAsync(hmean.data(), dsum.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  for (double& v : hmean) v /= (double)n_cov;
  PCDB_CUDA(cudaMemcpyAsync(dsum.p, hmean.data(), sizeof(double) * D, cudaMemcpyHostToDevice, st));
  k_cov_tile<<<dim3(D / 16, D / 16), 256, 0, st>>>(words, n_cov, stride, D, dsum.as<double>(), dcov.as<double>());
  PCDB_LAUNCH_CHECK();
  std::vector<double> cov((size_t)D * D), V;
  PCDB_CUDA(cudaMemcpyAsync(cov.data(), dcov.p, sizeof(double) * (size_t)D * D, cudaMemcpyDeviceToHost, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  free_tmp();
  for (double v : cov)
    if (!std::isfinite(v)) return PCDB_OK;  // non-finite codewords: no pre-filter
  std::vector<double> lead;  // the d leading axes, by rows
  if (D <= 512) {
    host_jacobi(cov, D, V, 8);
    std::vector<int> order(D);
    for (int i = 0; i < D; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cov[(size_t)a * D + a] > cov[(size_t)b * D + b]; });
    lead.resize((size_t)d * D);
    for (int c = 0; c < d; ++c)
      for (int k = 0; k < D; ++k) lead[(size_t)c * D + k] = V[(size_t)order[c] * D + k];
  } else {
    host_subspace(cov, D, d, lead);
  }
  // basis [D][d] in fp32 + its orthonormality defect ||P^T P - I||_2 <= max row sum of |P^T P - I| (evaluated in fp64
  // on the fp32 values the device multiplies with)
  std::vector<float> hb((size_t)D * d), hm(D);
  for (int c = 0; c < d; ++c)
    for (int k = 0; k < D; ++k) hb[(size_t)k * d + c] = (float)lead[(size_t)c * D + k];
  double defect = 0;
  for (int a = 0; a < d; ++a) {
    double row = 0;
    for (int b = 0; b < d; ++b) {
      double g = 0;
      for (int k = 0; k < D; ++k) g += (double)hb[(size_t)k * d + a] * (double)hb[(size_t)k * d + b];
      row += std::fabs(g - (a == b ? 1.0 : 0.0));
    }
    defect = std::max(defect, row);
  }
  if (!(defect < 1e-3)) return PCDB_OK;
  double mn2 = 0;
  for (int k = 0; k < D; ++k) {
    hm[k] = (float)hmean[k];
    mn2 += (double)hm[k] * hm[k];
  }
  PCDB_CUDA(pf.basis.ensure(sizeof(float) * (size_t)D * d));
  PCDB_CUDA(pf.mean.ensure(sizeof(float) * D));
  PCDB_CUDA(cudaMemcpyAsync(pf.basis.p, hb.data(), sizeof(float) * (size_t)D * d, cudaMemcpyHostToDevice, st));
  PCDB_CUDA(cudaMemcpyAsync(pf.mean.p, hm.data(), sizeof(float) * D, cudaMemcpyHostToDevice, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  pf.orth_defect = defect * 1.01 + 1e-12;
  pf.mean_norm = (float)std::sqrt(mn2);
  pf.d = d;
  pf.f = f;
  // ---- projected codebook operand
  DevBuf tmp;
  PCDB_CUDA(tmp.ensure(sizeof(float) * (size_t)cb.N * d));
  int rc = project_rows(ctx, words, cb.N, D, d, pf.mean.as<float>(), pf.basis.as<float>(), tmp.as<float>());
  if (rc == PCDB_OK) rc = prepare_operand(ctx, tmp.as<float>(), cb.N, d, false, pf.proj);
  unsigned* mx = reinterpret_cast<unsigned*>(ctx->ws.scalars.as<char>() + 160);
  if (rc == PCDB_OK && cudaMemsetAsync(mx, 0, 4, st) != cudaSuccess) rc = PCDB_E_CUDA;
  if (rc == PCDB_OK) {
    k_centered_norm<<<cdiv(cb.N * 32, 256), 256, 0, st>>>(words, cb.N, D, pf.mean.as<float>(), nullptr, mx);
    float rmax = 0;
    if (cudaMemcpyAsync(&rmax, mx, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      rc = PCDB_E_CUDA;
    pf.eta_c = (float)(1.01 * (double)(D + 3) * 5.9604644775390625e-08 * std::sqrt((double)d) * (double)rmax * 1.000001);
  }
  tmp.release();
  PCDB_TRY(rc);
  if (!pf.proj.ready) return PCDB_OK;
  // ---- sample operand (full-length rows)
  const int64_t n_s = cb.N / f;
  if (n_s < 4 * BN) return PCDB_OK;
  PCDB_CUDA(pf.s2g.ensure(sizeof(int) * (size_t)n_s));
  PCDB_CUDA(tmp.ensure(sizeof(float) * (size_t)n_s * D));
  k_sample_rows<<<cdiv(n_s * 32, 256), 256, 0, st>>>(words, cb.N, D, f, n_s, pf.s2g.as<int>(), tmp.as<float>());
  rc = cudaGetLastError() == cudaSuccess ? PCDB_OK : PCDB_E_CUDA;
  if (rc == PCDB_OK) rc = prepare_operand(ctx, tmp.as<float>(), n_s, D, false, pf.sample);
  tmp.release();
  PCDB_TRY(rc);
  pf.ready = pf.sample.ready && pf.proj.ready;
  return PCDB_OK;
}

}  // namespace

bool gemm_supported(const pcdb_ctx* ctx, int dist_type) { return ctx->cb.gemm_ready[dist_type == PCDB_DIST_CHISQUARED]; }

// fp16 operand copy of the codebook for one distance family (the rows, or their square roots for chi^2), |c|^2, error
// maxima and the codebook-side tensor map.  Called by pcdb_set_codebook for the context's DistanceType and lazily by the
// first search that uses the other one.
int gemm_prepare_codebook(pcdb_ctx* ctx, int dist_type) {
  Codebook_d& cb = ctx->cb;
  const int fam = dist_type == PCDB_DIST_CHISQUARED ? 1 : 0;
  cb.gemm_ready[fam] = false;
  cb.gemm_tried[fam] = true;
  GemmState* gs = state_of(ctx);
  if (fam == 0) gs->pca.release();
  if (cb.D % 16 != 0 || cb.D < BK || cb.N < 1) return PCDB_OK;  // scan path only
  PCDB_TRY(prepare_operand(ctx, cb.words.as<float>(), cb.N, cb.D, fam == 1, gs->op[fam]));
  cb.gemm_ready[fam] = gs->op[fam].ready;
  if (fam == 0 && cb.gemm_ready[0]) PCDB_TRY(pca_prepare(ctx));
  return PCDB_OK;
}

namespace {

// sweep geometry of one launch: slices, candidate capacity, grid
struct SweepPlan {
  GemmArgs g;
  int grid = 0;
  bool a_res = false;
};

// fills the part of GemmArgs that depends on (queries, operand rows, row length) only
int plan_sweep(pcdb_ctx* ctx, GemmState* gs, int64_t Q, int64_t n_rows, int Dh, bool a_res, int K, SweepPlan& p) {
  GemmArgs& g = p.g;
  std::memset(&g, 0, sizeof(g));
  g.Q = Q;
  g.N = n_rows;
  g.D = Dh;
  g.n_mpairs = (int)cdiv(Q, 2 * BM);
  g.n_ntiles = (int)cdiv(n_rows, BN);
  // codebook slices: small enough to stay L2-resident while every query-tile pair passes over them (20 MB of fp16
  // rows), and at least as many as it takes to give every CTA pair a unit when there are few queries
  const int max_pairs = std::max(1, ctx->sm_count / 2);
  const int64_t tile_bytes = (int64_t)BN * Dh * (int64_t)sizeof(__half);
  static const int64_t slice_mb = [] {  // tuning knob for experiments; the default is what profiles/ was measured with
    const char* e = getenv("PCDB_GEMM_SLICE_MB");
    const long v = e ? atol(e) : 0;
    return (int64_t)(v > 0 ? v : 20);
  }();
  // (the streaming-query variant for D = 1344 re-reads its query tile with every codebook tile and sits on the L2->SM
  // limit either way; it keeps the query-major sweep, whose L2 hit rate measured 99 %)
  int S = a_res ? (int)cdiv((int64_t)g.n_ntiles * tile_bytes, slice_mb << 20) : 1;
  if (g.n_mpairs < max_pairs) {
    // Few queries (one cloud): every (slice, query pair) unit runs at the same time, the kernel takes as long as the
    // CTA pair with the most tiles.  Units are dealt round-robin, so pick the slice count that minimises
    // rounds x tiles per slice — 2 query pairs on 74 CTA pairs want 37 slices (one round), not 38 (two).
    const int pairs = gs->max_clusters[a_res ? 1 : 0] > 0 ? gs->max_clusters[a_res ? 1 : 0] : max_pairs;
    long long best = -1;
    int best_s = S;
    for (int s = 1; s <= std::min(g.n_ntiles, 4 * pairs); ++s) {
      const long long cost = (long long)cdiv((int64_t)g.n_mpairs * s, pairs) * (long long)cdiv(g.n_ntiles, s);
      if (best < 0 || cost < best) {
        best = cost;
        best_s = s;
      }
    }
    S = best_s;
  }
  static const int s_env = [] { const char* e = getenv("PCDB_GEMM_S"); return e ? atoi(e) : 0; }();  // experiments
  if (s_env > 0) S = s_env;
  S = std::max(1, std::min(S, g.n_ntiles));
  g.tiles_per_split = (int)cdiv(g.n_ntiles, S);
  S = (int)cdiv(g.n_ntiles, g.tiles_per_split);
  g.n_splits = S;
  // Slices of one query tile run one after the other when there are more tile pairs than CTA pairs; with few queries
  // they run side by side, each starting from an empty bound and contributing its own descending staircase
  const int in_flight = std::min(S, (int)cdiv(max_pairs, g.n_mpairs));
  g.stage_cap = K <= 4 ? 64 : (K <= 8 ? 128 : CAND_CAP_MAX);
  g.cand_cap = g.stage_cap * std::max(1, in_flight);
  PCDB_CUDA(gs->stage.ensure(sizeof(int2) * (size_t)ctx->sm_count * EPI_THREADS * CAND_CAP_MAX));
  g.stage = gs->stage.as<int2>();
  p.grid = 2 * std::min(g.n_mpairs * S, max_pairs);  // CTA pairs (cluster of 2)
  p.a_res = a_res;
  return PCDB_OK;
}

// Exact kNN on the tensor cores.
//   Euclidean: one sweep with the running-bound filter, exact fp32 re-rank of the candidates.  Large codebooks and
//   batches take the PCA pre-filter (struct PcaFilter): the bound sweep runs over a 1/f sample of the codebook, a pooled
//   sweep over the projected operands collects every codeword that can still be among the K nearest, the exact functor
//   decides, and the few queries whose pool runs past its cap are searched by the plain sweep.
//   ChiSquared: the Hellinger sandwich, two sweeps over the sqrt-transformed operands.  Sweep 1 is the Euclidean
//   machinery on sqrt rows (its candidates, re-ranked with the exact chi^2 functor, give U = the K-th smallest chi^2
//   of K real codewords); sweep 2 collects EVERY row whose Hellinger value can still be <= U into a pooled list, and
//   the exact FLANN-order chi^2 of those rows decides.  The result is the exact scan's, bit for bit.
// depth: 0 = a caller's search, 1 = the search of the queries a depth-0 call could not settle (plain sweep, own buffers)
int knn_gemm_impl(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, bool use_ratio,
                  float ratio_thr, int depth) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  GemmState* gs = state_of(ctx);
  const bool chi = dist_type == PCDB_DIST_CHISQUARED;
  PCDB_CUDA(w.knn_idx.ensure(sizeof(int) * (Q * k + 1)));
  PCDB_CUDA(w.knn_dist.ensure(sizeof(float) * (Q * k + 1)));
  PCDB_CUDA(w.knn_cnt.ensure(sizeof(int) * (Q + 1)));
  if (Q == 0) return PCDB_OK;
  const int K = use_ratio ? k + 1 : k;
  const int D = cb.D;
  PcaFilter& pf = gs->pca;
  static const int pca_min_q = env_int("PCDB_GEMM_PCA_MIN_Q", 8192);
  bool pca = !chi && depth == 0 && pf.ready && pf.sample.n > K && Q >= pca_min_q &&
             !env_int("PCDB_GEMM_PCA_SKIP", 0);  // read per call: A/B runs against the plain sweep on one context
  if (pca && pf.backoff_left > 0) {  // recent batches overflowed their pools: plain sweep for now
    --pf.backoff_left;
    pca = false;
  }
  GemmOperand& op = pca ? pf.sample : gs->op[chi ? 1 : 0];  // operand of the bound sweep
  // query side: fp16 copy + norms + margins
  const int aug = (D + K_AUG + BK - 1) / BK <= KB_RES_MAX ? 1 : 0;
  const int Dh = D + (aug ? K_AUG : 0);
  PCDB_CUDA(w.feat_h.ensure(sizeof(__half) * (size_t)Q * Dh + 256));
  PCDB_CUDA(gs->qnorm_h.ensure(sizeof(float) * (Q + 1)));
  PCDB_CUDA(gs->qerr.ensure(sizeof(float) * (Q + 1)));
  PCDB_CUDA(gs->margin.ensure(sizeof(float) * (Q + 1)));
  if (chi) {
    PCDB_CUDA(gs->qn2.ensure(sizeof(float) * (Q + 1)));
    PCDB_CUDA(gs->qbad.ensure(sizeof(int) * (Q + 1)));
    PCDB_CUDA(gs->eps.ensure(sizeof(float) * (Q + 1)));
    k_prep_rows<false, true><<<cdiv(Q * 32, 256), 256, 0, st>>>(queries_d, Q, D, aug, w.feat_h.as<__half>(),
                                                               gs->qn2.as<float>(), gs->qnorm_h.as<float>(),
                                                               gs->qerr.as<float>(), gs->qbad.as<int>(), nullptr);
  } else {
    k_prep_rows<false, false><<<cdiv(Q * 32, 256), 256, 0, st>>>(queries_d, Q, D, aug, w.feat_h.as<__half>(), nullptr,
                                                                gs->qnorm_h.as<float>(), gs->qerr.as<float>(), nullptr,
                                                                nullptr);
  }
  PCDB_LAUNCH_CHECK();
  k_margin<<<cdiv(Q, 256), 256, 0, st>>>(gs->qnorm_h.as<float>(), gs->qerr.as<float>(), Q, D, aug, op.cmax_h,
                                         op.cerr_max, op.cmax2, chi ? 1 : 0, gs->margin.as<float>(),
                                         chi ? gs->eps.as<float>() : nullptr);
  PCDB_LAUNCH_CHECK();
  CUtensorMap map_a;
  PCDB_TRY(make_map(ctx, &map_a, w.feat_h.p, Q, Dh, BM));
  const bool a_res = aug != 0;
  SweepPlan sp;
  PCDB_TRY(plan_sweep(ctx, gs, Q, op.n, Dh, a_res, K, sp));
  GemmArgs& g = sp.g;
  const int grid = sp.grid;
  g.margin = gs->margin.as<float>();
  g.cnorm = op.cnorm.as<float>();
  const int S2 = 2;  // candidate lists: one per column half, shared by all slices
  const size_t nc = (size_t)Q * S2 * g.cand_cap;
  PCDB_CUDA(w.cand_idx.ensure(sizeof(int) * nc + 16));
  PCDB_CUDA(w.cand_apx.ensure(sizeof(float) * nc + 16));
  PCDB_CUDA(w.cand_cnt.ensure(sizeof(int) * ((size_t)Q * S2 + 1)));
  PCDB_CUDA(w.cand_thr.ensure(sizeof(float) * ((size_t)Q * S2 + 1)));
  PCDB_CUDA(w.scalars.ensure(256));
  PCDB_CUDA(cudaMemsetAsync(w.scalars.as<char>() + 64, 0, 8, st));
  g.cand_idx = w.cand_idx.as<int>();
  g.cand_apx = w.cand_apx.as<float>();
  g.cand_cnt = w.cand_cnt.as<int>();
  PCDB_CUDA(gs->bound.ensure(sizeof(float) * (Q + 1)));
  g.bound = gs->bound.as<float>();
  PCDB_CUDA(cudaMemsetAsync(w.cand_cnt.p, 0, sizeof(int) * (size_t)Q * S2, st));
  k_fill_f32<<<cdiv(Q, 256), 256, 0, st>>>(g.bound, Q, INFINITY);
  PCDB_LAUNCH_CHECK();
  cudaEvent_t e0 = ctx->ev[5], e1 = ctx->ev[6];
  if (depth == 0) {
    PCDB_CUDA(cudaEventRecord(e0, st));
    PCDB_CUDA(cudaEventRecord(ctx->ev_knn[0], st));
    ctx->stats.knn_prefilter_dim = pca ? pf.d : 0;
    ctx->stats.knn_prefilter_sample_rows = pca ? pf.sample.n : 0;
  }
  if (a_res)
    PCDB_TRY((launch_gemm_kt<true>(ctx, K, map_a, op.map_b, g, grid)));
  else
    PCDB_TRY((launch_gemm_kt<false>(ctx, K, map_a, op.map_b, g, grid)));
  if (depth == 0) PCDB_CUDA(cudaEventRecord(ctx->ev_knn[1], st));
  if (!chi && !pca && depth == 0) {
    PCDB_CUDA(cudaEventRecord(e1, st));
    ctx->gemm_events_valid = true;
  }
  if (depth == 0) pcdb_trace_point(ctx, pca ? "knn: sample sweep" : "knn: sweep");
  k_final_thr<<<cdiv(Q, 256), 256, 0, st>>>(g.bound, g.margin, Q, w.cand_thr.as<float>());
  PCDB_LAUNCH_CHECK();
  if (pca) {  // the re-rank reads the codebook: sample rows -> codebook rows
    k_remap_cands<<<cdiv((int64_t)nc, 256), 256, 0, st>>>(w.cand_idx.as<int>(), w.cand_cnt.as<int>(), Q, S2, g.cand_cap,
                                                          pf.s2g.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  // exact functor values of the candidates, the K best per query -> ws.knn_part_*
  PCDB_TRY(stage_knn_rerank(ctx, queries_d, Q, K, S2, g.cand_cap, dist_type));
  // queries that need another search: an overflowed candidate list (or, chi^2, a negative entry)
  PCDB_CUDA(gs->fb_flag.ensure(sizeof(int) * (Q + 2)));
  PCDB_CUDA(gs->fb_pos.ensure(sizeof(int) * (Q + 2)));
  k_overflow_flags<<<cdiv(Q + 1, 256), 256, 0, st>>>(w.cand_cnt.as<int>(), S2, Q, g.cand_cap,
                                                     chi ? gs->qbad.as<int>() : nullptr, gs->fb_flag.as<int>());
  PCDB_LAUNCH_CHECK();

  if (depth == 0) pcdb_trace_point(ctx, "knn: re-rank");
  if (chi || pca) {
    // ---- sweep 2: every row whose lower bound (Hellinger value / projected distance) can still be <= U, pooled
    PCDB_CUDA(gs->thr2.ensure(sizeof(float) * (Q + 1)));
    CUtensorMap map_a2 = map_a;
    const GemmOperand* op2 = &op;
    SweepPlan sp2 = sp;
    static const int q_cap_env = env_int("PCDB_GEMM_QCAP", 2048);
    if (chi) {
      k_chi_thr<<<cdiv(Q, 256), 256, 0, st>>>(w.knn_part_d.as<float>(), K, gs->qn2.as<float>(), gs->eps.as<float>(),
                                              gs->fb_flag.as<int>(), Q, D, gs->thr2.as<float>());
      PCDB_LAUNCH_CHECK();
    } else {
      // project the queries, build their fp16 operand, the error bound of that operand pair and the pool threshold
      const int d = pf.d, dh = d + K_AUG;
      PCDB_CUDA(pf.yq.ensure(sizeof(float) * (size_t)Q * d + 256));
      PCDB_CUDA(pf.yq_h.ensure(sizeof(__half) * (size_t)Q * dh + 256));
      for (DevBuf* b : {&pf.rnorm_q, &pf.qn2p, &pf.qnormp, &pf.qerrp, &pf.epsp, &pf.marginp})
        PCDB_CUDA(b->ensure(sizeof(float) * (Q + 1)));
      PCDB_TRY(project_rows(ctx, queries_d, Q, D, d, pf.mean.as<float>(), pf.basis.as<float>(), pf.yq.as<float>()));
      k_centered_norm<<<cdiv(Q * 32, 256), 256, 0, st>>>(queries_d, Q, D, pf.mean.as<float>(), pf.rnorm_q.as<float>(),
                                                         nullptr);
      PCDB_LAUNCH_CHECK();
      k_prep_rows<false, false><<<cdiv(Q * 32, 256), 256, 0, st>>>(pf.yq.as<float>(), Q, d, 1, pf.yq_h.as<__half>(),
                                                                  pf.qn2p.as<float>(), pf.qnormp.as<float>(),
                                                                  pf.qerrp.as<float>(), nullptr, nullptr);
      PCDB_LAUNCH_CHECK();
      k_margin<<<cdiv(Q, 256), 256, 0, st>>>(pf.qnormp.as<float>(), pf.qerrp.as<float>(), Q, d, 1, pf.proj.cmax_h,
                                             pf.proj.cerr_max, pf.proj.cmax2, 0, pf.marginp.as<float>(),
                                             pf.epsp.as<float>());
      PCDB_LAUNCH_CHECK();
      k_pca_thr<<<cdiv(Q, 256), 256, 0, st>>>(w.knn_part_d.as<float>(), K, pf.qn2p.as<float>(), pf.epsp.as<float>(),
                                              pf.rnorm_q.as<float>(), gs->fb_flag.as<int>(), Q, D, d, pf.orth_defect,
                                              pf.eta_c, gs->thr2.as<float>());
      PCDB_LAUNCH_CHECK();
      PCDB_TRY(make_map(ctx, &map_a2, pf.yq_h.p, Q, dh, BM));
      if (env_int("PCDB_EXP_POOL_NOTHR", 0)) {  // experiment: nothing passes the pooled sweep (kernel floor)
        k_fill_f32<<<cdiv(Q, 256), 256, 0, st>>>(gs->thr2.as<float>(), Q, -INFINITY);
        PCDB_LAUNCH_CHECK();
      }
      pcdb_trace_point(ctx, "knn: project queries");
      op2 = &pf.proj;
      PCDB_TRY(plan_sweep(ctx, gs, Q, pf.proj.n, dh, true, 8, sp2));  // K = 8: 128-entry private lists
    }
    PCDB_CUDA(gs->q_cnt.ensure(sizeof(int) * (Q + 2)));
    PCDB_CUDA(gs->q_off.ensure(sizeof(int) * (Q + 2)));
    PCDB_CUDA(gs->q_fill.ensure(sizeof(int) * (Q + 2)));
    unsigned long long* pool_count = reinterpret_cast<unsigned long long*>(w.scalars.as<char>() + 72);
    GemmArgs g2 = sp2.g;
    g2.margin = gs->thr2.as<float>();
    g2.cnorm = op2->cnorm.as<float>();
    g2.pool_count = pool_count;
    g2.q_cnt = gs->q_cnt.as<int>();
    g2.q_cap = pca ? q_cap_env : 0;
    // bit 0: passing chunks are parked in shared memory, bit 1: early hand-over of the accumulator (short rows only),
    // bit 2: experiment without read-back (wrong results; the hand-over's own cost).  Read per call: A/B runs.
    g2.stash = env_int("PCDB_GEMM_POOL_STASH", 3);
    unsigned long long total = 0;
    for (int attempt = 0;; ++attempt) {
      const int64_t want = std::max<int64_t>(gs->pool_cap, std::max<int64_t>((int64_t)1 << 22, Q * 64));
      PCDB_CUDA(gs->pool_rc.ensure(sizeof(int2) * (size_t)want));
      PCDB_CUDA(gs->pool_q.ensure(sizeof(int) * (size_t)want));
      gs->pool_cap = want;
      g2.pool_rc = gs->pool_rc.as<int2>();
      g2.pool_q = gs->pool_q.as<int>();
      g2.pool_cap = want;
      PCDB_CUDA(cudaMemsetAsync(pool_count, 0, 8, st));
      PCDB_CUDA(cudaMemsetAsync(gs->q_cnt.p, 0, sizeof(int) * (Q + 1), st));
      if (depth == 0) PCDB_CUDA(cudaEventRecord(ctx->ev_knn[2], st));
      if (sp2.a_res)
        PCDB_TRY((launch_gemm<true, 1, true>(ctx, map_a2, op2->map_b, g2, sp2.grid)));
      else
        PCDB_TRY((launch_gemm<false, 1, true>(ctx, map_a2, op2->map_b, g2, sp2.grid)));
      if (depth == 0) {
        PCDB_CUDA(cudaEventRecord(ctx->ev_knn[3], st));
        ctx->knn_sweep_events_valid = true;
      }
      PCDB_TRY(pcdb_read_small(ctx, &total, pool_count, 8));
      PCDB_TRY(pcdb_sync_reads(ctx));
      if ((int64_t)total <= gs->pool_cap) break;
      if (attempt >= 2 || total > 0x7fff0000ull)
        return ctx->fail(PCDB_E_CAPACITY, "candidate pool: %llu entries for %lld queries", total, (long long)Q);
      gs->pool_cap = (int64_t)(total + total / 4);  // grow and sweep again (first batches of a new workload only)
    }
    if (depth == 0) {
      PCDB_CUDA(cudaEventRecord(e1, st));
      ctx->gemm_events_valid = true;
    }
    ctx->stats.knn_candidates += (int64_t)total;
    if (depth == 0) pcdb_trace_point(ctx, "knn: pooled sweep");
    if (pca) {  // queries past their pool cap join the fallback list; their pool entries are dropped
      k_pool_cap_flags<<<cdiv(Q + 1, 256), 256, 0, st>>>(gs->q_cnt.as<int>(), Q, g2.q_cap, gs->fb_flag.as<int>());
      PCDB_LAUNCH_CHECK();
    }
    PCDB_TRY(stage_knn_chi_pool(ctx, queries_d, Q, K, (int64_t)total, gs->pool_rc.as<int2>(), gs->pool_q.as<int>(),
                                gs->q_cnt.as<int>(), gs->q_off.as<int>(), gs->q_fill.as<int>(), &gs->csr_row,
                                &gs->csr_d, dist_type, pca ? gs->fb_flag.as<int>() : nullptr));
  }
  PCDB_TRY(pcdb_cub_exclusive_sum_i32(ctx, gs->fb_flag.as<int>(), gs->fb_pos.as<int>(), Q + 1));
  PCDB_TRY(stage_knn_finish(ctx, Q, k, K, use_ratio, ratio_thr));
  if (depth == 0 && (chi || pca)) pcdb_trace_point(ctx, "knn: pool evaluation");

  // queries the search above could not settle: searched again (still on the GPU) and merged back
  int n_fb = 0;
  unsigned long long n_eval = 0;
  PCDB_TRY(pcdb_read_small(ctx, &n_fb, gs->fb_pos.as<int>() + Q, sizeof(int)));
  PCDB_TRY(pcdb_read_small(ctx, &n_eval, w.scalars.as<char>() + 64, sizeof(n_eval)));
  PCDB_TRY(pcdb_sync_reads(ctx));
  if (!chi && !pca) ctx->stats.knn_candidates += (int64_t)n_eval;
  if (!pca) ctx->stats.knn_fallback_queries += n_fb;
  else ctx->stats.knn_prefilter_resweep_queries += n_fb;  // handed to the plain sweep, not to the exact scan
  if (pca) {
    static const int backoff_env = env_int("PCDB_GEMM_PCA_BACKOFF", 1);
    if (backoff_env && (int64_t)n_fb * 4 > Q) {
      pf.backoff = std::min(256, std::max(16, 2 * pf.backoff));
      pf.backoff_left = pf.backoff;
    } else {
      pf.backoff = 0;
    }
  }
  if (n_fb > 0) {
    FallbackBufs& fb = gs->fb[depth ? 1 : 0];
    PCDB_CUDA(fb.q.ensure(sizeof(float) * (size_t)n_fb * D));
    PCDB_CUDA(fb.list.ensure(sizeof(int) * n_fb));
    PCDB_CUDA(fb.idx.ensure(sizeof(int) * ((size_t)Q * k + 1)));
    PCDB_CUDA(fb.dist.ensure(sizeof(float) * ((size_t)Q * k + 1)));
    PCDB_CUDA(fb.cnt.ensure(sizeof(int) * (Q + 1)));
    k_overflow_gather<<<cdiv(Q * 32, 256), 256, 0, st>>>(gs->fb_flag.as<int>(), gs->fb_pos.as<int>(), Q, D, queries_d,
                                                         fb.list.as<int>(), fb.q.as<float>());
    PCDB_LAUNCH_CHECK();
    // keep the results aside, search the remaining queries into the knn_* buffers, then merge back
    PCDB_CUDA(cudaMemcpyAsync(fb.idx.p, w.knn_idx.p, sizeof(int) * Q * k, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(fb.dist.p, w.knn_dist.p, sizeof(float) * Q * k, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(fb.cnt.p, w.knn_cnt.p, sizeof(int) * Q, cudaMemcpyDeviceToDevice, st));
    if (pca)  // the plain single sweep over the whole codebook; ITS overflows go to the exact scan
      PCDB_TRY(knn_gemm_impl(ctx, fb.q.as<float>(), n_fb, k, dist_type, use_ratio, ratio_thr, 1));
    else
      PCDB_TRY(stage_knn_scan(ctx, fb.q.as<float>(), n_fb, k, dist_type, use_ratio, ratio_thr));
    k_overflow_scatter<<<cdiv(n_fb, 128), 128, 0, st>>>(fb.list.as<int>(), n_fb, k, w.knn_idx.as<int>(),
                                                        w.knn_dist.as<float>(), w.knn_cnt.as<int>(), fb.idx.as<int>(),
                                                        fb.dist.as<float>(), fb.cnt.as<int>());
    PCDB_LAUNCH_CHECK();
    PCDB_CUDA(cudaMemcpyAsync(w.knn_idx.p, fb.idx.p, sizeof(int) * Q * k, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(w.knn_dist.p, fb.dist.p, sizeof(float) * Q * k, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(w.knn_cnt.p, fb.cnt.p, sizeof(int) * Q, cudaMemcpyDeviceToDevice, st));
    if (depth == 0) pcdb_trace_point(ctx, "knn: remaining queries");
  }
  return PCDB_OK;
}

}  // namespace

int stage_knn_gemm(pcdb_ctx* ctx, const float* queries_d, int64_t Q, int k, int dist_type, bool use_ratio,
                   float ratio_thr) {
  return knn_gemm_impl(ctx, queries_d, Q, k, dist_type, use_ratio, ratio_thr, 0);
}
