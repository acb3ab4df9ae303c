// comm.cu — the multi-GPU exchange steps of the path, one process (or thread) per GPU, NCCL over
// NVLink 5 / NVSwitch.  Everything stays on the device: the collectives run on the context's stream between this
// library's own kernels, and the only host round trips are the element counts that size the buffers.
//
//  * Row-sharded codebook (config C4, "codebooks too large for one GPU's share"): every rank holds the descriptor
//    rows [row_lo, row_hi) — fp32 for the exact functor plus the fp16 tensor-core operand, 99 % of a codebook's bytes —
//    and the COMPLETE vote tables (80 bytes per codeword).  One activation =
//        all-gather(v) of the ranks' query descriptors  ->  every rank searches ALL queries in ITS rows (tcgen05 GEMM +
//        exact re-rank, the same kernels as the replicated path)  ->  one all-to-all of (f32 distance, i32 global row)
//        x K per query, each rank receiving the n_ranks partial lists of ITS OWN queries  ->  per-query merge on the
//        device (ties -> lower global row, then the N<=k / distance-ratio semantics of activateKNN).
//    The owner of the queries then casts the votes from its replicated tables: no vote exchange, and the vote list is
//    the one an unsharded codebook produces, in the same order.
//    Replaces the single FLANN index of codebook/codebook.cpp:483-538 built at utils/flann_helper.cpp:21-70.
//  * Keypoint-sharded scene (config C5, one large scene on several GPUs): cloud and codebook replicated, rank r
//    describes / activates / votes for the r-th contiguous slice of the voxel-grid keypoints, the votes are
//    all-gathered(v) in rank order — keypoint order — and every rank runs the maxima search on the complete list.
//    Replaces the keypoint loop of ImplicitShapeModel::detect (implicit_shape_model.cpp:583-712).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host program such as PyTorch already loaded, else the
// system one), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "common.cuh"
#include "stages.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?");
    return &api;
  }
#define PCDB_SYM(field, name)                                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));           \
  if (!api.field && api.err.empty()) api.err = std::string("libnccl lacks ") + name;
  PCDB_SYM(GetUniqueId, "ncclGetUniqueId")
  PCDB_SYM(CommInitRank, "ncclCommInitRank")
  PCDB_SYM(CommDestroy, "ncclCommDestroy")
  PCDB_SYM(AllGather, "ncclAllGather")
  PCDB_SYM(Broadcast, "ncclBroadcast")
  PCDB_SYM(Send, "ncclSend")
  PCDB_SYM(Recv, "ncclRecv")
  PCDB_SYM(GroupStart, "ncclGroupStart")
  PCDB_SYM(GroupEnd, "ncclGroupEnd")
  PCDB_SYM(GetErrorString, "ncclGetErrorString")
  PCDB_SYM(GetVersion, "ncclGetVersion")
#undef PCDB_SYM
  if (!api.err.empty()) {
    dlclose(api.lib);
    api.lib = nullptr;
  }
  return &api;
}

struct CommState {
  ncclComm_t comm = nullptr;
  int rank = 0, n = 1;
  bool shard_keypoints = false;
  DevBuf cnt_d, q_all, send, recv, votes_all, key_local, key_all, key_sorted, ord, ord_sorted;
  std::vector<int64_t> counts;
  ~CommState() {
    DevBuf* all[] = {&cnt_d, &q_all, &send, &recv, &votes_all, &key_local, &key_all, &key_sorted, &ord, &ord_sorted};
    for (DevBuf* b : all) b->release();
  }
};

CommState* comm_of(pcdb_ctx* ctx) { return static_cast<CommState*>(ctx->comm_state); }

#define PCDB_NCCL(call)                                                                                  \
  do {                                                                                                   \
    ncclResult_t r__ = (call);                                                                           \
    if (r__ != ncclSuccess)                                                                              \
      return ctx->fail(PCDB_E_COMM, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                       nccl_api()->GetErrorString(r__));                                                 \
  } while (0)

// element counts of every rank (one int64 each): the only host round trip of an exchange
int gather_counts(pcdb_ctx* ctx, int64_t mine, std::vector<int64_t>& all) {
  CommState* c = comm_of(ctx);
  NcclApi* a = nccl_api();
  cudaStream_t st = ctx->stream;
  PCDB_CUDA(c->cnt_d.ensure(sizeof(int64_t) * (c->n + 1)));
  int64_t* d = c->cnt_d.as<int64_t>();
  PCDB_CUDA(cudaMemcpyAsync(d + c->n, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  PCDB_NCCL(a->AllGather(d + c->n, d, 1, ncclInt64, c->comm, st));
  all.assign(c->n, 0);
  PCDB_CUDA(cudaMemcpyAsync(all.data(), d, sizeof(int64_t) * c->n, cudaMemcpyDeviceToHost, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  return PCDB_OK;
}

// all-gather of differently sized byte ranges: one broadcast per root inside a group (NCCL fuses them)
int all_gather_v(pcdb_ctx* ctx, const void* send_d, void* recv_d, const std::vector<int64_t>& counts, size_t elem) {
  CommState* c = comm_of(ctx);
  NcclApi* a = nccl_api();
  PCDB_NCCL(a->GroupStart());
  size_t off = 0;
  for (int r = 0; r < c->n; ++r) {
    const size_t bytes = (size_t)counts[r] * elem;
    if (bytes) {
      void* slot = static_cast<char*>(recv_d) + off;  // non-roots pass a valid (unused) send pointer
      ncclResult_t rc = a->Broadcast(r == c->rank ? send_d : slot, slot, bytes, ncclInt8, r, c->comm, ctx->stream);
      if (rc != ncclSuccess) {
        a->GroupEnd();
        return ctx->fail(PCDB_E_COMM, "ncclBroadcast failed: %s", a->GetErrorString(rc));
      }
    }
    off += bytes;
  }
  PCDB_NCCL(a->GroupEnd());
  ctx->stats.comm_bytes += (int64_t)(off - (size_t)counts[c->rank] * elem);  // bytes this rank received over NVLink
  return PCDB_OK;
}

__global__ void k_pack_pairs(const int* __restrict__ idx, const float* __restrict__ dist, const int* __restrict__ cnt,
                             long long Q, int K, int2* out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Q * K) return;
  const long long q = t / K;
  const int j = (int)(t % K);
  const bool ok = j < cnt[q] && idx[t] >= 0;
  out[t] = ok ? make_int2(idx[t], __float_as_int(dist[t])) : make_int2(-1, 0x7f800000);
}

// Per-query merge of the n_ranks partial lists (any order inside a list), then activateKNN's semantics on the merged
// list (activation_strategy_knn.h:50-54,75-85).  recv: [n][Q][K] (row, distance bits), global row ids.
__global__ void k_merge_pairs(const int2* __restrict__ recv, int n, long long Q, int K, int k, int use_ratio,
                              float ratio_thr, int by_row, long long n_total, int* idx_out, float* dist_out,
                              int* cnt_out) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float bd[PCDB_MAX_K + 1];
  int bi[PCDB_MAX_K + 1];
  int cnt = 0;
  for (int s = 0; s < n; ++s)
    for (int j = 0; j < K; ++j) {
      const int2 e = recv[((long long)s * Q + q) * K + j];
      if (e.x < 0) continue;
      const float d = __int_as_float(e.y);
      const int idx = e.x;
#define PCDB_LESS(da, ia, db, ib) (by_row ? (ia) < (ib) : ((da) < (db) || ((da) == (db) && (ia) < (ib))))
      if (cnt == K && !PCDB_LESS(d, idx, bd[K - 1], bi[K - 1])) continue;
      int p = (cnt < K) ? cnt++ : K - 1;
      while (p > 0 && PCDB_LESS(d, idx, bd[p - 1], bi[p - 1])) {
        bd[p] = bd[p - 1];
        bi[p] = bi[p - 1];
        --p;
      }
#undef PCDB_LESS
      bd[p] = d;
      bi[p] = idx;
    }
  int use = min(cnt, k);
  if (!by_row && use_ratio && k == 1 && cnt >= 2) {
    if (__fdiv_rn(bd[0], bd[1]) > ratio_thr) use = 0;
  }
  for (int j = 0; j < k; ++j) {
    idx_out[q * k + j] = j < use ? bi[j] : -1;
    dist_out[q * k + j] = j < use ? bd[j] : __int_as_float(0x7fc00000);
  }
  cnt_out[q] = by_row ? (int)n_total : use;
}

// Keypoint sharding is BLOCK-CYCLIC: blocks of KP_BLOCK consecutive keypoints (voxel order: spatial neighbours, so a
// block still shares its staged neighbourhood) go to the ranks round-robin.  A contiguous split hands one rank the dense
// slab of a scene (the table) and the others wait for it: measured 12 ms vs 5.8 ms of work on two GPUs.
constexpr int KP_BLOCK = 32;
__host__ __device__ inline long long kp_global_index(long long local, int rank, int n) {
  return ((local / KP_BLOCK) * n + rank) * KP_BLOCK + local % KP_BLOCK;
}
__global__ void k_slice_cyclic(const float4* __restrict__ kp4, const int* __restrict__ kp_cloud, long long n_local,
                               int rank, int n, float4* kp4_out, int* kp_cloud_out) {
  long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= n_local) return;
  const long long g = kp_global_index(l, rank, n);
  kp4_out[l] = kp4[g];
  kp_cloud_out[l] = kp_cloud[g];
}
// sort key of a vote = global index of the keypoint that cast it (votes of one rank are already ascending in it)
__global__ void k_vote_keys(const int* __restrict__ vote_feat, const int* __restrict__ feat_kp, long long V, int rank,
                            int n, unsigned* key) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= V) return;
  key[v] = (unsigned)kp_global_index(feat_kp[vote_feat[v]], rank, n);
}
__global__ void k_iota(int* a, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int)i;
}
__global__ void k_permute_votes(const pcdb_vote* __restrict__ in, const int* __restrict__ order, long long V,
                                pcdb_vote* out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long v = t / 5;
  if (v >= V) return;
  const int part = (int)(t % 5);  // an 80-byte record = five 16-byte pieces
  reinterpret_cast<float4*>(out + v)[part] = reinterpret_cast<const float4*>(in + order[v])[part];
}

}  // namespace

bool comm_active(const pcdb_ctx* ctx) { return ctx->comm_state != nullptr; }
bool comm_codebook_sharded(const pcdb_ctx* ctx) { return ctx->comm_state && ctx->cb.N_table != ctx->cb.N; }
bool comm_keypoints_sharded(const pcdb_ctx* ctx) {
  return ctx->comm_state && static_cast<const CommState*>(ctx->comm_state)->shard_keypoints;
}

// ActivationStrategyKNN::activateKNN for this rank's queries against the row-sharded codebook (collective).
// Results in ws.knn_idx / knn_dist / knn_cnt, Q_local x k, as the unsharded search leaves them.
int stage_knn_sharded(pcdb_ctx* ctx, const float* queries_d, int64_t Q_local, int k, int dist_type, int mode,
                      bool use_ratio, float ratio_thr) {
  CommState* c = comm_of(ctx);
  NcclApi* a = nccl_api();
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const Codebook_d& cb = ctx->cb;
  const int D = cb.D;
  const int n = c->n;
  cudaEvent_t e0 = ctx->ev_comm[0], e1 = ctx->ev_comm[1], e2 = ctx->ev_comm[2], e3 = ctx->ev_comm[3];
  PCDB_CUDA(cudaEventRecord(e0, st));
  PCDB_TRY(gather_counts(ctx, Q_local, c->counts));
  int64_t Qtot = 0, my_off = 0;
  for (int r = 0; r < n; ++r) {
    if (r == c->rank) my_off = Qtot;
    Qtot += c->counts[r];
  }
  if (Qtot > 0x7fffff00ll / std::max(1, k + 1)) return ctx->fail(PCDB_E_INVALID, "too many queries in one sharded batch");
  PCDB_CUDA(w.knn_idx.ensure(sizeof(int) * (Q_local * k + 1)));
  PCDB_CUDA(w.knn_dist.ensure(sizeof(float) * (Q_local * k + 1)));
  PCDB_CUDA(w.knn_cnt.ensure(sizeof(int) * (Q_local + 1)));
  if (Qtot == 0) return PCDB_OK;
  // every rank sees every query
  PCDB_CUDA(c->q_all.ensure(sizeof(float) * (size_t)Qtot * D + 16));
  PCDB_TRY(all_gather_v(ctx, queries_d, c->q_all.p, c->counts, sizeof(float) * (size_t)D));
  PCDB_CUDA(cudaEventRecord(e1, st));
  // local search over this rank's rows: K best per query, no ratio test (it needs the GLOBAL second neighbour)
  const int K = use_ratio ? k + 1 : k;
  PCDB_TRY(pcdb_run_knn(ctx, c->q_all.as<float>(), Qtot, K, dist_type, mode, false, 0.f));
  PCDB_CUDA(c->send.ensure(sizeof(int2) * (size_t)Qtot * K + 16));
  PCDB_CUDA(c->recv.ensure(sizeof(int2) * (size_t)n * std::max<int64_t>(Q_local, 1) * K + 16));
  k_pack_pairs<<<cdiv(Qtot * K, 256), 256, 0, st>>>(w.knn_idx.as<int>(), w.knn_dist.as<float>(), w.knn_cnt.as<int>(),
                                                    Qtot, K, c->send.as<int2>());
  PCDB_LAUNCH_CHECK();
  PCDB_CUDA(cudaEventRecord(e2, st));
  // the exchange step: rank p receives, from every rank, the partial lists of ITS queries
  PCDB_NCCL(a->GroupStart());
  {
    int64_t off = 0;
    for (int p = 0; p < n; ++p) {
      const size_t sbytes = sizeof(int2) * (size_t)c->counts[p] * K;
      const size_t rbytes = sizeof(int2) * (size_t)Q_local * K;
      ncclResult_t rc = ncclSuccess;
      if (sbytes) rc = a->Send(c->send.as<int2>() + off * K, sbytes, ncclInt8, p, c->comm, st);
      if (rc == ncclSuccess && rbytes)
        rc = a->Recv(c->recv.as<int2>() + (size_t)p * Q_local * K, rbytes, ncclInt8, p, c->comm, st);
      if (rc != ncclSuccess) {
        a->GroupEnd();
        return ctx->fail(PCDB_E_COMM, "ncclSend/ncclRecv failed: %s", a->GetErrorString(rc));
      }
      off += c->counts[p];
    }
  }
  PCDB_NCCL(a->GroupEnd());
  ctx->stats.comm_bytes += (int64_t)(sizeof(int2) * (size_t)(n - 1) * Q_local * K);
  if (Q_local > 0) {
    const int by_row = cb.N_table <= k ? 1 : 0;  // activateKNN returns every codeword, in id order, when N <= k
    k_merge_pairs<<<cdiv(Q_local, 128), 128, 0, st>>>(c->recv.as<int2>(), n, Q_local, K, k, use_ratio ? 1 : 0,
                                                      ratio_thr, by_row, cb.N_table, w.knn_idx.as<int>(),
                                                      w.knn_dist.as<float>(), w.knn_cnt.as<int>());
    PCDB_LAUNCH_CHECK();
  }
  PCDB_CUDA(cudaEventRecord(e3, st));
  ctx->comm_events_valid = true;
  (void)my_off;
  return PCDB_OK;
}

// keypoint-sharded scene: keep this rank's blocks of the Q keypoints (kp4 / kp_cloud / kp_off), B == 1
int stage_slice_keypoints(pcdb_ctx* ctx, int B, int64_t Q, int64_t* Q_local_out) {
  CommState* c = comm_of(ctx);
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  if (B != 1) return ctx->fail(PCDB_E_UNSUPPORTED, "keypoint sharding handles one scene per call");
  const int64_t full_blocks = Q / KP_BLOCK, tail = Q % KP_BLOCK;
  int64_t n = (full_blocks / c->n) * KP_BLOCK;                 // whole rounds
  const int64_t left = full_blocks % c->n;                     // blocks of the last, incomplete round
  if (c->rank < left) n += KP_BLOCK;
  if (c->rank == left) n += tail;                              // the partial block follows the last full one
  if (n > 0) {
    PCDB_CUDA(w.kp_in.ensure(sizeof(float4) * (size_t)(n + 1)));
    PCDB_CUDA(w.feat_valid.ensure(sizeof(int) * (size_t)(n + 2)));
    k_slice_cyclic<<<cdiv(n, 256), 256, 0, st>>>(w.kp4.as<float4>(), w.kp_cloud.as<int>(), n, c->rank, c->n,
                                                 w.kp_in.as<float4>(), w.feat_valid.as<int>());
    PCDB_LAUNCH_CHECK();
    PCDB_CUDA(cudaMemcpyAsync(w.kp4.p, w.kp_in.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(w.kp_cloud.p, w.feat_valid.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
  }
  const long long off[2] = {0, (long long)n};
  PCDB_CUDA(cudaMemcpyAsync(w.kp_off.p, off, sizeof(off), cudaMemcpyHostToDevice, st));
  *Q_local_out = n;
  return PCDB_OK;
}

// keypoint-sharded scene: every rank ends up with the votes of all ranks in GLOBAL keypoint order — the order one GPU
// produces — by gathering (vote, keypoint index) and a stable radix sort on the index
int stage_gather_votes(pcdb_ctx* ctx, int B, int64_t F_local, int k, int64_t V_local, int64_t* V_out) {
  CommState* c = comm_of(ctx);
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  (void)F_local;
  (void)k;
  if (B != 1) return ctx->fail(PCDB_E_UNSUPPORTED, "keypoint sharding handles one scene per call");
  PCDB_CUDA(cudaEventRecord(ctx->ev_comm[2], st));
  PCDB_CUDA(c->key_local.ensure(sizeof(unsigned) * (size_t)(V_local + 1)));
  if (V_local > 0) {
    k_vote_keys<<<cdiv(V_local, 256), 256, 0, st>>>(w.vote_feat.as<int>(), w.feat_kp.as<int>(), V_local, c->rank, c->n,
                                                    c->key_local.as<unsigned>());
    PCDB_LAUNCH_CHECK();
  }
  PCDB_TRY(gather_counts(ctx, V_local, c->counts));
  int64_t V = 0;
  for (int r = 0; r < c->n; ++r) V += c->counts[r];
  if (V > 0x7fffff00ll) return ctx->fail(PCDB_E_INVALID, "more than 2^31 votes");
  PCDB_CUDA(c->votes_all.ensure(sizeof(pcdb_vote) * (size_t)(V + 1)));
  PCDB_CUDA(c->key_all.ensure(sizeof(unsigned) * (size_t)(V + 1)));
  PCDB_CUDA(c->key_sorted.ensure(sizeof(unsigned) * (size_t)(V + 1)));
  PCDB_CUDA(c->ord.ensure(sizeof(int) * (size_t)(V + 1)));
  PCDB_CUDA(c->ord_sorted.ensure(sizeof(int) * (size_t)(V + 1)));
  PCDB_TRY(all_gather_v(ctx, w.votes.p, c->votes_all.p, c->counts, sizeof(pcdb_vote)));
  PCDB_TRY(all_gather_v(ctx, c->key_local.p, c->key_all.p, c->counts, sizeof(unsigned)));
  PCDB_CUDA(w.votes.ensure(sizeof(pcdb_vote) * (size_t)(V + 1)));
  if (V > 0) {
    k_iota<<<cdiv(V, 256), 256, 0, st>>>(c->ord.as<int>(), V);
    PCDB_LAUNCH_CHECK();
    PCDB_TRY(pcdb_cub_sort_pairs_u32(ctx, c->key_all.as<unsigned>(), c->key_sorted.as<unsigned>(), c->ord.as<int>(),
                                     c->ord_sorted.as<int>(), V, 32));
    k_permute_votes<<<cdiv(V * 5, 256), 256, 0, st>>>(c->votes_all.as<pcdb_vote>(), c->ord_sorted.as<int>(), V,
                                                      w.votes.as<pcdb_vote>());
    PCDB_LAUNCH_CHECK();
  }
  const long long off[2] = {0, (long long)V};
  PCDB_CUDA(w.vote_off.ensure(sizeof(long long) * 2));
  PCDB_CUDA(cudaMemcpyAsync(w.vote_off.p, off, sizeof(off), cudaMemcpyHostToDevice, st));
  PCDB_TRY(stage_votes_unpack(ctx, B, V));
  PCDB_CUDA(cudaEventRecord(ctx->ev_comm[3], st));
  ctx->comm_events_valid = true;
  *V_out = V;
  return PCDB_OK;
}

float comm_last_exchange_ms(pcdb_ctx* ctx) {
  if (!ctx->comm_state || !ctx->comm_events_valid) return 0.f;
  float a = 0.f, b = 0.f;
  if (comm_codebook_sharded(ctx)) cudaEventElapsedTime(&a, ctx->ev_comm[0], ctx->ev_comm[1]);  // query all-gather
  cudaEventElapsedTime(&b, ctx->ev_comm[2], ctx->ev_comm[3]);                                   // top-k exchange + merge / vote gather
  return a + b;
}

extern "C" {

int pcdb_comm_unique_id(void* id_out, int32_t bytes) {
  NcclApi* a = nccl_api();
  if (!a->lib || !id_out || bytes < (int32_t)sizeof(ncclUniqueId)) return PCDB_E_COMM;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return PCDB_E_COMM;
  std::memset(id_out, 0, (size_t)bytes);
  std::memcpy(id_out, &id, sizeof(id));
  return PCDB_OK;
}

int pcdb_comm_init(pcdb_ctx* ctx, int32_t rank, int32_t n_ranks, const void* unique_id) {
  if (!ctx) return PCDB_E_INVALID;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !unique_id) return ctx->fail(PCDB_E_INVALID, "bad communicator arguments");
  if (ctx->comm_state) return ctx->fail(PCDB_E_STATE, "communicator already initialised");
  NcclApi* a = nccl_api();
  if (!a->lib) return ctx->fail(PCDB_E_COMM, "NCCL unavailable: %s", a->err.c_str());
  PCDB_CUDA(cudaSetDevice(ctx->device));
  CommState* c = new CommState();
  c->rank = rank;
  c->n = n_ranks;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = a->CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return ctx->fail(PCDB_E_COMM, "ncclCommInitRank failed: %s", a->GetErrorString(r));
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev_comm[i]);
  ctx->comm_state = c;
  ctx->comm_state_free = [](void* p) {
    CommState* s = static_cast<CommState*>(p);
    if (s->comm) nccl_api()->CommDestroy(s->comm);
    delete s;
  };
  return PCDB_OK;
}

int pcdb_comm_destroy(pcdb_ctx* ctx) {
  if (!ctx) return PCDB_E_INVALID;
  if (!ctx->comm_state) return PCDB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->comm_state_free(ctx->comm_state);
  ctx->comm_state = nullptr;
  for (int i = 0; i < 4; ++i)
    if (ctx->ev_comm[i]) {
      cudaEventDestroy(ctx->ev_comm[i]);
      ctx->ev_comm[i] = nullptr;
    }
  return PCDB_OK;
}

int pcdb_comm_info(pcdb_ctx* ctx, int32_t* rank_out, int32_t* n_ranks_out, int32_t* nccl_version_out) {
  if (!ctx) return PCDB_E_INVALID;
  CommState* c = comm_of(ctx);
  if (rank_out) *rank_out = c ? c->rank : 0;
  if (n_ranks_out) *n_ranks_out = c ? c->n : 1;
  if (nccl_version_out) {
    int v = 0;
    NcclApi* a = nccl_api();
    if (a->lib) a->GetVersion(&v);
    *nccl_version_out = v;
  }
  return PCDB_OK;
}

int pcdb_comm_shard_keypoints(pcdb_ctx* ctx, int32_t enable) {
  if (!ctx) return PCDB_E_INVALID;
  if (!ctx->comm_state) return ctx->fail(PCDB_E_STATE, "pcdb_comm_shard_keypoints before pcdb_comm_init");
  comm_of(ctx)->shard_keypoints = enable != 0;
  return PCDB_OK;
}

}  // extern "C"
