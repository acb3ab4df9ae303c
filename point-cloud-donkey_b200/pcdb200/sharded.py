"""Host-side plumbing for the multi-GPU modes of libpcdb200.

The exchange steps themselves — query all-gather, per-shard tcgen05 search, top-k all-to-all, merge; vote all-gather for
a keypoint-sharded scene — run INSIDE the library on the device (csrc/comm.cu, NCCL over NVLink).  What a host program
has to do is small: decide the row shards and hand the NCCL unique id from rank 0 to the other ranks.  torch.distributed
(any backend) is used for that hand-off here; MPI or a file would do as well."""
import numpy as np

from . import api
from .structs import Codebook


def shard_bounds(n_rows, world):
    """Contiguous row shards, sizes differing by at most one: rank r holds rows [b[r], b[r+1])."""
    base, rem = divmod(n_rows, world)
    return [r * base + min(r, rem) for r in range(world + 1)]


def interleave_codebook(cb, world):
    """The same codebook with its rows dealt cyclically over `world` contiguous shards (row r of `cb` lands in shard
    r % world), vote table and per-row arrays carried along.  Training appends codewords cloud by cloud, i.e. class by
    class (implicit_shape_model.cpp:447-475), so a contiguous shard of the untouched table holds a few classes only: a
    query of another class has no near word on that shard, its bound stays loose, the pools of the PCA pre-filter
    overflow there and the shard falls back to the plain sweep (measured on 2 GPUs at C4: 99 ms per step against 41 ms
    for the replicated table).  Dealt shards are statistically alike.  Activation, votes and labels of the permuted
    table equal those of `cb` (codeword ids, weights and votes move with their rows); only the winner among codewords at
    EXACTLY equal distance can change, since ties go to the lower row.  Returns (codebook, perm) with
    new row i = old row perm[i]."""
    n = cb.N
    perm = np.argsort(np.arange(n) % max(1, world), kind="stable")
    cnt = np.diff(cb.vote_off)[perm]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    # votes of new row i: the block of old row perm[i], in stored order
    vidx = np.repeat(cb.vote_off[:-1][perm] - off[:-1], cnt) + np.arange(int(off[-1]))
    out = Codebook(cb.words[perm], off, cb.vote_xyz[vidx], cb.vote_weight[vidx], cb.vote_class[vidx],
                   cb.vote_instance[vidx], cb.vote_bbox[vidx],
                   cb.vote_class_weight[vidx] if cb.vote_class_weight is not None and len(cb.vote_class_weight) else cb.vote_class_weight,
                   cb.kp_train[perm], cb.codeword_ids[perm] if cb.codeword_ids is not None and len(cb.codeword_ids) else cb.codeword_ids,
                   cb.sigma2, None if cb.codeword_weight is None else cb.codeword_weight[perm])
    return out, perm


def init_comm(ctx, dist):
    """Collective over an initialised torch.distributed group: creates ctx's NCCL communicator (pcdb_comm_init)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])
    return rank, world


def shard_codebook(ctx, cb, dist):
    """Row-shards `cb` over the ranks of ctx's communicator (descriptor rows sharded, vote tables replicated)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    b = shard_bounds(cb.N, world)
    ctx.set_codebook_sharded(cb, b[rank], b[rank + 1])
    return b[rank], b[rank + 1]
