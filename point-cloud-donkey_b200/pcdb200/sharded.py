"""Host-side plumbing for the multi-GPU modes of libpcdb200.

The exchange steps themselves — query all-gather, per-shard tcgen05 search, top-k all-to-all, merge; vote all-gather for
a keypoint-sharded scene — run INSIDE the library on the device (csrc/comm.cu, NCCL over NVLink).  What a host program
has to do is small: decide the row shards and hand the NCCL unique id from rank 0 to the other ranks.  torch.distributed
(any backend) is used for that hand-off here; MPI or a file would do as well."""
from . import api


def shard_bounds(n_rows, world):
    """Contiguous row shards, sizes differing by at most one: rank r holds rows [b[r], b[r+1])."""
    base, rem = divmod(n_rows, world)
    return [r * base + min(r, rem) for r in range(world + 1)]


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
