"""Row-sharded codebook: every rank holds rows [lo, hi) of the codebook (uploaded with
row_base = lo so row ids are global), all ranks see the same query batch.  One exchange step: all-gather of the
per-shard top-k (distance, row) lists, then a per-query merge on the device (pcdb_merge_topk; ties -> lower global
row).  The owner of each winning row casts its votes; votes are all-gathered for the maxima search.

torch.distributed is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

from .structs import VOTE_DTYPE


def shard_bounds(n_rows, world):
    """Contiguous row shards, sizes differing by at most one."""
    base, rem = divmod(n_rows, world)
    lo = [r * base + min(r, rem) for r in range(world + 1)]
    return lo


def _all_gather_np(arr, device):
    """All-gather equally shaped numpy arrays -> list of numpy arrays (rank order)."""
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def _all_gather_var(arr, device):
    """All-gather of 1-D byte buffers of different lengths (pad to the maximum)."""
    n = torch.tensor([arr.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = np.zeros(m, np.uint8)
    buf[: arr.shape[0]] = arr
    parts = _all_gather_np(buf, device)
    return [p[:s] for p, s in zip(parts, sizes)]


def sharded_knn(ctx, queries, k, dist_type, use_ratio=False, ratio_thr=0.95, device="cpu", mode=0):
    """ActivationStrategyKNN::activateKNN against a row-sharded codebook.  `ctx` must run WITHOUT the distance-ratio
    test (it needs the global second neighbour): it is applied here after the merge.  Returns (idx, dist, count)."""
    kk = k + 1 if use_ratio else k
    idx, dst, _ = ctx.knn(queries, k=kk, dist_type=dist_type, mode=mode)
    all_idx = np.stack(_all_gather_np(idx.astype(np.int32), device))
    all_dst = np.stack(_all_gather_np(np.nan_to_num(dst.astype(np.float32), nan=np.inf), device))
    midx, mdst = ctx.merge_topk(all_idx, all_dst)
    cnt = (midx[:, :k] >= 0).sum(1).astype(np.int32)
    if use_ratio and k == 1:
        ok = midx[:, 1] >= 0
        cnt = np.where(ok & (mdst[:, 0] / np.where(ok, mdst[:, 1], 1) > np.float32(ratio_thr)), 0, cnt).astype(np.int32)
    return midx[:, :k].copy(), mdst[:, :k].copy(), cnt


def sharded_cast_votes(ctx, row_lo, row_hi, feat_xyz, feat_lrf, feat_off, idx, dst, cnt, device="cpu"):
    """Each rank casts the votes of the winning rows it owns; all ranks receive all votes (rank-major order inside a
    cloud — the reference's vote order is arbitrary too, voting/voting.cpp:73-76)."""
    own = (idx >= row_lo) & (idx < row_hi)
    masked = np.where(own, idx, -1).astype(np.int32)
    votes, voff = ctx.cast_votes(feat_xyz, feat_lrf, feat_off, masked, dst, cnt)
    B = len(feat_off) - 1
    parts = _all_gather_var(votes.view(np.uint8).reshape(-1), device)
    offs = _all_gather_np(voff.astype(np.int64), device)
    per_rank = [p.view(VOTE_DTYPE) for p in parts]
    out, out_off = [], [0]
    for b in range(B):
        for r, v in enumerate(per_rank):
            out.append(v[offs[r][b]:offs[r][b + 1]])
        out_off.append(out_off[-1] + sum(int(offs[r][b + 1] - offs[r][b]) for r in range(len(per_rank))))
    votes_all = np.concatenate(out) if out else np.zeros(0, VOTE_DTYPE)
    return votes_all, np.asarray(out_off, np.int64)


def sharded_scene_votes(ctx, prm, xyz, normals, rgb, device="cpu"):
    """One large scene on several GPUs: the cloud and the codebook are replicated, the
    KEYPOINTS are sharded contiguously over the ranks.  Every rank computes the voxel-grid keypoints of the whole scene
    (cheap, identical everywhere), the reference frames / descriptors / activation / votes of its own keypoint slice, and
    the votes are all-gathered in rank order — which is keypoint order, so every rank ends up with exactly the vote list
    a single GPU produces.  Returns (votes, vote_off) for one cloud; feed them to ctx.find_maxima."""
    rank, world = dist.get_rank(), dist.get_world_size()
    xyz = np.ascontiguousarray(xyz, np.float32)
    finite = np.isfinite(xyz).all(1)                       # removeNaNFromPointCloud (implicit_shape_model.cpp:611)
    pts = xyz[finite]
    col = (np.zeros(len(xyz), np.uint32) if rgb is None else np.asarray(rgb, np.uint32))[finite]
    nrm = np.ascontiguousarray(normals, np.float32)[finite]
    kp, kr, _ = ctx.voxel_keypoints(pts, col, [0, len(pts)], prm.leaf_size)
    cuts = shard_bounds(len(kp), world)
    lo, hi = cuts[rank], cuts[rank + 1]
    surf = np.isfinite(nrm).all(1)                         # filterNormals (:1040-1068)
    sx, sn, sc = pts[surf], nrm[surf], col[surf]
    soff = [0, len(sx)]
    kx, kc = kp[lo:hi], kr[lo:hi]
    lrf = ctx.shot_lrf(sx, soff, kx, [0, len(kx)], prm.lrf_radius)
    ok = np.isfinite(lrf[:, 0]) & np.isfinite(lrf[:, 3]) & np.isfinite(lrf[:, 6])      # features.cpp:64-76
    kx, kc, lrf = kx[ok], kc[ok], lrf[ok]
    desc = ctx.shot_describe(prm.feature_type, sx, sn, sc, soff, kx, kc, lrf, [0, len(kx)], prm.feature_radius)
    ok = ~np.isnan(desc).any(1)                            # removeNaNFeatures (:1276-1308)
    kx, lrf, desc = kx[ok], lrf[ok], desc[ok]
    if len(kx):
        idx, dst, cnt = ctx.knn(desc, k=prm.knn_k, dist_type=prm.distance_type)
        votes, _ = ctx.cast_votes(kx, lrf, [0, len(kx)], idx, dst, cnt)
    else:
        votes = np.zeros(0, VOTE_DTYPE)
    parts = _all_gather_var(votes.view(np.uint8).reshape(-1), device)
    allv = np.concatenate([p.view(VOTE_DTYPE) for p in parts]) if parts else np.zeros(0, VOTE_DTYPE)
    return allv, np.array([0, len(allv)], np.int64)
