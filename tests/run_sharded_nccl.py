#!/usr/bin/env python
"""Multi-GPU exchange paths of libpcdb200 over real NCCL, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/run_sharded_nccl.py [--words 0] [--out result.json]

Section 1 — row-sharded codebook (config C4): every rank uploads descriptor rows [lo, hi) of a CSHOT-1344 codebook plus
the complete vote tables (pcdb_set_codebook_sharded) and classifies ITS OWN test clouds; the query all-gather, the
top-k all-to-all and the merge all happen inside the library on the device.  Checked on every rank against an
unsharded context on the same GPU: identical rows, bit-identical distances (k = 2, the k = 1 distance-ratio test,
Euclidean and ChiSquared), byte-identical vote lists, maxima and labels.
Section 2 — keypoint-sharded scene (config C5): all ranks pass the same cluttered scene, the votes are all-gathered
inside the library; votes and maxima must equal the single-GPU fused path byte for byte.

torch.distributed (gloo) is used only to hand out the NCCL unique id and to reduce the pass/fail flags.
--words 0 builds the full 1.07 M x 1344 codebook of the C4 bench (about a minute of set-up per rank)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))


def shard_bounds(n_rows, world):
    base, rem = divmod(n_rows, world)
    return [r * base + min(r, rem) for r in range(world + 1)]


def new_comm(ctx, rank, world, dist, api):
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--words", type=int, default=20000, help="codebook size (0 = the C4 bench's 1.07 M words)")
    ap.add_argument("--clouds", type=int, default=8, help="test clouds per rank")
    ap.add_argument("--scene-points", type=int, default=4096, help="points per scene object")
    ap.add_argument("--chi-scan-check", type=int, default=1,
                    help="compare the sharded ChiSquared search with the unsharded exact SCAN on this many clouds' queries")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from pcdb200 import api, synth
    from pcdb200.structs import DIST_CHISQUARED, DIST_EUCLIDEAN, KNN_SCAN

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")

    full = api.Context(device=local)
    t0 = time.time()
    wl, prm, cb = bench.build_world("c4", args.words if args.words > 0 else 0, full, rank_log=(rank == 0))
    setup_s = time.time() - t0
    bounds = shard_bounds(cb.N, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    sh = api.Context(prm, device=local)
    new_comm(sh, rank, world, dist, api)
    sh.set_codebook_sharded(cb, lo, hi)
    info = sh.comm_info()

    # this rank's own test clouds (different on every rank, different counts too: ragged all-gather)
    n_te = args.clouds + (rank % 2)
    te = [(rank * 5 + c) % wl["n_classes"] for c in range(n_te)]
    xt, nt, ct, ot = synth.make_clouds(te, [9000 + rank * 100 + i for i in range(n_te)], wl["P"], scale=wl["scale"],
                                       jitter=0.002)
    tx, tl, td, toff = full.compute_features(xt, nt, ct, ot)
    res = {"world": world, "codewords": int(cb.N), "D": int(cb.D), "rows_this_rank": [int(lo), int(hi)],
           "queries_rank0": int(td.shape[0]), "nccl_version": info["nccl_version"], "setup_seconds": setup_s}
    ok = {}

    def same(a, b):
        return bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
                    and np.array_equal(a[2], b[2]))

    # ---- 1a. activation: k = 2, Euclidean
    ok["knn_k2_l2"] = same(sh.knn(td, k=2, dist_type=DIST_EUCLIDEAN), full.knn(td, k=2, dist_type=DIST_EUCLIDEAN))
    # ---- 1b. the distance-ratio test needs the GLOBAL second neighbour
    rp = prm.copy()
    rp.use_distance_ratio, rp.distance_ratio_threshold = 1, 0.8
    sh.set_params(rp), full.set_params(rp)
    a, b = sh.knn(td, k=1, dist_type=DIST_EUCLIDEAN), full.knn(td, k=1, dist_type=DIST_EUCLIDEAN)
    ok["knn_ratio_l2"] = same(a, b) and bool((b[2] == 0).any()) and bool((b[2] == 1).any())
    sh.set_params(prm), full.set_params(prm)
    # ---- 1c. ChiSquared (tensor-core sandwich on every shard) vs the unsharded sandwich and vs the unsharded exact scan
    a = sh.knn(td, k=2, dist_type=DIST_CHISQUARED)
    ok["knn_k2_chi2"] = same(a, full.knn(td, k=2, dist_type=DIST_CHISQUARED))
    nq = int(toff[min(args.chi_scan_check, n_te)])
    # every rank must take part in the collective even when it checks fewer queries: run the scan on the full context only
    s_idx, s_dst, s_cnt = full.knn(td[:nq], k=2, dist_type=DIST_CHISQUARED, mode=KNN_SCAN)
    ok["knn_k2_chi2_vs_scan"] = same((a[0][:nq], a[1][:nq], a[2][:nq]), (s_idx, s_dst, s_cnt))

    # ---- 1d. the fused path: this rank's clouds through the sharded codebook
    torch.cuda.synchronize()
    dist.barrier()
    sh.reset_stats()
    t1 = time.perf_counter()
    labels, mx, moff = sh.classify_batch(xt, nt, ct, ot)
    dt = time.perf_counter() - t1
    st = sh.stats()
    votes, voff = sh.get_votes(n_te, int(st["n_votes"]) + 16)
    rlabels, rmx, rmoff = full.classify_batch(xt, nt, ct, ot)
    rvotes, rvoff = full.get_votes(n_te, len(votes) + 16)
    ok["labels"] = bool(np.array_equal(labels, rlabels))
    ok["maxima"] = bool(np.array_equal(moff, rmoff) and mx.tobytes() == rmx.tobytes())
    ok["votes"] = bool(np.array_equal(voff, rvoff) and votes.tobytes() == rvotes.tobytes() and len(votes) > 0)
    res.update({"sharded_classify_seconds_rank0": dt, "exchange_ms_rank0": st["comm_ms"],
                "nvlink_bytes_received_rank0": int(st["comm_bytes"]), "labels": labels.tolist(), "truth": te,
                "knn_fallback_queries": int(st["knn_fallback_queries"])})
    # an empty batch on one rank must not break the collective
    e = np.zeros((0, 3), np.float32)
    if rank == world - 1:
        l0, _, _ = sh.classify_batch(e, e, np.zeros(0, np.uint32), np.array([0], np.int64))
        ok["empty_rank"] = len(l0) == 0
    else:
        l1, _, _ = sh.classify_batch(xt, nt, ct, ot)
        ok["empty_rank"] = bool(np.array_equal(l1, rlabels))

    # ---- 2. one scene, keypoints sharded over the ranks, votes all-gathered in the library
    sprm = synth.workload_params("c4", knn_k=1, single_object_mode=0, min_votes_threshold=3)
    full.set_params(sprm)
    sx_, sn_, sc_, _ = synth.make_scene([0, 1, 2, 3, 4, 5], 21, args.scene_points, scale=wl["scale"],
                                        plane_points=30000, clutter_points=5000)
    soff_ = np.array([0, len(sx_)], np.int64)
    _, rmx2, rmoff2 = full.classify_batch(sx_, sn_, sc_, soff_)
    rv2, rvoff2 = full.get_votes(1, int(1e6))
    new_comm(full, rank, world, dist, api)
    full.comm_shard_keypoints(True)
    torch.cuda.synchronize()
    dist.barrier()
    full.reset_stats()
    t2 = time.perf_counter()
    _, smx, smoff = full.classify_batch(sx_, sn_, sc_, soff_)
    dt_scene = time.perf_counter() - t2
    sst = full.stats()
    sv, svoff = full.get_votes(1, len(rv2) + 16)
    ok["scene"] = bool(np.array_equal(svoff, rvoff2) and sv.tobytes() == rv2.tobytes() and len(sv) > 0
                       and np.array_equal(smoff, rmoff2) and smx.tobytes() == rmx2.tobytes())
    full.comm_shard_keypoints(False)
    res.update({"scene_points": int(len(sx_)), "scene_votes": int(len(sv)), "scene_maxima": int(smoff[1]),
                "scene_seconds_rank0": dt_scene, "scene_keypoints_this_rank": int(sst["n_keypoints"]),
                "scene_vote_gather_ms_rank0": sst["comm_ms"]})

    names = sorted(ok)
    flags = torch.tensor([int(bool(ok[n])) for n in names])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["identical_to_unsharded"] = {n: bool(f) for n, f in zip(names, flags.tolist())}
    res["this_rank"] = {n: bool(ok[n]) for n in names}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f)
    elif not all(ok.values()):
        print("rank %d: %s" % (rank, {n: bool(ok[n]) for n in names}), file=sys.stderr, flush=True)
    dist.barrier()
    sh.close()
    full.close()
    dist.destroy_process_group()
    return 0 if all(res["identical_to_unsharded"].values()) else 1


if __name__ == "__main__":
    sys.exit(main())
