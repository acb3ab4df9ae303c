#!/usr/bin/env python
"""Row-sharded codebook over real NCCL: run with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/run_sharded_nccl.py [--out result.json]
Every rank uploads its row shard of a CSHOT-1344 codebook (row_base = first global row), all ranks see the same
query batch; the exchange step is one all-gather of the per-shard top-k lists followed by pcdb_merge_topk on the
device, then the owners cast the votes and all ranks receive them.  Checked on every rank against the unsharded
codebook on the same GPU: identical rows, bit-identical distances, the same vote multiset, the same labels; on rank 0
also against the oracle.  A second section shards the KEYPOINTS of one cluttered scene over the ranks (C5) and checks
that the all-gathered votes and the maxima equal the single-GPU fused path bit for bit.  (The world-size-2 gloo twin of this script is tests/test_sharding_cpu.py.)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--train-per-class", type=int, default=6)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pcdb200 import api, sharded, synth, train

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    os.environ.pop("NCCL_DEBUG", None)  # any level >= VERSION prints a banner on stdout; keep it to the JSON line
    if os.environ.get("PCDB_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["PCDB_NCCL_DEBUG"]
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    wl = synth.WORKLOADS["c4"]
    prm = synth.workload_params("c4", knn_k=2)
    n_cls, P = 6, 4096
    full = api.Context(prm, device=local)
    tr_cls = [c for c in range(n_cls) for _ in range(args.train_per_class)]
    x, n, c, o = synth.make_clouds(tr_cls, [4000 + i for i in range(len(tr_cls))], P, scale=wl["scale"])
    fx, fl, fd, foff = full.compute_features(x, n, c, o)
    bbs = np.stack([train.aabb(x[o[i]:o[i + 1]]) for i in range(len(tr_cls))])
    cb = train.train_codebook(full, prm, fx, fl, fd, foff, tr_cls, list(range(len(tr_cls))), bbs, n_cls)
    full.set_codebook(cb)
    bounds = sharded.shard_bounds(cb.N, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    shard = api.Context(prm, cb.rows(lo, hi), device=local, row_base=lo)

    te = [c % n_cls for c in range(12)]
    xt, nt, ct, ot = synth.make_clouds(te, [9000 + i for i in range(len(te))], P, scale=wl["scale"])
    tx, tl, td, toff = full.compute_features(xt, nt, ct, ot)

    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    idx, dst, cnt = sharded.sharded_knn(shard, td, 2, prm.distance_type, device=dev)
    votes, voff = sharded.sharded_cast_votes(shard, lo, hi, tx, tl, toff, idx, dst, cnt, device=dev)
    mx, moff, _, _ = full.find_maxima(votes, voff)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    ridx, rdst, rcnt = full.knn(td, k=2)
    ok_knn = bool(np.array_equal(idx, ridx) and np.array_equal(dst.view(np.uint32), rdst.view(np.uint32))
                  and np.array_equal(cnt, rcnt))
    rvotes, rvoff = full.cast_votes(tx, tl, toff, ridx, rdst, rcnt)
    ok_votes = bool(np.array_equal(voff, rvoff))
    for b in range(len(toff) - 1):
        a = np.sort(votes[voff[b]:voff[b + 1]].view(np.uint8).reshape(-1, 80), axis=0)
        r = np.sort(rvotes[rvoff[b]:rvoff[b + 1]].view(np.uint8).reshape(-1, 80), axis=0)
        ok_votes = ok_votes and bool(np.array_equal(a, r))
    rmx, rmoff, _, _ = full.find_maxima(rvotes, rvoff)
    labels = np.array([mx["class_id"][moff[b]] if moff[b + 1] > moff[b] else -1 for b in range(len(te))])
    rlabels = np.array([rmx["class_id"][rmoff[b]] if rmoff[b + 1] > rmoff[b] else -1 for b in range(len(te))])
    ok_labels = bool(np.array_equal(labels, rlabels))
    ok_oracle = None
    if rank == 0:
        from oracle import oracle_py as orc
        orc.set_num_threads(os.cpu_count() or 1)
        m = orc.Model(prm, cb)
        oidx, odst, ocnt = m.knn(td[:64], k=2)
        ok_oracle = bool(np.array_equal(idx[:64], oidx) and np.array_equal(cnt[:64], ocnt))

    # C5: one scene, keypoints sharded over the ranks, votes all-gathered; must equal the single-GPU fused path
    sprm = synth.workload_params("c4", knn_k=1, single_object_mode=0, min_votes_threshold=3)
    full.set_params(sprm)
    sx_, sn_, sc_, truth = synth.make_scene([0, 1, 2, 3, 4, 5], 21, P, scale=wl["scale"], plane_points=30000,
                                            clutter_points=5000)
    soff_ = np.array([0, len(sx_)], np.int64)
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    sv, svoff = sharded.sharded_scene_votes(full, sprm, sx_, sn_, sc_, device=dev)
    smx, smoff, _, _ = full.find_maxima(sv, svoff)
    torch.cuda.synchronize()
    dt_scene = time.perf_counter() - t1
    _, rmx2, rmoff2 = full.classify_batch(sx_, sn_, sc_, soff_)
    rv2, rvoff2 = full.get_votes(1, len(sv) + 16)
    ok_scene = bool(np.array_equal(svoff, rvoff2) and sv.tobytes() == rv2.tobytes() and len(sv) > 0
                    and np.array_equal(smoff, rmoff2) and np.array_equal(smx["class_id"], rmx2["class_id"])
                    and np.array_equal(smx["weight"].view(np.uint32), rmx2["weight"].view(np.uint32)))
    full.set_params(prm)

    flags = torch.tensor([int(ok_knn), int(ok_votes), int(ok_labels), int(ok_scene)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res = {"world": world, "codewords": int(cb.N), "D": int(cb.D), "queries": int(td.shape[0]), "k": 2,
           "knn_identical": bool(flags[0].item()), "votes_identical": bool(flags[1].item()),
           "labels_identical": bool(flags[2].item()), "knn_vs_oracle_first64": ok_oracle,
           "scene_keypoint_sharding_identical_to_one_gpu": bool(flags[3].item()), "scene_points": int(len(sx_)),
           "scene_votes": int(len(sv)), "scene_maxima": int(smoff[1]), "scene_seconds_rank0": dt_scene,
           "labels": labels.tolist(), "truth": te, "sharded_path_seconds_rank0": dt}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()
    shard.close()
    full.close()
    ok = (res["knn_identical"] and res["votes_identical"] and res["labels_identical"] and ok_oracle in (None, True)
          and res["scene_keypoint_sharding_identical_to_one_gpu"])
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
