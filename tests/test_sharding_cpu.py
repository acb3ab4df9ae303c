"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: test-set sharding needs no exchange; the row-sharded
codebook path (all-gather of per-shard top-k + merge, owner casts votes) must reproduce the unsharded result.
The GPU calls are replaced by the oracle (test infrastructure) — the collective logic under test is the product's."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardCtx:
    """One codebook shard served by the oracle, with global row ids (stands in for api.Context on CPU)."""

    def __init__(self, orc, prm, cb_full, lo, hi):
        self.orc, self.lo, self.hi = orc, lo, hi
        self.cb = cb_full.rows(lo, hi)
        self.m = orc.Model(prm, self.cb)

    def knn(self, q, k=None, dist_type=None, mode=0):
        idx, d, c = self.m.knn(q, k=k, dist_type=dist_type)
        return np.where(idx >= 0, idx + self.lo, -1).astype(np.int32), d, c

    def merge_topk(self, i, d):
        return self.orc.merge_topk(i, d)

    # stage entry points used by the keypoint-sharded scene path
    def voxel_keypoints(self, xyz, rgb, off, leaf):
        return self.orc.voxel_keypoints(xyz, rgb, off, leaf)

    def shot_lrf(self, sx, soff, kx, koff, radius):
        return self.orc.shot_lrf(sx, soff, kx, koff, radius)

    def shot_describe(self, ft, sx, sn, sc, soff, kx, kc, lrf, koff, radius):
        return self.orc.shot_describe(ft, sx, sn, sc, soff, kx, kc, lrf, koff, radius)

    def find_maxima(self, votes, voff):
        return self.m.find_maxima(votes, voff)

    def cast_votes(self, fx, fl, foff, idx, dst, cnt):
        # the oracle has no mask support: cast per (feature, j) for owned rows only
        local = np.where(idx >= 0, idx - self.lo, -1).astype(np.int32)
        votes, off = [], [0]
        from pcdb200.structs import VOTE_DTYPE
        B = len(foff) - 1
        for b in range(B):
            for f in range(int(foff[b]), int(foff[b + 1])):
                for j in range(int(cnt[f])):
                    if local[f, j] < 0:
                        continue
                    v, _ = self.m.cast_votes(fx[f:f + 1], fl[f:f + 1], [0, 1], local[f:f + 1, j:j + 1], dst[f:f + 1, j:j + 1],
                                             np.ones(1, np.int32))
                    votes.append(v)
            off.append(sum(len(v) for v in votes))
        return (np.concatenate(votes) if votes else np.zeros(0, VOTE_DTYPE)), np.asarray(off, np.int64)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    import torch.distributed as dist
    from oracle import oracle_py as orc
    from pcdb200 import sharded, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm = synth.workload_params("c2", knn_k=2)
        tr_cls = [0, 0, 1, 1, 2]
        xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [300 + i for i in range(5)], 900)
        fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
        bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(5)])
        cb = orc.train(prm, fx, fl, fd, foff, tr_cls, list(range(5)), bb, 3)
        xt, nt, rt, ot = synth.make_clouds([0, 1, 2], [700, 701, 702], 900)
        tx, tl, td, toff = orc.compute_features(prm, xt, nt, rt, ot)
        bounds = sharded.shard_bounds(cb.N, world)
        ctx = OracleShardCtx(orc, prm, cb, bounds[rank], bounds[rank + 1])
        idx, dst, cnt = sharded.sharded_knn(ctx, td, 2, prm.distance_type)
        full = orc.Model(prm, cb)
        ridx, rdst, rcnt = full.knn(td, k=2)
        ok = np.array_equal(idx, ridx) and np.array_equal(dst.view(np.uint32), rdst.view(np.uint32)) and np.array_equal(cnt, rcnt)
        votes, voff = sharded.sharded_cast_votes(ctx, bounds[rank], bounds[rank + 1], tx, tl, toff, idx, dst, cnt)
        rvotes, rvoff = full.cast_votes(tx, tl, toff, ridx, rdst, rcnt)
        ok = ok and np.array_equal(voff, rvoff)
        # same multiset of votes per cloud (order inside a cloud is rank-major)
        for b in range(len(toff) - 1):
            a = np.sort(votes[voff[b]:voff[b + 1]].view(np.uint8).reshape(-1, 80), axis=0)
            r = np.sort(rvotes[rvoff[b]:rvoff[b + 1]].view(np.uint8).reshape(-1, 80), axis=0)
            ok = ok and np.array_equal(a, r)
        mx, moff, _, _ = full.find_maxima(votes, voff)
        rmx, rmoff, _, _ = full.find_maxima(rvotes, rvoff)
        ok = ok and np.array_equal(moff, rmoff) and np.array_equal(mx["class_id"], rmx["class_id"])
        # test-set sharding: ranks classify disjoint contiguous shards, labels concatenate to the unsharded answer
        lab_all, _, _ = full.classify_batch(xt, nt, rt, ot, want_maxima=False)
        B = len(ot) - 1
        cuts = sharded.shard_bounds(B, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        lab, _, _ = full.classify_batch(xt[ot[lo]:ot[hi]], nt[ot[lo]:ot[hi]], rt[ot[lo]:ot[hi]], ot[lo:hi + 1] - ot[lo],
                                        want_maxima=False) if hi > lo else (np.zeros(0, np.int32), None, None)
        ok = ok and np.array_equal(lab, lab_all[lo:hi])
        # one scene, keypoints sharded over the ranks (C5): the gathered votes are the single-rank votes, bit for bit
        sprm = synth.workload_params("c2", knn_k=2, single_object_mode=0, min_votes_threshold=3)
        sx_, sn_, sc_, _ = synth.make_scene([0, 1, 2], 5, 900, plane_points=1500, clutter_points=300)
        sx_[7] = np.nan
        whole = OracleShardCtx(orc, sprm, cb, 0, cb.N)
        sv, svoff = sharded.sharded_scene_votes(whole, sprm, sx_, sn_, sc_)
        fx1, fl1, fd1, fo1 = orc.compute_features(sprm, sx_, sn_, sc_, [0, len(sx_)])
        i1, d1, c1 = orc.Model(sprm, cb).knn(fd1, k=2)
        rv, rvo = orc.Model(sprm, cb).cast_votes(fx1, fl1, fo1, i1, d1, c1)
        ok = ok and np.array_equal(svoff, rvo) and sv.tobytes() == rv.tobytes() and len(sv) > 0
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world2_sharded_codebook_and_test_set():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds():
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    from pcdb200 import sharded
    assert sharded.shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert sharded.shard_bounds(3, 8)[-1] == 3 and sharded.shard_bounds(0, 2) == [0, 0, 0]
