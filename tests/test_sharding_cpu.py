"""World-size-2 gloo tests (CPU) of the multi-GPU exchange protocol.  csrc/comm.cu runs the protocol on the device over
NCCL (tests/run_sharded_nccl.py checks that on real GPUs); here the SAME steps — ragged query all-gather, per-shard
search of all queries for K = k (+1 with the ratio test), exchange of the (distance, global row) lists, per-query merge
with ties -> lower row, then activateKNN's N <= k and distance-ratio semantics — are restated on the host with the
oracle standing in for the per-shard search, and must reproduce the unsharded oracle bit for bit.  Test-set sharding
(the default multi-GPU mode) needs no exchange and is checked too."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gather_np(arr, dist, torch):
    """Ragged all-gather of equally typed numpy arrays (first axis ragged)."""
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, np.ascontiguousarray(arr))
    return objs


def protocol_knn(shard_model, row_lo, n_total, queries, k, dist_type, use_ratio, ratio_thr, dist, torch):
    """Host model of comm.cu:stage_knn_sharded for this rank's `queries`."""
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = _gather_np(queries, dist, torch)                       # every rank sees every query
    counts = [len(p) for p in parts]
    q_all = np.concatenate(parts)
    K = k + 1 if use_ratio else k
    idx, dst, cnt = shard_model.knn(q_all, k=K, dist_type=dist_type)  # no ratio test on a shard
    j = np.arange(K)[None, :]
    valid = (j < cnt[:, None]) & (idx >= 0)
    idx = np.where(valid, idx + row_lo, -1).astype(np.int32)       # global row ids
    dst = np.where(valid, dst, np.inf).astype(np.float32)
    lists = _gather_np(np.stack([idx.astype(np.int64), dst.view(np.uint32).astype(np.int64)], -1), dist, torch)
    o = sum(counts[:rank])
    mine = np.stack([p[o:o + counts[rank]] for p in lists])         # [world][Q_local][K][2] — what the all-to-all delivers
    Q = counts[rank]
    out_i = np.full((Q, k), -1, np.int32)
    out_d = np.full((Q, k), np.nan, np.float32)
    out_c = np.zeros(Q, np.int32)
    by_row = n_total <= k
    for q in range(Q):
        cand = [(np.array([d], np.uint32).view(np.float32)[0], int(i))
                for s in range(world) for i, d in mine[s, q] if i >= 0]
        cand.sort(key=(lambda t: t[1]) if by_row else (lambda t: (t[0], t[1])))
        cand = cand[:K]
        use = min(len(cand), k)
        if not by_row and use_ratio and k == 1 and len(cand) >= 2:
            if np.float32(cand[0][0]) / np.float32(cand[1][0]) > np.float32(ratio_thr):
                use = 0
        for t in range(use):
            out_d[q, t], out_i[q, t] = cand[t]
        out_c[q] = n_total if by_row else use
    return out_i, out_d, out_c


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    import torch
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
        # every rank has its OWN test clouds (ragged: rank 0 two clouds, rank 1 three)
        n_te = 2 + rank
        xt, nt, rt, ot = synth.make_clouds(list(range(n_te)), [700 + 10 * rank + i for i in range(n_te)], 900)
        tx, tl, td, toff = orc.compute_features(prm, xt, nt, rt, ot)
        bounds = sharded.shard_bounds(cb.N, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        shard = orc.Model(prm, cb.rows(lo, hi))
        full = orc.Model(prm, cb)
        ok = True
        for dist_type in (0, 1):
            idx, dst, cnt = protocol_knn(shard, lo, cb.N, td, 2, dist_type, False, 0.0, dist, torch)
            ridx, rdst, rcnt = full.knn(td, k=2, dist_type=dist_type)
            ok = ok and np.array_equal(idx, ridx) and np.array_equal(dst.view(np.uint32), rdst.view(np.uint32)) \
                and np.array_equal(cnt, rcnt)
        # the k = 1 distance-ratio test needs the GLOBAL second neighbour: applied after the merge, never on a shard
        rp = prm.copy()
        rp.use_distance_ratio, rp.distance_ratio_threshold, rp.knn_k = 1, 0.8, 1
        full_r = orc.Model(rp, cb)
        idx, dst, cnt = protocol_knn(shard, lo, cb.N, td, 1, 0, True, 0.8, dist, torch)
        ridx, rdst, rcnt = full_r.knn(td, k=1, dist_type=0)
        ok = ok and np.array_equal(idx, ridx) and np.array_equal(cnt, rcnt) and 0 < int(rcnt.sum()) < len(rcnt)
        ok = ok and np.array_equal(dst.view(np.uint32)[rcnt > 0], rdst.view(np.uint32)[rcnt > 0])
        # the merged activation feeds the owner's (replicated) vote tables: votes equal the unsharded ones, in order
        idx, dst, cnt = protocol_knn(shard, lo, cb.N, td, 2, 0, False, 0.0, dist, torch)
        votes, voff = full.cast_votes(tx, tl, toff, idx, dst, cnt)
        ridx, rdst, rcnt = full.knn(td, k=2, dist_type=0)
        rvotes, rvoff = full.cast_votes(tx, tl, toff, ridx, rdst, rcnt)
        ok = ok and np.array_equal(voff, rvoff) and votes.tobytes() == rvotes.tobytes() and len(votes) > 0
        # N_total <= k: every codeword, in id order (activation_strategy_knn.h:50-54)
        tiny = cb.rows(0, 3)
        tb = sharded.shard_bounds(3, world)
        tshard = orc.Model(prm, tiny.rows(tb[rank], tb[rank + 1])) if tb[rank + 1] > tb[rank] else None

        class _Empty:
            def knn(self, qq, k=None, dist_type=None):
                return (np.full((len(qq), k), -1, np.int32), np.full((len(qq), k), np.nan, np.float32),
                        np.zeros(len(qq), np.int32))
        idx, dst, cnt = protocol_knn(tshard or _Empty(), tb[rank], 3, td[:5], 4, 0, False, 0.0, dist, torch)
        ridx, rdst, rcnt = orc.Model(prm, tiny).knn(td[:5], k=4, dist_type=0)
        ok = ok and np.array_equal(idx[:, :3], ridx[:, :3]) and np.array_equal(cnt, rcnt) \
            and np.array_equal(dst[:, :3].view(np.uint32), rdst[:, :3].view(np.uint32))
        # test-set sharding: ranks classify disjoint contiguous shards, labels concatenate to the unsharded answer
        xa, na, ra, oa = synth.make_clouds([0, 1, 2], [800, 801, 802], 900)
        lab_all, _, _ = full.classify_batch(xa, na, ra, oa, want_maxima=False)
        cuts = sharded.shard_bounds(3, world)
        a, b = cuts[rank], cuts[rank + 1]
        lab, _, _ = full.classify_batch(xa[oa[a]:oa[b]], na[oa[a]:oa[b]], ra[oa[a]:oa[b]], oa[a:b + 1] - oa[a],
                                        want_maxima=False) if b > a else (np.zeros(0, np.int32), None, None)
        ok = ok and np.array_equal(lab, lab_all[a:b])
        # one scene, keypoints sharded BLOCK-CYCLICALLY over the ranks (C5, comm.cu: blocks of 32 keypoints round-robin
        # for load balance): every vote travels with the global index of its keypoint and a stable sort on that index
        # restores the single-rank vote list, bit for bit
        sprm = synth.workload_params("c2", knn_k=2, single_object_mode=0, min_votes_threshold=3)
        sx_, sn_, sc_, _ = synth.make_scene([0, 1, 2], 5, 900, plane_points=1500, clutter_points=300)
        sx_[7] = np.nan
        sm = orc.Model(sprm, cb)
        fin = np.isfinite(sx_).all(1)
        pts, col, nr = sx_[fin], sc_[fin], sn_[fin]
        kp, kr, _ = orc.voxel_keypoints(pts, col, [0, len(pts)], sprm.leaf_size)
        BS = 32
        mine = np.array([g for g in range(len(kp)) if (g // BS) % world == rank], np.int64)
        surf = np.isfinite(nr).all(1)
        sxx, snn, scc, soff = pts[surf], nr[surf], col[surf], [0, int(surf.sum())]
        kx, kc, gidx = kp[mine], kr[mine], mine
        lrf = orc.shot_lrf(sxx, soff, kx, [0, len(kx)], sprm.lrf_radius)
        good = np.isfinite(lrf[:, 0]) & np.isfinite(lrf[:, 3]) & np.isfinite(lrf[:, 6])
        kx, kc, lrf, gidx = kx[good], kc[good], lrf[good], gidx[good]
        desc = orc.shot_describe(sprm.feature_type, sxx, snn, scc, soff, kx, kc, lrf, [0, len(kx)], sprm.feature_radius)
        good = ~np.isnan(desc).any(1)
        kx, lrf, desc, gidx = kx[good], lrf[good], desc[good], gidx[good]
        i1, d1, c1 = sm.knn(desc, k=2)
        v_mine, per_feat = sm.cast_votes(kx, lrf, np.arange(len(kx) + 1), i1, d1, c1)   # one "cloud" per feature
        keys = np.repeat(gidx, np.diff(per_feat))
        parts = _gather_np(v_mine, dist, torch)
        kparts = _gather_np(keys, dist, torch)
        allv, allk = np.concatenate(parts), np.concatenate(kparts)
        sv = allv[np.argsort(allk, kind="stable")]
        fx1, fl1, fd1, fo1 = orc.compute_features(sprm, sx_, sn_, sc_, [0, len(sx_)])
        i1, d1, c1 = sm.knn(fd1, k=2)
        rv, rvo = sm.cast_votes(fx1, fl1, fo1, i1, d1, c1)
        ok = ok and sv.tobytes() == rv.tobytes() and len(sv) > 0 and len(kp) > 2 * BS
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world2_exchange_protocol_matches_the_unsharded_oracle():
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


def test_interleaved_codebook_is_the_same_model():
    """sharded.interleave_codebook deals the rows cyclically over the shards (training appends codewords class by class,
    so contiguous shards of the untouched table hold a few classes each): the permuted table must be the same model —
    same distances, rows mapped through the permutation, same votes, same labels — and every contiguous shard of it
    must see every class."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    from oracle import oracle_py as orc
    from pcdb200 import sharded, synth
    prm = synth.workload_params("c2", knn_k=2)
    tr_cls = [0, 0, 1, 1, 2, 2]
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [300 + i for i in range(6)], 900)
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(6)])
    cb = orc.train(prm, fx, fl, fd, foff, tr_cls, list(range(6)), bb, 3)
    world = 3
    cbi, perm = sharded.interleave_codebook(cb, world)
    assert cbi.N == cb.N and sorted(perm.tolist()) == list(range(cb.N))
    assert np.array_equal(cbi.words, cb.words[perm]) and np.array_equal(cbi.kp_train, cb.kp_train[perm])
    for i in (0, 1, cb.N // 2, cb.N - 1):  # the vote block of a row travels with it
        a0, a1 = int(cbi.vote_off[i]), int(cbi.vote_off[i + 1])
        b0, b1 = int(cb.vote_off[perm[i]]), int(cb.vote_off[perm[i] + 1])
        assert a1 - a0 == b1 - b0 and np.array_equal(cbi.vote_xyz[a0:a1], cb.vote_xyz[b0:b1]) \
            and np.array_equal(cbi.vote_class[a0:a1], cb.vote_class[b0:b1])
    # contiguous shards of the dealt table see every class; those of the original table do not
    b = sharded.shard_bounds(cb.N, world)
    row_class = np.array([cb.vote_class[cb.vote_off[r]] for r in range(cb.N)])
    for r in range(world):
        assert len(set(row_class[perm][b[r]:b[r + 1]].tolist())) == 3
    assert any(len(set(row_class[b[r]:b[r + 1]].tolist())) < 3 for r in range(world))
    xt, nt, rt, ot = synth.make_clouds([0, 1, 2], [700, 701, 702], 900)
    m0, m1 = orc.Model(prm, cb), orc.Model(prm, cbi)
    tx, tl, td, toff = orc.compute_features(prm, xt, nt, rt, ot)
    i0, d0, c0 = m0.knn(td, k=2)
    i1, d1, c1 = m1.knn(td, k=2)
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32)) and np.array_equal(c0, c1)
    untied = d0[:, 0] != d0[:, 1]
    assert untied.mean() > 0.9 and np.array_equal(perm[i1[untied]], i0[untied])
    l0, _, _ = m0.classify_batch(xt, nt, rt, ot, want_maxima=False)
    l1, _, _ = m1.classify_batch(xt, nt, rt, ot, want_maxima=False)
    assert np.array_equal(l0, l1) and l0.tolist() == [0, 1, 2]


def test_interleave_codebook_csr_with_empty_rows_and_optional_arrays():
    """Rows without votes, a world that does not divide N, absent optional arrays: every row keeps its own vote block."""
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    from pcdb200 import sharded
    from pcdb200.structs import Codebook
    rng = np.random.default_rng(5)
    for n_rows, world, with_opt in ((53, 4, True), (7, 8, False), (64, 2, True), (1, 3, False)):
        cnt = rng.integers(0, 4, n_rows)
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        V = int(off[-1])
        tag = np.repeat(np.arange(n_rows), cnt)                     # every vote remembers its row
        cb = Codebook(rng.random((n_rows, 16), np.float32), off, np.stack([tag, tag, tag], 1), tag + 0.5, tag % 3,
                      tag, np.tile(np.arange(7, dtype=np.float32), (V, 1)) + tag[:, None],
                      (tag + 0.25) if with_opt else None, rng.random((n_rows, 3), np.float32),
                      np.arange(n_rows) if with_opt else None, np.ones(3), (np.arange(n_rows) + 2.0) if with_opt else None)
        out, perm = sharded.interleave_codebook(cb, world)
        assert sorted(perm.tolist()) == list(range(n_rows)) and out.N == n_rows
        assert np.array_equal(perm % world, np.sort(perm % world))  # grouped by residue = contiguous shards of a deal
        assert int(out.vote_off[-1]) == V and np.array_equal(np.diff(out.vote_off), cnt[perm])
        for i in range(n_rows):
            a0, a1 = int(out.vote_off[i]), int(out.vote_off[i + 1])
            assert (out.vote_instance[a0:a1] == perm[i]).all() and (out.vote_xyz[a0:a1, 0] == perm[i]).all()
            assert np.allclose(out.vote_weight[a0:a1], perm[i] + 0.5) and (out.vote_bbox[a0:a1, 0] == perm[i]).all()
            if with_opt:
                assert np.allclose(out.vote_class_weight[a0:a1], perm[i] + 0.25)
        assert np.array_equal(out.words, cb.words[perm])
        if with_opt:
            assert np.array_equal(out.codeword_ids, perm) and np.allclose(out.codeword_weight, perm + 2.0)
        else:
            assert out.vote_class_weight is None or len(out.vote_class_weight) == 0
