"""CPU tests of the oracle (oracle/pcd_oracle.cpp): the pins against real reference code / a real FLANN build, and the
known-answer tests SURVEY.md 8c lists (the reference ships no tests of its own)."""
import os

import numpy as np
import pytest

from pcdb200 import synth
from pcdb200.structs import (DIST_CHISQUARED, DIST_EUCLIDEAN, FEATURE_CSHOT, FEATURE_SHOT, VOTE_DTYPE, Codebook,
                             default_params)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- pins --------------------------------------------------------------------------------------------------------
def test_lab_matches_reference_golden(orc):
    """RGB->CIELab and the colour distance are bit-exact with the reference's own color_conversion.cpp."""
    g = np.load(os.path.join(GOLD, "lab_golden.npz"))
    lab = orc.rgb_to_lab_normalized(g["rgb"])
    assert np.array_equal(lab.view(np.uint32), g["lab"].view(np.uint32))
    d = orc.color_distance(lab, lab[g["perm"]])
    assert np.array_equal(d.view(np.uint32), g["dist"].view(np.uint32))
    srgb, sxyz = orc.lab_luts()
    assert np.array_equal(srgb, g["srgb_lut"]) and np.array_equal(sxyz, g["sxyz_lut"])


def test_lab_matches_live_reference_build(orc):
    """Same pin against oracle/_ref/libref_color.so when it travelled with the snapshot."""
    import ctypes as C
    ref = orc.ref_color_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference not mounted)")
    rgb = np.random.default_rng(3).integers(0, 1 << 24, 2000, dtype=np.uint32).astype(np.uint32)
    out = np.empty((2000, 3), np.float32)
    ref.ref_rgb_to_lab_normalized(rgb.ctypes.data_as(C.c_void_p), C.c_int64(2000), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(orc.rgb_to_lab_normalized(rgb).view(np.uint32), out.view(np.uint32))


@pytest.mark.parametrize("name", ["d352", "d1344", "d30"])
def test_l2_functor_matches_flann_golden(orc, name):
    """Squared-L2 values (4-wide grouping + tail) and the exact-kNN order are bit-exact with a real FLANN build."""
    f = np.load(os.path.join(GOLD, "flann_golden.npz"))
    W, Q = f[name + "_words"], f[name + "_queries"]
    gi, gd = f[name + "_l2_idx"], f[name + "_l2_dist"]
    for j in range(gi.shape[1]):
        od = orc.distance(Q, W[gi[:, j]], DIST_EUCLIDEAN)
        assert np.array_equal(od.view(np.uint32), gd[:, j].view(np.uint32))
    if W.shape[1] in (352, 1344):
        prm = default_params(knn_k=3, feature_type=FEATURE_CSHOT if W.shape[1] == 1344 else FEATURE_SHOT)
        cb = _dummy_codebook(W)
        m = orc.Model(prm, cb)
        idx, dist, cnt = m.knn(Q, k=3, dist_type=DIST_EUCLIDEAN)
        assert np.array_equal(idx, gi) and np.array_equal(dist.view(np.uint32), gd.view(np.uint32))
        assert (cnt == 3).all()


def _dummy_codebook(W, n_classes=2):
    N = W.shape[0]
    return Codebook(W, np.arange(N + 1), np.zeros((N, 3)), np.ones(N), np.zeros(N), np.zeros(N),
                    np.tile(np.array([1, 0, 0, 0, 1, 1, 1], np.float32), (N, 1)), np.ones(N), np.zeros((N, 3)),
                    np.arange(N), np.ones(n_classes))


# ---- A.1 voxel grid ----------------------------------------------------------------------------------------------
def _voxel_numpy(xyz, rgb, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = xyz.min(0), xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    out, cols = [], []
    s = 0
    while s < len(order):
        e = s
        acc = np.zeros(3, np.float32)
        c = np.zeros(3, np.float32)
        while e < len(order) and idx[order[e]] == idx[order[s]]:
            acc = acc + xyz[order[e]]
            v = int(rgb[order[e]])
            c = c + np.array([(v >> 16) & 255, (v >> 8) & 255, v & 255], np.float32)
            e += 1
        n = np.float32(e - s)
        out.append(acc / n)
        cc = (c / n).astype(np.uint32)
        cols.append((cc[0] << 16) | (cc[1] << 8) | cc[2])
        s = e
    return np.array(out, np.float32), np.array(cols, np.uint32)


def test_voxel_keypoints_known_answer(orc):
    xyz, _, rgb, off = synth.make_clouds([0, 1], [1, 2], 700)
    kp, kr, koff = orc.voxel_keypoints(xyz, rgb, off, 0.1)
    for b in range(2):
        e_kp, e_rgb = _voxel_numpy(xyz[off[b]:off[b + 1]], rgb[off[b]:off[b + 1]], 0.1)
        assert np.array_equal(kp[koff[b]:koff[b + 1]].view(np.uint32), e_kp.view(np.uint32))
        assert np.array_equal(kr[koff[b]:koff[b + 1]], e_rgb)


def test_voxel_keypoints_edge_cases(orc):
    # empty batch entry, single point, NaN points ignored
    xyz = np.array([[0.01, 0.02, 0.03], [np.nan, 0, 0], [0.5, 0.5, 0.5], [0.51, 0.5, 0.5]], np.float32)
    off = np.array([0, 0, 1, 4], np.int64)
    kp, kr, koff = orc.voxel_keypoints(xyz, np.zeros(4, np.uint32), off, 0.1)
    assert koff.tolist() == [0, 0, 1, 2]
    assert np.allclose(kp[0], xyz[0]) and np.allclose(kp[1], (xyz[2] + xyz[3]) / 2)


# ---- A.2 radius search ---------------------------------------------------------------------------------------------
def test_radius_neighbours_vs_numpy(orc):
    xyz, _, _, off = synth.make_clouds([0, 1], [3, 4], 900)
    kp, _, koff = orc.voxel_keypoints(xyz, None, off, 0.15)
    r = 0.22
    noff, idx, d2 = orc.radius_neighbours(xyz, off, kp, koff, r)
    r2 = np.float32(r * r)
    for b in range(2):
        pts = xyz[off[b]:off[b + 1]]
        for q in range(koff[b], koff[b + 1]):
            diff = kp[q] - pts
            dd = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
            sel = np.nonzero(dd < r2)[0]
            order = np.lexsort((sel, dd[sel]))
            assert np.array_equal(idx[noff[q]:noff[q + 1]], sel[order].astype(np.int32))
            assert np.array_equal(d2[noff[q]:noff[q + 1]], dd[sel][order])


# ---- A.3 LRF -------------------------------------------------------------------------------------------------------
def _lrf_numpy(pts, kp, R):
    """Independent restatement of SURVEY A.3 with numpy.linalg.eigh: float32 membership test, float64 covariance,
    majority sign with the 5-around-the-median tie rule on the (d^2, index)-sorted valid neighbours."""
    diff = kp - pts
    dd = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    sel = np.nonzero((dd < np.float32(R * R)) & ~((pts == kp).all(1)))[0]
    sel = sel[np.lexsort((sel, dd[sel]))]
    v = (pts[sel] - kp).astype(np.float64)
    w = R - np.sqrt(dd[sel].astype(np.float64))
    cov = (v[:, :, None] * v[:, None, :] * w[:, None, None]).sum(0) / w.sum()
    _, V = np.linalg.eigh(cov)

    def fix(a):
        s = 2 * ((v @ a) >= 0).sum() - len(v)
        if s == 0:
            m = len(v) // 2
            s = 1 if ((v[m - 2:m + 3] @ a) > 0).sum() >= 3 else -1
        return a if s > 0 else -a

    x, z = fix(V[:, 2].copy()), fix(V[:, 0].copy())
    return np.stack([x, np.cross(z, x), z])


def test_lrf_known_answer(orc):
    """Hand-checkable neighbourhood: spreads 0.5 > 0.4 > 0.3 along x, y, z, so x-axis ~ +/-e_x (largest weighted
    variance) and z-axis ~ +/-e_z; two extra points on the +x/+z side break the sign symmetry.  Checked against an
    independent numpy eigh restatement."""
    pts = np.array([[0.5, 0, 0], [-0.5, 0, 0], [0, 0.4, 0], [0, -0.4, 0], [0, 0, 0.3], [0, 0, -0.3],
                    [0.2, 0.01, 0.05], [0.15, -0.01, 0.06], [0, 0, 0]], np.float32)
    lrf = orc.shot_lrf(pts, [0, len(pts)], np.zeros((1, 3), np.float32), [0, 1], 1.0)[0].reshape(3, 3)
    expect = _lrf_numpy(pts, np.zeros(3, np.float32), 1.0)
    assert np.allclose(lrf, expect, atol=1e-6)
    assert lrf[0, 0] > 0.9 and lrf[2, 2] > 0.9
    assert np.allclose(lrf @ lrf.T, np.eye(3), atol=1e-6)
    assert np.isclose(np.linalg.det(lrf.astype(np.float64)), 1.0, atol=1e-6)
    # a real cloud: every finite frame matches the numpy restatement (sign ties excluded by construction)
    xyz, _, _, off = synth.make_clouds([1], [77], 1200)
    kp, _, koff = orc.voxel_keypoints(xyz, None, off, 0.2)
    got = orc.shot_lrf(xyz, off, kp, koff, 0.3)
    for q in range(len(kp)):
        if np.isfinite(got[q]).all():
            assert np.allclose(got[q].reshape(3, 3), _lrf_numpy(xyz, kp[q], float(np.float32(0.3))), atol=2e-5)


def test_lrf_too_few_neighbours_is_nan(orc):
    pts = np.array([[0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1], [0.05, 0.05, 0]], np.float32)
    lrf = orc.shot_lrf(pts, [0, 4], np.zeros((1, 3), np.float32), [0, 1], 1.0)
    assert np.isnan(lrf).all()


# ---- A.4 SHOT ------------------------------------------------------------------------------------------------------
def _volume_index(x, y, z, dist, r):
    bit4 = 1 if (y > 0 or (y == 0 and x < 0)) else 0
    bit3 = (1 - bit4) if (x > 0 or (x == 0 and y > 0)) else bit4
    di = ((bit4 << 3) + (bit3 << 2)) << 1
    if x * y > 0 or x == 0:
        di += 0 if abs(x) >= abs(y) else 4
    else:
        di += 4 if abs(x) > abs(y) else 0
    di += 1 if z > 0 else 0
    di += 2 if dist > r / 2 else 0
    return di


def test_shot_known_answer_five_volume_centres(orc):
    """Five neighbours, each at the exact centre of a different SHOT volume, normals = z axis (bin distance exactly
    10): every interpolation residual is 0, each point adds weight 4 to one bin -> five bins = 1/sqrt(5)."""
    r = 1.0
    specs = [(0.25, np.pi / 4, 0), (0.75, np.pi / 4, 2), (0.25, 3 * np.pi / 4, 4), (0.75, 3 * np.pi / 4, 6),
             (0.75, np.pi / 4, 7)]
    pts, bins = [], []
    for rad, inc, sel in specs:
        az = -7 * np.pi / 8 + sel * np.pi / 4
        p = np.array([rad * np.sin(inc) * np.cos(az), rad * np.sin(inc) * np.sin(az), rad * np.cos(inc)])
        pts.append(p)
        bins.append(_volume_index(p[0], p[1], p[2], rad, r) * 11 + 10)
    assert len(set(bins)) == 5
    pts = np.array(pts, np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (5, 1))
    lrf = np.eye(3, dtype=np.float32).reshape(1, 9)
    desc = orc.shot_describe(FEATURE_SHOT, pts, nrm, None, [0, 5], np.zeros((1, 3), np.float32), None, lrf, [0, 1], r)[0]
    expect = np.zeros(352, np.float32)
    expect[bins] = 1 / np.sqrt(5)
    assert np.allclose(desc, expect, atol=2e-6)


def test_shot_invariants_and_rigid_motion(orc):
    xyz, nrm, rgb, off = synth.make_clouds([2], [9], 1500)
    kp, kr, koff = orc.voxel_keypoints(xyz, rgb, off, 0.15)
    lrf = orc.shot_lrf(xyz, off, kp, koff, 0.3)
    ok = np.isfinite(lrf).all(1)
    assert ok.sum() > 20
    for ft, dim in ((FEATURE_SHOT, 352), (FEATURE_CSHOT, 1344)):
        d = orc.shot_describe(ft, xyz, nrm, rgb, off, kp, kr, lrf, koff, 0.4)
        assert d.shape == (len(kp), dim)
        assert np.allclose(np.linalg.norm(d[ok].astype(np.float64), axis=1), 1, atol=1e-5) and (d[ok] >= 0).all()
        # rigid motion of the whole cloud (keypoints moved along): descriptors unchanged to float rounding
        R = synth._rand_rot(np.random.default_rng(5)).astype(np.float64)
        t = np.array([0.3, -0.2, 0.1])
        mv = lambda a: (a.astype(np.float64) @ R.T + t).astype(np.float32)  # noqa: E731
        xyz2, kp2, nrm2 = mv(xyz), mv(kp), (nrm.astype(np.float64) @ R.T).astype(np.float32)
        lrf2 = orc.shot_lrf(xyz2, off, kp2, koff, 0.3)
        d2 = orc.shot_describe(ft, xyz2, nrm2, rgb, off, kp2, kr, lrf2, koff, 0.4)
        # points on the search-radius boundary may flip membership under float rounding: compare the bulk
        err = np.abs(d2[ok] - d[ok]).max(1)
        assert np.median(err) < 1e-4 and (err < 5e-3).mean() > 0.9
    # permutation of the surface points: same neighbour sets, same (d2, index-free) order -> same descriptor
    perm = np.random.default_rng(6).permutation(len(xyz))
    d_a = orc.shot_describe(FEATURE_SHOT, xyz, nrm, rgb, off, kp, kr, lrf, koff, 0.4)
    d_b = orc.shot_describe(FEATURE_SHOT, xyz[perm], nrm[perm], rgb[perm], off, kp, kr, lrf, koff, 0.4)
    assert np.allclose(d_a[ok], d_b[ok], atol=1e-6)


def test_descriptor_nan_when_under_five_neighbours(orc):
    pts = np.array([[0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (3, 1))
    d = orc.shot_describe(FEATURE_SHOT, pts, nrm, None, [0, 3], np.zeros((1, 3), np.float32), None,
                          np.eye(3, dtype=np.float32).reshape(1, 9), [0, 1], 1.0)
    assert np.isnan(d).all()


# ---- A.6 kNN -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dist_type", [DIST_EUCLIDEAN, DIST_CHISQUARED])
def test_knn_vs_numpy_bruteforce(orc, dist_type):
    rng = np.random.default_rng(2)
    W = rng.random((500, 352), dtype=np.float32) ** 3
    W /= np.linalg.norm(W, axis=1, keepdims=True)
    Q = W[:40] + 0.05 * rng.random((40, 352), dtype=np.float32)
    m = orc.Model(default_params(knn_k=4), _dummy_codebook(W))
    idx, dist, cnt = m.knn(Q, k=4, dist_type=dist_type)
    Wd, Qd = W.astype(np.float64), Q.astype(np.float64)
    if dist_type == DIST_EUCLIDEAN:
        full = ((Qd[:, None] - Wd[None]) ** 2).sum(-1)
    else:
        s = Qd[:, None] + Wd[None]
        full = np.where(s > 0, (Qd[:, None] - Wd[None]) ** 2 / np.where(s > 0, s, 1), 0).sum(-1)
    ref = np.argsort(full, axis=1, kind="stable")[:, :4]
    assert (idx == ref).mean() > 0.99  # float32 vs float64 near-ties may swap
    assert np.allclose(dist, np.take_along_axis(full, idx.astype(np.int64), 1), rtol=2e-5)
    assert (np.diff(dist, axis=1) >= 0).all() and (cnt == 4).all()


def test_knn_small_codebook_and_ratio(orc):
    rng = np.random.default_rng(4)
    W = rng.random((3, 352), dtype=np.float32)
    Q = rng.random((5, 352), dtype=np.float32)
    m = orc.Model(default_params(knn_k=4), _dummy_codebook(W))
    idx, dist, cnt = m.knn(Q, k=4)  # N <= k: every row, in row order (activation_strategy_knn.h:50-54)
    assert (cnt == 3).all() and (idx[:, :3] == np.arange(3)).all() and (idx[:, 3] == -1).all()
    W = rng.random((50, 352), dtype=np.float32)
    prm = default_params(knn_k=1, use_distance_ratio=1, distance_ratio_threshold=0.8)
    m = orc.Model(prm, _dummy_codebook(W))
    idx, dist, cnt = m.knn(Q, k=1)
    m2 = orc.Model(default_params(knn_k=2), _dummy_codebook(W))
    _, d2, _ = m2.knn(Q, k=2)
    assert np.array_equal(cnt, (d2[:, 0] / d2[:, 1] <= np.float32(0.8)).astype(np.int32))


# ---- A.7 votes / A.8 maxima ---------------------------------------------------------------------------------------
def test_self_classification_collapses_on_bbox_centre(orc, small_world):
    """A cloud voted against a codebook trained on itself: all distances 0, all votes on the bbox centre, one maximum
    per cloud, weight 1.0 after normalisation."""
    w = small_world
    xyz, nrm, rgb, off, tr_cls = w["train"]
    m = orc.Model(w["prm"], w["cb"])
    fx, fl, fd, foff = w["feats"]
    idx, dist, cnt = m.knn(fd)
    assert (dist[:, 0] == 0).all() and (cnt == 1).all()
    votes, voff = m.cast_votes(fx, fl, foff, idx, dist, cnt)
    assert len(votes) == len(fx)
    for b in range(len(tr_cls)):
        centre = orc.aabb(xyz[off[b]:off[b + 1]])[:3]
        v = votes[voff[b]:voff[b + 1]]
        assert np.abs(v["position"] - centre).max() < 1e-4
        assert (v["class_id"] == tr_cls[b]).all() and (v["instance_id"] == b).all()
    labels, mx, moff = m.classify_batch(xyz, nrm, rgb, off)
    assert labels.tolist() == tr_cls
    assert (np.diff(moff) == 1).all() and np.allclose(mx["weight"], 1.0)


def test_vote_rotation_matches_matrix_form(orc, small_world):
    """The quaternion route of Utils::rotateBack equals v0*X + v1*Y + v2*Z to ~1e-6."""
    w = small_world
    xt, nt, rt, ot, _ = w["test"]
    m = orc.Model(w["prm"], w["cb"])
    fx, fl, fd, foff = orc.compute_features(w["prm"], xt, nt, rt, ot)
    idx, dist, cnt = m.knn(fd)
    votes, voff = m.cast_votes(fx, fl, foff, idx, dist, cnt)
    cb = w["cb"]
    sig = cb.sigma2
    n = 0
    for f in range(len(fx)):
        row = idx[f, 0]
        v0 = int(cb.vote_off[row])
        if dist[f, 0] > 2 * sig[cb.vote_class[v0]]:
            continue
        R = fl[f].reshape(3, 3).astype(np.float64)
        expect = fx[f] + cb.vote_xyz[v0].astype(np.float64) @ R
        assert np.abs(votes[n]["position"] - expect).max() < 2e-5
        assert votes[n]["codeword_id"] == cb.codeword_ids[row]
        n += 1
    assert n == len(votes) and n > 0


def _blob_votes(centres, n_per, sigma, rng, cls=0):
    votes = np.zeros(len(centres) * n_per, VOTE_DTYPE)
    for i, c in enumerate(centres):
        sl = slice(i * n_per, (i + 1) * n_per)
        votes["position"][sl] = (np.asarray(c) + rng.normal(scale=sigma, size=(n_per, 3))).astype(np.float32)
    votes["weight"] = 1.0
    votes["class_id"] = cls
    votes["bbox_quat"][:, 0] = 1
    votes["bbox_size"] = 1
    return votes


def test_meanshift_two_blobs(orc):
    rng = np.random.default_rng(8)
    centres = [(0.0, 0.0, 0.0), (2.0, 0.5, -1.0)]
    votes = _blob_votes(centres, 200, 0.05, rng)
    prm = default_params(bandwidth=0.3)
    m = orc.Model(prm, _dummy_codebook(np.zeros((4, 352), np.float32)))
    mx, moff, mi, mw = m.find_maxima(votes, [0, len(votes)])
    assert len(mx) == 2 and np.isclose(mx["weight"].sum(), 1.0, atol=1e-6)
    got = mx["position"][np.argsort(mx["position"][:, 0])]
    assert np.abs(got - np.array(centres)).max() < 0.02
    assert sorted(mx["n_votes"].tolist()) == [200, 200]
    # member lists partition the votes of well separated blobs
    assert len(np.unique(mi)) == 400


def test_find_maxima_thresholds(orc):
    rng = np.random.default_rng(9)
    votes = np.concatenate([_blob_votes([(0, 0, 0)], 100, 0.02, rng, cls=0),
                            _blob_votes([(3, 0, 0)], 10, 0.02, rng, cls=1)])
    cbk = _dummy_codebook(np.zeros((4, 352), np.float32))
    m = orc.Model(default_params(bandwidth=0.3), cbk)
    mx, _, _, _ = m.find_maxima(votes, [0, len(votes)])
    assert mx["class_id"].tolist() == [0, 1] and mx["weight"][0] > mx["weight"][1]
    m.set_params(default_params(bandwidth=0.3, best_k=1))
    assert len(m.find_maxima(votes, [0, len(votes)])[0]) == 1
    m.set_params(default_params(bandwidth=0.3, min_threshold=-0.5))
    assert len(m.find_maxima(votes, [0, len(votes)])[0]) == 1
    m.set_params(default_params(bandwidth=0.3, min_votes_threshold=50))
    assert len(m.find_maxima(votes, [0, len(votes)])[0]) == 1
    # empty batch entries
    mx, moff, _, _ = m.find_maxima(votes[:0], [0, 0, 0])
    assert len(mx) == 0 and moff.tolist() == [0, 0, 0]


def test_oracle_end_to_end_labels(orc, small_world):
    w = small_world
    xt, nt, rt, ot, te_cls = w["test"]
    m = orc.Model(w["prm"], w["cb"])
    labels, mx, moff = m.classify_batch(xt, nt, rt, ot)
    assert (labels == np.array(te_cls)).mean() >= 0.75
    assert m.last_counts["votes"] > 0.2 * m.last_counts["features"]  # the sigma^2 filter keeps a healthy share


def test_merge_topk(orc):
    rng = np.random.default_rng(10)
    S, Q, k = 3, 20, 4
    d = np.sort(rng.random((S, Q, k)).astype(np.float32), axis=2)
    i = rng.integers(0, 1000, (S, Q, k)).astype(np.int32)
    i[2, :, 3] = -1
    idx, dist = orc.merge_topk(i, d)
    for q in range(Q):
        c = sorted((float(d[s, q, j]), int(i[s, q, j])) for s in range(S) for j in range(k) if i[s, q, j] >= 0)[:k]
        assert [x[1] for x in c] == idx[q].tolist() and np.allclose([x[0] for x in c], dist[q])


# ---- normals ---------------------------------------------------------------------------------------------
def test_pca_normals_known_answers(orc):
    """computePointNormalMod: a noisy plane gives +-z, curvature ~ 0; every point agrees with numpy.linalg.eigh of the
    same neighbourhood; method 0 orients towards the origin, method 1 away from the centroid."""
    rng = np.random.default_rng(3)
    pts = np.zeros((1500, 3), np.float32)
    pts[:, :2] = rng.uniform(-1, 1, (1500, 2))
    pts[:, 2] = 2.0 + rng.normal(0, 1e-4, 1500)
    off = np.array([0, 1500], np.int64)
    prm = default_params(normal_radius=0.2, consistent_normals_method=0)
    n, cv = orc.compute_normals(prm, pts, off)
    assert np.isfinite(n).all() and (np.abs(n[:, 2]) > 0.9999).all() and (n[:, 2] < 0).all()  # towards the origin
    assert (cv < 2e-3).all()  # float moment sums at z = 2: cancellation noise, as in the reference
    x, _, _, o = synth.make_clouds([1], [17], 3000)
    n0, cv0 = orc.compute_normals(prm, x, o)
    for i in (0, 100, 2000):
        nb = x[((x - x[i]) ** 2).sum(1) < np.float32(0.2 ** 2)].astype(np.float64)
        w, V = np.linalg.eigh(np.cov(nb.T, bias=True))
        assert abs(abs(np.dot(n0[i], V[:, 0])) - 1) < 1e-4 and abs(cv0[i] - w[0] / w.sum()) < 1e-3
        assert np.dot(n0[i], -x[i]) >= 0
    prm1 = default_params(normal_radius=0.2, consistent_normals_method=1)
    n1, _ = orc.compute_normals(prm1, x, o)
    assert (((x - x.mean(0)) * n1).sum(1) >= 0).all()  # pointing away from the centroid


def test_shot_frame_normals_and_repair_loop(orc):
    """Method 2: normal = inverted z axis of the SHOT frame at NormalRadius; the reference's repair loop rewrites the
    FIRST n_invalid points (normal_orientation.cpp:96-106), kept as written."""
    x, _, _, o = synth.make_clouds([1], [17], 2000)
    far = np.array([[9, 9, 9], [9.01, 9, 9], [9, 9.01, 9.01]], np.float32)  # 3 points: PCA defined, no SHOT frame
    xyz = np.concatenate([x, far]).astype(np.float32)
    off = np.array([0, len(xyz)], np.int64)
    prm = default_params(normal_radius=0.1, consistent_normals_method=2)
    n2, cv2 = orc.compute_normals(prm, xyz, off)
    kp = xyz.copy()
    lrf = orc.shot_lrf(xyz, off, kp, [0, len(kp)], float(np.float32(0.1)))
    valid = np.isfinite(lrf[:, 0])
    assert (~valid).sum() == 3
    body = np.arange(3, 2000)
    assert np.array_equal(n2[body], -lrf[body, 6:9])
    raw, _ = orc.compute_normals(default_params(normal_radius=0.1, consistent_normals_method=0), xyz, off)
    assert np.allclose(np.abs((n2[:3] * raw[:3]).sum(1)), 1, atol=1e-6) and (cv2[:3] == 0).all()


# ---- the approximate-mode cost stand-in ------------------------------------------------------------
def test_kd_forest_is_a_search_over_the_same_rows(orc):
    """FLANN-like randomized kd-forest: with unlimited checks it must return the exact neighbours; with 128 checks its
    distances can only be >= the exact ones and most queries still find the true neighbour on clustered data."""
    rng = np.random.default_rng(5)
    centres = rng.random((40, 64), dtype=np.float32)
    W = (centres[rng.integers(0, 40, 4000)] + 0.05 * rng.standard_normal((4000, 64))).astype(np.float32)
    Q = (centres[rng.integers(0, 40, 200)] + 0.05 * rng.standard_normal((200, 64))).astype(np.float32)
    N = W.shape[0]
    cb = Codebook(W, np.arange(N + 1), np.zeros((N, 3)), np.ones(N), np.zeros(N), np.zeros(N),
                  np.tile(np.array([1, 0, 0, 0, 1, 1, 1], np.float32), (N, 1)), np.ones(N), np.zeros((N, 3)),
                  np.arange(N), np.ones(2))
    m = orc.Model(default_params(), cb)
    ie, de, _ = m.knn(Q, k=2, dist_type=DIST_EUCLIDEAN)
    assert m.set_approximate(4, 10 ** 9) >= 0
    ia, da, _ = m.knn(Q, k=2, dist_type=DIST_EUCLIDEAN)
    assert np.array_equal(ia, ie) and np.allclose(da, de, rtol=1e-6)
    m.set_approximate(4, 128)
    ib, db, _ = m.knn(Q, k=2, dist_type=DIST_EUCLIDEAN)
    assert (db >= de * (1 - 1e-6)).all() and (ib[:, 0] == ie[:, 0]).mean() > 0.5
    m.set_approximate(0)
    ic, _, _ = m.knn(Q, k=2, dist_type=DIST_EUCLIDEAN)
    assert np.array_equal(ic, ie)


# ---- committed whole-path fixture (tests/golden/path_golden.npz) ---------------------------------------------------------
def _path_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden",
                                                                              "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.path_world(), np.load(os.path.join(os.path.dirname(__file__), "golden", "path_golden.npz"))


def test_oracle_reproduces_the_committed_path_fixture(orc):
    """The oracle is deterministic and has not drifted: every stage output of the seeded world equals the committed
    file bit for bit (features, codebook, activation, votes, maxima, labels)."""
    (prm, tr_cls, (xyz, nrm, rgb, off), te_cls, (xt, nt, rt, ot)), G = _path_golden()
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    assert np.array_equal(foff, G["train_feat_off"])
    bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = orc.train(prm, fx, fl, fd, foff, tr_cls, list(range(len(tr_cls))), bb, 3)
    assert cb.words.tobytes() == G["codebook_words"].tobytes() and cb.sigma2.tobytes() == G["codebook_sigma2"].tobytes()
    m = orc.Model(prm, cb)
    tx, tl, td, toff = orc.compute_features(prm, xt, nt, rt, ot)
    assert np.array_equal(toff, G["feat_off"]) and tx.tobytes() == G["feat_xyz"].tobytes()
    assert tl.tobytes() == G["feat_lrf"].tobytes() and td.tobytes() == G["feat_desc"].tobytes()
    idx, dist, cnt = m.knn(td, k=2)
    assert np.array_equal(idx, G["knn_idx"]) and dist.tobytes() == G["knn_dist"].tobytes()
    votes, voff = m.cast_votes(tx, tl, toff, idx, dist, cnt)
    assert np.array_equal(voff, G["vote_off"]) and votes.tobytes() == G["votes"].tobytes()
    labels, mx, moff = m.classify_batch(xt, nt, rt, ot)
    assert labels.tolist() == G["labels"].tolist() == te_cls
    assert np.array_equal(moff, G["maxima_off"]) and mx.tobytes() == G["maxima"].tobytes()


def test_pruned_exact_search_equals_the_linear_scan(orc, small_world):
    """oracle_py.knn_exact_pruned (sgemm proposals + FLANN-order functor, used by bench.py to check >= 64 labels at the
    1 M-word scale) must return exactly what the linear scan returns: rows, distance bits, counts — both functors — and
    classify_batch_pruned the same labels as classify_batch."""
    prm, cb = small_world["prm"], small_world["cb"]
    m = orc.Model(prm, cb)
    _, _, fd, _ = small_world["feats"]
    xt, nt, rt, ot, _ = small_world["test"]
    q = orc.compute_features(prm, xt, nt, rt, ot)[2][:300]
    q = np.concatenate([q, fd[:20]])  # exact hits (distance 0) included
    for dist_type in (0, 1):
        for k in (1, 3):
            a = orc.knn_exact_pruned(m, q, k=k, dist_type=dist_type)
            b = m.knn(q, k=k, dist_type=dist_type)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
            assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    lab = orc.classify_batch_pruned(m, xt, nt, rt, ot)
    ref, _, _ = m.classify_batch(xt, nt, rt, ot, want_maxima=False)
    assert np.array_equal(lab, ref)


def _votes(centres, n_per, cls, rng, sigma=0.01, inst=0):
    from pcdb200.structs import VOTE_DTYPE
    v = np.zeros(len(centres) * n_per, VOTE_DTYPE)
    for i, c in enumerate(centres):
        v["position"][i * n_per:(i + 1) * n_per] = (np.asarray(c) + rng.normal(scale=sigma, size=(n_per, 3))).astype(np.float32)
    v["weight"] = 1.0
    v["class_id"] = cls
    v["instance_id"] = inst
    v["bbox_quat"][:, 0] = 1.0
    v["bbox_size"] = 1.0
    return v


def test_cross_class_merge_filter_known_answer(orc):
    """MaxFilterType "Merge" (maxima_handler.cpp:296-383): a class with a large search distance subsumes the maxima of
    classes with smaller ones around it; same-class maxima inside the group are merged by weighted averaging
    (mergeMaxima :386-443) and only the heaviest candidate of the group survives."""
    from pcdb200.structs import Codebook, default_params
    rng = np.random.default_rng(3)
    votes = np.concatenate([_votes([(0, 0, 0)], 50, 0, rng),                      # class 0: search distance 0.5
                            _votes([(0.2, 0, 0), (-0.2, 0, 0)], 40, 1, rng),      # class 1: 0.1 -> two maxima, 80 votes
                            _votes([(5, 5, 5)], 10, 1, rng)])                     # far away: untouched
    off = np.array([0, len(votes)], np.int64)
    W = np.zeros((2, 352), np.float32)
    cb = Codebook(W, np.arange(3), np.zeros((2, 3)), np.ones(2), np.zeros(2), np.zeros(2), np.zeros((2, 7)), np.ones(2),
                  np.zeros((2, 3)), np.arange(2), np.ones(2))
    prm = default_params(bandwidth=0.3, radius_type=1, radius_factor=1.0, max_filter_type=2, single_object_mode=0,
                         ms_kernel=1)
    m = orc.Model(prm, cb)
    m.set_class_dimensions([0.5, 0.1], [0.5, 0.1])
    mx, moff, mi, mw = m.find_maxima(votes, off)
    assert len(mx) == 2
    # the merged class-1 candidate (80 votes) outweighs the class-0 maximum (50 votes) and sits between its two parts
    assert mx["class_id"].tolist() == [1, 1] and mx["n_votes"].tolist() == [80, 10]
    assert np.allclose(mx["position"][0], [0, 0, 0], atol=0.01) and np.allclose(mx["position"][1], [5, 5, 5], atol=0.01)
    assert np.isclose(mx["weight"].sum(), 1.0)
    assert sorted(mi[mx["vote_begin"][0]:mx["vote_begin"][0] + 80].tolist()) == list(range(50, 130))
    # without the filter: three class-1 maxima and the class-0 one
    prm.max_filter_type = 0
    m.set_params(prm)
    assert sorted(m.find_maxima(votes, off)[0]["n_votes"].tolist()) == [10, 40, 40, 50]
    # "Simple" with a per-class radius type uses the search distance of the class processed before the last (here
    # class 0's 0.5: MaximaHandler::m_radius as iFindMaxima left it): the class-0 maximum suppresses both neighbours
    prm.max_filter_type = 1
    m.set_params(prm)
    mx = m.find_maxima(votes, off)[0]
    assert sorted(zip(mx["class_id"].tolist(), mx["n_votes"].tolist())) == [(0, 50), (1, 10)]


def test_organized_normals_known_answer(orc):
    """Integral-image normals of an exactly planar organized cloud are the plane normal, pointing at the sensor; the
    10-pixel image border, NaN holes and the pixels next to a depth jump stay NaN (PCL's IntegralImageNormalEstimation,
    AVERAGE_3D_GRADIENT, border policy IGNORE; implicit_shape_model.cpp:948-966)."""
    H, W = 50, 70
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    nrm = np.array([0.2, -0.1, -1.0]); nrm /= np.linalg.norm(nrm)
    x, y = (u - W / 2) * 0.01, (v - H / 2) * 0.01
    z = (1.5 * nrm[2] - nrm[0] * x - nrm[1] * y) / nrm[2]   # plane n . p = 1.5 n_z: z = 1.5 on the optical axis
    xyz = np.stack([x, y, z], -1).astype(np.float32)
    n = orc.compute_normals_organized(xyz)
    inner = n[10:-10, 10:-10]
    assert np.isnan(n[:10]).all() and np.isnan(n[:, :10]).all() and np.isnan(n[-10:]).all() and np.isnan(n[:, -10:]).all()
    assert np.allclose(inner, nrm, atol=2e-5)
    xyz[25, 35] = np.nan
    xyz[30:, 50:, 2] += 0.5                      # a depth jump
    n = orc.compute_normals_organized(xyz)
    assert np.isnan(n[25, 35]).all() and np.isnan(n[24:27, 34:37]).all()
    assert np.isnan(n[30, 50]).all() and np.isnan(n[29, 49]).all()
    assert np.allclose(n[15, 15], nrm, atol=2e-5) and np.allclose(n[38, 58], nrm, atol=2e-5)


def _ransac_votes(rng, n_in, n_out, centre, cls, R=None, t=None, noise=0.0):
    """Votes of one maximum: n_in whose scene keypoint is the rigid image R kt + t of the training keypoint (+ noise),
    n_out whose scene keypoint is unrelated."""
    v = _votes([centre], n_in + n_out, cls, rng, sigma=0.01)
    kt = rng.uniform(-0.5, 0.5, size=(n_in + n_out, 3))
    if R is None:
        ang = 0.7
        R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    if t is None:
        t = np.array([0.3, -0.2, 0.1])
    ks = kt @ R.T + t + rng.normal(scale=noise, size=kt.shape) if noise > 0 else kt @ R.T + t
    ks[n_in:] = rng.uniform(-2, 2, size=(n_out, 3))
    v["keypoint_training"] = kt.astype(np.float32)
    v["keypoint"] = ks.astype(np.float32)
    return v


def test_ransac_vote_filtering_known_answer(orc):
    """Voting.RansacVoteFiltering (voting.cpp:110-127,356-433): only the votes whose (training keypoint -> scene
    keypoint) correspondence agrees with the dominant rigid transform stay in the maximum; a maximum whose transform is
    the identity, or that has no consistent transform at all, is dropped (as written in the reference)."""
    from pcdb200.structs import Codebook, default_params
    rng = np.random.default_rng(11)
    a = _ransac_votes(rng, 60, 40, (0, 0, 0), 0, noise=0.002)            # 60 % inliers
    b = _ransac_votes(rng, 50, 0, (3, 0, 0), 1, R=np.eye(3), t=np.zeros(3))  # identity transform: dropped
    c = _ransac_votes(rng, 0, 30, (0, 3, 0), 1)                          # no consistent transform
    votes = np.concatenate([a, b, c])
    off = np.array([0, len(votes)], np.int64)
    W = np.zeros((2, 352), np.float32)
    cb = Codebook(W, np.arange(3), np.zeros((2, 3)), np.ones(2), np.zeros(2), np.zeros(2), np.zeros((2, 7)), np.ones(2),
                  np.zeros((2, 3)), np.arange(2), np.ones(2))
    prm = default_params(bandwidth=0.3, single_object_mode=0, min_votes_threshold=4)
    m = orc.Model(prm, cb)
    assert sorted(m.find_maxima(votes, off)[0]["n_votes"].tolist()) == [30, 50, 100]
    prm.ransac_vote_filtering = 1
    prm.ransac_inlier_threshold = 0.02
    m.set_params(prm)
    mx, moff, mi, mw = m.find_maxima(votes, off)
    # the random maximum c may keep a few accidental inliers of its best 3-point fit, never a real consensus
    big = mx[mx["n_votes"] >= 10]
    assert len(big) == 1 and big["class_id"][0] == 0 and big["n_votes"][0] == 60
    members = mi[big["vote_begin"][0]:big["vote_begin"][0] + 60]
    assert sorted(members.tolist()) == list(range(60))            # exactly the inlier votes
    assert np.isclose(big["raw_weight"][0], mw[big["vote_begin"][0]:big["vote_begin"][0] + 60].sum(), rtol=1e-5)
    assert not np.any((mx["class_id"] == 1) & (mx["n_votes"] >= 40))  # the identity-pose maximum is gone
    # per-class threshold types scale the threshold by the learned class dimension (voting.cpp:112-122)
    prm.ransac_threshold_type = 1
    prm.ransac_inlier_threshold = 0.04
    m.set_params(prm)
    m.set_class_dimensions([0.5, 1.0], [2.0, 1.0])   # 0.04 * 0.5 = the same 0.02 for class 0
    mx2 = m.find_maxima(votes, off)[0]
    assert mx2[mx2["n_votes"] >= 10]["n_votes"].tolist() == [60]


def test_single_object_max_types_known_answer(orc, small_world):
    """SingleObjectMaxType VotingSpaceVotes collects every vote of a class at the centroid; ModelRadiusVotes those
    within the farthest point's distance; BandwidthVotes those within Voting.Bandwidth."""
    prm = small_world["prm"].copy()
    cb = small_world["cb"]
    xt, nt, rt, ot, _ = small_world["test"]
    x, n, r, o = xt[ot[0]:ot[1]], nt[ot[0]:ot[1]], rt[ot[0]:ot[1]], np.array([0, ot[1] - ot[0]], np.int64)
    counts = {}
    for mt in (0, 1, 2, 3):
        prm.single_object_mode, prm.single_object_max_type = 1, mt
        m = orc.Model(prm, cb)
        _, mx, moff = m.classify_batch(x, n, r, o)
        counts[mt] = mx
    centroid = x.astype(np.float32).sum(0, dtype=np.float32) / np.float32(len(x))
    for mt in (1, 2, 3):
        assert np.allclose(counts[mt]["position"], centroid, atol=1e-5)
        assert len(set(counts[mt]["class_id"].tolist())) == len(counts[mt])  # one maximum per voted class
    # (the farthest vote itself fails the strict d^2 < h^2 test of the radius search: VotingSpaceVotes misses it)
    assert counts[3]["n_votes"].sum() >= counts[1]["n_votes"].sum() > 0
    assert counts[2]["n_votes"].sum() >= counts[3]["n_votes"].sum() - len(counts[2])
