"""ctypes binding of the CPU oracle (oracle/pcd_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under point-cloud-donkey_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_ROOT, "point-cloud-donkey_b200"))
from pcdb200.structs import (  # noqa: E402
    MAXIMUM_DTYPE, VOTE_DTYPE, Codebook, Params, f32, i32, i64, ptr, u32,
)

_LIB = None
_REF = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "pcd_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_model_create.restype = C.c_void_p
    return _LIB


def ref_color_lib():
    """oracle/_ref/libref_color.so: the reference's own color_conversion.cpp (None if never built)."""
    global _REF
    so = os.path.join(_HERE, "_ref", "libref_color.so")
    if _REF is None and os.path.exists(so):
        _REF = C.CDLL(so)
    return _REF


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle error %d: %s" % (rc, lib().orc_last_error().decode()))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


F, I64, I32, U32 = C.c_float, C.c_int64, C.c_int32, C.c_uint32


def voxel_keypoints(xyz, rgb, cloud_off, leaf):
    xyz, rgb, cloud_off = f32(xyz), u32(rgb), i64(cloud_off)
    B = len(cloud_off) - 1
    cap = xyz.shape[0]
    kp = np.empty((cap, 3), np.float32)
    kr = np.empty(cap, np.uint32)
    off = np.empty(B + 1, np.int64)
    _check(lib().orc_voxel_keypoints(ptr(xyz, F), ptr(rgb, U32), ptr(cloud_off, I64), B, C.c_float(leaf),
                                     ptr(kp, F), ptr(kr, U32), ptr(off, I64), I64(cap)))
    n = int(off[-1])
    return kp[:n].copy(), kr[:n].copy(), off


def radius_neighbours(surf_xyz, surf_off, kp_xyz, kp_off, radius):
    surf_xyz, surf_off, kp_xyz, kp_off = f32(surf_xyz), i64(surf_off), f32(kp_xyz), i64(kp_off)
    B = len(surf_off) - 1
    Q = kp_xyz.shape[0]
    cap = max(1, int(sum(int(kp_off[b + 1] - kp_off[b]) * int(surf_off[b + 1] - surf_off[b]) for b in range(B))))
    noff = np.zeros(Q + 1, np.int64)
    idx = np.empty(cap, np.int32)
    d2 = np.empty(cap, np.float32)
    _check(lib().orc_radius_neighbours(ptr(surf_xyz, F), ptr(surf_off, I64), ptr(kp_xyz, F), ptr(kp_off, I64), B,
                                       C.c_double(radius), ptr(noff, I64), ptr(idx, I32), ptr(d2, F), I64(cap)))
    n = int(noff[-1])
    return noff, idx[:n].copy(), d2[:n].copy()


def shot_lrf(surf_xyz, surf_off, kp_xyz, kp_off, radius):
    surf_xyz, surf_off, kp_xyz, kp_off = f32(surf_xyz), i64(surf_off), f32(kp_xyz), i64(kp_off)
    out = np.empty((kp_xyz.shape[0], 9), np.float32)
    _check(lib().orc_shot_lrf(ptr(surf_xyz, F), ptr(surf_off, I64), ptr(kp_xyz, F), ptr(kp_off, I64),
                              len(surf_off) - 1, C.c_double(radius), ptr(out, F)))
    return out


def shot_describe(feature_type, surf_xyz, surf_normals, surf_rgb, surf_off, kp_xyz, kp_rgb, kp_lrf, kp_off, radius):
    surf_xyz, surf_normals, surf_rgb, surf_off = f32(surf_xyz), f32(surf_normals), u32(surf_rgb), i64(surf_off)
    kp_xyz, kp_rgb, kp_lrf, kp_off = f32(kp_xyz), u32(kp_rgb), f32(kp_lrf), i64(kp_off)
    D = 1344 if feature_type == 1 else 352
    out = np.empty((kp_xyz.shape[0], D), np.float32)
    _check(lib().orc_shot_describe(feature_type, ptr(surf_xyz, F), ptr(surf_normals, F), ptr(surf_rgb, U32),
                                   ptr(surf_off, I64), ptr(kp_xyz, F), ptr(kp_rgb, U32), ptr(kp_lrf, F),
                                   ptr(kp_off, I64), len(surf_off) - 1, C.c_double(radius), ptr(out, F)))
    return out


def compute_normals(prm, xyz, cloud_off):
    """ImplicitShapeModel::computeNormals for unorganized clouds: (normals (P,3), curvature (P,)), NaN where the
    reference yields NaN; uses prm.normal_radius / prm.consistent_normals_method."""
    xyz, cloud_off = f32(xyz), i64(cloud_off)
    nrm = np.empty((xyz.shape[0], 3), np.float32)
    curv = np.empty(xyz.shape[0], np.float32)
    _check(lib().orc_compute_normals(C.byref(prm), ptr(xyz, F), ptr(cloud_off, I64), len(cloud_off) - 1, ptr(nrm, F),
                                     ptr(curv, F)))
    return nrm, curv


def compute_normals_organized(xyz_hw3):
    """ImplicitShapeModel::computeNormals for an organized cloud (height x width x 3): IntegralImageNormalEstimation,
    AVERAGE_3D_GRADIENT, depth-change factor 0.02, smoothing size 10, normals towards the sensor origin."""
    xyz = f32(xyz_hw3)
    h, w = xyz.shape[0], xyz.shape[1]
    nrm = np.empty((h, w, 3), np.float32)
    _check(lib().orc_compute_normals_organized(ptr(xyz, F), w, h, ptr(nrm, F)))
    return nrm


def compute_features(prm, xyz, normals, rgb, cloud_off):
    xyz, normals, rgb, cloud_off = f32(xyz), f32(normals), u32(rgb), i64(cloud_off)
    B = len(cloud_off) - 1
    cap = xyz.shape[0]
    fx = np.empty((cap, 3), np.float32)
    fl = np.empty((cap, 9), np.float32)
    fd = np.empty((cap, prm.dim), np.float32)
    off = np.empty(B + 1, np.int64)
    _check(lib().orc_compute_features(C.byref(prm), ptr(xyz, F), ptr(normals, F), ptr(rgb, U32), ptr(cloud_off, I64),
                                      B, ptr(fx, F), ptr(fl, F), ptr(fd, F), ptr(off, I64), I64(cap)))
    n = int(off[-1])
    return fx[:n].copy(), fl[:n].copy(), fd[:n].copy(), off


def distance(a, b, dist_type):
    a, b = f32(a), f32(b)
    out = np.empty(a.shape[0], np.float32)
    _check(lib().orc_distance(ptr(a, F), ptr(b, F), I64(a.shape[0]), a.shape[1], dist_type, ptr(out, F)))
    return out


def rgb_to_lab_normalized(rgb):
    rgb = u32(rgb)
    out = np.empty((rgb.shape[0], 3), np.float32)
    _check(lib().orc_rgb_to_lab_normalized(ptr(rgb, U32), I64(rgb.shape[0]), ptr(out, F)))
    return out


def color_distance(lab, lab_ref):
    lab, lab_ref = f32(lab), f32(lab_ref)
    out = np.empty(lab.shape[0], np.float32)
    _check(lib().orc_color_distance(ptr(lab, F), ptr(lab_ref, F), I64(lab.shape[0]), ptr(out, F)))
    return out


def lab_luts():
    a, b = np.empty(256, np.float32), np.empty(4000, np.float32)
    _check(lib().orc_lab_luts(ptr(a, F), ptr(b, F)))
    return a, b


def aabb(xyz):
    xyz = f32(xyz)
    out = np.empty(10, np.float32)
    _check(lib().orc_aabb(ptr(xyz, F), I64(xyz.shape[0]), ptr(out, F)))
    return out


def train(prm, feat_xyz, feat_lrf, feat_desc, feat_off, cloud_class, cloud_instance, cloud_bbox10, n_classes):
    """ImplicitShapeModel::train + Codebook::activate on precomputed training features -> Codebook."""
    feat_xyz, feat_lrf, feat_desc, feat_off = f32(feat_xyz), f32(feat_lrf), f32(feat_desc), i64(feat_off)
    cloud_class, cloud_instance, cloud_bbox10 = u32(cloud_class), u32(cloud_instance), f32(cloud_bbox10)
    N, V = I64(0), I64(0)
    _check(lib().orc_train(C.byref(prm), ptr(feat_xyz, F), ptr(feat_lrf, F), ptr(feat_desc, F), ptr(feat_off, I64),
                           len(feat_off) - 1, ptr(cloud_class, U32), ptr(cloud_instance, U32), ptr(cloud_bbox10, F),
                           n_classes, C.byref(N), C.byref(V)))
    N, V, D = N.value, V.value, prm.dim
    words = np.empty((N, D), np.float32)
    vote_off = np.empty(N + 1, np.int64)
    vxyz = np.empty((V, 3), np.float32)
    vw = np.empty(V, np.float32)
    vc = np.empty(V, np.uint32)
    vi = np.empty(V, np.uint32)
    vb = np.empty((V, 7), np.float32)
    vcw = np.empty(V, np.float32)
    kpt = np.empty((N, 3), np.float32)
    ids = np.empty(N, np.int32)
    s2 = np.empty(n_classes, np.float32)
    _check(lib().orc_train_fetch(ptr(words, F), ptr(vote_off, I64), ptr(vxyz, F), ptr(vw, F), ptr(vc, U32),
                                 ptr(vi, U32), ptr(vb, F), ptr(vcw, F), ptr(kpt, F), ptr(ids, I32), ptr(s2, F),
                                 n_classes))
    return Codebook(words, vote_off, vxyz, vw, vc, vi, vb, vcw, kpt, ids, s2)


class Model:
    """Oracle-side model: same method names as pcdb200.Context so parity tests read symmetrically."""

    def __init__(self, prm: Params, cb: Codebook):
        self.prm = prm.copy()
        self.cb = cb
        self.h = C.c_void_p(lib().orc_model_create(
            C.byref(self.prm), ptr(cb.words, F), I64(cb.N), cb.D, ptr(cb.vote_off, I64), ptr(cb.vote_xyz, F),
            ptr(cb.vote_weight, F), ptr(cb.vote_class, U32), ptr(cb.vote_instance, U32), ptr(cb.vote_bbox, F),
            ptr(cb.vote_class_weight, F), ptr(cb.kp_train, F), ptr(cb.codeword_ids, I32),
            ptr(cb.codeword_weight, F), ptr(cb.sigma2, F), cb.n_classes))

    def set_params(self, prm):
        self.prm = prm.copy()
        lib().orc_model_set_params(self.h, C.byref(self.prm))

    def set_class_dimensions(self, first_dim, second_dim):
        """Voting::m_dimensions_map: per class (mean object radius, mean median bounding-box side)."""
        a, b = f32(first_dim), f32(second_dim)
        lib().orc_model_set_class_dimensions(self.h, ptr(a, F), ptr(b, F), len(a))

    def close(self):
        if self.h:
            lib().orc_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn(self, queries, k=None, dist_type=None, mode=0):
        queries = f32(queries)
        k = self.prm.knn_k if k is None else k
        dist_type = self.prm.distance_type if dist_type is None else dist_type
        Q = queries.shape[0]
        idx = np.empty((Q, k), np.int32)
        dist = np.empty((Q, k), np.float32)
        cnt = np.empty(Q, np.int32)
        _check(lib().orc_knn(self.h, ptr(queries, F), I64(Q), k, dist_type, mode, ptr(idx, I32), ptr(dist, F),
                             ptr(cnt, I32)))
        return idx, dist, cnt

    def cast_votes(self, feat_xyz, feat_lrf, feat_off, knn_idx, knn_dist, knn_count):
        feat_xyz, feat_lrf, feat_off = f32(feat_xyz), f32(feat_lrf), i64(feat_off)
        knn_idx, knn_dist, knn_count = i32(knn_idx), f32(knn_dist), i32(knn_count)
        B = len(feat_off) - 1
        k = knn_idx.shape[1]
        maxv = int(np.max(np.diff(self.cb.vote_off))) if self.cb.N else 0
        cap = max(1, int(knn_count.sum()) * max(1, maxv))
        votes = np.zeros(cap, VOTE_DTYPE)
        voff = np.empty(B + 1, np.int64)
        _check(lib().orc_cast_votes(self.h, ptr(feat_xyz, F), ptr(feat_lrf, F), ptr(feat_off, I64), B,
                                    ptr(knn_idx, I32), ptr(knn_dist, F), ptr(knn_count, I32), k,
                                    votes.ctypes.data_as(C.c_void_p), ptr(voff, I64), I64(cap)))
        return votes[: int(voff[-1])].copy(), voff

    def find_maxima(self, votes, vote_off):
        votes = np.ascontiguousarray(votes, dtype=VOTE_DTYPE)
        vote_off = i64(vote_off)
        B = len(vote_off) - 1
        cap = max(1, votes.shape[0] * 2 + 16)
        mx = np.zeros(cap, MAXIMUM_DTYPE)
        moff = np.empty(B + 1, np.int64)
        _check(lib().orc_find_maxima(self.h, votes.ctypes.data_as(C.c_void_p), ptr(vote_off, I64), B,
                                     mx.ctypes.data_as(C.c_void_p), ptr(moff, I64), I64(cap)))
        n = I64(0)
        mi = np.empty(max(1, votes.shape[0] * 4 + 16), np.int64)
        mw = np.empty(mi.shape[0], np.float32)
        _check(lib().orc_get_maximum_votes(ptr(mi, I64), ptr(mw, F), I64(mi.shape[0]), C.byref(n)))
        return mx[: int(moff[-1])].copy(), moff, mi[: n.value].copy(), mw[: n.value].copy()

    def set_approximate(self, trees=4, checks=128, seed=1):
        """FLANN-like randomized kd-forest for the activation (the reference's default: KDTreeIndexParams(4),
        SearchParams(128)).  COST STAND-IN ONLY: neighbour sets depend on random draws and are never used for parity.
        trees=0 switches back to the exact search.  Returns the index build time in ms (the "flann" bucket)."""
        f = lib().orc_model_set_approximate
        f.restype = C.c_double
        return f(self.h, int(trees), int(checks), C.c_uint32(seed))

    def classify_batch(self, xyz, normals, rgb, cloud_off, want_maxima=True):
        xyz, normals, rgb, cloud_off = f32(xyz), f32(normals), u32(rgb), i64(cloud_off)
        B = len(cloud_off) - 1
        labels = np.empty(B, np.int32)
        cap = max(16, xyz.shape[0])
        mx = np.zeros(cap, MAXIMUM_DTYPE) if want_maxima else None
        moff = np.zeros(B + 1, np.int64) if want_maxima else None
        times = np.zeros(7, np.float64)
        counts = np.zeros(8, np.int64)
        _check(lib().orc_classify_batch(self.h, ptr(xyz, F), ptr(normals, F), ptr(rgb, U32), ptr(cloud_off, I64), B,
                                        ptr(labels, I32), None if mx is None else mx.ctypes.data_as(C.c_void_p),
                                        ptr(moff, I64), I64(cap), ptr(times, C.c_double), ptr(counts, I64)))
        self.last_times = dict(zip("complete features keypoints normals flann voting maxima".split(), times))
        self.last_counts = dict(zip("keypoints features nbr_lrf nbr_shot votes".split(), counts[:5]))
        if want_maxima:
            return labels, mx[: int(moff[-1])].copy(), moff
        return labels, None, None


def merge_topk(cand_idx, cand_dist):
    cand_idx, cand_dist = i32(cand_idx), f32(cand_dist)
    S, Q, k = cand_idx.shape
    idx = np.empty((Q, k), np.int32)
    dist = np.empty((Q, k), np.float32)
    _check(lib().orc_merge_topk(ptr(cand_idx, I32), ptr(cand_dist, F), S, I64(Q), k, ptr(idx, I32), ptr(dist, F)))
    return idx, dist


def knn_exact_pruned(model, queries, k=1, dist_type=0, chunk=320):
    """The exact search of Model.knn by a cheaper route, for label checks over many clouds at bench scale: a BLAS sgemm
    (squared-L2 expansion; for chi^2 the Hellinger sandwich H^2 <= chi^2 <= 2 H^2 on sqrt rows) proposes every row that
    can still be among the k best given a rigorous bound on the fp32 rounding, and the FLANN-order functor
    (orc_distance) decides among them by (distance, row).  Same result as the linear scan by construction; the CPU test
    suite checks it against Model.knn.  (k+1)-th neighbour / ratio test are not handled: callers pass plain k."""
    cb = model.cb
    W = cb.words
    Q = f32(queries)
    D = W.shape[1]
    chi = dist_type == 1
    if chi and ((W < 0).any() or (Q < 0).any()):
        return model.knn(Q, k=k, dist_type=dist_type)
    cache = getattr(model, "_pruned_cache", None)
    if cache is None or cache[0] != chi:
        Wt = np.sqrt(W) if chi else W
        wn = (Wt.astype(np.float64) ** 2).sum(1)
        cache = (chi, np.ascontiguousarray(Wt.T), wn.astype(np.float32), float(np.sqrt(wn.max())), float(wn.max()))
        model._pruned_cache = cache
    _, WtT, wn32, wmax, wn_max = cache
    idx_out = np.full((len(Q), k), -1, np.int32)
    dst_out = np.full((len(Q), k), np.nan, np.float32)
    cnt_out = np.zeros(len(Q), np.int32)
    u = 2.0 ** -24
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    except Exception:
        limit = None
    for s in range(0, len(Q), chunk):
        q = Q[s:s + chunk]
        qt = np.sqrt(q) if chi else q
        qn = (qt.astype(np.float64) ** 2).sum(1)
        a = qt @ WtT                                               # fp32 sgemm, [chunk, N]
        a *= np.float32(-2.0)
        a += wn32[None, :]                                         # |c|^2 - 2 q.c (the per-query |q|^2 is a constant)
        # |fl(q.c) - q.c| <= gamma_D |q||c| for any summation order (gamma_D = D u / (1 - D u)), doubled in the distance;
        # three fp32 roundings of values up to |c|^2 + 2|q||c|; the fp32 sqrt adds 2^-23 (|s|+|t|)^2
        eps = (2.0 * (D * u / (1 - D * u)) * np.sqrt(qn) * wmax * 1.01 + 4 * u * (wn_max + 2 * np.sqrt(qn) * wmax)
               + (2.0 ** -22) * (np.sqrt(qn) + wmax) ** 2)
        kth = a.min(axis=1) if k == 1 else np.partition(a, k - 1, axis=1)[:, k - 1]
        for i in range(len(q)):
            if chi:
                # U = k-th smallest exact chi^2 among the k approx-nearest rows bounds the k-th smallest overall
                near = np.argpartition(a[i], k - 1)[:k] if k > 1 else np.array([int(np.argmin(a[i]))])
                U = float(np.sort(distance(W[near], np.repeat(q[i:i + 1], len(near), 0), 1))[k - 1])
                thr = U * (1 + 1.01 * (D + 8) * u) - qn[i] + eps[i]
            else:
                thr = float(kth[i]) + 2 * eps[i]
            cand = np.nonzero(a[i] <= np.float32(thr) + np.float32(1e-30))[0]
            ex = distance(W[cand], np.repeat(q[i:i + 1], len(cand), 0), dist_type)  # functor(codeword, query)
            order = np.lexsort((cand, ex))[:k]
            n = len(order)
            idx_out[s + i, :n] = cand[order]
            dst_out[s + i, :n] = ex[order]
            cnt_out[s + i] = n
    del limit
    return idx_out, dst_out, cnt_out


def classify_batch_pruned(model, xyz, normals, rgb, cloud_off):
    """Model.classify_batch (labels only) with knn_exact_pruned as the activation: features, vote casting and the maxima
    search are the oracle's own stage functions.  For parameter sets without the distance-ratio test."""
    prm = model.prm
    assert not prm.use_distance_ratio
    cloud_off = i64(cloud_off)
    fx, fl, fd, foff = compute_features(prm, xyz, normals, rgb, cloud_off)
    idx, dst, cnt = knn_exact_pruned(model, fd, k=prm.knn_k, dist_type=prm.distance_type)
    votes, voff = model.cast_votes(fx, fl, foff, idx, dst, cnt)
    mx, moff, _, _ = model.find_maxima(votes, voff)
    B = len(cloud_off) - 1
    return np.array([int(mx["class_id"][moff[b]]) if moff[b + 1] > moff[b] else -1 for b in range(B)], np.int32)
