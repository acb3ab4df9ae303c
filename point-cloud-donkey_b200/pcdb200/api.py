"""ctypes binding of libpcdb200.so (include/pcdb200.h).

The library is the product; this module only marshals numpy arrays through the C-ABI.  There is no fallback of
any kind: a missing library or a machine without an sm_100 GPU raises.
"""
import ctypes as C
import os

import numpy as np

from .structs import (MAXIMUM_DTYPE, VOTE_DTYPE, Codebook, Params, Stats, f32, i32, i64, ptr, u32)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libpcdb200.so")
_LIB = None

F, I64, I32, U32, D = C.c_float, C.c_int64, C.c_int32, C.c_uint32, C.c_double

# every symbol include/pcdb200.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "pcdb_abi_version", "pcdb_create", "pcdb_destroy", "pcdb_last_error", "pcdb_default_params", "pcdb_set_params",
    "pcdb_set_stream", "pcdb_set_codebook", "pcdb_voxel_keypoints", "pcdb_radius_neighbours", "pcdb_shot_lrf",
    "pcdb_shot_describe", "pcdb_compute_normals", "pcdb_compute_features", "pcdb_knn", "pcdb_distance_pairs", "pcdb_cast_votes", "pcdb_find_maxima",
    "pcdb_get_maximum_votes", "pcdb_get_votes", "pcdb_get_last_sizes", "pcdb_get_maxima", "pcdb_compute_normals_organized", "pcdb_classify_batch", "pcdb_classify_batch_d", "pcdb_merge_topk",
    "pcdb_get_stats", "pcdb_reset_stats", "pcdb_comm_unique_id", "pcdb_comm_init", "pcdb_comm_destroy", "pcdb_comm_info",
    "pcdb_set_codebook_sharded", "pcdb_comm_shard_keypoints", "pcdb_set_class_dimensions",
]
COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it and hands the bytes to the other ranks)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    rc = lib().pcdb_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise PcdbError(rc, "NCCL unavailable (libnccl.so.2 could not be loaded)")
    return bytes(buf)


class PcdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pcdb error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libpcdb200.so is not built (%s); run __graft_entry__.build() — there is no CPU fallback" % LIB_PATH)
        _LIB = C.CDLL(LIB_PATH)
        _LIB.pcdb_last_error.restype = C.c_char_p
        _LIB.pcdb_last_error.argtypes = [C.c_void_p]
        _LIB.pcdb_destroy.argtypes = [C.c_void_p]
    return _LIB


class Context:
    """One pcdb_ctx (one GPU).  Method names follow the reference hooks (one per stage) so parity tests read symmetrically."""

    def __init__(self, prm: Params = None, cb: Codebook = None, device=0, row_base=0):
        self.h = C.c_void_p()
        rc = lib().pcdb_create(C.byref(self.h), int(device))
        if rc != 0:
            raise PcdbError(rc, lib().pcdb_last_error(None).decode())
        self.prm = None
        self.cb = None
        if prm is not None:
            self.set_params(prm)
        if cb is not None:
            self.set_codebook(cb, row_base)

    def _check(self, rc):
        if rc != 0:
            raise PcdbError(rc, lib().pcdb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            lib().pcdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, prm):
        self.prm = prm.copy()
        self._check(lib().pcdb_set_params(self.h, C.byref(self.prm)))

    def set_class_dimensions(self, first_dim, second_dim):
        """Voting::m_dimensions_map: per class (mean object radius, mean median bounding-box side)."""
        a, b = f32(first_dim), f32(second_dim)
        self._check(lib().pcdb_set_class_dimensions(self.h, ptr(a, F), ptr(b, F), len(a)))

    def set_stream(self, cuda_stream_ptr):
        self._check(lib().pcdb_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_codebook(self, cb: Codebook, row_base=0):
        self.cb = cb
        self._check(lib().pcdb_set_codebook(
            self.h, ptr(cb.words, F), I64(cb.N), cb.D, ptr(cb.vote_off, I64), ptr(cb.vote_xyz, F),
            ptr(cb.vote_weight, F), ptr(cb.vote_class, U32), ptr(cb.vote_instance, U32), ptr(cb.vote_bbox, F),
            ptr(cb.vote_class_weight, F), ptr(cb.kp_train, F), ptr(cb.codeword_ids, I32),
            ptr(cb.codeword_weight, F), ptr(cb.sigma2, F), cb.n_classes, I64(row_base)))

    # ---- multi-GPU -------------------------------------------------------------------------------------------
    def comm_init(self, rank, n_ranks, unique_id):
        """Collective: ncclCommInitRank on this context's device."""
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._check(lib().pcdb_comm_init(self.h, int(rank), int(n_ranks), buf))

    def comm_destroy(self):
        self._check(lib().pcdb_comm_destroy(self.h))

    def comm_info(self):
        r, n, v = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(lib().pcdb_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v)))
        return {"rank": r.value, "n_ranks": n.value, "nccl_version": v.value}

    def comm_shard_keypoints(self, enable=True):
        self._check(lib().pcdb_comm_shard_keypoints(self.h, 1 if enable else 0))

    def set_codebook_sharded(self, cb: Codebook, row_lo, row_hi):
        """Descriptor rows [row_lo, row_hi) on this rank, complete vote tables.  `cb` may hold all
        rows (the shard is sliced out here) — only the slice is uploaded."""
        self.cb = cb
        words = np.ascontiguousarray(cb.words[row_lo:row_hi])
        self._check(lib().pcdb_set_codebook_sharded(
            self.h, ptr(words, F), I64(row_lo), I64(row_hi), I64(cb.N), cb.D, ptr(cb.vote_off, I64),
            ptr(cb.vote_xyz, F), ptr(cb.vote_weight, F), ptr(cb.vote_class, U32), ptr(cb.vote_instance, U32),
            ptr(cb.vote_bbox, F), ptr(cb.vote_class_weight, F), ptr(cb.kp_train, F), ptr(cb.codeword_ids, I32),
            ptr(cb.codeword_weight, F), ptr(cb.sigma2, F), cb.n_classes))

    # ---- stage-level entry points ---------------------------------------------------------------------------
    def voxel_keypoints(self, xyz, rgb, cloud_off, leaf):
        xyz, rgb, cloud_off = f32(xyz), u32(rgb), i64(cloud_off)
        B = len(cloud_off) - 1
        cap = max(1, xyz.shape[0])
        kp = np.empty((cap, 3), np.float32)
        kr = np.empty(cap, np.uint32)
        off = np.empty(B + 1, np.int64)
        self._check(lib().pcdb_voxel_keypoints(self.h, ptr(xyz, F), ptr(rgb, U32), ptr(cloud_off, I64), B,
                                               C.c_float(leaf), ptr(kp, F), ptr(kr, U32), ptr(off, I64), I64(cap)))
        n = int(off[-1])
        return kp[:n].copy(), kr[:n].copy(), off

    def radius_neighbours(self, surf_xyz, surf_off, kp_xyz, kp_off, radius):
        surf_xyz, surf_off, kp_xyz, kp_off = f32(surf_xyz), i64(surf_off), f32(kp_xyz), i64(kp_off)
        B = len(surf_off) - 1
        Q = kp_xyz.shape[0]
        cap = max(1, int(sum(int(kp_off[b + 1] - kp_off[b]) * int(surf_off[b + 1] - surf_off[b]) for b in range(B))))
        noff = np.zeros(Q + 1, np.int64)
        idx = np.empty(cap, np.int32)
        d2 = np.empty(cap, np.float32)
        self._check(lib().pcdb_radius_neighbours(self.h, ptr(surf_xyz, F), ptr(surf_off, I64), ptr(kp_xyz, F),
                                                 ptr(kp_off, I64), B, D(radius), ptr(noff, I64), ptr(idx, I32),
                                                 ptr(d2, F), I64(cap)))
        n = int(noff[-1])
        return noff, idx[:n].copy(), d2[:n].copy()

    def shot_lrf(self, surf_xyz, surf_off, kp_xyz, kp_off, radius):
        surf_xyz, surf_off, kp_xyz, kp_off = f32(surf_xyz), i64(surf_off), f32(kp_xyz), i64(kp_off)
        out = np.empty((kp_xyz.shape[0], 9), np.float32)
        self._check(lib().pcdb_shot_lrf(self.h, ptr(surf_xyz, F), ptr(surf_off, I64), ptr(kp_xyz, F),
                                        ptr(kp_off, I64), len(surf_off) - 1, D(radius), ptr(out, F)))
        return out

    def shot_describe(self, feature_type, surf_xyz, surf_normals, surf_rgb, surf_off, kp_xyz, kp_rgb, kp_lrf, kp_off,
                      radius):
        surf_xyz, surf_normals, surf_rgb, surf_off = f32(surf_xyz), f32(surf_normals), u32(surf_rgb), i64(surf_off)
        kp_xyz, kp_rgb, kp_lrf, kp_off = f32(kp_xyz), u32(kp_rgb), f32(kp_lrf), i64(kp_off)
        dim = 1344 if feature_type == 1 else 352
        out = np.empty((kp_xyz.shape[0], dim), np.float32)
        self._check(lib().pcdb_shot_describe(self.h, feature_type, ptr(surf_xyz, F), ptr(surf_normals, F),
                                             ptr(surf_rgb, U32), ptr(surf_off, I64), ptr(kp_xyz, F), ptr(kp_rgb, U32),
                                             ptr(kp_lrf, F), ptr(kp_off, I64), len(surf_off) - 1, D(radius),
                                             ptr(out, F)))
        return out

    def compute_normals(self, xyz, cloud_off):
        """ImplicitShapeModel::computeNormals: (normals (P,3), curvature (P,)) with the context's NormalRadius /
        ConsistentNormalsMethod."""
        xyz, cloud_off = f32(xyz), i64(cloud_off)
        nrm = np.empty((xyz.shape[0], 3), np.float32)
        curv = np.empty(xyz.shape[0], np.float32)
        self._check(lib().pcdb_compute_normals(self.h, ptr(xyz, F), ptr(cloud_off, I64), len(cloud_off) - 1,
                                               ptr(nrm, F), ptr(curv, F)))
        return nrm, curv

    def compute_normals_organized(self, xyz_hw3):
        """Organized branch of ImplicitShapeModel::computeNormals: (height, width, 3) points -> (height, width, 3)
        normals (IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT)."""
        xyz = f32(xyz_hw3)
        h, w = xyz.shape[0], xyz.shape[1]
        nrm = np.empty((h, w, 3), np.float32)
        self._check(lib().pcdb_compute_normals_organized(self.h, ptr(xyz, F), w, h, ptr(nrm, F)))
        return nrm

    def compute_features(self, xyz, normals, rgb, cloud_off):
        xyz, normals, rgb, cloud_off = f32(xyz), f32(normals), u32(rgb), i64(cloud_off)
        B = len(cloud_off) - 1
        cap = max(1, xyz.shape[0])
        fx = np.empty((cap, 3), np.float32)
        fl = np.empty((cap, 9), np.float32)
        fd = np.empty((cap, self.prm.dim), np.float32)
        off = np.empty(B + 1, np.int64)
        self._check(lib().pcdb_compute_features(self.h, ptr(xyz, F), ptr(normals, F), ptr(rgb, U32),
                                                ptr(cloud_off, I64), B, ptr(fx, F), ptr(fl, F), ptr(fd, F),
                                                ptr(off, I64), I64(cap)))
        n = int(off[-1])
        return fx[:n].copy(), fl[:n].copy(), fd[:n].copy(), off

    def knn(self, queries, k=None, dist_type=None, mode=0):
        queries = f32(queries)
        k = self.prm.knn_k if k is None else k
        dist_type = self.prm.distance_type if dist_type is None else dist_type
        Q = queries.shape[0]
        idx = np.empty((Q, k), np.int32)
        dist = np.empty((Q, k), np.float32)
        cnt = np.empty(Q, np.int32)
        self._check(lib().pcdb_knn(self.h, ptr(queries, F), I64(Q), k, dist_type, mode, ptr(idx, I32), ptr(dist, F),
                                   ptr(cnt, I32)))
        return idx, dist, cnt

    def distance_pairs(self, a, b, dist_type):
        a, b = f32(a), f32(b)
        out = np.empty(a.shape[0], np.float32)
        self._check(lib().pcdb_distance_pairs(self.h, ptr(a, F), ptr(b, F), I64(a.shape[0]), a.shape[1], dist_type,
                                              ptr(out, F)))
        return out

    def cast_votes(self, feat_xyz, feat_lrf, feat_off, knn_idx, knn_dist, knn_count):
        feat_xyz, feat_lrf, feat_off = f32(feat_xyz), f32(feat_lrf), i64(feat_off)
        knn_idx, knn_dist, knn_count = i32(knn_idx), f32(knn_dist), i32(knn_count)
        B = len(feat_off) - 1
        k = knn_idx.shape[1]
        maxv = int(np.max(np.diff(self.cb.vote_off))) if self.cb.N else 0
        cap = max(1, int(knn_count.sum()) * max(1, maxv))
        votes = np.zeros(cap, VOTE_DTYPE)
        voff = np.empty(B + 1, np.int64)
        self._check(lib().pcdb_cast_votes(self.h, ptr(feat_xyz, F), ptr(feat_lrf, F), ptr(feat_off, I64), B,
                                          ptr(knn_idx, I32), ptr(knn_dist, F), ptr(knn_count, I32), k,
                                          votes.ctypes.data_as(C.c_void_p), ptr(voff, I64), I64(cap)))
        return votes[: int(voff[-1])].copy(), voff

    def find_maxima(self, votes, vote_off):
        votes = np.ascontiguousarray(votes, dtype=VOTE_DTYPE)
        vote_off = i64(vote_off)
        B = len(vote_off) - 1
        cap = max(1, votes.shape[0] + 16)
        mx = np.zeros(cap, MAXIMUM_DTYPE)
        moff = np.empty(B + 1, np.int64)
        self._check(lib().pcdb_find_maxima(self.h, votes.ctypes.data_as(C.c_void_p), ptr(vote_off, I64), B,
                                           mx.ctypes.data_as(C.c_void_p), ptr(moff, I64), I64(cap)))
        mi, mw = self.get_maximum_votes()
        return mx[: int(moff[-1])].copy(), moff, mi, mw

    def get_maximum_votes(self):
        n = I64(0)
        rc = lib().pcdb_get_maximum_votes(self.h, None, None, I64(0), C.byref(n))
        if n.value == 0:
            return np.empty(0, np.int64), np.empty(0, np.float32)
        mi = np.empty(n.value, np.int64)
        mw = np.empty(n.value, np.float32)
        self._check(lib().pcdb_get_maximum_votes(self.h, ptr(mi, I64), ptr(mw, F), I64(n.value), C.byref(n)))
        del rc
        return mi, mw

    def get_votes(self, B, capacity):
        votes = np.zeros(max(1, capacity), VOTE_DTYPE)
        voff = np.empty(B + 1, np.int64)
        self._check(lib().pcdb_get_votes(self.h, votes.ctypes.data_as(C.c_void_p), ptr(voff, I64), I64(capacity)))
        return votes[: int(voff[-1])].copy(), voff

    def classify_batch(self, xyz, normals, rgb, cloud_off, want_maxima=True):
        """ImplicitShapeModel::detect for B clouds (host buffers in, host buffers out)."""
        xyz, normals, rgb, cloud_off = f32(xyz), f32(normals), u32(rgb), i64(cloud_off)
        B = len(cloud_off) - 1
        labels = np.empty(B, np.int32)
        cap = max(16, xyz.shape[0])
        mx = np.zeros(cap, MAXIMUM_DTYPE) if want_maxima else None
        moff = np.zeros(B + 1, np.int64) if want_maxima else None
        times = np.zeros(7, np.float64)
        self._check(lib().pcdb_classify_batch(
            self.h, ptr(xyz, F), ptr(normals, F), ptr(rgb, U32), ptr(cloud_off, I64), B, ptr(labels, I32),
            None if mx is None else mx.ctypes.data_as(C.c_void_p), ptr(moff, I64), I64(cap), ptr(times, D)))
        self.last_times = dict(zip("complete features keypoints normals flann voting maxima".split(), times))
        if want_maxima:
            return labels, mx[: int(moff[-1])].copy(), moff
        return labels, None, None

    def classify_batch_device(self, xyz_ptr, normals_ptr, rgb_ptr, cloud_off, labels_ptr):
        """Device-resident inputs (raw device pointers, e.g. torch.Tensor.data_ptr()); labels stay on the device."""
        cloud_off = i64(cloud_off)
        self._check(lib().pcdb_classify_batch_d(self.h, C.c_void_p(xyz_ptr),
                                                C.c_void_p(normals_ptr) if normals_ptr else None,
                                                C.c_void_p(rgb_ptr) if rgb_ptr else None, ptr(cloud_off, I64),
                                                len(cloud_off) - 1, C.c_void_p(labels_ptr)))

    def merge_topk(self, cand_idx, cand_dist):
        cand_idx, cand_dist = i32(cand_idx), f32(cand_dist)
        S, Q, k = cand_idx.shape
        idx = np.empty((Q, k), np.int32)
        dist = np.empty((Q, k), np.float32)
        self._check(lib().pcdb_merge_topk(self.h, ptr(cand_idx, I32), ptr(cand_dist, F), S, I64(Q), k, ptr(idx, I32),
                                          ptr(dist, F)))
        return idx, dist

    def stats(self):
        s = Stats()
        self._check(lib().pcdb_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def reset_stats(self):
        self._check(lib().pcdb_reset_stats(self.h))
