"""Generates the committed golden fixtures.  Run in the build container only (needs /root/reference for the colour
pin and cv2 for the FLANN pin):   python tests/golden/make_golden.py

lab_golden.npz    RGB -> normalised CIELab + colour distance computed by the REFERENCE's own
                  third_party/pcl_color_conversion/color_conversion.cpp (compiled into oracle/_ref by oracle/Makefile).
flann_golden.npz  squared-L2 exact nearest neighbours + distances computed by a real FLANN build
                  (cv2.flann_Index, linear index = exhaustive search with FLANN's own distance functors).
lzf_golden.npz    a binary_compressed PCD payload compressed by the REFERENCE's vendored liblzf 3.6
                  (third_party/liblzf-3.6/lzf_c.c, compiled into oracle/_ref/libref_lzf.so): pins the LZF decoder of
                  host/io_formats.h against the real encoder.
path_golden.npz   REGRESSION fixture of the whole path, produced by the oracle itself (not a pin against the reference:
                  nothing in the reference produces vectors here): keypoints, LRFs, descriptors, activation, votes,
                  maxima and labels of a small seeded world (pcdb200.synth).  It freezes today's oracle, so that a later
                  edit of the oracle or of a kernel that moves any stage shows up against a committed file, on the CPU
                  (tests/test_oracle.py) and on the GPU (tests/test_gpu_parity.py).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_lab():
    ref = orc.ref_color_lib()
    assert ref is not None, "oracle/_ref/libref_color.so missing: run make -C oracle with /root/reference mounted"
    rng = np.random.default_rng(7)
    rgb = np.concatenate([
        rng.integers(0, 1 << 24, 4096, dtype=np.uint32),
        np.array([0, 0xFFFFFF, 0xFF0000, 0x00FF00, 0x0000FF, 0x808080, 0x010101, 0xFEFEFE], np.uint32),
    ]).astype(np.uint32)
    lab = np.empty((rgb.shape[0], 3), np.float32)
    ref.ref_rgb_to_lab_normalized(rgb.ctypes.data_as(C.c_void_p), C.c_int64(rgb.shape[0]), lab.ctypes.data_as(C.c_void_p))
    perm = rng.permutation(rgb.shape[0])
    dist = np.empty(rgb.shape[0], np.float32)
    lab_ref = np.ascontiguousarray(lab[perm])
    ref.ref_color_distance(lab.ctypes.data_as(C.c_void_p), lab_ref.ctypes.data_as(C.c_void_p), C.c_int64(rgb.shape[0]),
                           dist.ctypes.data_as(C.c_void_p))
    srgb = np.empty(256, np.float32)
    sxyz = np.empty(4000, np.float32)
    ref.ref_lab_luts(srgb.ctypes.data_as(C.c_void_p), sxyz.ctypes.data_as(C.c_void_p))
    np.savez_compressed(os.path.join(HERE, "lab_golden.npz"), rgb=rgb, lab=lab, perm=perm.astype(np.int64), dist=dist,
                        srgb_lut=srgb, sxyz_lut=sxyz)
    print("lab_golden: %d colours" % rgb.shape[0])


def make_flann():
    import cv2
    rng = np.random.default_rng(11)
    out = {}
    for name, D in (("d352", 352), ("d1344", 1344), ("d30", 30)):
        # SHOT-like rows: sparse non-negative, L2-normalised
        base = rng.random((600, D), dtype=np.float32) ** 4
        base *= rng.random((600, D)) < 0.3
        base /= np.maximum(np.linalg.norm(base, axis=1, keepdims=True), 1e-12)
        base = base.astype(np.float32)
        q = base[:64] + 0.02 * rng.random((64, D), dtype=np.float32)
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
        words = np.ascontiguousarray(base[64:])
        # this OpenCV build exposes FLANN's L2 functor for float data (distance id 1); its chi^2 functor is not
        # compiled in, so chi^2 stays unpinned (oracle header says so)
        index = cv2.flann_Index(words, dict(algorithm=0), 1)  # FLANN_INDEX_LINEAR, FLANN_DIST_L2
        idx, d = index.knnSearch(q, 3, params=dict(checks=-1))
        out[f"{name}_l2_idx"] = idx.astype(np.int32)
        out[f"{name}_l2_dist"] = d.astype(np.float32)
        out[f"{name}_words"] = words
        out[f"{name}_queries"] = q
    np.savez_compressed(os.path.join(HERE, "flann_golden.npz"), **out)
    print("flann_golden written")


def make_lzf():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_lzf.so")
    assert os.path.exists(so), "oracle/_ref/libref_lzf.so missing: run make -C oracle with /root/reference mounted"
    lz = C.CDLL(so)
    lz.lzf_compress.restype = C.c_uint
    lz.lzf_compress.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    from pcdb200 import synth
    xyz, nrm, rgb, _ = synth.make_clouds([4], [123], 3000)
    rgb[:800] = rgb[0]
    n = len(xyz)
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2], rgb.astype(np.uint32), nrm[:, 0], nrm[:, 1], nrm[:, 2], np.zeros(n, np.float32)]
    raw = b"".join(np.ascontiguousarray(c).tobytes() for c in cols)   # PCD binary_compressed: structure of arrays
    out = C.create_string_buffer(len(raw) * 2)
    m = lz.lzf_compress(raw, len(raw), out, len(raw) * 2)
    assert 0 < m < len(raw)
    np.savez_compressed(os.path.join(HERE, "lzf_golden.npz"), raw=np.frombuffer(raw, np.uint8),
                        compressed=np.frombuffer(out.raw[:m], np.uint8), points=np.int64(n))
    print("lzf_golden: %d -> %d bytes by the reference's liblzf" % (len(raw), m))


def path_world():
    """Inputs of path_golden.npz, regenerated from seeds by the tests (only the outputs are stored)."""
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
    from pcdb200 import synth
    prm = synth.workload_params("c2", knn_k=2, single_object_mode=0, min_votes_threshold=2)
    tr_cls = [c for c in range(3) for _ in range(2)]
    train = synth.make_clouds(tr_cls, [31 + i for i in range(len(tr_cls))], 1200)
    te_cls = [0, 1, 2]
    test = synth.make_clouds(te_cls, [77 + i for i in range(len(te_cls))], 1200)
    return prm, tr_cls, train, te_cls, test


def make_path():
    prm, tr_cls, (xyz, nrm, rgb, off), te_cls, (xt, nt, rt, ot) = path_world()
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = orc.train(prm, fx, fl, fd, foff, tr_cls, list(range(len(tr_cls))), bb, 3)
    m = orc.Model(prm, cb)
    tx, tl, td, toff = orc.compute_features(prm, xt, nt, rt, ot)
    idx, dist, cnt = m.knn(td, k=2)
    votes, voff = m.cast_votes(tx, tl, toff, idx, dist, cnt)
    labels, mx, moff = m.classify_batch(xt, nt, rt, ot)
    np.savez_compressed(os.path.join(HERE, "path_golden.npz"), train_feat_off=foff, codebook_words=cb.words,
                        codebook_sigma2=cb.sigma2, feat_xyz=tx, feat_lrf=tl, feat_desc=td, feat_off=toff, knn_idx=idx,
                        knn_dist=dist, knn_cnt=cnt, vote_off=voff, votes=votes.view(np.uint8).reshape(-1, 80),
                        labels=labels, maxima=mx.view(np.uint8).reshape(len(mx), -1), maxima_off=moff)
    print("path_golden: %d test features, %d votes, %d maxima, labels %s" % (len(tx), len(votes), len(mx), labels))


if __name__ == "__main__":
    if os.path.isdir("/root/reference"):
        make_lab()
        make_flann()
        make_lzf()
    make_path()
