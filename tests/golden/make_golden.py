"""Generates the committed golden fixtures.  Run in the build container only (needs /root/reference for the colour
pin and cv2 for the FLANN pin):   python tests/golden/make_golden.py

lab_golden.npz    RGB -> normalised CIELab + colour distance computed by the REFERENCE's own
                  third_party/pcl_color_conversion/color_conversion.cpp (compiled into oracle/_ref by oracle/Makefile).
flann_golden.npz  squared-L2 exact nearest neighbours + distances computed by a real FLANN build
                  (cv2.flann_Index, linear index = exhaustive search with FLANN's own distance functors).
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


if __name__ == "__main__":
    make_lab()
    make_flann()
