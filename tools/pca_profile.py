#!/usr/bin/env python
"""Runs the Euclidean activation of ~11.7 k real SHOT-352 queries against the 1.07 M-word C3 descriptor codebook twice
(the PCA pre-filter path) and prints the library's timing; meant to be wrapped in ncu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pcdb200 import api
from pcdb200.structs import KNN_GEMM
import test_gpu_scale as tgs
ctx, prm, cb, Q = tgs._descriptor_codebook(api, "c3", 1_070_000, int(sys.argv[1]) if len(sys.argv) > 1 else 40)
for it in range(3):
    t0 = time.perf_counter()
    ctx.reset_stats()
    r = ctx.knn(Q, k=1, mode=KNN_GEMM)
    st = ctx.stats()
    print("Q=%d e2e %.2f ms, sweeps %.2f ms, candidates/query %.1f" % (Q.shape[0], (time.perf_counter() - t0) * 1e3, st["knn_gemm_ms"], st["knn_candidates"] / Q.shape[0]))
