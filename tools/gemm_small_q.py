#!/usr/bin/env python
"""Activation latency for few queries (one cloud) against a large codebook: pcdb_knn in GEMM mode, timed end to end
(host buffers) and by the library's own CUDA events, for several slice counts (PCDB_GEMM_S, read at first use => one
process per setting).  usage: python tools/gemm_small_q.py [N] [D]"""
import os, subprocess, sys, json

def child(N, D, Qs):
    import time
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-donkey_b200"))
    from pcdb200 import api, synth
    from pcdb200.structs import KNN_GEMM
    rng = np.random.default_rng(1)
    words = np.abs(rng.standard_normal((N, D), dtype=np.float32))
    words /= np.linalg.norm(words, axis=1, keepdims=True)
    prm = synth.workload_params("c3")
    ctx = api.Context(prm)
    from pcdb200.structs import Codebook
    z = np.zeros
    bbox = z((N, 7), np.float32)
    bbox[:, 0] = 1
    cb = Codebook(words, np.arange(N + 1, dtype=np.int64), z((N, 3), np.float32), np.ones(N, np.float32),
                  z(N, np.uint32), z(N, np.uint32), bbox, np.ones(N, np.float32), z((N, 3), np.float32),
                  np.arange(N, dtype=np.int32), np.ones(1, np.float32))
    ctx.set_codebook(cb)
    out = {}
    for Q in Qs:
        q = words[rng.integers(0, N, Q)] + 0.05 * rng.standard_normal((Q, D), dtype=np.float32)
        for _ in range(3):
            ctx.knn(q, k=1, mode=KNN_GEMM)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            ctx.knn(q, k=1, mode=KNN_GEMM)
            ts.append((time.perf_counter() - t0) * 1e3)
        out[Q] = {"e2e_ms_median": float(np.median(ts)), "gemm_ms": ctx.stats().get("knn_gemm_ms")}
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), [int(x) for x in sys.argv[4].split(",")])
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 1070260
        D = int(sys.argv[2]) if len(sys.argv) > 2 else 352
        for S in ["0", "12", "24", "37", "49", "74", "148", "296"]:
            env = dict(os.environ, PCDB_GEMM_S=S)
            r = subprocess.run([sys.executable, __file__, "--child", str(N), str(D), "200,300,672,2048"], env=env,
                               capture_output=True, text=True)
            print("S=%s" % S, r.stdout.strip() or r.stderr[-400:])
