#!/usr/bin/env python
"""A/B runs of the PCA pre-filter on a descriptor codebook of 1.07 M words, one context, identical results required.
  c3 (SHOT-352, d = 112): the pooled sweep with PCDB_GEMM_POOL_STASH = 3 (default: passing chunks parked in shared
     memory + early hand-over of the accumulator), 1 (parking only), 0 (neither), each also with a threshold nothing
     passes (PCDB_EXP_POOL_NOTHR=1: the kernel's floor; with STASH=4 the epilogue does not even read the accumulator:
     what the hand-over alone costs), and the plain sweep (PCDB_GEMM_PCA_SKIP=1).  (Two rejected kernel variants were measured with earlier
     versions of this script: profiles/r02_pool_variants_c3_*_rejected.json.)
  c4 (CSHOT-1344: basis by block power iteration, streaming bound sweep over the sample): the pre-filter against the
     plain streaming sweep.
Prints one JSON document.  usage: python tools/pool_variants.py [c3|c4] [test clouds]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_test = int(sys.argv[2]) if len(sys.argv) > 2 else (1024 if wl == "c3" else 512)
from pcdb200 import api
from pcdb200.structs import KNN_GEMM
import test_gpu_scale as tgs
t0 = time.time()
ctx, prm, cb, Q = tgs._descriptor_codebook(api, wl, 1_070_000, n_test)
out = {"workload": wl, "queries": int(Q.shape[0]), "words": int(cb.N), "D": int(Q.shape[1]), "setup_s": round(time.time() - t0, 1),
       "variants": []}
ref = None
if wl == "c3":
    variants = [dict(), dict(PCDB_GEMM_POOL_STASH=1), dict(PCDB_GEMM_POOL_STASH=0), dict(PCDB_EXP_POOL_NOTHR=1),
                dict(PCDB_GEMM_POOL_STASH=1, PCDB_EXP_POOL_NOTHR=1), dict(PCDB_GEMM_POOL_STASH=4, PCDB_EXP_POOL_NOTHR=1),
                dict(PCDB_GEMM_PCA_SKIP=1), dict()]
else:
    variants = [dict(), dict(PCDB_GEMM_PCA_SKIP=1), dict()]
KEYS = ("PCDB_GEMM_POOL_STASH", "PCDB_EXP_POOL_NOTHR", "PCDB_GEMM_PCA_SKIP")
DEFAULTS = {"PCDB_GEMM_POOL_STASH": 3}
for v in variants:
    for k in KEYS:
        os.environ[k] = str(v.get(k, DEFAULTS.get(k, 0)))
    rec = {"env": v, "runs": []}
    for it in range(3):
        ctx.reset_stats()
        t0 = time.perf_counter()
        r = ctx.knn(Q, k=1, mode=KNN_GEMM)
        st = ctx.stats()
        rec["runs"].append({"call_ms": round((time.perf_counter() - t0) * 1e3, 2), "activation_ms": round(st["knn_ms"], 3),
                            "gemm_ms": round(st["knn_gemm_ms"], 3), "bound_sweep_ms": round(st["knn_bound_sweep_ms"], 3),
                            "pool_sweep_ms": round(st["knn_pool_sweep_ms"], 3), "pooled_per_query": round(st["knn_candidates"] / Q.shape[0], 2),
                            "resweep_queries": int(st["knn_prefilter_resweep_queries"]), "prefilter_dim": int(st["knn_prefilter_dim"])})
    if not v.get("PCDB_EXP_POOL_NOTHR") and v.get("PCDB_GEMM_POOL_STASH", 3) != 4:
        if ref is None:
            ref = r
        rec["identical_to_first_variant"] = bool(np.array_equal(ref[0], r[0]) and np.array_equal(ref[1].view(np.uint32), r[1].view(np.uint32)) and np.array_equal(ref[2], r[2]))
    out["variants"].append(rec)
    print("variant", v, rec["runs"][-1], rec.get("identical_to_first_variant"), file=sys.stderr, flush=True)
print(json.dumps(out))
