"""GPU parity at the size the bench runs (pytest -m gpu): the tensor-core activation (tcgen05 GEMM + exact re-rank;
for ChiSquared the Hellinger sandwich) against the exact CUDA-core scan on REAL descriptors — >= 10 k SHOT-352 queries x
the full 1.07 M-word C3 codebook and >= 2 k CSHOT-1344 queries x 1.07 M words — rows and distance bits identical, plus
a sample against the CPU oracle.  The codebooks are the descriptors of the workloads' own training clouds (clustering
"None": codewords ARE training descriptors, implicit_shape_model.cpp:447-475); the vote tables are not needed here."""
import os
import sys
import time

import numpy as np
import pytest

from pcdb200 import synth
from pcdb200.structs import DIST_CHISQUARED, DIST_EUCLIDEAN, KNN_GEMM, KNN_SCAN, Codebook

pytestmark = pytest.mark.gpu


def _descriptor_codebook(api, name, n_words, n_test_clouds):
    wl = synth.WORKLOADS[name]
    prm = synth.workload_params(name)
    ctx = api.Context(prm)
    probe = synth.make_clouds(list(range(4)), [10_000 + i for i in range(4)], wl["P"], scale=wl["scale"])
    per_cloud = max(1.0, ctx.compute_features(*probe)[0].shape[0] / 4)
    n_clouds = int(np.ceil(1.05 * n_words / per_cloud))
    rows = []
    for s in range(0, n_clouds, 256):
        m = min(256, n_clouds - s)
        cl = synth.make_clouds([(s + i) % wl["n_classes"] for i in range(m)], [1_000_000 + s + i for i in range(m)],
                               wl["P"], scale=wl["scale"])
        rows.append(ctx.compute_features(*cl)[2])
    W = np.concatenate(rows)[:n_words]
    assert W.shape[0] >= 0.97 * n_words
    te = synth.make_clouds([i % wl["n_classes"] for i in range(n_test_clouds)],
                           [50_000_000 + i for i in range(n_test_clouds)], wl["P"], scale=wl["scale"])
    Q = ctx.compute_features(*te)[2]
    N = W.shape[0]
    cb = Codebook(W, np.arange(N + 1), np.zeros((N, 3)), np.ones(N), np.zeros(N), np.zeros(N),
                  np.tile(np.array([1, 0, 0, 0, 1, 1, 1], np.float32), (N, 1)), np.ones(N), np.zeros((N, 3)),
                  np.arange(N), np.ones(2))
    ctx.set_codebook(cb)
    return ctx, prm, cb, Q


@pytest.fixture(scope="module")
def api():
    from pcdb200 import api as _api
    return _api


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) \
        and np.array_equal(a[2], b[2])


def test_c3_headline_scale_gemm_equals_scan(api, orc):
    t0 = time.time()
    ctx, prm, cb, Q = _descriptor_codebook(api, "c3", 1_070_000, 40)
    assert cb.N >= 1_000_000 and Q.shape[0] >= 10_000 and Q.shape[1] == 352
    t1 = time.time()
    ctx.reset_stats()
    a = ctx.knn(Q, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)
    st = ctx.stats()
    b = ctx.knn(Q, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(a, b), "%d of %d rows differ" % ((a[0] != b[0]).sum(), a[0].size)
    if os.environ.get("PCDB_GEMM_PCA", "1") != "0":
        # >= 8192 queries against >= 262144 words take the PCA pre-filter (sample sweep -> pooled projected sweep -> exact
        # evaluation): its pool holds tens of rows per query where the plain sweep keeps ~1
        assert st["knn_candidates"] / Q.shape[0] > 5
    a4 = ctx.knn(Q[:4096], k=4, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)   # few queries: the plain single sweep
    b4 = ctx.knn(Q[:4096], k=4, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(a4, b4)
    a3 = ctx.knn(Q[:9000], k=3, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)   # the pre-filter with K > 1
    b3 = ctx.knn(Q[:9000], k=3, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(a3, b3)
    prm_r = prm.copy()
    prm_r.use_distance_ratio = 1                                             # ratio test: K = k + 1 inside
    ctx.set_params(prm_r)
    ar = ctx.knn(Q[:9000], k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)
    br = ctx.knn(Q[:9000], k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(ar, br)
    ctx.set_params(prm)
    # ChiSquared through the Hellinger sandwich (two tensor-core sweeps + pooled exact re-rank) vs the exact chi^2 scan
    ctx.reset_stats()
    c = ctx.knn(Q, k=1, dist_type=DIST_CHISQUARED, mode=KNN_GEMM)
    stc = ctx.stats()
    d = ctx.knn(Q, k=1, dist_type=DIST_CHISQUARED, mode=KNN_SCAN)
    assert _same(c, d), "%d of %d chi^2 rows differ" % ((c[0] != d[0]).sum(), c[0].size)
    c2 = ctx.knn(Q[:2048], k=3, dist_type=DIST_CHISQUARED, mode=KNN_GEMM)
    d2 = ctx.knn(Q[:2048], k=3, dist_type=DIST_CHISQUARED, mode=KNN_SCAN)
    assert _same(c2, d2)
    # a sample against the CPU oracle (exact search over 1.07 M words costs the host seconds per query batch)
    m = orc.Model(prm, cb)
    o = m.knn(Q[:24], k=1, dist_type=DIST_EUCLIDEAN)
    assert np.array_equal(a[0][:24], o[0]) and np.array_equal(a[1][:24].view(np.uint32), o[1].view(np.uint32))
    o = m.knn(Q[:24], k=1, dist_type=DIST_CHISQUARED)
    assert np.array_equal(c[0][:24], o[0]) and np.array_equal(c[1][:24].view(np.uint32), o[1].view(np.uint32))
    print("C3 scale: %d queries x %d words; L2 candidates/query %.2f (fallback %d); chi^2 pooled candidates/query %.1f "
          "(fallback %d); set-up %.0f s, checks %.0f s"
          % (Q.shape[0], cb.N, st["knn_candidates"] / Q.shape[0], st["knn_fallback_queries"],
             stc["knn_candidates"] / Q.shape[0], stc["knn_fallback_queries"], t1 - t0, time.time() - t1))
    ctx.close()


def test_c4_headline_scale_gemm_equals_scan(api, orc):
    ctx, prm, cb, Q = _descriptor_codebook(api, "c4", 1_070_000, 56)
    assert cb.N >= 1_000_000 and Q.shape[0] >= 8_192 and Q.shape[1] == 1344
    # >= 8192 queries: the PCA pre-filter for long rows (basis by block power iteration, streaming bound sweep over the
    # sample, pooled sweep over the 112-dimensional projections) against the exact scan
    Qa = Q[:10_240]
    ctx.reset_stats()
    aa = ctx.knn(Qa, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)
    sta = ctx.stats()
    ba = ctx.knn(Qa, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(aa, ba), "%d of %d rows differ (pre-filter)" % ((aa[0] != ba[0]).sum(), aa[0].size)
    if os.environ.get("PCDB_GEMM_PCA", "1") != "0" and os.environ.get("PCDB_GEMM_PCA_WIDE", "1") != "0":
        assert sta["knn_prefilter_dim"] > 0 and sta["knn_candidates"] / Qa.shape[0] > 5
    a2 = ctx.knn(Qa[:8_500], k=2, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)   # the pre-filter with K = 2
    b2 = ctx.knn(Qa[:8_500], k=2, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(a2, b2)
    Q = Q[:4096]                                                             # few queries: the plain streaming sweep
    a = ctx.knn(Q, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_GEMM)
    b = ctx.knn(Q, k=1, dist_type=DIST_EUCLIDEAN, mode=KNN_SCAN)
    assert _same(a, b), "%d of %d rows differ" % ((a[0] != b[0]).sum(), a[0].size)
    ctx.reset_stats()
    c = ctx.knn(Q[:2048], k=2, dist_type=DIST_CHISQUARED, mode=KNN_GEMM)
    stc = ctx.stats()
    d = ctx.knn(Q[:2048], k=2, dist_type=DIST_CHISQUARED, mode=KNN_SCAN)
    assert _same(c, d), "%d of %d chi^2 rows differ" % ((c[0] != d[0]).sum(), c[0].size)
    m = orc.Model(prm, cb)
    o = m.knn(Q[:8], k=1, dist_type=DIST_EUCLIDEAN)
    assert np.array_equal(a[0][:8], o[0]) and np.array_equal(a[1][:8].view(np.uint32), o[1].view(np.uint32))
    print("C4 scale: %d queries x %d words x 1344 (pre-filter: %d queries, %.1f pooled rows per query, %d swept again); "
          "chi^2 pooled candidates/query %.1f (fallback %d)"
          % (Q.shape[0], cb.N, Qa.shape[0], sta["knn_candidates"] / Qa.shape[0], sta["knn_prefilter_resweep_queries"],
             stc["knn_candidates"] / 2048, stc["knn_fallback_queries"]))
    ctx.close()
