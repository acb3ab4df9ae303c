#!/usr/bin/env python
"""Would the PCA pre-filter pay for ChiSquared?  CPU experiment (numpy + the oracle's SHOT-352 descriptors, no GPU).

For non-negative rows  H^2 <= chi^2 <= 2 H^2  with H^2 = |sqrt(q) - sqrt(c)|^2, and for an orthonormal basis P the
projected distance |P^T sqrt(q) - P^T sqrt(c)|^2 <= H^2.  A projected sweep therefore has to pool every codeword whose
projected H^2 is <= U, U an upper bound of the K-th smallest chi^2.  This script counts those codewords per query for
  U = the exact nearest chi^2 (the best any bound can do) and U = the nearest chi^2 inside a 1/16 sample (what one sweep
  over a sample gives), d = 112 principal axes of the square-root rows,
next to the Euclidean analogue on the same descriptors (projected squared L2 <= nearest squared L2 / nearest inside the
sample), which is what the shipped pre-filter pools.  Pool sizes grow about linearly with the codebook size: the ratio
between the two families at equal N is the figure of interest.  Output: one JSON document (profiles/r02_chi2_prefilter_experiment.json
is a committed run).  usage: python tools/chi2_prefilter_experiment.py [train clouds per class = 6] [test clouds = 6]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))
import numpy as np
from oracle import oracle_py as orc
from pcdb200 import synth


def main():
    per_class = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_test = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    wl = synth.WORKLOADS["c3"]
    prm = synth.workload_params("c3")
    t0 = time.time()
    tr_cls = [c for c in range(wl["n_classes"]) for _ in range(per_class)]
    rows = []
    for s in range(0, len(tr_cls), 40):
        x, n, c, o = synth.make_clouds(tr_cls[s:s + 40], [1_000_000 + s + i for i in range(len(tr_cls[s:s + 40]))], wl["P"],
                                       scale=wl["scale"], jitter=0.002)
        rows.append(orc.compute_features(prm, x, n, c, o)[2])
    W = np.concatenate(rows).astype(np.float32)
    x, n, c, o = synth.make_clouds([i % wl["n_classes"] for i in range(n_test)], [50_000_000 + i for i in range(n_test)],
                                   wl["P"], scale=wl["scale"])
    Q = orc.compute_features(prm, x, n, c, o)[2].astype(np.float32)
    N, D, M = W.shape[0], W.shape[1], Q.shape[0]
    out = {"N": int(N), "D": int(D), "Q": int(M), "descriptors_s": round(time.time() - t0, 1), "d": 112, "sample": "1/16"}
    assert (W >= 0).all() and (Q >= 0).all()

    def pca(rows_, d):
        mu = rows_.mean(0, keepdims=True)
        sub = (rows_ - mu).astype(np.float64)
        ev, vec = np.linalg.eigh(sub.T @ sub / len(sub))
        return mu, vec[:, np.argsort(ev)[::-1][:d]].astype(np.float32)

    rng = np.random.default_rng(1)
    samp = np.sort(rng.choice(N, N // 16, replace=False))
    pct = [50, 90, 99, 100]

    def summarise(cnt):
        return {"mean": float(cnt.mean()), "percentiles": dict(zip(map(str, pct), np.percentile(cnt, pct).tolist())),
                "mean_fraction_of_codebook": float(cnt.mean() / N)}

    # ---- Euclidean
    wn = (W * W).sum(1)
    d2 = (Q * Q).sum(1, keepdims=True) + wn[None, :] - 2.0 * Q @ W.T
    nn = np.maximum(d2.min(1), 0)
    nn_s = np.maximum(d2[:, samp].min(1), 0)
    mu, P = pca(W, 112)
    Wp, Qp = (W - mu) @ P, (Q - mu) @ P
    lb = (Qp * Qp).sum(1, keepdims=True) + (Wp * Wp).sum(1)[None, :] - 2.0 * Qp @ Wp.T - 1e-5
    out["euclidean"] = {"bound_over_nearest_median": float(np.median(nn_s / np.maximum(nn, 1e-12))),
                        "pool_vs_exact_nearest": summarise((lb <= nn[:, None]).sum(1)),
                        "pool_vs_sample_bound": summarise((lb <= nn_s[:, None]).sum(1))}
    # ---- ChiSquared
    chi_nn = np.full(M, np.inf, np.float32)
    chi_nn_s = np.full(M, np.inf, np.float32)
    in_s = np.zeros(N, bool)
    in_s[samp] = True
    for s in range(0, N, 2048):
        w = W[s:s + 2048]
        num = (Q[:, None, :] - w[None, :, :]) ** 2
        den = Q[:, None, :] + w[None, :, :]
        chi = np.where(den > 0, num / np.where(den > 0, den, 1), 0).sum(2)
        chi_nn = np.minimum(chi_nn, chi.min(1))
        if in_s[s:s + 2048].any():
            chi_nn_s = np.minimum(chi_nn_s, chi[:, in_s[s:s + 2048]].min(1))
    SW, SQ = np.sqrt(W), np.sqrt(Q)
    h2 = (SQ * SQ).sum(1, keepdims=True) + (SW * SW).sum(1)[None, :] - 2.0 * SQ @ SW.T
    mu, P = pca(SW, 112)
    Wp, Qp = (SW - mu) @ P, (SQ - mu) @ P
    lbh = (Qp * Qp).sum(1, keepdims=True) + (Wp * Wp).sum(1)[None, :] - 2.0 * Qp @ Wp.T - 1e-5
    out["chisquared"] = {"bound_over_nearest_median": float(np.median(chi_nn_s / np.maximum(chi_nn, 1e-12))),
                         "nearest_chi2_over_its_H2_median": float(np.median(chi_nn / np.maximum(h2.min(1), 1e-12))),
                         "full_length_H2_pool_vs_exact_nearest": summarise((h2 - 1e-5 <= chi_nn[:, None]).sum(1)),
                         "pool_vs_exact_nearest": summarise((lbh <= chi_nn[:, None]).sum(1)),
                         "pool_vs_sample_bound": summarise((lbh <= chi_nn_s[:, None]).sum(1))}
    e, c2 = out["euclidean"], out["chisquared"]
    out["chi2_pool_over_euclidean_pool"] = {
        "vs_exact_nearest": c2["pool_vs_exact_nearest"]["mean"] / max(e["pool_vs_exact_nearest"]["mean"], 1e-9),
        "vs_sample_bound": c2["pool_vs_sample_bound"]["mean"] / max(e["pool_vs_sample_bound"]["mean"], 1e-9)}
    out["seconds"] = round(time.time() - t0, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
