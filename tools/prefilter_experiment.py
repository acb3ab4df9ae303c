#!/usr/bin/env python
"""Can the exact activation (2 Q N D flop) be made cheaper by a SOUND pre-filter?  Offline experiment on the C3 workload
(real SHOT-352 descriptors of the synthetic ModelNet40-shaped world, ~1.07 M codewords), run on the GPU box with torch:

  A. PCA lower bounds.  For an orthonormal basis P (d columns), |P^T q - P^T c|^2 <= |q - c|^2, so a d-dimensional GEMM
     (2 Q N d flop) can discard every codeword whose projected distance exceeds the exact nearest distance.  Reported per
     d: how many codewords per query survive (a) against the TRUE nearest distance (the best any such filter can do) and
     (b) against the distance of the projected-space argmin re-evaluated exactly (what a two-pass kernel would have).
  B. Triangle inequality on k-means cells: a cell with centre m and radius r can be skipped for q when
     |q - m| - r > (nearest distance).  Reported: the fraction of (query, cell) pairs that can be skipped.

Output: one JSON document (candidate-count percentiles and a flop model) on stdout; profiles/r02_prefilter_experiment.json
is a committed run.  usage: python tools/prefilter_experiment.py [--queries 4096] [--words 0]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "point-cloud-donkey_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--words", type=int, default=0)
    ap.add_argument("--workload", default="c3")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from pcdb200 import api
    bench.claim_stdout()
    ctx = api.Context(device=0)
    wl, prm, cb = bench.build_world(args.workload, args.words, ctx, rank_log=True)
    # queries: descriptors of unseen test clouds
    x, n, c, o, _ = bench.test_batch(wl, 32, 0, 7)
    desc = ctx.compute_features(x, n, c, o)[2]
    Q = min(args.queries, desc.shape[0])
    desc = desc[:Q]
    idx, dist, cnt = ctx.knn(desc, k=1)          # exact nearest codeword (the library's tensor-core path)
    dev = torch.device("cuda", 0)
    W = torch.from_numpy(cb.words).to(dev)
    Qd = torch.from_numpy(desc).to(dev)
    d_star = torch.from_numpy(dist[:, 0].copy()).to(dev)
    N, D = W.shape
    out = {"workload": args.workload, "N": int(N), "D": int(D), "Q": int(Q),
           "nearest_distance_percentiles": np.percentile(dist[:, 0], [1, 10, 50, 90, 99]).tolist()}
    # distance spread: how far is a typical codeword compared with the nearest one
    samp = W[torch.randint(0, N, (4096,), device=dev)]
    dd = torch.cdist(Qd[:512], samp) ** 2
    out["random_codeword_distance_percentiles"] = np.percentile(dd.cpu().numpy(), [1, 10, 50, 90, 99]).tolist()

    # ---- A. PCA lower bounds
    mu = W.mean(0, keepdim=True)
    sub = W[torch.randperm(N, device=dev)[:200000]] - mu
    cov = (sub.T @ sub).double() / sub.shape[0]
    evals, evecs = torch.linalg.eigh(cov)
    order = torch.argsort(evals, descending=True)
    evecs = evecs[:, order].float()
    out["pca_energy"] = {}
    out["pca"] = {}
    total = float(evals.sum())
    pct = [50, 90, 99, 99.9, 100]
    for d in (16, 32, 64, 128, 192):
        P = evecs[:, :d].contiguous()
        out["pca_energy"][str(d)] = float(evals[order][:d].sum()) / total
        Wp = (W - mu) @ P
        Qp = (Qd - mu) @ P
        wn = (Wp * Wp).sum(1)
        n_true, n_two = [], []
        for s in range(0, Q, 256):
            qp = Qp[s:s + 256]
            lb = (qp * qp).sum(1, keepdim=True) + wn[None, :] - 2.0 * qp @ Wp.T      # projected squared distances
            lb = lb - 1e-4                                                           # fp32 slack keeps the bound sound
            n_true.append((lb <= d_star[s:s + 256, None]).sum(1))
            arg = lb.argmin(1)
            u = ((Qd[s:s + 256] - W[arg]) ** 2).sum(1)                               # exact distance of the projected argmin
            n_two.append((lb <= u[:, None]).sum(1))
        n_true = torch.cat(n_true).float().cpu().numpy()
        n_two = torch.cat(n_two).float().cpu().numpy()
        flop_filter = 2.0 * Q * N * d
        rec = {"candidates_vs_true_nearest": {"mean": float(n_true.mean()), "percentiles": dict(zip(map(str, pct), np.percentile(n_true, pct).tolist()))},
               "candidates_two_pass": {"mean": float(n_two.mean()), "percentiles": dict(zip(map(str, pct), np.percentile(n_two, pct).tolist()))},
               "flop_model": {"exact_2QND": 2.0 * Q * N * D, "filter_2QNd": flop_filter,
                              "rerank_two_pass": float(n_two.sum()) * 2.0 * D,
                              "total_over_exact": (flop_filter + float(n_two.sum()) * 2.0 * D) / (2.0 * Q * N * D)}}
        out["pca"][str(d)] = rec
    # ---- A'. single-sweep variant: the bound U comes from the exact nearest distance inside a random 1/f SAMPLE of the
    # codebook (a full-D sweep over N/f rows), then ONE projected sweep pools every codeword with LB <= U
    out["sample_bound"] = {}
    for f in (8, 16, 32):
        perm = torch.randperm(N, device=dev)[: N // f]
        Ws = W[perm]
        wsn = (Ws * Ws).sum(1)
        U = torch.empty(Q, device=dev)
        for s in range(0, Q, 256):
            qq = Qd[s:s + 256]
            dd2 = (qq * qq).sum(1, keepdim=True) + wsn[None, :] - 2.0 * qq @ Ws.T
            arg = dd2.argmin(1)
            U[s:s + 256] = ((qq - Ws[arg]) ** 2).sum(1)
        rec = {"U_over_nearest_percentiles": dict(zip(map(str, pct), np.percentile((U / d_star.clamp(min=1e-12)).cpu().numpy(), pct).tolist()))}
        for d in (96, 128, 160):
            P = evecs[:, :d].contiguous()
            Wp = (W - mu) @ P
            Qp = (Qd - mu) @ P
            wn = (Wp * Wp).sum(1)
            cnts = []
            for s in range(0, Q, 256):
                qp = Qp[s:s + 256]
                lb = (qp * qp).sum(1, keepdim=True) + wn[None, :] - 2.0 * qp @ Wp.T
                cnts.append((lb <= (U[s:s + 256, None] * 1.002 + 2e-3)).sum(1))     # with a realistic fp16 margin
            cn = torch.cat(cnts).float().cpu().numpy()
            rec["d%d" % d] = {"mean": float(cn.mean()), "mean_capped_2048": float(np.minimum(cn, 2048).mean()),
                              "frac_over_1024": float((cn > 1024).mean()), "frac_over_2048": float((cn > 2048).mean()),
                              "frac_over_4096": float((cn > 4096).mean()),
                              "percentiles": dict(zip(map(str, pct), np.percentile(cn, pct).tolist()))}
        out["sample_bound"]["1/%d" % f] = rec
    # ---- B. triangle inequality on k-means cells
    K = 2048
    cen = W[torch.randperm(N, device=dev)[:K]].clone()
    for _ in range(8):
        assign = torch.empty(N, dtype=torch.long, device=dev)
        for s in range(0, N, 65536):
            assign[s:s + 65536] = torch.cdist(W[s:s + 65536], cen).argmin(1)
        sums = torch.zeros_like(cen).index_add_(0, assign, W)
        cnts = torch.bincount(assign, minlength=K).clamp(min=1).float()[:, None]
        cen = sums / cnts
    rad = torch.zeros(K, device=dev)
    dist_to_cen = torch.empty(N, device=dev)
    for s in range(0, N, 65536):
        dist_to_cen[s:s + 65536] = (W[s:s + 65536] - cen[assign[s:s + 65536]]).norm(dim=1)
    rad = torch.zeros(K, device=dev).scatter_reduce_(0, assign, dist_to_cen, "amax")
    qc = torch.cdist(Qd, cen)                                  # |q - m|
    skip = (qc - rad[None, :]) > d_star.sqrt()[:, None]        # Euclidean (not squared) triangle inequality
    sizes = torch.bincount(assign, minlength=K).float()
    out["triangle_kmeans"] = {"cells": K, "pairs_skippable": float(skip.float().mean()),
                              "codewords_skippable": float((skip.float() * sizes[None, :]).sum() / (Q * N)),
                              "median_cell_radius": float(rad.median()), "median_nearest": float(d_star.sqrt().median()),
                              "median_query_to_cell": float(qc.median())}
    bench.emit(out)


if __name__ == "__main__":
    main()
