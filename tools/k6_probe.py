#!/usr/bin/env python
"""Times the clustering stage (K6) on one B200: cluster() with and without silhouette
trimming, and the reads x reads tile.  Prints one JSON line per measurement."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from myrio_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100_000)
ap.add_argument("--leaves", type=int, default=20_000)
ap.add_argument("--clusters", type=int, default=4)
ap.add_argument("--tile", type=int, default=8192)
ap.add_argument("--only", default="tile,cluster,silhouette")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = api.Context(0)
tree = synth.make_tree(a.leaves, 1002, root_seed=1002)
reads = synth.make_reads(tree, a.reads, 2002)
q = api.MyrSeqs(ctx, reads)
only = a.only.split(",")
if "tile" in only:
    sub = api.MyrSeqs(ctx, reads.subset(np.arange(min(a.tile, reads.n))))
    c6 = sub.compute_kmer_counts(6, 0.2).into_normalized_l2()
    for rep in range(a.reps):
        ctx.prof_reset(); ctx.prof_enable(True)
        t0 = time.perf_counter()
        out = api.pairwise(c6, c6)
        dt = time.perf_counter() - t0
        ms, n = ctx.prof_query("k6_")
        ctx.prof_enable(False)
    print(json.dumps({"what": "pairwise tile", "rows": c6.n_rows, "cols": c6.n_rows, "nnz": c6.nnz,
                      "kernel_ms": ms, "call_s": dt, "pairs_per_s_kernel": c6.n_rows ** 2 / (ms * 1e-3)}), flush=True)
for name, sil in (("cluster", None), ("silhouette", 1.4)):
    if name not in only:
        continue
    for rep in range(a.reps):
        ctx.prof_reset(); ctx.prof_enable(True)
        t0 = time.perf_counter()
        out = api.cluster(q, api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=a.clusters,
                                                      silhouette_trimming=sil))
        dt = time.perf_counter() - t0
        res = {p: ctx.prof_query(p) for p in ("k6_pair_store", "k6_pair_argmax", "k6_pair_silhouette", "k6_silsum", "k6_slab_schedule_count", "k6_slab_schedule_fill", "k6_slab", "k6_recenter", "k1_", "")}
        ctx.prof_enable(False)
    sizes = [int(len(c)) for c in out.clusters]
    print(json.dumps({"what": name, "reads": reads.n, "clusters": sizes, "rejected": int(len(out.rejected)),
                      "iters": out.n_iters, "call_s": dt, "kernels_ms": {k: v for k, v in res.items()}}), flush=True)
