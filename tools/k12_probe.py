#!/usr/bin/env python
"""Times K1 (read k-mer counting) and K2 (leaf k-mer counting) on one B200; one JSON line."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from myrio_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100_000)
ap.add_argument("--leaves", type=int, default=100_000)
ap.add_argument("--k", type=int, default=18)
a = ap.parse_args()
ctx = api.Context(0)
tree = synth.make_tree(a.leaves, 1002, root_seed=1002)
reads = synth.make_reads(tree, a.reads, 2002)
q = api.MyrSeqs(ctx, reads)
store = api.TaxTreeStore(ctx, tree)
res = {"env": os.environ.get("MYRIO_KC_THREADS", ""), "k": a.k}
for rep in range(3):
    ctx.prof_reset(); ctx.prof_enable(True)
    t0 = time.perf_counter(); c = q.compute_kmer_counts(a.k, 0.1); ctx.sync(); t_r = time.perf_counter() - t0
    k1 = ctx.prof_query("k1_")[0]
    ctx.prof_reset()
    t0 = time.perf_counter(); lc = store._count(a.k, 32, 0); ctx.sync(); t_l = time.perf_counter() - t0
    k2 = ctx.prof_query("k2_")[0]
    ctx.prof_enable(False)
_, _, _, hck = c.download()
leaf_kmers = float(2 * (tree.lengths() - a.k + 1).clip(0).sum())
res.update({"k1_ms": k1, "k1_call_ms": t_r * 1e3, "read_kmers": float(hck.sum()), "read_kmers_per_s_kernel": float(hck.sum()) / (k1 * 1e-3),
            "k2_ms": k2, "k2_call_ms": t_l * 1e3, "leaf_kmers": leaf_kmers, "leaf_kmers_per_s_kernel": leaf_kmers / (k2 * 1e-3),
            "leaf_kmers_per_s_call": leaf_kmers / t_l})
print(json.dumps(res), flush=True)
