#!/usr/bin/env python
"""Full-size parity / property checks of BASELINE.json configs 2-5 on one B200.

Each check prints one JSON line.  The CUDA path (through the C ABI) is compared with
the CPU oracle on a random sample of queries at the FULL tree size, plus
size-independent properties (a leaf's own vector scores ~1 against itself and ranks
first; shard merge == unsharded result).

  python tools/scale_checks.py [--only k_sweep,four_trees,million,cluster]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (checker only)
from myrio_b200 import api, synth  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def sample_parity(ctx, tree, ttc, leaf_counts, reads, k, x, n_sample, seed, leaf_id_base=0):
    """GPU top-x of all reads; oracle on a sample of them against ALL leaves."""
    q = api.MyrSeqs(ctx, reads).compute_kmer_counts(k, 0.1)
    qn = q.into_normalized_l2()
    t0 = time.perf_counter()
    res = api.TaxTreeResults.from_compute_tree(qn, ttc, api.SIM_COSINE, x)
    dt = time.perf_counter() - t0
    lro, lkeys, lvals, _ = leaf_counts.download(aux=False)
    oln = oracle.normalize(oracle.Csr(lro, lkeys, lvals))
    qro, qkeys, qvals, _ = qn.download(aux=False)
    oq = oracle.Csr(qro, qkeys, qvals)
    rng = np.random.default_rng(seed)
    idx = rng.choice(reads.n, min(n_sample, reads.n), replace=False)
    mism = 0
    for i in idx:
        kq, vq = oq.row(int(i))
        s = oracle.score_all(oln, kq, vq)
        es, ei = oracle.topx(s, x, id_base=leaf_id_base)
        mism += int((ei != res.top_ids[i]).any() or (bits(es) != bits(res.top_scores[i])).any())
    return {"queries": reads.n, "leaves": tree.n, "sampled": int(len(idx)), "topx_mismatches": mism,
            "score_topx_s": dt, "scores_per_s": reads.n * tree.n / dt, "stats": ctx.score_stats()}


def check_k_sweep(ctx, args):
    """configs[3]: k = 12/18/24/31 on the 100k-leaf tree: index size vs scoring bandwidth."""
    tree = synth.make_tree(args.leaves, 1002, root_seed=1002)
    reads = synth.make_reads(tree, args.reads, 2002)
    store = api.TaxTreeStore(ctx, tree)
    for k in (12, 18, 24, 31):
        t0 = time.perf_counter()
        store.compute_and_overwrite_kmer_counts(k, 32)
        ttc = api.TaxTreeCompute.from_store_tree(store, k, 32)
        t_build = time.perf_counter() - t0
        r = sample_parity(ctx, tree, ttc, store.kmer_store_counts_vec[0], reads, k, 100, args.sample, 7 + k)
        info = ttc.info()
        r.update({"check": "k_sweep", "k": k, "count_and_index_s": t_build, "index": info,
                  "index_bytes": info["n_postings"] * 8 + info["n_unique_total"] * (32 + 12),
                  "posting_GBps": r["stats"]["postings_touched"] * 8 / r["score_topx_s"] / 1e9})
        print(json.dumps(r), flush=True)
        del ttc


def check_four_trees(ctx, args):
    """configs[2]: four gene trees (50k leaves each), ~1 kb reads with ~5 % error, every read
    scored against all four trees."""
    genes = [("matK", 850, 1003), ("rbcL", 1400, 1004), ("ITS", 650, 1005), ("trnH-psbA", 500, 1006)]
    trees, ttcs, stores = [], [], []
    for name, n, seed in genes:
        tr = synth.make_tree(args.gene_leaves, seed, mean_len=n, sd_len=int(0.06 * n), min_len=int(0.7 * n),
                             max_len=int(1.35 * n))
        st = api.TaxTreeStore(ctx, tr, gene=name)
        ttcs.append(api.TaxTreeCompute.from_store_tree(st, 18, 32))
        trees.append(tr)
        stores.append(st)
    per = args.reads // 4
    parts = [synth.make_reads(tr, per, 2003 + i, q_shift=3) for i, tr in enumerate(trees)]
    codes = np.concatenate([p.codes for p in parts])
    qual = np.concatenate([p.qual for p in parts])
    lens = np.concatenate([p.lengths() for p in parts])
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    reads = synth.SeqSet(codes, off, qual)
    err = float(np.mean(np.power(10.0, -qual.astype(np.float64) / 10.0)))
    total_s, total_scores = 0.0, 0.0
    for g, (tr, ttc, st) in enumerate(zip(trees, ttcs, stores)):
        r = sample_parity(ctx, tr, ttc, st.kmer_store_counts_vec[0], reads, 18, 100, args.sample // 4 + 1, 11 + g)
        total_s += r["score_topx_s"]
        total_scores += reads.n * tr.n
        r.update({"check": "four_trees", "gene": genes[g][0], "mean_read_len": float(lens.mean()),
                  "mean_phred_error": err})
        print(json.dumps(r), flush=True)
    print(json.dumps({"check": "four_trees_total", "reads": reads.n, "scores": total_scores, "seconds": total_s,
                      "scores_per_s": total_scores / total_s}), flush=True)
    return reads


def check_cluster(ctx, args):
    """configs[2]/[4] clustering: mixed reads of four genes, k=6, C=4; assignments vs the oracle,
    silhouette trimming on a subsample (the oracle's O(N^2) pass bounds the size)."""
    genes = [(850, 1003), (1400, 1004), (650, 1005), (500, 1006)]
    trees = [synth.make_tree(2000, s, mean_len=n, sd_len=int(0.06 * n), min_len=int(0.7 * n), max_len=int(1.35 * n))
             for n, s in genes]
    for n_reads, sil in ((args.cluster_reads, None), (args.silhouette_reads, 1.4)):
        per = n_reads // 4
        parts = [synth.make_reads(tr, per, 3003 + i) for i, tr in enumerate(trees)]
        order = np.random.default_rng(5).permutation(per * 4)
        codes = np.concatenate([p.codes for p in parts])
        qual = np.concatenate([p.qual for p in parts])
        lens = np.concatenate([p.lengths() for p in parts])
        off = np.zeros(len(lens) + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        reads = synth.SeqSet(codes, off, qual).subset(order)
        truth = (order // per).astype(np.int64)
        m = api.MyrSeqs(ctx, reads)
        t0 = time.perf_counter()
        out = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=4,
                                                      silhouette_trimming=sil))
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
        oassign, oiters, _ = oracle.cluster(oc, 20, None, 4, silhouette=sil)
        t_cpu = time.perf_counter() - t0
        kept = out.assign >= 0
        # purity of the clustering against the gene of origin
        purity = 0.0
        for c in range(4):
            sel = out.assign == c
            if sel.any():
                purity += np.bincount(truth[sel], minlength=4).max()
        pairs = None
        if sil is not None:
            sizes = np.bincount(out.assign[out.assign >= 0], minlength=4).astype(np.float64)
            pairs = float((sizes * (sizes + sizes.max())).sum())  # upper bound of own+neighbour pairs
        print(json.dumps({"check": "cluster", "reads": reads.n, "silhouette": sil,
                          "assign_mismatches": int((out.assign != oassign).sum()), "n_iters": out.n_iters,
                          "oracle_iters": oiters, "gpu_s": t_gpu, "oracle_s": t_cpu, "oracle_threads": oracle.max_threads(),
                          "purity": purity / max(1, kept.sum()), "rejected": int((~kept).sum()),
                          "pair_similarities_upper_bound": pairs}), flush=True)


def check_million(ctx, args):
    """configs[4] on ONE GPU: 1M-leaf tree (10 GB of postings), a query sample against all leaves,
    and the two-shard merge == unsharded property."""
    t0 = time.perf_counter()
    tree = synth.make_tree(args.million, 1005, root_seed=1005)
    t_gen = time.perf_counter() - t0
    reads = synth.make_reads(tree, args.million_reads, 2005)
    t0 = time.perf_counter()
    store = api.TaxTreeStore(ctx, tree)
    store.compute_and_append_kmer_counts(18, 32)
    ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32)
    t_build = time.perf_counter() - t0
    r = sample_parity(ctx, tree, ttc, store.kmer_store_counts_vec[0], reads, 18, 100, args.sample, 99)
    r.update({"check": "million", "tree_gen_s": t_gen, "upload_count_index_s": t_build, "index": ttc.info()})
    print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="k_sweep,four_trees,cluster,million")
    ap.add_argument("--leaves", type=int, default=100_000)
    ap.add_argument("--reads", type=int, default=20_000)
    ap.add_argument("--gene-leaves", type=int, default=50_000)
    ap.add_argument("--sample", type=int, default=24)
    ap.add_argument("--cluster-reads", type=int, default=40_000)
    ap.add_argument("--silhouette-reads", type=int, default=4_000)
    ap.add_argument("--million", type=int, default=1_000_000)
    ap.add_argument("--million-reads", type=int, default=2_000)
    args = ap.parse_args()
    ctx = api.Context(0)
    only = set(args.only.split(","))
    if "k_sweep" in only:
        check_k_sweep(ctx, args)
    if "four_trees" in only:
        check_four_trees(ctx, args)
    if "cluster" in only:
        check_cluster(ctx, args)
    if "million" in only:
        check_million(ctx, args)


if __name__ == "__main__":
    main()
