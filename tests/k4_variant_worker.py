"""Worker of test_gpu_parity.py::test_k4_kernel_variants: MYRIO_K4_COUNT is read once per process, so each K4
variant (count mode forced on / off) runs in its own process.  Cases chosen to hit the ODD paths of count mode:
k = 6 (reads and leaves full of repeated k-mers: odd query values and odd postings everywhere), leaves with
ambiguity codes (non-unit counts), a cluster-centroid query (its smallest value is not its common value), JT."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from myrio_b200 import api, synth  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def check(ctx, iupac, base_off, q_g, q_o, k, sim_g, sim_o, tile_leaves, x, seed=0):
    st = api.TaxTreeStore(ctx, iupac=iupac, base_off=base_off)
    ttc = api.TaxTreeCompute.from_store_tree(st, k, 32, seed=seed, tile_leaves=tile_leaves)
    oln = oracle.normalize(oracle.count_leaves(iupac, base_off, k, 32, seed=seed))
    res = api.TaxTreeResults.from_compute_tree(q_g, ttc, sim_g, x, full_scores=True)
    for i in range(q_o.n_rows):
        kq, vq = q_o.row(i)
        s = oracle.score_all(oln, kq, vq, sim=sim_o)
        assert (bits(s) == bits(res.scores[i])).all(), f"k={k} tile={tile_leaves} row {i}"
        ts, ti = oracle.topx(s, x)
        assert (ti == res.top_ids[i]).all() and (bits(ts) == bits(res.top_scores[i])).all(), f"k={k} top-x {i}"


def main():
    ctx = api.Context(0)
    rng = np.random.default_rng(7)
    tree = synth.make_tree(2600, 1501)
    reads = synth.make_reads(tree, 48, 1502)
    m = api.MyrSeqs(ctx, reads)
    acgt = synth.tree_to_iupac(tree)
    amb = acgt.copy()
    pos = rng.choice(len(amb), len(amb) // 150, replace=False)
    amb[pos] = rng.choice(np.array([3, 5, 6, 9, 10, 12, 15], np.uint8), len(pos))
    for k in (18, 12, 6):
        qg = m.compute_kmer_counts(k, 0.1).into_normalized_l2()
        qo = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, k, 0.1))
        for tl in (0, 256, 2048):
            check(ctx, acgt, tree.base_off, qg, qo, k, api.SIM_COSINE, oracle.SIM_COSINE, tl, 40)
        check(ctx, amb, tree.base_off, qg, qo, k, api.SIM_JACARD_TANIMOTO, oracle.SIM_JACARD_TANIMOTO, 0, 40, seed=5)
    # centroid query at k = 18: thousands of keys, many distinct values
    near = synth.make_reads(tree, 150, 1503, leaf_ids=np.arange(900, 960))
    counts = api.MyrSeqs(ctx, near).compute_kmer_counts(18, 0.1)
    cen = api.compute_cluster_centroid(counts, np.arange(near.n))
    ck, cv = oracle.centroid(oracle.count_reads(near.codes, near.qual, near.base_off, 18, 0.1), np.arange(near.n))
    qo = oracle.Csr(np.array([0, len(ck)], np.uint64), ck, cv)
    check(ctx, acgt, tree.base_off, cen, qo, 18, api.SIM_COSINE, oracle.SIM_COSINE, 0, 100)
    check(ctx, amb, tree.base_off, cen, qo, 18, api.SIM_COSINE, oracle.SIM_COSINE, 512, 100, seed=9)
    # more than 63 dense keys per (query, tile): the leaves themselves as queries (bit-sliced counters flush mid-walk)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, keep_leaf_vecs=True)
    res = api.TaxTreeResults.from_compute_tree(ttc.leaf_vecs, ttc, api.SIM_COSINE, 10, q_first=300, q_count=64,
                                               full_scores=True)
    oln = oracle.normalize(oracle.count_leaves(acgt, tree.base_off, 18, 32))
    for j in range(64):
        kq, vq = oln.row(300 + j)
        assert (bits(oracle.score_all(oln, kq, vq)) == bits(res.scores[j])).all(), f"self-scores {j}"
    print(f"K4_VARIANT_OK count={os.environ.get('MYRIO_K4_COUNT')}")


if __name__ == "__main__":
    main()
