"""BASELINE.json configs[0] end to end on the GPU against the oracle, bit for bit: `myrio tree new -k 18`
followed by `myrio run` on 10 000 mixed reads (myrio-cli/src/app/process/tree.rs:116-151, run.rs:104-469),
composed from the C-ABI calls the way INTEGRATION.md wires them into the reference:

  tree new : leaf k-mer counts at k=18 for every tree                     (tax/store.rs:265-285)
  run      : fingerprints at cluster_k=6 (explicit leaf samples)           (tax/compute.rs:51-77)
             pass 1  cluster(reads, init = global fingerprints)            (run.rs:216-231)
             naive centroids of the pass-1 clusters at k=6                 (run.rs:231-241)
             local fingerprints re-weighted by exp(1 + alpha * sim)        (run.rs:243-262)
             pass 2  cluster(reads, init = re-weighted fingerprints, silhouette) (run.rs:264-276)
             per cluster: read counts at search_k=18, centroid = the query (run.rs:301-309)
             per (tree, query): all leaf scores + cut(100)                 (tax/results.rs:82-93, 310-329)

Everything the device computes is compared with the oracle: counts, fingerprints (local, global and
re-weighted), both cluster assignments, silhouette scores, queries, every score row and the top-100."""
import numpy as np
import pytest

import oracle
from myrio_b200 import synth

pytestmark = pytest.mark.gpu

N_READS = (7000, 3000)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def csr_from_rows(rows):
    ro = np.zeros(len(rows) + 1, np.uint64)
    ro[1:] = np.cumsum([len(k) for k, _ in rows])
    return ro, np.concatenate([k for k, _ in rows]), np.concatenate([v for _, v in rows]).astype(np.float32)


def test_config1_tree_new_and_run_match_oracle():
    from myrio_b200 import api
    ctx = api.Context(0)
    rng = np.random.default_rng(1001)
    trees = [synth.make_tree(2000, 1001), synth.make_tree(500, 1011, mean_len=850, sd_len=50, min_len=600, max_len=1100)]
    parts = [synth.make_reads(t, n, 2001 + i) for i, (t, n) in enumerate(zip(trees, N_READS))]
    codes = np.concatenate([p.codes for p in parts])
    qual = np.concatenate([p.qual for p in parts])
    lens = np.concatenate([p.lengths() for p in parts])
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    order = rng.permutation(len(lens))
    reads = synth.SeqSet(codes, off, qual).subset(order)
    gene_of_read = (order >= N_READS[0]).astype(np.int64)
    samples = [[rng.choice(t.n, 32, replace=False) for _ in range(4)] for t in trees]

    # ---- tree new -k 18 (+ the normalised leaf vectors / index of `run`)
    stores, ttcs, o_leaves = [], [], []
    for t in trees:
        st = api.TaxTreeStore(ctx, t)
        st.compute_and_append_kmer_counts(18, 32)
        lro, lk, lv, resc = st.kmer_store_counts_vec[0].download()
        ol = oracle.count_leaves(synth.tree_to_iupac(t), t.base_off, 18, 32)
        assert (lro == ol.row_off).all() and (lk == ol.keys).all() and (lv == ol.vals).all() and (resc == ol.aux).all()
        stores.append(st)
        ttcs.append(api.TaxTreeCompute.from_store_tree(st, 18, 32))
        o_leaves.append(oracle.normalize(ol))

    # ---- fingerprints at cluster_k = 6
    g_init_rows, o_init_rows, g_locals, o_locals = [], [], [], []
    for t, st, smp in zip(trees, stores, samples):
        loc, glob = api.fingerprints(st, 6, smp, 32)
        _, gk, gv, _ = glob.download(aux=False)
        oc6 = oracle.count_leaves(synth.tree_to_iupac(t), t.base_off, 6, 32)
        scaled = oracle.Csr(oc6.row_off, oc6.keys, (oc6.vals.astype(np.float32) * np.repeat(
            oc6.aux, np.diff(oc6.row_off).astype(np.int64))).astype(np.float32))
        local = [oracle.centroid(scaled, ids) for ids in smp]
        lro, lkk, lvv = csr_from_rows(local)
        ek, ev = oracle.centroid(oracle.Csr(lro, lkk, lvv), np.arange(len(local)))
        assert (gk == ek).all() and (bits(gv) == bits(ev)).all()
        g_init_rows.append((gk, gv))
        o_init_rows.append((ek, ev))
        g_locals.append(api.stack_rows(ctx, loc))
        o_locals.append(oracle.Csr(lro, lkk, lvv))
    iro, ik, iv = csr_from_rows(o_init_rows)
    o_init = oracle.Csr(iro, ik, iv)
    g_init = api.SFVecs.upload(ctx, iro, ik, iv)

    # ---- pass 1
    m = api.MyrSeqs(ctx, reads)
    oc6 = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    o1, o1_iters, _ = oracle.cluster(oc6, 20, o_init, 2)
    g1 = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, intial_centroids=g_init,
                                                 expected_nb_of_clusters=2))
    assert (g1.assign == o1).all() and g1.n_iters == o1_iters

    # ---- naive centroids of the pass-1 clusters, then pass 2 with silhouette trimming
    gc6 = m.compute_kmer_counts(6, 0.2)
    g_rows, o_rows = [], []
    for c in range(2):
        members = np.nonzero(o1 == c)[0]
        assert len(members) > 0
        _, gk, gv, _ = api.compute_cluster_centroid(gc6, members).download(aux=False)
        ek, ev = oracle.centroid(oc6, members)
        assert (gk == ek).all() and (bits(gv) == bits(ev)).all()
        # run.rs:243-262: cluster c is paired with tree c; its local fingerprints are re-weighted by
        # exp(1 + alpha * sim(local, naive query)), summed and normalised -> initial centroid of pass 2
        rk, rv = oracle.reweight_fingerprints(o_locals[c], ek, ev, 2.5)
        naive = api.compute_cluster_centroid(gc6, members)
        _, wk, wv, _ = api.reweight_fingerprints(g_locals[c], naive, api.SIM_COSINE, 2.5).download(aux=False)
        assert (wk == rk).all() and (bits(wv) == bits(rv)).all()
        g_rows.append((wk, wv))
        o_rows.append((rk, rv))
    cro, ck, cv = csr_from_rows(o_rows)
    o2, o2_iters, o2_sil = oracle.cluster(oc6, 20, oracle.Csr(cro, ck, cv), 2, silhouette=1.4)
    g2 = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20,
                                                 intial_centroids=api.SFVecs.upload(ctx, cro, ck, cv),
                                                 expected_nb_of_clusters=2, silhouette_trimming=1.4))
    assert (g2.assign == o2).all() and g2.n_iters == o2_iters
    both = ~np.isnan(o2_sil)
    assert (np.isnan(g2.silhouette) == np.isnan(o2_sil)).all()
    assert (bits(g2.silhouette[both]) == bits(o2_sil[both])).all()
    assert (o2 == -2).sum() > 0  # silhouette trimming rejected something

    # ---- one query per cluster at search_k = 18, scored against every tree
    gc18 = m.compute_kmer_counts(18, 0.1)
    oc18 = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1)
    for c in range(2):
        members = np.nonzero(o2 == c)[0]
        q = api.compute_cluster_centroid(gc18, members)
        _, qk, qv, _ = q.download(aux=False)
        ek, ev = oracle.centroid(oc18, members)
        assert (qk == ek).all() and (bits(qv) == bits(ev)).all()
        gene = int(np.bincount(gene_of_read[members], minlength=2).argmax())
        best_per_tree = []
        for t, ttc, ol in zip(trees, ttcs, o_leaves):
            res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 100, full_scores=True)
            s = oracle.score_all(ol, ek, ev)
            assert (bits(s) == bits(res.scores[0])).all()
            ts, ti = oracle.topx(s, 100)
            assert (ti == res.top_ids[0]).all() and (bits(ts) == bits(res.top_scores[0])).all()
            best_per_tree.append(float(res.top_scores[0][0]))
        assert int(np.argmax(best_per_tree)) == gene  # the cluster is identified with its own gene's tree
