"""Full-size parity of BASELINE.json configs[1]-[4] on one B200 (driver-run: `pytest -m gpu`).

configs[0] runs end to end in test_gpu_config1_pipeline.py.  Here the CUDA path (through the C
ABI) processes the WHOLE workload of each configuration; the CPU oracle -- which would need hours
for 1e10..1e12 merge-joins -- checks a seeded random SAMPLE of it bit for bit (k-mer counts of
sampled leaves and reads, the complete top-100 of sampled queries against ALL leaves, cluster
assignments of all reads, silhouette scores of sampled reads), and size-independent properties
cover the rest (shard merge == unsharded result, the trimming rule recomputed from the scores).

  configs[1]/[3]  100 k-leaf tree x 100 k reads, k = 12 / 18 / 24 / 31        test_k_sweep_100k_leaves
  configs[2]      four gene trees x 50 k leaves, 200 k ~1 kb reads, 5 % error test_four_trees_200k_reads
  configs[4]      1 M-leaf tree as 8 shards: seed -> max -> sweep -> merge     test_million_leaves_eight_shards
                  read-read clustering of >= 100 k reads, silhouette on        test_cluster_100k_reads
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from myrio_b200 import synth

pytestmark = pytest.mark.gpu

# MYRIO_TEST_SCALE < 1 shrinks every size (development runs); the driver runs the full sizes
SCALE = float(os.environ.get("MYRIO_TEST_SCALE", "1.0"))


def sz(n, lo=64):
    return max(lo, int(n * SCALE))


@pytest.fixture(scope="module")
def api():
    from myrio_b200 import api as _api
    return _api


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rows_of(csr, idx):
    """Sub-batch of a Csr (rows idx, in that order)."""
    idx = np.asarray(idx, np.int64)
    lens = (csr.row_off[idx + 1] - csr.row_off[idx]).astype(np.int64)
    off = np.zeros(len(idx) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    pos = synth._ragged_positions(csr.row_off[idx].astype(np.int64), lens)
    return oracle.Csr(off, csr.keys[pos], csr.vals[pos], None if csr.aux is None else csr.aux[idx])


def check_leaf_counts_sample(tree, gpu_counts, k, sample, seed, leaf_id_base=0):
    """K2 at full size: the oracle recounts a sample of the leaves (32 resamples each)."""
    lro, lkeys, lvals, rescale = gpu_counts
    idx = np.sort(np.random.default_rng(seed).choice(tree.n, min(sample, tree.n), replace=False))
    sub = tree.subset(idx)
    exp = oracle.count_leaves(synth.tree_to_iupac(sub), sub.base_off, k, 32)
    got = rows_of(oracle.Csr(lro, lkeys, lvals, rescale), idx)
    assert (got.row_off == exp.row_off).all()
    assert (got.keys == exp.keys).all() and (got.vals == exp.vals).all()
    assert (bits(got.aux) == bits(exp.aux)).all()


def check_read_counts_sample(reads, gpu_counts, k, cutoff, sample, seed):
    """K1 at full size: the oracle recounts a sample of the reads."""
    ro, keys, vals, hck = gpu_counts
    idx = np.sort(np.random.default_rng(seed).choice(reads.n, min(sample, reads.n), replace=False))
    sub = reads.subset(idx)
    exp = oracle.count_reads(sub.codes, sub.qual, sub.base_off, k, cutoff)
    got = rows_of(oracle.Csr(ro, keys, vals, hck), idx)
    assert (got.row_off == exp.row_off).all() and (got.keys == exp.keys).all()
    assert (bits(got.vals) == bits(exp.vals)).all() and (got.aux == exp.aux).all()


def check_topx_sample(oln, oq, top_scores, top_ids, x, sample, seed, leaf_id_base=0):
    """K4/K5 at full size: sampled queries against ALL leaves with the reference's merge-join."""
    idx = np.random.default_rng(seed).choice(oq.n_rows, min(sample, oq.n_rows), replace=False)
    for i in idx:
        kq, vq = oq.row(int(i))
        s = oracle.score_all(oln, kq, vq)
        es, ei = oracle.topx(s, x, id_base=leaf_id_base)
        assert (ei == top_ids[i]).all(), f"query {i}: top-{x} ids differ"
        assert (bits(es) == bits(top_scores[i])).all(), f"query {i}: top-{x} score bits differ"
    return len(idx)


# ------------------------------------------------------------------ configs[1] and configs[3]

@pytest.fixture(scope="module")
def tree_100k():
    tree = synth.make_tree(sz(100_000, 3000), 1002, root_seed=1002)
    reads = synth.make_reads(tree, sz(100_000, 2000), 2002)
    return tree, reads


@pytest.mark.parametrize("k", [18, 12, 24, 31])
def test_k_sweep_100k_leaves(api, ctx, tree_100k, k):
    """configs[1] (k = 18) and configs[3] (k = 12 / 24 / 31): every read of the 100 k-read batch against the
    100 k-leaf tree; >= 256 sampled queries per k checked against the oracle on all leaves."""
    tree, reads = tree_100k
    store = api.TaxTreeStore(ctx, tree)
    store.compute_and_overwrite_kmer_counts(k, 32)
    counts = store.kmer_store_counts_vec[0].download()
    check_leaf_counts_sample(tree, counts, k, 512, 100 + k)
    ttc = api.TaxTreeCompute.from_store_tree(store, k, 32)
    info = ttc.info()
    assert info["n_leaves"] == tree.n and info["n_postings"] == len(counts[1])

    q = api.MyrSeqs(ctx, reads).compute_kmer_counts(k, 0.1)
    qc = q.download()
    check_read_counts_sample(reads, qc, k, 0.1, 2048, 200 + k)
    qn = q.into_normalized_l2()
    res = api.TaxTreeResults.from_compute_tree(qn, ttc, api.SIM_COSINE, 100)

    oln = oracle.normalize(oracle.Csr(counts[0], counts[1], counts[2]))
    oq = oracle.normalize(oracle.Csr(qc[0], qc[1], qc[2]))
    gro, gkeys, gvals, _ = qn.download(aux=False)
    assert (gkeys == oq.keys).all() and (bits(gvals) == bits(oq.vals)).all()  # into_normalized_l2 of every read
    n = check_topx_sample(oln, oq, res.top_scores, res.top_ids, 100, 256, 300 + k)
    assert n >= min(256, reads.n)
    # size-independent property: the list of every query is sorted (score desc, payload id asc) and in range
    ts, ti = res.top_scores, res.top_ids.astype(np.int64)
    assert (ts[:, :-1] >= ts[:, 1:]).all()
    tie = ts[:, :-1] == ts[:, 1:]
    assert (ti[:, :-1][tie] < ti[:, 1:][tie]).all()
    assert (ti < tree.n).all()


# ------------------------------------------------------------------ configs[2]

def test_four_trees_200k_reads(api, ctx):
    """configs[2]: matK / rbcL / ITS / trnH-psbA trees of 50 k leaves each, 200 k mixed ~1 kb reads at ~5 %
    error: clustering into 4 clusters at k = 6 (all assignments vs the oracle), then every read scored
    against all four trees at k = 18 (sampled queries per tree vs the oracle on all leaves)."""
    trees = [synth.make_gene_tree(sz(50_000, 1500), n, s) for _, n, s in synth.FOUR_GENES]
    reads = synth.make_mixed_reads(trees, sz(200_000, 2000), 2003)
    assert 0.035 < reads.meta["mean_phred_error"] < 0.065
    m = api.MyrSeqs(ctx, reads)

    # stage 3 first, as `myrio run` does: cluster() k=6, t1=0.2, t2=20, C=4 (clustering.rs:55-233)
    out = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=4))
    oc6 = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    oassign, oiters, _ = oracle.cluster(oc6, 20, None, 4)
    assert (out.assign == oassign).all() and out.n_iters == oiters

    q = m.compute_kmer_counts(18, 0.1)
    qc = q.download()
    check_read_counts_sample(reads, qc, 18, 0.1, 2048, 31)
    qn = q.into_normalized_l2()
    oq = oracle.normalize(oracle.Csr(qc[0], qc[1], qc[2]))
    for g, tree in enumerate(trees):
        store = api.TaxTreeStore(ctx, tree, gene=synth.FOUR_GENES[g][0])
        ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32)
        counts = store.kmer_store_counts_vec[0].download()
        check_leaf_counts_sample(tree, counts, 18, 256, 40 + g)
        res = api.TaxTreeResults.from_compute_tree(qn, ttc, api.SIM_COSINE, 100)
        oln = oracle.normalize(oracle.Csr(counts[0], counts[1], counts[2]))
        # half of the sample from the tree's own reads (long hit lists), half from the other genes (few hits)
        own = np.nonzero(reads.meta["gene"] == g)[0]
        rng = np.random.default_rng(50 + g)
        idx = np.concatenate([rng.choice(own, min(40, len(own)), replace=False),
                              rng.choice(reads.n, min(40, reads.n), replace=False)])
        for i in idx:
            kq, vq = oq.row(int(i))
            s = oracle.score_all(oln, kq, vq)
            es, ei = oracle.topx(s, 100)
            assert (ei == res.top_ids[i]).all() and (bits(es) == bits(res.top_scores[i])).all()
        # reads of this gene find this gene: best score of an own read beats a foreign read's on average
        best = res.top_scores[:, 0]
        assert best[own].mean() > 4 * best[reads.meta["gene"] != g].mean()
        del ttc, store


# ------------------------------------------------------------------ configs[4]

def test_million_leaves_eight_shards(api):
    """configs[4] scoring on ONE GPU: the 1 M-leaf tree cut into 8 payload-id shards, one context per shard
    (as 8 ranks hold them).  Seed pass on every shard, MAX of the bounds (the all-reduce), sweep with the
    exchanged bound, merge of the 8 composite lists (K7) -- compared with the oracle's top-100 over all
    1 M leaves on sampled queries, and with the plain exact-list exchange on all queries."""
    import torch
    from myrio_b200 import _lib
    n_shards, x = 8, 100
    L_total = sz(1_000_000, 8000)
    tree = synth.make_tree(L_total, 1005, root_seed=1005)
    reads = synth.make_reads(tree, sz(20_000, 500), 2005)
    Q = reads.n
    per = (L_total + n_shards - 1) // n_shards
    base = api.Context(0)
    qn = api.MyrSeqs(base, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    qro, qk, qv, _ = qn.download(aux=False)
    oq = oracle.Csr(qro, qk, qv)
    sample = np.random.default_rng(77).choice(Q, min(48, Q), replace=False)
    oracle_rows = np.zeros((len(sample), L_total), np.float32)

    shards = []
    for r in range(n_shards):
        lo, hi = r * per, min(L_total, (r + 1) * per)
        c = api.Context(0)
        sub = tree.subset(np.arange(lo, hi))
        st = api.TaxTreeStore(c, sub, leaf_id_base=lo)
        ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32)
        counts = st.kmer_store_counts_vec[0].download()
        if r in (0, n_shards - 1):
            check_leaf_counts_sample(sub, counts, 18, 128, 60 + r)
        oln = oracle.normalize(oracle.Csr(counts[0], counts[1], counts[2]))
        for j, i in enumerate(sample):
            kq, vq = oq.row(int(i))
            oracle_rows[j, lo:hi] = oracle.score_all(oln, kq, vq)
        del oln, counts
        st.kmer_store_counts_vec = []  # the index holds what scoring needs; free the 1.6 GB of counts per shard
        shards.append((c, st, ttc, api.SFVecs.upload(c, qro, qk, qv), lo))
    assert sum(s[2].info()["n_leaves"] for s in shards) == L_total

    dev = torch.device("cuda", 0)
    bounds = torch.zeros((n_shards, Q), dtype=torch.int32, device=dev)
    for r, (c, st, sh, qq, lo) in enumerate(shards):
        _lib.check(c._L.myrio_score_topx_seed_async(c.h, sh.h, qq.h, 0, Q, x, 0, C.c_void_p(bounds[r].data_ptr())))
        c.sync()
    gbound = bounds.max(dim=0).values.contiguous()  # == all-reduce(MAX) over the ranks
    tops = torch.zeros((n_shards, Q, x), dtype=torch.int64, device=dev)
    for r, (c, st, sh, qq, lo) in enumerate(shards):
        _lib.check(c._L.myrio_score_topx_sweep_async(c.h, sh.h, Q, x, 0, C.c_void_p(gbound.data_ptr()),
                                                     C.c_void_p(tops[r].data_ptr())))
        c.sync()
    out_s = torch.zeros((Q, x), dtype=torch.float32, device=dev)
    out_i = torch.zeros((Q, x), dtype=torch.int32, device=dev)
    _lib.check(base._L.myrio_topx_merge_composites_async(base.h, C.c_void_p(tops.data_ptr()), n_shards, Q, x,
                                                         C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr())))
    base.sync()
    two_s, two_i = out_s.cpu().numpy(), out_i.cpu().numpy().view(np.uint32)

    # (1) sampled queries: the oracle's cut(100) over all 1 M leaves
    for j, i in enumerate(sample):
        es, ei = oracle.topx(oracle_rows[j], x)
        assert (ei == two_i[i]).all(), f"query {i}"
        assert (bits(es) == bits(two_s[i])).all(), f"query {i}"

    # (2) all queries: exact shard-local top-x lists, all-gathered and merged (the plain exchange)
    parts_s = torch.zeros((n_shards, Q, x), dtype=torch.float32, device=dev)
    parts_i = torch.zeros((n_shards, Q, x), dtype=torch.int32, device=dev)
    for r, (c, st, sh, qq, lo) in enumerate(shards):
        _lib.check(c._L.myrio_score_topx_async(c.h, sh.h, qq.h, 0, Q, x, 0, C.c_void_p(parts_s[r].data_ptr()),
                                               C.c_void_p(parts_i[r].data_ptr())))
        c.sync()
    ms = torch.zeros((Q, x), dtype=torch.float32, device=dev)
    mi = torch.zeros((Q, x), dtype=torch.int32, device=dev)
    _lib.check(base._L.myrio_topx_merge_async(base.h, C.c_void_p(parts_s.data_ptr()), C.c_void_p(parts_i.data_ptr()),
                                              n_shards, Q, x, C.c_void_p(ms.data_ptr()), C.c_void_p(mi.data_ptr())))
    base.sync()
    assert (mi.cpu().numpy().view(np.uint32) == two_i).all()
    assert (bits(ms.cpu().numpy()) == bits(two_s)).all()
    # the exchanged bound pruned something on the shards that do not hold a query's clade
    assert (tops == 0).any()


def test_cluster_100k_reads(api, ctx):
    """configs[4] clustering: read-read clustering of 100 k mixed reads of four genes (k = 6, C = 4) with
    silhouette trimming 1.4.  All assignments and the iteration count vs the oracle; the O(N^2) silhouette
    pass is checked on sampled reads (the oracle computes inner() for the sample, same f32 order), and the
    trimming rule (clustering.rs:244-266) is recomputed from the scores of ALL reads."""
    trees = [synth.make_gene_tree(2000, n, s) for _, n, s in synth.FOUR_GENES]
    parts = [synth.make_reads(t, sz(25_000, 500), 3003 + i) for i, t in enumerate(trees)]
    reads = synth.concat(parts)
    reads = reads.subset(np.random.default_rng(5).permutation(reads.n))
    m = api.MyrSeqs(ctx, reads)
    ssdcf = 1.4
    out = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=4,
                                                  silhouette_trimming=ssdcf))
    oc6 = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    mask = np.zeros(reads.n, bool)
    mask[np.random.default_rng(9).choice(reads.n, min(96, reads.n), replace=False)] = True
    oassign, oiters, osil = oracle.cluster(oc6, 20, None, 4, silhouette=ssdcf, silhouette_sample=mask)
    assert out.n_iters == oiters
    trimmed = out.assign == -2
    assert (out.assign[~trimmed] == oassign[~trimmed]).all()  # final assignment of every read (:219-233)
    checked = mask & ~np.isnan(osil)
    assert checked.sum() >= min(64, reads.n // 2)
    assert (bits(out.silhouette[checked]) == bits(osil[checked])).all()
    # trimming: s < mean_c - ssdcf * std_c, population std, sequential f32 sums in member order
    f = np.float32
    n_trim = 0
    for c in range(4):
        members = np.nonzero(oassign == c)[0]
        if len(members) == 0:
            continue
        s = out.silhouette[members].astype(np.float32)
        n = f(len(members))
        mean = f(np.add.accumulate(s, dtype=np.float32)[-1]) / n
        d = (s - mean).astype(np.float32)
        sq = f(np.add.accumulate((d * d).astype(np.float32), dtype=np.float32)[-1])
        std = np.sqrt(f(f(1.0) / n) * sq).astype(np.float32)
        cut = f(mean - f(std * f(ssdcf)))
        exp_trim = s < cut
        assert (trimmed[members] == exp_trim).all(), f"cluster {c}"
        n_trim += int(exp_trim.sum())
    assert n_trim == int(trimmed.sum()) and n_trim > 0
