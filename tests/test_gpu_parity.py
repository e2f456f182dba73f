"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs.  Integer/byte/index results must be bit-exact; f32 scores
are compared bit-exactly too (the kernels keep the reference's summation order,
which is stricter than the 1e-6 relative tolerance north_star allows)."""
import numpy as np
import pytest

import oracle
from myrio_b200 import synth

pytestmark = pytest.mark.gpu


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


def seqset(seqs, quals=None):
    codes = np.concatenate([np.array(["ACGT".index(c) for c in s], np.uint8) for s in seqs]) if seqs else np.zeros(0, np.uint8)
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    q = None
    if quals is not None:
        q = np.concatenate([np.array(x, np.uint8) for x in quals]) if quals else np.zeros(0, np.uint8)
    return synth.SeqSet(codes, off, q)


def assert_reads_parity(api, ctx, reads, k, cutoff):
    got = api.MyrSeqs(ctx, reads).compute_kmer_counts(k, cutoff)
    ro, keys, vals, hck = got.download()
    exp = oracle.count_reads(reads.codes, reads.qual, reads.base_off, k, cutoff)
    assert (ro == exp.row_off).all()
    assert (keys == exp.keys).all()
    assert (bits(vals) == bits(exp.vals)).all()
    assert (hck == exp.aux).all()
    return got, exp


# ---------------------------------------------------------------- K1 reads

def test_reference_kat_simfunc(api, ctx):
    # myrio-core/src/similarity.rs:217-232 through the GPU path
    f32_min = float(np.finfo(np.float32).min)
    s = seqset(["ACTG", "ACTC"], [[0] * 4, [0] * 4])
    got = api.MyrSeqs(ctx, s).compute_kmer_counts(2, f32_min)
    ro, keys, vals, hck = got.download()
    assert hck.tolist() == [6, 6]
    a = (keys[ro[0]:ro[1]], vals[ro[0]:ro[1]])
    b = (keys[ro[1]:ro[2]], vals[ro[1]:ro[2]])
    cos = oracle.similarity(oracle.SIM_COSINE, False, a[0], a[1], b[0], b[1])
    assert abs(cos - np.float32(2 / 3)) < np.finfo(np.float32).eps
    nro, nkeys, nvals, _ = got.into_normalized_l2().download()
    tile = api.pairwise(got.into_normalized_l2(), got.into_normalized_l2())
    assert abs(tile[0, 1] - cos) < np.finfo(np.float32).eps
    assert tile[0, 1] == tile[1, 0]
    # overlap(kc1, kc2) == 0.5 on the raw counts (similarity.rs:226)
    raw = api.SFVecs.upload(ctx, ro, keys, vals)
    assert abs(api.pairwise(raw, raw, api.SIM_OVERLAP)[0, 1] - np.float32(0.5)) < np.finfo(np.float32).eps


@pytest.mark.parametrize("k", [2, 6, 12, 18, 24, 31, 32])
def test_count_reads_k_sweep(api, ctx, k):
    tree = synth.make_tree(50, 5)
    reads = synth.make_reads(tree, 300, 6)
    cutoff = 0.1 if k >= 12 else 0.2
    assert_reads_parity(api, ctx, reads, k, cutoff)


def test_msb_first_key_order_matches_oracle(api):
    # the alternative packing of a k-mer key (myrio_ctx_set_kmer_bit_order): reads, ACGT and ambiguous leaves,
    # and scores on top of them, against the oracle in the same mode
    c = api.Context(0)
    c.set_kmer_bit_order(True)
    oracle.set_kmer_msb_first(True)
    try:
        tree = synth.make_tree(90, 19)
        reads = synth.make_reads(tree, 40, 23)
        for k, cut in ((6, 0.2), (18, 0.1), (32, 0.05)):
            got = api.MyrSeqs(c, reads).compute_kmer_counts(k, cut)
            ro, keys, vals, hck = got.download()
            oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, k, cut)
            assert (ro == oc.row_off).all() and (keys == oc.keys).all() and (vals == oc.vals).all()
            assert (hck == oc.aux).all()
        iupac = synth.tree_to_iupac(tree).copy()
        rng = np.random.default_rng(3)
        iupac[rng.choice(len(iupac), 40, replace=False)] = 15  # N
        store = api.TaxTreeStore(c, iupac=iupac, base_off=tree.base_off)
        store.compute_and_append_kmer_counts(18, 32, seed=9)
        lro, lk, lv, resc = store.kmer_store_counts_vec[0].download()
        ol = oracle.count_leaves(iupac, tree.base_off, 18, 32, seed=9)
        assert (lro == ol.row_off).all() and (lk == ol.keys).all() and (lv == ol.vals).all() and (resc == ol.aux).all()
        ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32, tile_leaves=32)
        q = api.MyrSeqs(c, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
        res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 5, full_scores=True)
        oln = oracle.normalize(ol)
        oqn = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
        for i in range(reads.n):
            kk, vv = oqn.row(i)
            assert (bits(oracle.score_all(oln, kk, vv)) == bits(res.scores[i])).all()
    finally:
        oracle.set_kmer_msb_first(False)


def test_count_reads_edge_cases(api, ctx):
    rng = np.random.default_rng(1)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))

    seqs = ["ACGT", "A", rnd(17), rnd(18), rnd(19), "A" * 200, "T" * 64, rnd(31), rnd(32), rnd(33), rnd(64),
            rnd(65), rnd(700), "ACGT" * 100]
    quals = [[40] * len(s) for s in seqs]
    quals[3] = [1] * 18          # every window fails
    quals[12] = list(rng.integers(1, 41, 700))
    s = seqset(seqs, quals)
    for k in (2, 18, 32):
        assert_reads_parity(api, ctx, s, k, 0.1)
    # cutoff extremes
    assert_reads_parity(api, ctx, s, 18, float(np.finfo(np.float32).min))
    assert_reads_parity(api, ctx, s, 18, 1.0)


def test_count_reads_empty_batch(api, ctx):
    s = seqset([], [])
    got = api.MyrSeqs(ctx, s).compute_kmer_counts(18, 0.1)
    assert got.n_rows == 0 and got.nnz == 0


def test_count_reads_long_reads_use_every_size_class(api, ctx):
    rng = np.random.default_rng(2)
    lens = [400, 600, 1100, 2300, 4500, 8800, 20000]   # 1k..16k shared-memory classes + global fallback
    codes = rng.integers(0, 4, sum(lens), dtype=np.uint8)
    qual = rng.integers(20, 41, sum(lens)).astype(np.uint8)
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    assert_reads_parity(api, ctx, synth.SeqSet(codes, off, qual), 18, 0.1)


def test_count_reads_errors(api, ctx):
    s = seqset(["ACGTACGT"], [[40] * 8])
    m = api.MyrSeqs(ctx, s)
    for k in (0, 1, 33):
        with pytest.raises(api.MyrioError) as e:
            m.compute_kmer_counts(k, 0.1)
        assert e.value.code == 1  # MYRIO_ERR_INVALID_K (the reference panics)
    bad = seqset(["ACGTACGT"], [[40] * 7 + [128]])
    with pytest.raises(api.MyrioError) as e:
        api.MyrSeqs(ctx, bad).compute_kmer_counts(2, 0.1)
    assert e.value.code == 5  # MYRIO_ERR_BAD_QUALITY


# --------------------------------------------------------------- K2 leaves

def assert_leaves_parity(api, ctx, iupac, base_off, k, R, seed=0, base=0):
    st = api.TaxTreeStore(ctx, iupac=iupac, base_off=base_off, leaf_id_base=base)
    st.compute_and_append_kmer_counts(k, R, seed)
    ro, keys, vals, rescale = st.kmer_store_counts_vec[0].download()
    exp = oracle.count_leaves(iupac, base_off, k, R, seed, base)
    assert (ro == exp.row_off).all()
    assert (keys == exp.keys).all()
    assert (vals == exp.vals).all()
    assert (bits(rescale) == bits(exp.aux)).all()
    return st, exp


@pytest.mark.parametrize("k", [2, 6, 18, 32])
def test_count_leaves_acgt(api, ctx, k):
    tree = synth.make_tree(200, 9)
    assert_leaves_parity(api, ctx, synth.tree_to_iupac(tree), tree.base_off, k, 32)


@pytest.mark.parametrize("R", [0, 1, 3, 32, 100])
def test_count_leaves_resample_counts(api, ctx, R):
    tree = synth.make_tree(20, 10)
    assert_leaves_parity(api, ctx, synth.tree_to_iupac(tree), tree.base_off, 12, R)


def test_count_leaves_gaps_ambiguity_and_sentinels(api, ctx):
    rng = np.random.default_rng(3)
    tree = synth.make_tree(30, 11)
    iu = synth.tree_to_iupac(tree).copy()
    # sprinkle gaps (0) and ambiguity codes over some leaves
    for leaf in range(0, 30, 3):
        b, e = int(tree.base_off[leaf]), int(tree.base_off[leaf + 1])
        pos = rng.integers(b, e, 12)
        iu[pos[:6]] = 0
        iu[pos[6:]] = rng.choice([3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15], 6)
    # an all-gap leaf, a too-short leaf, a heavily ambiguous leaf
    extra = [np.zeros(40, np.uint8), np.array([8, 4, 2], np.uint8), np.full(60, 15, np.uint8)]
    iu2 = np.concatenate([iu] + extra)
    off = np.concatenate([tree.base_off, tree.base_off[-1] + np.cumsum([40, 3, 60]).astype(np.uint64)])
    for k, seed in ((6, 1), (18, 2)):
        assert_leaves_parity(api, ctx, iu2, off, k, 32, seed=seed, base=1000)


def test_count_leaves_long_leaf(api, ctx):
    rng = np.random.default_rng(4)
    lens = [700, 3000, 9000, 30000]
    iu = synth.DNA_TO_IUPAC[rng.integers(0, 4, sum(lens))]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    assert_leaves_parity(api, ctx, iu, off, 18, 32)


# -------------------------------------------------- normalise + K3 index

def test_normalize_and_index_match_oracle(api, ctx):
    tree = synth.make_tree(700, 12)
    st, exp = assert_leaves_parity(api, ctx, synth.tree_to_iupac(tree), tree.base_off, 18, 32)
    nro, nkeys, nvals, _ = st.kmer_store_counts_vec[0].into_normalized_l2().download()
    oln = oracle.normalize(exp)
    assert (bits(nvals) == bits(oln.vals)).all()
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, tile_leaves=256)
    info = ttc.info()
    assert info["n_tiles"] == 3 and info["n_postings"] == len(exp.keys)
    for t in range(info["n_tiles"]):
        lo, hi = t * 256, min(700, (t + 1) * 256)
        sub = oracle.Csr(oln.row_off[lo:hi + 1] - oln.row_off[lo], oln.keys[int(oln.row_off[lo]):int(oln.row_off[hi])],
                         oln.vals[int(oln.row_off[lo]):int(oln.row_off[hi])])
        uk, of, pl, pv = oracle.invert(sub)
        guk, gof, gpl, gpv = ttc.export_tile(t)
        assert (guk == uk).all() and (gof == of).all() and (gpl == pl).all() and (bits(gpv) == bits(pv)).all()


# ----------------------------------------------------- K4 / K5 scoring

def build_case(api, ctx, n_leaves, n_reads, seed, k=18, tile_leaves=0):
    tree = synth.make_tree(n_leaves, seed)
    reads = synth.make_reads(tree, n_reads, seed + 1)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, k, 32, tile_leaves=tile_leaves)
    q = api.MyrSeqs(ctx, reads).compute_kmer_counts(k, 0.1).into_normalized_l2()
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, k, 32))
    oqn = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, k, 0.1))
    return tree, reads, ttc, q, oln, oqn


@pytest.mark.parametrize("tile_leaves", [64, 1000, 0])
def test_score_all_and_topx_bit_exact(api, ctx, tile_leaves):
    tree, reads, ttc, q, oln, oqn = build_case(api, ctx, 1500, 40, 21, tile_leaves=tile_leaves)
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 100, full_scores=True)
    for i in range(reads.n):
        k, v = oqn.row(i)
        s = oracle.score_all(oln, k, v)
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, 100)
        assert (ti == res.top_ids[i]).all()
        assert (bits(ts) == bits(res.top_scores[i])).all()
    st = ctx.score_stats()
    assert st["query_keys"] == int(oqn.row_off[-1])
    assert st["key_hits"] > 0 and st["postings_touched"] >= st["key_hits"]


def test_topx_bound_pruning_switch_gives_the_same_topx(api, ctx):
    # myrio_ctx_set_topx_bound_pruning: finishing only the scores that can reach a query's running bound, or
    # every score, is the same top-x bit for bit (multi-tile tree, so the sweep runs with a bound)
    tree, reads, ttc, q, oln, oqn = build_case(api, ctx, 3000, 60, 33, tile_leaves=256)
    a = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 25)
    ctx.set_topx_bound_pruning(False)
    try:
        b = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 25)
    finally:
        ctx.set_topx_bound_pruning(True)
    assert (a.top_ids == b.top_ids).all() and (bits(a.top_scores) == bits(b.top_scores)).all()
    for i in range(0, reads.n, 5):
        k, v = oqn.row(i)
        ts, ti = oracle.topx(oracle.score_all(oln, k, v), 25)
        assert (ti == a.top_ids[i]).all() and (bits(ts) == bits(a.top_scores[i])).all()


def test_score_jacard_tanimoto_and_k12(api, ctx):
    tree, reads, ttc, q, oln, oqn = build_case(api, ctx, 400, 12, 31, k=12)
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_JACARD_TANIMOTO, 7, full_scores=True)
    for i in range(reads.n):
        k, v = oqn.row(i)
        s = oracle.score_all(oln, k, v, sim=oracle.SIM_JACARD_TANIMOTO)
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, 7)
        assert (ti == res.top_ids[i]).all()


@pytest.mark.parametrize("k", [6, 18])
def test_score_overlap_and_direct_path(api, ctx, k):
    # Overlap (similarity.rs:172-182) runs as the reference's merge-join over the leaf vectors;
    # the same kernel reproduces the inverted-index results for Cosine / JacardTanimoto
    tree, reads, ttc, q, oln, oqn = build_case(api, ctx, 700, 24, 33, k=k, tile_leaves=128)
    x = 9
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_OVERLAP, x, full_scores=True)
    for i in range(reads.n):
        kk, vv = oqn.row(i)
        s = oracle.score_all(oln, kk, vv, sim=oracle.SIM_OVERLAP)
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, x)
        assert (ti == res.top_ids[i]).all() and (bits(ts) == bits(res.top_scores[i])).all()
    for sim in (api.SIM_COSINE, api.SIM_JACARD_TANIMOTO):
        a = api.TaxTreeResults.from_compute_tree(q, ttc, sim, x, full_scores=True)
        b = api.TaxTreeResults.from_compute_tree(q, ttc, sim, x, full_scores=True, direct=True)
        assert (bits(a.scores) == bits(b.scores)).all()
        assert (a.top_ids == b.top_ids).all() and (bits(a.top_scores) == bits(b.top_scores)).all()


def test_direct_topx_ties_and_short_rows(api, ctx):
    # duplicates give exact ties (smallest payload id first); x larger than the leaf count pads
    keys = [[1, 2, 3], [1, 2, 3], [7, 8], [1, 2, 3], [9], [2, 3, 4]]
    row_off = np.zeros(len(keys) + 1, np.uint64)
    row_off[1:] = np.cumsum([len(v) for v in keys])
    kk = np.concatenate([np.array(v, np.uint64) for v in keys])
    counts = api.SFVecs.upload(ctx, row_off, kk, np.ones(len(kk), np.uint16))
    ttc = api.TaxTreeCompute.from_counts(counts, 100)
    on = oracle.normalize(oracle.Csr(row_off, kk, np.ones(len(kk), np.float32)))
    qv = api.SFVecs.upload(ctx, np.array([0, 3], np.uint64), np.array([1, 2, 3], np.uint64),
                           np.ones(3, np.float32)).into_normalized_l2()
    qro, qk, qvv, _ = qv.download(aux=False)
    for sim, osim in ((api.SIM_OVERLAP, oracle.SIM_OVERLAP), (api.SIM_COSINE, oracle.SIM_COSINE)):
        s = oracle.score_all(on, qk, qvv, sim=osim)
        for x in (2, 4, 10):
            r = api.TaxTreeResults.from_compute_tree(qv, ttc, sim, x, direct=True)
            ts, ti = oracle.topx(s, x, id_base=100)
            assert (r.top_ids[0] == ti).all() and (bits(r.top_scores[0]) == bits(ts)).all()


def test_topx_ties_zero_rows_and_padding(api, ctx):
    # handcrafted leaf vectors: duplicates give exact ties, unrelated leaves give zeros
    keys = [[1, 2, 3], [1, 2, 3], [7, 8], [1, 2, 3], [9], [2, 3, 4]]
    row_off = np.zeros(len(keys) + 1, np.uint64)
    row_off[1:] = np.cumsum([len(x) for x in keys])
    k = np.concatenate([np.array(x, np.uint64) for x in keys])
    v = np.ones(len(k), np.uint16)
    counts = api.SFVecs.upload(ctx, row_off, k, v)
    ttc = api.TaxTreeCompute.from_counts(counts, leaf_id_base=100, tile_leaves=4)
    q = api.SFVecs.upload(ctx, np.array([0, 3, 4], np.uint64), np.array([1, 2, 3, 50], np.uint64),
                          np.array([1, 1, 1, 1], np.float32)).into_normalized_l2()
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 8, full_scores=True)
    oln = oracle.normalize(oracle.Csr(row_off, k, v))
    for i, (qk, qv) in enumerate([([1, 2, 3], [1, 1, 1]), ([50], [1])]):
        qv = np.array(qv, np.float32) / np.sqrt(np.float32(len(qv)))
        s = oracle.score_all(oln, np.array(qk, np.uint64), qv.astype(np.float32))
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, 8, id_base=100)
        assert (ti == res.top_ids[i]).all() and (bits(ts) == bits(res.top_scores[i])).all()
    assert res.top_ids[0][:3].tolist() == [100, 101, 103]      # ties: payload_id ascending
    assert res.top_ids[1][:6].tolist() == [100, 101, 102, 103, 104, 105] and res.top_ids[1][6] == 0xFFFFFFFF


def test_topx_merge_matches_single_shard(api, ctx):
    import ctypes as C
    import torch
    tree, reads, ttc, q, oln, oqn = build_case(api, ctx, 900, 16, 41)
    full = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 20)
    # two shards of the same leaves: [0,450) and [450,900)
    x, Q = 20, reads.n
    parts_s = torch.zeros((2, Q, x), dtype=torch.float32, device="cuda")
    parts_i = torch.zeros((2, Q, x), dtype=torch.int32, device="cuda")
    keep = []
    for p, (lo, hi) in enumerate([(0, 450), (450, 900)]):
        sub = tree.subset(np.arange(lo, hi))
        st = api.TaxTreeStore(ctx, sub, leaf_id_base=lo)
        shard = api.TaxTreeCompute.from_store_tree(st, 18, 32, tile_leaves=128)
        keep.append((st, shard))
        from myrio_b200 import _lib
        _lib.check(ctx._L.myrio_score_topx_async(ctx.h, shard.h, q.h, 0, Q, x, 0, C.c_void_p(parts_s[p].data_ptr()),
                                                 C.c_void_p(parts_i[p].data_ptr())))
    out_s = torch.zeros((Q, x), dtype=torch.float32, device="cuda")
    out_i = torch.zeros((Q, x), dtype=torch.int32, device="cuda")
    _lib.check(ctx._L.myrio_topx_merge_async(ctx.h, C.c_void_p(parts_s.data_ptr()), C.c_void_p(parts_i.data_ptr()),
                                             2, Q, x, C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr())))
    ctx.sync()
    assert (out_i.cpu().numpy().view(np.uint32) == full.top_ids).all()
    assert (bits(out_s.cpu().numpy()) == bits(full.top_scores)).all()


def test_two_phase_sharded_topx_with_bound_exchange(api):
    # three shards (a context each, as three ranks would have): seed pass on every shard, the bounds are
    # maxed like the all-reduce does, sweep with the global bound, merge of the composite lists ==
    # the unsharded top-x, although a shard's own list may now be shorter than x
    import ctypes as C
    import torch
    from myrio_b200 import _lib
    base = api.Context(0)
    tree, reads, ttc, q, oln, oqn = build_case(api, base, 1200, 24, 43, tile_leaves=64)
    x, Q = 20, reads.n
    full = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, x)
    qro, qk, qv, _ = q.download(aux=False)
    shards = []
    for lo, hi in ((0, 400), (400, 800), (800, 1200)):
        c = api.Context(0)
        st = api.TaxTreeStore(c, tree.subset(np.arange(lo, hi)), leaf_id_base=lo)
        shards.append((c, st, api.TaxTreeCompute.from_store_tree(st, 18, 32, tile_leaves=64),
                       api.SFVecs.upload(c, qro, qk, qv)))
    bounds = torch.zeros((3, Q), dtype=torch.int32, device="cuda")
    for r, (c, st, sh, qq) in enumerate(shards):
        _lib.check(c._L.myrio_score_topx_seed_async(c.h, sh.h, qq.h, 0, Q, x, 0, C.c_void_p(bounds[r].data_ptr())))
        c.sync()
    gbound = bounds.max(dim=0).values.contiguous()
    assert (gbound > bounds).any()  # the exchange raises some shard's bound
    tops = torch.zeros((3, Q, x), dtype=torch.int64, device="cuda")
    for r, (c, st, sh, qq) in enumerate(shards):
        _lib.check(c._L.myrio_score_topx_sweep_async(c.h, sh.h, Q, x, 0, C.c_void_p(gbound.data_ptr()),
                                                     C.c_void_p(tops[r].data_ptr())))
        c.sync()
    assert (tops == 0).any()  # pruned by another shard's bound: shorter than x somewhere
    out_s = torch.zeros((Q, x), dtype=torch.float32, device="cuda")
    out_i = torch.zeros((Q, x), dtype=torch.int32, device="cuda")
    _lib.check(base._L.myrio_topx_merge_composites_async(base.h, C.c_void_p(tops.data_ptr()), 3, Q, x,
                                                         C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr())))
    base.sync()
    assert (out_i.cpu().numpy().view(np.uint32) == full.top_ids).all()
    assert (bits(out_s.cpu().numpy()) == bits(full.top_scores)).all()
    # protocol errors: sweep without seed, seed that does not fit one batch is reported, not silently wrong
    c0, _, sh0, _ = shards[0]
    assert c0._L.myrio_score_topx_sweep_async(c0.h, sh0.h, Q, x, 0, None, C.c_void_p(tops[0].data_ptr())) == _lib.ERR_BAD_ARG


# ------------------------------------------------------------ K6 cluster

def two_gene_reads(n_each, seed):
    t1, t2 = synth.make_tree(80, seed), synth.make_tree(80, seed + 1)
    r1, r2 = synth.make_reads(t1, n_each, seed + 2), synth.make_reads(t2, n_each, seed + 3)
    codes = np.concatenate([r1.codes, r2.codes])
    qual = np.concatenate([r1.qual, r2.qual])
    off = np.concatenate([r1.base_off, r2.base_off[1:] + r1.base_off[-1]])
    return synth.SeqSet(codes, off, qual)


def test_pairwise_tile_bit_exact(api, ctx):
    reads = two_gene_reads(40, 51)
    oc = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2))
    g = api.MyrSeqs(ctx, reads).compute_kmer_counts(6, 0.2).into_normalized_l2()
    for sim in (api.SIM_COSINE, api.SIM_JACARD_TANIMOTO, api.SIM_OVERLAP):
        assert (bits(api.pairwise(g, g, sim)) == bits(oracle.pairwise(oc, oc, sim))).all()
    # large k takes the merge-join fallback
    oc18 = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
    g18 = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    assert (bits(api.pairwise(g18, g18)) == bits(oracle.pairwise(oc18, oc18))).all()


@pytest.mark.parametrize("silhouette", [None, 1.4])
@pytest.mark.parametrize("with_init", [False, True])
def test_cluster_matches_oracle(api, ctx, silhouette, with_init):
    reads = two_gene_reads(150, 61)
    m = api.MyrSeqs(ctx, reads)
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    init_g = init_o = None
    if with_init:
        # centroids of the first/last 20 reads as supplied initial centroids
        ck1, cv1 = oracle.centroid(oc, np.arange(0, 20))
        ck2, cv2 = oracle.centroid(oc, np.arange(280, 300))
        ro = np.array([0, len(ck1), len(ck1) + len(ck2)], np.uint64)
        init_o = oracle.Csr(ro, np.concatenate([ck1, ck2]), np.concatenate([cv1, cv2]))
        init_g = api.SFVecs.upload(ctx, ro, init_o.keys, init_o.vals)
    for expected in (2, 3):
        oassign, oiters, osil = oracle.cluster(oc, 20, init_o, expected, silhouette=silhouette)
        out = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, intial_centroids=init_g,
                                                      expected_nb_of_clusters=expected,
                                                      silhouette_trimming=silhouette))
        assert (out.assign == oassign).all()
        assert out.n_iters == oiters
        if silhouette is not None:
            both = ~np.isnan(osil)
            assert (np.isnan(out.silhouette) == np.isnan(osil)).all()
            assert (bits(out.silhouette[both]) == bits(osil[both])).all()


@pytest.mark.parametrize("sim", ["jacard", "overlap"])
def test_cluster_other_similarities_match_oracle(api, ctx, sim):
    gsim, osim = {"jacard": (api.SIM_JACARD_TANIMOTO, oracle.SIM_JACARD_TANIMOTO),
                  "overlap": (api.SIM_OVERLAP, oracle.SIM_OVERLAP)}[sim]
    reads = two_gene_reads(90, 67)
    m = api.MyrSeqs(ctx, reads)
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    for silhouette in (None, 1.4):
        oassign, oiters, osil = oracle.cluster(oc, 20, None, 2, silhouette=silhouette, sim=osim)
        out = api.cluster(m, api.ClusteringParameters(k=6, t1=0.2, t2=20, similarity=gsim,
                                                      expected_nb_of_clusters=2, silhouette_trimming=silhouette))
        assert (out.assign == oassign).all() and out.n_iters == oiters
        if silhouette is not None:
            both = ~np.isnan(osil)
            assert (bits(out.silhouette[both]) == bits(osil[both])).all()


@pytest.mark.parametrize("world", [2, 3])
def test_cluster_row_sharded_ranks_agree_with_oracle(api, world):
    # myrio_cluster_sharded with `world` ranks emulated as threads on ONE GPU (a context each); the
    # all-gather callback is a barrier-synchronised copy between the ranks' host buffers
    import ctypes as C
    import threading
    reads = two_gene_reads(101, 91)
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 6, 0.2)
    oassign, oiters, osil = oracle.cluster(oc, 20, None, 3, silhouette=1.4)
    barrier = threading.Barrier(world)
    stage = {}
    results, errors = [None] * world, []

    def make_cb(rank):
        def cb(user, host_buf, nbytes):
            nbytes = int(nbytes)
            arr = np.ctypeslib.as_array(C.cast(host_buf, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
            stage[rank] = arr[rank * nbytes:(rank + 1) * nbytes].copy()
            barrier.wait(timeout=120)
            for r in range(world):
                arr[r * nbytes:(r + 1) * nbytes] = stage[r]
            barrier.wait(timeout=120)
            return 0
        return cb

    def run(rank):
        try:
            c = api.Context(0)
            m = api.MyrSeqs(c, reads)
            prm = api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=3, silhouette_trimming=1.4)
            results[rank] = api.cluster(m, prm, shard=(rank, world, make_cb(rank)))
        except Exception as e:  # pragma: no cover
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    both = ~np.isnan(osil)
    for out in results:
        assert (out.assign == oassign).all() and out.n_iters == oiters
        assert (bits(out.silhouette[both]) == bits(osil[both])).all()


def test_cluster_errors(api, ctx):
    reads = two_gene_reads(10, 71)
    m = api.MyrSeqs(ctx, reads)
    with pytest.raises(api.MyrioError) as e:
        api.cluster(m, api.ClusteringParameters(expected_nb_of_clusters=0))
    assert e.value.code == 2
    with pytest.raises(api.MyrioError) as e:
        api.cluster(m, api.ClusteringParameters(t2=10 ** 9, expected_nb_of_clusters=2))
    assert e.value.code == 7


def test_compute_cluster_centroid_search_k(api, ctx):
    # run.rs:301-309: the query of a cluster is the centroid of its reads at search_k
    reads = two_gene_reads(30, 81)
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1)
    g = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1)
    ids = np.arange(0, 30)
    ck, cv = oracle.centroid(oc, ids)
    ro, keys, vals, _ = api.compute_cluster_centroid(g, ids).download(aux=False)
    assert ro.tolist() == [0, len(ck)] and (keys == ck).all() and (bits(vals) == bits(cv)).all()


# ------------------------------------------- long queries, dense keys, properties

def test_fingerprints_match_oracle(api, ctx):
    # tax/compute.rs:51-77 with explicit leaf samples: local fingerprint = normalize(sum over the sampled
    # leaves of count * rescale) at cluster_k, global = normalize(sum of the local ones).  Leaves with
    # ambiguity codes make the rescale factors (and the summed values) non-integer f32.
    rng = np.random.default_rng(17)
    tree = synth.make_tree(120, 41)
    iupac = synth.tree_to_iupac(tree).copy()
    amb = rng.choice(len(iupac), 60, replace=False)
    iupac[amb] = rng.choice(np.array([3, 5, 6, 9, 10, 12, 7, 15], np.uint8), 60)  # R Y S W K M ... N
    store = api.TaxTreeStore(ctx, iupac=iupac, base_off=tree.base_off)
    samples = [rng.choice(120, 16, replace=False) for _ in range(5)] + [np.array([7], np.int64)]
    local, glob = api.fingerprints(store, 6, samples, 32, seed=3)
    oc = oracle.count_leaves(iupac, tree.base_off, 6, 32, seed=3)
    scaled = oracle.Csr(oc.row_off, oc.keys, (oc.vals.astype(np.float32) *
                                              np.repeat(oc.aux, np.diff(oc.row_off).astype(np.int64))).astype(np.float32))
    assert (np.abs(oc.aux - 1 / 32) > 1e-9).any()  # at least one non-trivial rescale factor
    rows = []
    for ids, fp in zip(samples, local):
        ek, ev = oracle.centroid(scaled, ids)
        _, gk, gv, _ = fp.download(aux=False)
        assert (gk == ek).all() and (bits(gv) == bits(ev)).all()
        rows.append((ek, ev))
    ro = np.zeros(len(rows) + 1, np.uint64)
    ro[1:] = np.cumsum([len(r[0]) for r in rows])
    lb = oracle.Csr(ro, np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows]))
    ek, ev = oracle.centroid(lb, np.arange(len(rows)))
    _, gk, gv, _ = glob.download(aux=False)
    assert (gk == ek).all() and (bits(gv) == bits(ev)).all()


def test_leaves_as_queries_self_score_and_long_hit_lists(api, ctx):
    # every leaf vector scored against the tree: ~1266 keys per query, all present in the
    # query's own tile (hit list drained several times), top-1 must be the leaf itself
    tree = synth.make_tree(2500, 91)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, keep_leaf_vecs=True)
    q = ttc.leaf_vecs
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 5, q_first=1000, q_count=200,
                                               full_scores=True)
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
    for j in range(200):  # every query: a hit list that mixes ordered and split batches is a rare event
        k, v = oln.row(1000 + j)
        s = oracle.score_all(oln, k, v)
        assert (bits(s) == bits(res.scores[j])).all(), j
        ts, ti = oracle.topx(s, 5)
        assert (ti == res.top_ids[j]).all() and (bits(ts) == bits(res.top_scores[j])).all(), j
    assert (res.top_ids[:, 0] == np.arange(1000, 1200)).all()
    # ~1266 sequential f32 additions: the self-score is 1 only up to the rounding the reference has too
    assert np.allclose(res.top_scores[:, 0], 1.0, atol=1e-4)


def test_cluster_centroid_query_against_tree(api, ctx):
    # run.rs:301-309 + results.rs:82-93: the query is the centroid of a cluster's reads at
    # search_k (thousands of keys), scored against every leaf
    tree = synth.make_tree(1800, 93)
    reads = synth.make_reads(tree, 120, 94, leaf_ids=np.arange(600, 640))
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32)
    counts = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1)
    cen = api.compute_cluster_centroid(counts, np.arange(reads.n))
    res = api.TaxTreeResults.from_compute_tree(cen, ttc, api.SIM_COSINE, 100, full_scores=True)
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1)
    ck, cv = oracle.centroid(oc, np.arange(reads.n))
    assert len(ck) > 3000
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
    s = oracle.score_all(oln, ck, cv)
    assert (bits(s) == bits(res.scores[0])).all()
    ts, ti = oracle.topx(s, 100)
    assert (ti == res.top_ids[0]).all() and (bits(ts) == bits(res.top_scores[0])).all()
    assert 600 <= ti[0] < 640
    # many tiles x a long query: the presence bit-matrix of the parallel plan-fill kernel does not
    # fit in shared memory, the per-key sequential plan-fill kernel takes over -- same bits
    ttc4 = api.TaxTreeCompute.from_store_tree(st, 18, 32, tile_leaves=4)
    res4 = api.TaxTreeResults.from_compute_tree(cen, ttc4, api.SIM_COSINE, 100, full_scores=True)
    assert (bits(res4.scores) == bits(res.scores)).all() and (res4.top_ids == res.top_ids).all()


def test_dense_and_sparse_tilings_agree(api, ctx):
    # the same tree indexed with tiles that allow dense keys (<= 1024 leaves) and with tiles
    # that do not (> 1024) must give identical bits; also non-ACGT-style counts (values that
    # are not all unit[leaf]) must demote keys to the sparse form, not change results
    tree = synth.make_tree(3000, 95)
    reads = synth.make_reads(tree, 30, 96)
    st = api.TaxTreeStore(ctx, tree)
    q = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    outs = []
    for tl in (1024, 512, 2048, 4096):
        ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, tile_leaves=tl)
        outs.append(api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 50, full_scores=True))
    for o in outs[1:]:
        assert (bits(o.scores) == bits(outs[0].scores)).all()
        assert (o.top_ids == outs[0].top_ids).all() and (bits(o.top_scores) == bits(outs[0].top_scores)).all()
    # perturbed counts: bump a few counts so that conserved keys carry non-unit values
    ro, keys, vals, _ = st.kmer_store_counts_vec[0].download(aux=False)
    vals = vals.copy()
    rng = np.random.default_rng(3)
    vals[rng.integers(0, len(vals), len(vals) // 50)] = 64
    pert = api.SFVecs.upload(ctx, ro, keys, vals)
    ttc = api.TaxTreeCompute.from_counts(pert)
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 20, full_scores=True)
    oln = oracle.normalize(oracle.Csr(ro, keys, vals))
    qro, qk, qv, _ = q.download(aux=False)
    for i in range(0, 30, 3):
        s = oracle.score_all(oln, qk[int(qro[i]):int(qro[i + 1])], qv[int(qro[i]):int(qro[i + 1])])
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, 20)
        assert (ti == res.top_ids[i]).all()


def test_topx_many_candidates_chunked_merge(api, ctx):
    # x = 1024 with many tiles: more than 4096 candidates per query reach the merge kernel
    tree = synth.make_tree(9000, 97)
    reads = synth.make_reads(tree, 6, 98)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32)
    q = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 1024, full_scores=True)
    for i in range(reads.n):
        ts, ti = oracle.topx(res.scores[i], 1024)
        assert (ti == res.top_ids[i]).all() and (bits(ts) == bits(res.top_scores[i])).all()
    with pytest.raises(api.MyrioError):
        api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 1025)


# ------------------------------------------------------------- FASTQ ingestion

def test_fastq_ingestion_matches_oracle_and_prepacked_path(api, ctx):
    tree = synth.make_tree(40, 101)
    reads = synth.make_reads(tree, 400, 102, min_length=1, min_mean_qual=0.0)  # unfiltered: the device filters
    seq = synth.ASCII[reads.codes]
    qual = (reads.qual + 33).astype(np.uint8)
    m, keep = api.MyrSeqs.from_fastq_records(ctx, seq, qual, reads.base_off, 200, 17.0, 40)
    ocodes, oqual, ooff, okeep = oracle.fastq_preprocess(seq, qual, reads.base_off, 200, 17.0, 40)
    assert (keep == okeep.astype(bool)).all() and 0 < keep.sum() < reads.n
    got = m.compute_kmer_counts(18, 0.1).download()
    exp = oracle.count_reads(ocodes, oqual, ooff, 18, 0.1)
    assert (got[0] == exp.row_off).all() and (got[1] == exp.keys).all()
    assert (bits(got[2]) == bits(exp.vals)).all() and (got[3] == exp.aux).all()
    # same result as packing on the host
    sub = reads.subset(np.nonzero(keep)[0])
    ref = api.MyrSeqs(ctx, sub).compute_kmer_counts(18, 0.1).download()
    assert all((a == b).all() for a, b in zip(got[:2], ref[:2]))
    # a non-ACGT base fails the whole batch like read_fastq
    bad = seq.copy()
    bad[int(reads.base_off[7]) + 3] = ord("N")
    with pytest.raises(api.MyrioError) as e:
        api.MyrSeqs.from_fastq_records(ctx, bad, qual, reads.base_off)
    assert e.value.code == 9 and "record 7" in str(e.value)
    # empty batch
    m0, k0 = api.MyrSeqs.from_fastq_records(ctx, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert m0.n == 0 and len(k0) == 0


def test_fasta_ingestion_matches_oracle_and_prepacked_path(api, ctx):
    # tax/store.rs:166-195: Seq::<Iupac>::from_str per record on the device; ambiguity codes, gaps,
    # an empty record, and records with an invalid byte (skipped like the reference, store.rs:181-187)
    rng = np.random.default_rng(23)
    tree = synth.make_tree(60, 111)
    iupac = synth.tree_to_iupac(tree).copy()
    pos = rng.choice(len(iupac), 300, replace=False)
    iupac[pos] = rng.choice(np.array([0, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15], np.uint8), 300)
    text = synth.iupac_to_ascii(iupac).copy()
    off = tree.base_off.copy()
    bad = {5: ord("a"), 17: ord("Z"), 40: ord("U")}
    for r, ch in bad.items():
        text[int(off[r]) + 11] = ch
    # an empty record in the middle (header followed directly by the next header)
    off = np.concatenate([off[:31], off[30:]])
    st, keep = api.TaxTreeStore.from_fasta_records(ctx, text, off)
    ocodes, ooff, okeep = oracle.fasta_to_iupac(text, off)
    assert (keep == okeep).all() and keep.sum() == len(off) - 1 - len(bad)
    assert not keep[5] and not keep[17] and keep[30] and not keep[41]  # record 40 moved to 41
    assert st.n_leaves == keep.sum()
    for k in (6, 18):
        got = st._count(k, 32, 3).download()
        exp = oracle.count_leaves(ocodes, ooff, k, 32, seed=3)
        assert (got[0] == exp.row_off).all() and (got[1] == exp.keys).all() and (got[2] == exp.vals).all()
        assert (bits(got[3]) == bits(exp.aux)).all()
    # same bits as packing on the host
    ref = api.TaxTreeStore(ctx, iupac=ocodes, base_off=ooff)._count(18, 32, 3).download()
    assert all((a == b).all() for a, b in zip(got[:3], ref[:3]))
    s0, k0 = api.TaxTreeStore.from_fasta_records(ctx, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert s0.n_leaves == 0 and len(k0) == 0


def test_reference_kat_from_unsorted_pairs_through_the_gpu(api, ctx):
    # data/sparse.rs:623-633 (pairs half): the GPU's sort + sum of (key, value) pairs is the gather / radix sort /
    # run reduction behind myrio_compute_cluster_centroid; six one-entry rows summed = from_unsorted_pairs, then L2
    pairs = [(10, 1.0), (2, 4.0), (1, 1.0), (10, 7.0), (10, 2.0), (3, 9.0)]
    ro = np.arange(7, dtype=np.uint64)
    rows = api.SFVecs.upload(ctx, ro, np.array([p[0] for p in pairs], np.uint64), np.array([p[1] for p in pairs], np.float32))
    _, k, v, _ = api.compute_cluster_centroid(rows, np.arange(6)).download(aux=False)
    ek, ev = oracle.sort_reduce_pairs_f32(pairs)
    assert k.tolist() == [1, 2, 3, 10] == ek.tolist()
    norm = oracle.normalize(oracle.Csr(np.array([0, 4], np.uint64), ek, ev))
    assert (bits(v) == bits(norm.vals)).all()


@pytest.mark.parametrize("count_mode,ub,plan,masks", [("1", "1", "2", "1"), ("1", "0", "2", "1"), ("0", "1", "2", "1"),
                                                      ("1", "1", "1", "1"), ("1", "1", "2", "0")])
def test_k4_kernel_variants(count_mode, ub, plan, masks):
    # the K4 kernels (count mode with / without the bound pass of the fused top-x, add-as-you-go), both
    # generations of the plan kernels and the plan with / without dense-key masks, forced through the odd-term
    # cases, each in its own process (the switches are read once per process)
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MYRIO_K4_COUNT=count_mode, MYRIO_K4_UB=ub, MYRIO_K4A_PLAN=plan, MYRIO_K4A_MASKS=masks)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "k4_variant_worker.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"K4_VARIANT_OK count={count_mode}" in r.stdout
