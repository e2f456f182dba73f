"""GPU parity of the post-scoring pass (myrio_results_from_compute_tree: K8a branch statistics, K8b per-branch
leaf selection on the device; trim / cut / pooling / confidence in the library's host code) against the oracle's
restatement of tax/results.rs:96-381 and run.rs:405-466, through the C ABI.  Every field is compared bit for bit."""
import numpy as np
import pytest

import oracle
from myrio_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from myrio_b200 import api as _api
    return _api


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


FIELDS = ("kind", "arg", "src_node", "nb_leaves", "conf_state", "force_keep", "payload_src_id")
F32_FIELDS = ("mean", "max", "pool_score", "conf", "payload_score")


def assert_same_tree(got, exp, what=""):
    for f in FIELDS:
        assert (getattr(got, f) == getattr(exp, f)).all(), f"{what}: {f}"
    for f in F32_FIELDS:
        assert (bits(getattr(got, f)) == bits(getattr(exp, f))).all(), f"{what}: {f}"


def build(api, ctx, n_leaves, seed, k=18):
    tree = synth.make_tree(n_leaves, seed)
    tax = synth.taxonomy(tree)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, k, 32, keep_leaf_vecs=True)
    txn = api.Taxonomy(ctx, tax["root_rank"], tax["kind"], tax["arg"], tree.n, tax["names"])
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, k, 32))
    return tree, tax, ttc, txn, oln


def test_results_match_oracle_on_the_configs0_tree(api):
    # BASELINE.json configs[0]: 2 000-leaf ITS tree; queries = reads and one cluster centroid (run.rs:301-309)
    ctx = api.Context(0)
    tree, tax, ttc, txn, oln = build(api, ctx, 2000, 1001)
    reads = synth.make_reads(tree, 60, 2001)
    counts = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1)
    qn = counts.into_normalized_l2()
    oc = oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1)
    oqn = oracle.normalize(oc)
    res = api.results_from_compute_tree(qn, ttc, txn)
    assert len(res) == reads.n
    for i in range(reads.n):
        kq, vq = oqn.row(i)
        exp = oracle.results_from_scores(tax["root_rank"], tax["kind"], tax["arg"], oracle.score_all(oln, kq, vq))
        assert_same_tree(res[i], exp, f"read {i}")
        assert len(res[i].payload_score) <= 100 and res[i].conf_state.max() == 2
        assert_same_tree(res[i].cut(7), exp.cut(7), f"read {i} cut(7)")     # run.rs:392 nb_best_display
    # the cluster's query: centroid of reads drawn around one clade
    near = synth.make_reads(tree, 80, 2002, leaf_ids=np.arange(700, 730))
    cen = api.compute_cluster_centroid(api.MyrSeqs(ctx, near).compute_kmer_counts(18, 0.1), np.arange(near.n))
    ck, cv = oracle.centroid(oracle.count_reads(near.codes, near.qual, near.base_off, 18, 0.1), np.arange(near.n))
    s = oracle.score_all(oln, ck, cv)
    for prm in (api.SearchParameters(), api.SearchParameters(max_amount_of_leaves_per_branch=3, nb_best_analysis=10),
                api.SearchParameters(max_amount_of_leaves_per_branch=1, nb_best_analysis=2000, lambda_leaf=5.0,
                                     lambda_branch=7.5, mu=0.3, gamma=1.5, delta=0.5, epsilon=0.8)):
        got = api.results_from_compute_tree(cen, ttc, txn, api.SIM_COSINE, prm)[0]
        exp = oracle.results_from_scores(tax["root_rank"], tax["kind"], tax["arg"], s,
                                         prm.max_amount_of_leaves_per_branch, prm.nb_best_analysis, prm.lambda_leaf,
                                         prm.lambda_branch, prm.mu, prm.gamma, prm.delta, prm.epsilon)
        assert_same_tree(got, exp, str(prm))
    # the best genus holds the clade the reads came from
    best = got.branches_at_rank(api.RANKS["Genus"])
    top = best[np.argmax(got.pool_score[best])]
    leaves_below = [int(p) for p in got.payload_src_id]
    assert any(700 <= p < 730 for p in leaves_below) and got.conf_state[top] == 2


def test_results_other_similarities_ties_and_batches(api):
    ctx = api.Context(0)
    tree, tax, ttc, txn, oln = build(api, ctx, 900, 1301)
    reads = synth.make_reads(tree, 24, 1302)
    qn = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    oqn = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
    for gsim, osim in ((api.SIM_JACARD_TANIMOTO, oracle.SIM_JACARD_TANIMOTO), (api.SIM_OVERLAP, oracle.SIM_OVERLAP)):
        res = api.results_from_compute_tree(qn, ttc, txn, gsim, q_first=5, q_count=9)
        for j in range(9):
            kq, vq = oqn.row(5 + j)
            s = oracle.score_all(oln, kq, vq, sim=osim)
            assert_same_tree(res[j], oracle.results_from_scores(7, tax["kind"], tax["arg"], s), f"sim {gsim} q {j}")
    # ties: duplicated leaves (identical sequences => identical scores) inside and across genera
    dup = tree.subset(np.concatenate([np.arange(300), np.arange(300)]))
    dup.meta = synth.make_tree(600, 1303).meta  # a 600-leaf taxonomy over the 2 x 300 duplicated sequences
    tax2 = synth.taxonomy(dup)
    st2 = api.TaxTreeStore(ctx, dup)
    ttc2 = api.TaxTreeCompute.from_store_tree(st2, 18, 32)
    txn2 = api.Taxonomy(ctx, 7, tax2["kind"], tax2["arg"], dup.n)
    oln2 = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(dup), dup.base_off, 18, 32))
    res2 = api.results_from_compute_tree(qn, ttc2, txn2, api.SIM_COSINE,
                                         api.SearchParameters(max_amount_of_leaves_per_branch=2, nb_best_analysis=9))
    n_ties = 0
    for i in range(reads.n):
        kq, vq = oqn.row(i)
        s = oracle.score_all(oln2, kq, vq)
        exp = oracle.results_from_scores(7, tax2["kind"], tax2["arg"], s, 2, 9)
        assert_same_tree(res2[i], exp, f"ties q {i}")
        n_ties += int(len(set(bits(exp.payload_score).tolist())) < len(exp.payload_score))
    assert n_ties > 0
    # bad arguments are reported
    with pytest.raises(api.MyrioError):
        api.Taxonomy(ctx, 7, [0, 1, 1], [2, 0, 0], 2)          # payload id twice
    with pytest.raises(api.MyrioError):
        api.Taxonomy(ctx, 7, [0, 1], [2, 0], 2)                # the list ends inside the root
    with pytest.raises(api.MyrioError):
        api.results_from_compute_tree(qn, ttc, txn2)           # taxonomy of another tree size


def test_in_database_confidence_two_genes(api):
    # run.rs:405-466 over two gene trees queried with reads of one clade
    ctx = api.Context(0)
    per_gene, o_off, o_name, o_pool, o_state, o_conf, ids = [], [0], [], [], [], [], {}
    for g, seed in enumerate((1401, 1402)):
        tree, tax, ttc, txn, oln = build(api, ctx, 1200, seed)
        near = synth.make_reads(tree, 50, 1410 + g, leaf_ids=np.arange(100, 140))
        cen = api.compute_cluster_centroid(api.MyrSeqs(ctx, near).compute_kmer_counts(18, 0.1), np.arange(near.n))
        got = api.results_from_compute_tree(cen, ttc, txn)[0]
        per_gene.append((got, tax["names"]))
        ck, cv = oracle.centroid(oracle.count_reads(near.codes, near.qual, near.base_off, 18, 0.1), np.arange(near.n))
        exp = oracle.results_from_scores(7, tax["kind"], tax["arg"], oracle.score_all(oln, ck, cv))
        assert_same_tree(got, exp, f"gene {g}")
        depth = got.depths()
        for i in np.nonzero((exp.kind == 0) & (depth == 7 - 2))[0]:
            o_name.append(ids.setdefault(tax["names"][int(exp.src_node[i])], len(ids)))
            o_pool.append(exp.pool_score[i])
            o_state.append(exp.conf_state[i])
            o_conf.append(exp.conf[i])
        o_off.append(len(o_name))
    got = api.in_database_confidence(per_gene)
    exp = oracle.in_database_confidence(o_off, o_name, o_pool, o_state, o_conf)
    assert bits([got])[0] == bits([exp])[0] and 0.0 < got < 1.0
