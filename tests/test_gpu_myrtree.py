"""GPU half of the `.myrtree` path: `myrio tree new -k 18` (K2 on the device, counts written back into the store,
file saved) followed by `myrio run` loading the file (cached counts -> leaf vectors -> index -> scores), through
the C ABI, against the oracle.  tax/store.rs:53-93, 265-308; tax/compute.rs:80-98."""
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


def test_tree_new_cache_counts_then_run_from_file(api, tmp_path):
    ctx = api.Context(0)
    rng = np.random.default_rng(3)
    tree = synth.make_tree(700, 1201)
    iupac = synth.tree_to_iupac(tree).copy()
    pos = rng.choice(len(iupac), 200, replace=False)
    iupac[pos] = rng.choice(np.array([0, 3, 5, 10, 15], np.uint8), 200)  # gaps + ambiguity codes in some leaves
    tax = synth.taxonomy(tree)
    f = api.MyrTreeFile.new("ITS", tax["root_rank"], tax["kind"], tax["arg"], tax["names"], iupac, tree.base_off)

    # `myrio tree new -k 6 -k 18`: compute_and_append_kmer_counts per k, then save_to_file (tree.rs:142-149)
    store = api.TaxTreeStore.from_file(ctx, f)
    assert store.n_leaves == tree.n and store.k_precomputed == []
    for k in (6, 18):
        store.compute_and_append_kmer_counts(k, 32, seed=11)
    path = tmp_path / "its.myrtree"
    store.save_to_file(str(path), 6)

    # the file holds exactly the oracle's counts, in payload order, for both k
    o = oracle.store_decode(open(path, "rb").read())
    assert o.k_precomputed == [6, 18] and (o.seq_codes == iupac).all() and o.names == tax["names"]
    for i, k in enumerate((6, 18)):
        e = oracle.count_leaves(iupac, tree.base_off, k, 32, seed=11)
        c = o.counts[i]
        assert (c.row_off == e.row_off).all() and (c.keys == e.keys).all() and (c.vals == e.vals).all()
        assert (bits(c.aux) == bits(e.aux)).all()

    # `myrio run`: load_from_file, cached counts picked for search_k (compute.rs:80-85: no recount), scores + cut(50)
    ctx2 = api.Context(0)
    loaded = api.TaxTreeStore.load_from_file(ctx2, str(path))
    assert loaded.k_precomputed == [6, 18] and loaded.gene == "ITS"
    launches = ctx2.launch_count
    ttc = api.TaxTreeCompute.from_store_tree(loaded, 18, 32)
    assert len(loaded.k_precomputed) == 2  # nothing appended: the cache was hit
    reads = synth.make_reads(tree, 40, 1202)
    q = api.MyrSeqs(ctx2, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 50, full_scores=True)
    assert ctx2.launch_count > launches
    oln = oracle.normalize(oracle.count_leaves(iupac, tree.base_off, 18, 32, seed=11))
    oqn = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
    for i in range(reads.n):
        kq, vq = oqn.row(i)
        s = oracle.score_all(oln, kq, vq)
        assert (bits(s) == bits(res.scores[i])).all()
        ts, ti = oracle.topx(s, 50)
        assert (ti == res.top_ids[i]).all() and (bits(ts) == bits(res.top_scores[i])).all()

    # a k that is not cached: --cache-counts appends it and the file grows by it (compute.rs:86-91)
    api.TaxTreeCompute.from_store_tree(loaded, 12, 32, seed=11)
    assert loaded.k_precomputed == [6, 18, 12]
    loaded.save_to_file(str(path), 3)
    again = api.MyrTreeFile.load_from_file(str(path))
    assert again.k_precomputed == [6, 18, 12]
    e12 = oracle.count_leaves(iupac, tree.base_off, 12, 32, seed=11)
    ro, keys, vals, resc = again.counts_host(2)
    assert (ro == e12.row_off).all() and (keys == e12.keys).all() and (vals == e12.vals).all()
    # overwrite (no --cache-counts, compute.rs:92-96) leaves one entry; shrink clears the cache (tree.rs:190-205)
    loaded.compute_and_overwrite_kmer_counts(18, 32, seed=11)
    assert loaded.file.k_precomputed == [18]
    loaded.file.shrink()
    assert loaded.file.k_precomputed == []
