"""Pins the oracle's restatement of the post-scoring pass (oracle/results_oracle.c: tax/results.rs:96-381,
run.rs:405-466) against a second, independent statement written here step by step in numpy float32 on a tree
small enough to follow by hand.  The reference has no test for this code, so this is what stands
between the oracle and a typo.  CPU only.  expf / powf come from libm, as Rust's f32::exp / f32::powf do."""
import ctypes

import numpy as np

import oracle

f32 = np.float32
_libm = ctypes.CDLL("libm.so.6")
_libm.expf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def expf(x):
    return f32(_libm.expf(ctypes.c_float(float(x))))


def powf(x, y):
    return f32(_libm.powf(ctypes.c_float(float(x)), ctypes.c_float(float(y))))


def softmax_pooling(scores, lam):  # results.rs:160-173
    num, den = f32(0), f32(0)
    for s in scores:
        s = f32(s)
        e = expf(f32(lam) * s)
        num = f32(num + f32(s * e))
        den = f32(den + e)
    return f32(num / den)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# root (Family) -> G1 {0, 1, 2}, G2 {3, 4}, and one leaf (5) directly under the family
KIND = [0, 0, 1, 1, 1, 0, 1, 1, 1]
ARG = [3, 3, 0, 1, 2, 2, 3, 4, 5]
SCORES = np.array([0.5, 0.9, 0.7, 0.2, 0.9, 0.1], np.float32)


def test_small_tree_step_by_step():
    r = oracle.results_from_scores(3, KIND, ARG, SCORES, max_amount_of_leaves_per_branch=2, nb_best_analysis=3)
    # cut(3) of the trimmed leaves [G1: .9 .7 | G2: .9 .2 | root: .1] -> {.9, .9, .7}: the root's own leaf and G2's .2 go
    assert r.kind.tolist() == [0, 0, 1, 1, 0, 1]
    assert r.arg.tolist() == [2, 2, 0, 1, 1, 2]
    assert r.src_node.tolist() == [0, 1, 3, 4, 5, 7]          # names: G1's leaves 1 and 2 (best first), G2's leaf 4
    assert r.payload_src_id.tolist() == [1, 2, 4]
    assert (bits(r.payload_score) == bits([0.9, 0.7, 0.9])).all()
    # branch statistics over ALL leaves, nested sums (results.rs:96-143)
    g1 = f32(f32(f32(f32(0) + SCORES[0]) + SCORES[1]) + SCORES[2])
    g2 = f32(f32(f32(0) + SCORES[3]) + SCORES[4])
    root = f32(f32(f32(f32(0) + g1) + g2) + SCORES[5])
    assert r.nb_leaves.tolist() == [6, 3, 0, 0, 2, 0]
    assert (bits(r.mean[[0, 1, 4]]) == bits([root / f32(6), g1 / f32(3), g2 / f32(2)])).all()
    assert (bits(r.max[[0, 1, 4]]) == bits([0.9, 0.9, 0.9])).all()
    # pooling (results.rs:175-205): genus rank first, then the family
    p1 = softmax_pooling([softmax_pooling([0.9, 0.7], 2.0) * f32(f32(2) / f32(f32(2) + f32(1.1)))], 20.0)
    p2 = softmax_pooling([softmax_pooling([0.9], 2.0) * f32(f32(1) / f32(f32(1) + f32(1.1)))], 20.0)
    pr = softmax_pooling([p1, p2], 20.0)
    assert (bits(r.pool_score[[0, 1, 4]]) == bits([pr, p1, p2])).all()
    # confidence at the genus rank (results.rs:207-233); the family is the root rank: none
    n = f32(2)
    mean = f32(f32(f32(0) + p1) + p2) / n
    sq = f32(f32(f32(0) + f32(f32(p1 - mean) * f32(p1 - mean))) + f32(f32(p2 - mean) * f32(p2 - mean)))
    sd = max(f32(np.sqrt(f32(f32(f32(1) / n) * sq))), f32(1e-3))
    x = f32(f32(p1 - p2) / sd)
    conf = f32(f32(powf(p1, 0.55) * powf(x, 2.5)) / f32(f32(0.9) + powf(x, 2.5)))
    assert p1 > p2
    assert r.conf_state.tolist() == [0, 2, 0, 0, 0, 0]
    assert bits(r.conf[1:2])[0] == bits([conf])[0]
    assert r.force_keep.tolist() == [0, 1, 0, 0, int(p2 > f32(0.2) and conf < f32(0.3)), 0]


def test_cut_honours_force_keep_and_ties_keep_order():
    r = oracle.results_from_scores(3, KIND, ARG, SCORES, max_amount_of_leaves_per_branch=2, nb_best_analysis=3)
    # run.rs:392 cut(nb_best_display = 1): only the best leaf, but the force_keep genus stays (with no leaf if need be)
    c = r.cut(1)
    assert c.payload_src_id.tolist() == [1]                   # tie .9 / .9: the earlier leaf in DFS order wins
    assert c.kind.tolist() == [0, 0, 1] and c.arg.tolist() == [1, 1, 0]
    assert (bits(c.pool_score) == bits(r.pool_score[[0, 1, 2]])).all()  # BRes cloned
    # all scores equal: trim keeps the first `max` leaves of a branch in child order, cut the first in DFS order
    flat = oracle.results_from_scores(3, KIND, ARG, np.full(6, 0.25, np.float32), 2, 3)
    assert flat.payload_src_id.tolist() == [0, 1, 3]


def test_single_genus_and_species_root():
    # one genus under the family: confidence Some(None) at the genus rank (results.rs:209-214)
    r = oracle.results_from_scores(3, [0, 0, 1, 1], [1, 2, 0, 1], np.array([0.3, 0.6], np.float32))
    assert r.conf_state.tolist() == [0, 1, 0, 0] and r.force_keep.tolist() == [0, 0, 0, 0]
    assert r.payload_src_id.tolist() == [1, 0]                # leaves re-ordered best first
    # root rank Species: Rank::collect_range_inclusive(Genus, Species) is empty -> no pooling at all
    s = oracle.results_from_scores(1, [0, 1, 1], [2, 0, 1], np.array([0.3, 0.6], np.float32))
    assert s.pool_score.tolist() == [0.0, 0.0, 0.0] and s.conf_state.tolist() == [0, 0, 0]


def test_in_database_confidence_step_by_step():
    # two genes; gene 0 ranks genus "A" first (conf .8), gene 1 ranks "B" first and holds "A" third
    off = [0, 3, 7]
    name = [10, 11, 12, 11, 13, 10, 14]        # A B C | B D A E
    pool = [0.7, 0.5, 0.1, 0.6, 0.4, 0.3, 0.2]
    state = [2, 0, 0, 2, 0, 0, 0]
    conf = [0.8, 0, 0, 0.4, 0, 0, 0]
    got = oracle.in_database_confidence(off, name, pool, state, conf)
    take = f32(8)
    mismatch = f32(f32(0) + f32(f32(f32(2) / take) * max(powf(0.8, 0.7), f32(0.12))))   # "A" is third in gene 1
    mismatch = f32(mismatch / f32(2))                                                  # gene 1 has no later gene
    csum = f32(f32(f32(0) + f32(0.8)) + f32(0.4))
    lam = f32(f32(0.5) + f32(f32(2) * csum))
    conf_score = f32(f32(1) - softmax_pooling([0.8, 0.4], lam))
    y = powf(conf_score, 0.8)
    unknown = f32(f32(1) / f32(f32(1) + expf(f32(f32(-8) * f32(f32(f32(mismatch + y) / f32(2)) - f32(0.45))))))
    assert bits([got])[0] == bits([f32(f32(1) - unknown)])[0]
