"""Pins the CPU oracle against every known-answer test the reference holds for
the hot path and against the golden fixtures in tests/golden/.
CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from myrio_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = np.finfo(np.float32).eps  # assert_float_eq!: |lhs-rhs| < f32::EPSILON (myrio-core/src/lib.rs:11-17)
F32_MIN = np.float32(-3.4028234663852886e38)


def dna(s):
    return np.array(["ACGT".index(c) for c in s], np.uint8)


def count_one(seq, qual, k, cutoff):
    b = dna(seq)
    off = np.array([0, len(b)], np.uint64)
    return oracle.count_reads(b, np.array(qual, np.uint8), off, k, cutoff)


def test_phred_table_matches_reference_literals():
    # myrio-core/src/constants.rs:28 rounded to f32 (tests/golden/make_golden.py)
    g = np.array(json.load(open(os.path.join(HERE, "golden/phred_correct_f32.json")))["f32_bits"],
                 np.uint32)
    t = oracle.phred_correct_table()
    assert (t.view(np.uint32) == g).all()


def test_kmer_key_is_lsb_first():
    # upstream bio-seq: AACTT 2-mers -> [0, 4, 13, 15] (assumed encoding)
    s = dna("AACTT")
    assert [oracle.kmer_key(s[i:i + 2], 2) for i in range(4)] == [0, 4, 13, 15]


def test_simfunc_test():
    # myrio-core/src/similarity.rs:217-232
    a = count_one("ACTG", [0] * 4, 2, F32_MIN)
    b = count_one("ACTC", [0] * 4, 2, F32_MIN)
    assert int(a.aux[0]) == 6 and int(b.aux[0]) == 6
    cos = oracle.similarity(oracle.SIM_COSINE, False, a.keys, a.vals, b.keys, b.vals)
    assert abs(cos - np.float32(2.0 / 3.0)) < EPS
    ov = oracle.similarity(oracle.SIM_OVERLAP, False, a.keys, a.vals, b.keys, b.vals)
    assert abs(ov - np.float32(0.5)) < EPS
    an, bn = oracle.normalize(a), oracle.normalize(b)
    cpn = oracle.similarity(oracle.SIM_COSINE, True, an.keys, an.vals, bn.keys, bn.vals)
    assert abs(cpn - cos) < EPS


def test_from_unsorted_singles():
    # myrio-core/src/data/sparse.rs:629-632
    k, v = oracle.sort_reduce_u16([5, 5, 3, 1, 5, 3, 3, 2, 1, 2, 2])
    assert k.tolist() == [1, 2, 3, 5] and v.tolist() == [2, 3, 3, 3]
    k, v = oracle.sort_reduce_f32([5, 5, 3, 1, 5, 3, 3, 2, 1, 2, 2])
    assert k.tolist() == [1, 2, 3, 5] and v.tolist() == [2.0, 3.0, 3.0, 3.0]
    k, v = oracle.sort_reduce_f32([])
    assert len(k) == 0


def test_dot_method():
    # myrio-core/src/data/sparse.rs:644-656
    one = lambda n: np.ones(n, np.float32)
    assert oracle.dot([1, 3, 5], one(3), [1, 2, 3, 4, 5], one(5)) == 3.0
    assert oracle.dot([10], one(1), [1, 2, 3, 4, 5], one(5)) == 0.0
    assert oracle.dot([4, 8], one(2), [1, 6, 8], one(3)) == 1.0


def test_merge_and_apply_add_and_merge_keys():
    # myrio-core/src/data/sparse.rs:610-620 and :594-599
    ck, cv = oracle.add([1, 2, 3], np.array([1, 2, 3], np.float32), [2, 3, 4],
                        np.array([2, 3, 4], np.float32))
    assert ck.tolist() == [1, 2, 3, 4] and cv.tolist() == [1.0, 4.0, 6.0, 4.0]
    ck, _ = oracle.add([1, 2, 6, 7], np.ones(4, np.float32), [2, 3, 4, 9, 10, 11, 12, 13],
                       np.ones(8, np.float32))
    assert ck.tolist() == [1, 2, 3, 4, 6, 7, 9, 10, 11, 12, 13]


def test_norm_l2():
    # myrio-core/src/data/sparse.rs:636-641: norm_l2 of [1,2,5] is sqrt(30)
    c = oracle.Csr(np.array([0, 3], np.uint64), np.array([1, 2, 6], np.uint64),
                   np.array([1.0, 2.0, 5.0], np.float32))
    n = oracle.normalize(c)
    expect = np.array([1.0, 2.0, 5.0], np.float32) / np.sqrt(np.float32(30.0))
    assert (n.vals == expect).all()


def test_both_strands_and_quality_gate():
    # data/myrseq.rs:95-105: every passing window emits its key and its revcomp's key
    seq = "ACGTTGCAAC"
    b = dna(seq)
    k = 3
    good = count_one(seq, [40] * 10, k, 0.1)
    rc = (3 - b[::-1]).astype(np.uint8)
    keys = [oracle.kmer_key(b[i:i + k], k) for i in range(8)] + \
           [oracle.kmer_key(rc[i:i + k], k) for i in range(8)]
    uk, cnt = np.unique(np.array(keys, np.uint64), return_counts=True)
    assert good.keys.tolist() == uk.tolist() and good.vals.tolist() == cnt.astype(np.float32).tolist()
    assert int(good.aux[0]) == 16
    # one Q=2 base (P=0.369): windows covering it have conf 0.369*0.9999^2 > 0.1 -> kept;
    # Q=0 (P=0) kills the 3 windows covering it, on both strands
    bad = count_one(seq, [40] * 4 + [0] + [40] * 5, k, 0.1)
    assert int(bad.aux[0]) == 2 * (8 - 3)
    # too short: nothing
    short = count_one("AC", [40, 40], 3, 0.1)
    assert short.row_off[-1] == 0 and int(short.aux[0]) == 0


def test_invalid_k_and_quality():
    with pytest.raises(ValueError):
        count_one("ACGT", [1] * 4, 1, 0.1)   # myrio-proc/src/lib.rs:70 panics
    with pytest.raises(ValueError):
        count_one("ACGT", [1] * 4, 33, 0.1)
    with pytest.raises(ValueError):
        count_one("ACGT", [200] * 4, 2, 0.1)  # table index out of bounds


def test_leaf_counts_acgt():
    # tax/mod.rs:70-148 on an ACGT-only leaf: every count is 32*c, rescale = 1/32
    tree = synth.make_tree(8, 7)
    lc = oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32)
    assert (lc.vals % 32 == 0).all() and (lc.aux == np.float32(1 / 32)).all()
    for l in range(tree.n):
        b = tree.seq(l)
        rc = (3 - b[::-1]).astype(np.uint8)
        nb = len(b) - 18 + 1
        keys = [oracle.kmer_key(b[i:i + 18], 18) for i in range(nb)] + \
               [oracle.kmer_key(rc[i:i + 18], 18) for i in range(nb)]
        uk, cnt = np.unique(np.array(keys, np.uint64), return_counts=True)
        k, v = lc.row(l)
        assert k.tolist() == uk.tolist() and v.tolist() == (32 * cnt).tolist()


def test_leaf_sentinels_and_gaps():
    # :99-102 fewer than k bases after dropping gaps -> ({0:1}, 1.0)
    codes = np.array([8, 0, 4, 0, 2], np.uint8)  # A - C - G
    lc = oracle.count_leaves(codes, np.array([0, 5], np.uint64), 4, 32)
    assert lc.keys.tolist() == [0] and lc.vals.tolist() == [1] and lc.aux[0] == 1.0
    # gaps are dropped before windowing
    lc2 = oracle.count_leaves(codes, np.array([0, 5], np.uint64), 3, 32)
    acg = oracle.kmer_key(dna("ACG"), 3)
    cgt = oracle.kmer_key(dna("CGT"), 3)
    assert sorted(lc2.keys.tolist()) == sorted([acg, cgt]) and lc2.vals.tolist() == [32, 32]


def test_leaf_ambiguity_cutoff():
    # an N resolves 4 ways over 32 resamples; k-mers seen <= (32+5)>>3 = 4 times are stripped
    codes = np.array([8, 4, 15, 2, 1, 8], np.uint8)  # A C N G T A
    lc = oracle.count_leaves(codes, np.array([0, 6], np.uint64), 2, 32, seed=5)
    assert (lc.vals > 4).all()
    assert int(lc.vals.astype(np.uint32).sum()) <= 32 * 2 * 5
    tot = np.float32(2 * 5) / np.float32(int(lc.vals.astype(np.uint32).sum()))
    assert lc.aux[0] == tot


def test_topx_order_and_padding():
    s = np.array([0.5, 0.9, 0.5, 0.0, 0.9, 0.1], np.float32)
    ts, ti = oracle.topx(s, 4)
    assert ti.tolist() == [1, 4, 0, 2] and ts.tolist() == [np.float32(0.9)] * 2 + [0.5, 0.5]
    ts, ti = oracle.topx(s[:2], 4, id_base=10)
    assert ti.tolist() == [11, 10, 0xFFFFFFFF, 0xFFFFFFFF]


def test_invert_roundtrip():
    tree = synth.make_tree(40, 3)
    ln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 12, 32))
    uk, of, pl, pv = oracle.invert(ln)
    assert (np.diff(uk.astype(np.int64)) > 0).all()
    # rebuild leaf rows from postings
    order = np.lexsort((np.repeat(uk, np.diff(of.astype(np.int64))), pl))
    assert (np.repeat(uk, np.diff(of.astype(np.int64)))[order] == ln.keys).all()
    assert (pv[order] == ln.vals).all()


def test_score_and_cluster_recover_structure():
    # not a reference KAT: sanity of the restated pipeline on synthetic data
    t1 = synth.make_tree(60, 11)
    t2 = synth.make_tree(60, 12)
    r1 = synth.make_reads(t1, 40, 21)
    r2 = synth.make_reads(t2, 40, 22)
    ln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(t1), t1.base_off, 18, 32))
    rc = oracle.normalize(oracle.count_reads(r1.codes, r1.qual, r1.base_off, 18, 0.1))
    hits = 0
    for i in range(10):
        k, v = rc.row(i)
        s = oracle.score_all(ln, k, v)
        ts, ti = oracle.topx(s, 5)
        hits += int(r1.meta["src_leaf"][i] in ti.tolist())
    assert hits >= 7
    # clustering of a two-gene mix at k=6
    codes = np.concatenate([r1.codes, r2.codes])
    qual = np.concatenate([r1.qual, r2.qual])
    off = np.concatenate([r1.base_off, r2.base_off[1:] + r1.base_off[-1]])
    cc = oracle.count_reads(codes, qual, off, 6, 0.2)
    assign, iters, sil = oracle.cluster(cc, 20, None, 2, silhouette=1.4)
    a, b = assign[:40], assign[40:]
    kept_a, kept_b = a[a >= 0], b[b >= 0]
    assert len(kept_a) > 20 and len(kept_b) > 20
    assert (kept_a == kept_a[0]).mean() > 0.9 and (kept_b == kept_b[0]).mean() > 0.9
    assert kept_a[0] != kept_b[0]
    assert 1 <= iters <= 20


def test_fastq_preprocess_restatement():
    # io/mod.rs:46-48 + data/myrseq.rs:113-131
    recs = [("ACGT" * 20, "I" * 80),            # Q40, len 80 -> kept
            ("ACGT" * 5, "I" * 20),             # too short (min_length 50)
            ("ACGT" * 20, "+" * 80),            # Q10 < 12 -> dropped
            ("", ""),                           # empty -> dropped
            ("TTGCA" * 12, "5" * 59 + " "),     # one byte below '!' saturates to Q0; mean still >= 12
            ("ACGT" * 20, "I" * 79 + "~")]      # Q93 max > max_qual 60
    seq = np.frombuffer("".join(r[0] for r in recs).encode(), np.uint8)
    qual = np.frombuffer("".join(r[1] for r in recs).encode(), np.uint8)
    off = np.zeros(len(recs) + 1, np.uint64)
    off[1:] = np.cumsum([len(r[0]) for r in recs])
    codes, q, koff, keep = oracle.fastq_preprocess(seq, qual, off, 50, 12.0, 60)
    assert keep.tolist() == [1, 0, 0, 0, 1, 0]
    assert koff.tolist() == [0, 80, 140]
    assert codes[:4].tolist() == [0, 1, 2, 3] and q[:80].tolist() == [40] * 80
    assert q[139] == 0 and q[138] == ord("5") - 33
    with pytest.raises(ValueError):
        oracle.fastq_preprocess(np.frombuffer(b"ACGN", np.uint8), np.frombuffer(b"IIII", np.uint8),
                                np.array([0, 4], np.uint64), 1, 0.0, 255)


def test_kmer_key_bit_orders():
    """The key packing is an assumption about the un-vendored bio-seq fork: both orders are
    implemented; MSB-first is LSB-first with its 2-bit groups reversed, and counts do not depend on it."""
    bases = np.array([0, 0, 1, 3, 3], np.uint8)  # AACTT
    try:
        assert [oracle.kmer_key(bases[i:i + 2], 2) for i in range(4)] == [0, 4, 13, 15]   # upstream bio-seq's own example
        oracle.set_kmer_msb_first(True)
        assert [oracle.kmer_key(bases[i:i + 2], 2) for i in range(4)] == [0, 1, 7, 15]
        rng = np.random.default_rng(5)
        seq = rng.integers(0, 4, 300).astype(np.uint8)
        qual = np.full(300, 30, np.uint8)
        off = np.array([0, 300], np.uint64)
        for k in (5, 18, 32):
            oracle.set_kmer_msb_first(False)
            a = oracle.count_reads(seq, qual, off, k, 0.1)
            oracle.set_kmer_msb_first(True)
            b = oracle.count_reads(seq, qual, off, k, 0.1)

            def rev_groups(key):
                out = 0
                for j in range(k):
                    out |= ((int(key) >> (2 * j)) & 3) << (2 * (k - 1 - j))
                return out
            ma = {rev_groups(x): float(v) for x, v in zip(a.keys, a.vals)}
            mb = {int(x): float(v) for x, v in zip(b.keys, b.vals)}
            assert ma == mb and (np.diff(b.keys.astype(object)) > 0).all()
    finally:
        oracle.set_kmer_msb_first(False)


def test_from_unsorted_pairs_kat():
    # myrio-core/src/data/sparse.rs:623-633, second half: from_unsorted_pairs
    k, v = oracle.sort_reduce_pairs_f32([(10, 1.0), (2, 4.0), (1, 1.0), (10, 7.0), (10, 2.0), (3, 9.0)])
    assert k.tolist() == [1, 2, 3, 10] and v.tolist() == [1.0, 4.0, 9.0, 10.0]
    k, v = oracle.sort_reduce_pairs_f32([])
    assert len(k) == 0 and len(v) == 0


def test_fasta_to_iupac_restatement():
    # tax/store.rs:178-188: bio-seq Iupac codes (A=8 C=4 G=2 T=1, ambiguity = OR, '-' = X = 0)
    recs = ["ACGT", "RYSWKMBDHVN-", "ACGU", "", "acgt", "NNAC"]
    seq = np.frombuffer("".join(recs).encode(), np.uint8)
    off = np.zeros(len(recs) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in recs])
    codes, koff, keep = oracle.fasta_to_iupac(seq, off)
    assert keep.tolist() == [True, True, False, True, False, True]
    assert koff.tolist() == [0, 4, 16, 16, 20]
    assert codes[:4].tolist() == [8, 4, 2, 1]
    assert codes[4:16].tolist() == [8 | 2, 4 | 1, 4 | 2, 8 | 1, 2 | 1, 8 | 4, 7, 11, 13, 14, 15, 0]
    assert codes[16:].tolist() == [15, 15, 8, 4]
    assert (synth.iupac_to_ascii(codes) == np.frombuffer(b"ACGTRYSWKMBDHVN-NNAC", np.uint8)).all()


def test_rust_sys_crate_declares_every_header_symbol():
    # the -sys crate cannot be compiled here (no cargo/rustc): at least keep its extern block complete
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "myrio_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(myrio_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(root, "rust", "myrio-b200-sys", "src", "lib.rs")).read()
    in_rust = set(re.findall(r"pub fn (myrio_[a-z0-9_]+)\s*\(", rs))
    assert declared - in_rust == set(), sorted(declared - in_rust)
