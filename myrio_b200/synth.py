"""Deterministic synthetic barcode-shaped data.

Trees: ranks p>c>o>f>g>s>leaf under one root; sequences evolve down the tree by
substitutions (+ indels on the upper ranks) with per-edge divergence
{p 12 %, c 8 %, o 6 %, f 4 %, g 2.5 %, s 1 %, leaf 0.3 %} and one conserved
block (160 bp, 1/10 rate; mimics 5.8S inside ITS and gives long posting lists).
Alphabet ACGT only (the reference's IUPAC resampling is OS-seeded).

Reads: recipe of the reference's simulator (myrio-exp/src/simseq/mod.rs:150-255):
window of ~0.6 n +- 0.15 n (min 150) on a random strand, block-wise Q scores
from the observed CDFs, per-base errors with p = 10^(-Q/10) split
deletion/insertion/SNP = 0.3/0.3/0.4, then MyrSeq::pre_process
(len >= 50, mean Q >= 12; myrio-core/src/data/myrseq.rs:113-131).

Pure numpy; no GPU, no oracle.  Everything is a function of the seeds.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Observed cumulative frequencies of Q (0..40) and of Q-block sizes (1..14) in
# nanopore barcode runs -- the empirical numbers of
# myrio-exp/src/simseq/constants.rs:3-27 (data, rounded to 4 decimals here).
_Q_CDF = np.array([
    0.0, 0.0003, 0.0057, 0.0244, 0.0569, 0.0983, 0.1430, 0.1856, 0.2248, 0.2612, 0.2956, 0.3286,
    0.3606, 0.3922, 0.4235, 0.4549, 0.4867, 0.5189, 0.5516, 0.5849, 0.6190, 0.6538, 0.6892,
    0.7252, 0.7616, 0.7978, 0.8332, 0.8664, 0.8963, 0.9221, 0.9431, 0.9597, 0.9723, 0.9815,
    0.9880, 0.9925, 0.9955, 0.9975, 0.9987, 0.9995, 1.0])
_QBLOCK_CDF = np.array([0.0, 0.3446, 0.5788, 0.7326, 0.8400, 0.8991, 0.9354, 0.9584, 0.9732,
                        0.9828, 0.9891, 0.9935, 0.9965, 0.9986, 1.0])
_QINNER_CDF = np.array([0.05, 0.15, 0.30, 0.70, 0.85, 0.95, 1.0])  # simseq/mod.rs:44-45, x=0 at 3

DIVERGENCE = {"p": 0.12, "c": 0.08, "o": 0.06, "f": 0.04, "g": 0.025, "s": 0.01, "leaf": 0.003}
RANKS = ["p", "c", "o", "f", "g", "s", "leaf"]


@dataclass
class SeqSet:
    """Ragged batch of sequences, one code per byte (2-bit Dna codes A0 C1 G2 T3)."""
    codes: np.ndarray      # uint8 [total]
    base_off: np.ndarray   # uint64 [n+1]
    qual: np.ndarray | None = None        # uint8 [total] Phred (reads only)
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return len(self.base_off) - 1

    def lengths(self) -> np.ndarray:
        return np.diff(self.base_off.astype(np.int64))

    def seq(self, i: int) -> np.ndarray:
        return self.codes[int(self.base_off[i]):int(self.base_off[i + 1])]

    def subset(self, idx) -> "SeqSet":
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths()[idx]
        off = np.zeros(len(idx) + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        pos = _ragged_positions(self.base_off[:-1].astype(np.int64)[idx], lens)
        return SeqSet(self.codes[pos], off, None if self.qual is None else self.qual[pos],
                      dict(self.meta))


def _ragged_positions(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Flat source positions of the concatenation of [starts[i], starts[i]+lens[i])."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    out_off = np.zeros(len(lens) + 1, np.int64)
    out_off[1:] = np.cumsum(lens)
    row = np.repeat(np.arange(len(lens)), lens)
    return starts[row] + (np.arange(total, dtype=np.int64) - out_off[row])


def _level_sizes(n_leaves: int) -> list[int]:
    r = max(n_leaves, 1) ** (1.0 / 7.0)
    sizes = [max(1, int(round(r ** (i + 1)))) for i in range(6)]
    sizes = [min(s, n_leaves) for s in sizes]
    for i in range(1, 6):
        sizes[i] = max(sizes[i], sizes[i - 1])
    return sizes + [n_leaves]


def make_tree(n_leaves: int, seed: int, *, root_seed: int | None = None, shard: int = 0,
              mean_len: int = 650, sd_len: int = 40, min_len: int = 450, max_len: int = 900,
              conserved_len: int = 160, conserved_rate: float = 0.1,
              indel_fraction: float = 0.10) -> SeqSet:
    """One reference tree ("gene"); leaves in payload_id order (= FASTA order,
    myrio-core/src/tax/store.rs:190-192).  `root_seed` fixes the ancestral
    sequence so that several shards (multi-GPU weak scaling) are subtrees of one
    tree; `seed`/`shard` drive everything below the root."""
    root_rng = np.random.default_rng([root_seed if root_seed is not None else seed, 0xA11CE])
    rng = np.random.default_rng([seed, shard, 0xBEEF])
    W = max_len + 64
    root_len = min(W - 32, mean_len + 3 * sd_len)
    seqs = np.zeros((1, W), np.uint8)
    seqs[0, :root_len] = root_rng.integers(0, 4, root_len, dtype=np.uint8)
    cons = np.zeros((1, W), bool)
    c0 = int(root_len * 0.42)
    cons[0, c0:c0 + conserved_len] = True
    lens = np.array([root_len], np.int64)

    sizes = _level_sizes(n_leaves)
    parents = []
    for li, rank in enumerate(RANKS):
        n_child = sizes[li]
        n_par = seqs.shape[0]
        par = (np.arange(n_child, dtype=np.int64) * n_par) // n_child
        parents.append(par)
        div = DIVERGENCE[rank]
        new_seqs = np.empty((n_child, W), np.uint8)
        new_cons = np.empty((n_child, W), bool)
        new_lens = lens[par].copy()
        chunk = max(1, (1 << 24) // W)
        for b in range(0, n_child, chunk):
            e = min(n_child, b + chunk)
            s = seqs[par[b:e]].copy()
            c = cons[par[b:e]]
            rate = np.where(c, np.float32(div * conserved_rate), np.float32(div))
            mut = rng.random(s.shape, dtype=np.float32) < rate
            shift = rng.integers(1, 4, s.shape, dtype=np.uint8)
            s = np.where(mut, (s + shift) & 3, s).astype(np.uint8)
            new_seqs[b:e] = s
            new_cons[b:e] = c
        if n_child <= 4096 and indel_fraction > 0:
            # indels on the upper ranks only (python loop over <= 4096 nodes)
            for j in range(n_child):
                L = int(new_lens[j])
                n_ev = rng.poisson(div * indel_fraction * L)
                for _ in range(int(n_ev)):
                    pos = int(rng.integers(0, L))
                    ln = int(rng.integers(1, 4))
                    if new_cons[j, pos]:
                        continue  # the conserved block stays in frame
                    if rng.random() < 0.5 and L - ln > min_len:
                        new_seqs[j, pos:L - ln] = new_seqs[j, pos + ln:L]
                        new_cons[j, pos:L - ln] = new_cons[j, pos + ln:L]
                        L -= ln
                    elif L + ln < W - 8:
                        new_seqs[j, pos + ln:L + ln] = new_seqs[j, pos:L].copy()
                        new_cons[j, pos + ln:L + ln] = new_cons[j, pos:L].copy()
                        new_seqs[j, pos:pos + ln] = rng.integers(0, 4, ln, dtype=np.uint8)
                        new_cons[j, pos:pos + ln] = False
                        L += ln
                new_lens[j] = L
        seqs, cons, lens = new_seqs, new_cons, new_lens

    # per-leaf amplicon length ~ N(mean, sd) clipped, trimmed from both ends
    want = np.clip(np.rint(rng.normal(mean_len, sd_len, n_leaves)), min_len, max_len).astype(np.int64)
    want = np.minimum(want, lens)
    start = (rng.random(n_leaves) * (lens - want + 1)).astype(np.int64)
    base_off = np.zeros(n_leaves + 1, np.uint64)
    base_off[1:] = np.cumsum(want)
    pos = _ragged_positions(start + np.arange(n_leaves, dtype=np.int64) * W, want)
    codes = seqs.reshape(-1)[pos]
    meta = {"kind": "tree", "seed": seed, "shard": shard, "level_sizes": sizes,
            "parents": parents}
    return SeqSet(codes, base_off, None, meta)


def _sample_cdf(rng, cdf: np.ndarray, n: int) -> np.ndarray:
    """Index i such that cdf[i-1] <= u < cdf[i] (cdf[0] may be 0 -> never chosen)."""
    return np.searchsorted(cdf, rng.random(n), side="right").astype(np.int64)


def make_reads(tree: SeqSet, n_reads: int, seed: int, *, q_shift: int = 0,
               min_window: int = 150, min_length: int = 50, min_mean_qual: float = 12.0,
               leaf_ids: np.ndarray | None = None, window_frac: float = 0.6, window_sd: float = 0.15,
               max_window: int | None = None) -> SeqSet:
    """Simulated reads of random leaves of `tree` (see module docstring).  Returns
    exactly n_reads reads that pass pre_process; meta['src_leaf'] holds the source."""
    rng = np.random.default_rng([seed, 0xF00D])
    out_codes, out_qual, out_len, out_src = [], [], [], []
    need = n_reads
    fwd_ratio = rng.uniform(0.35, 0.65)
    tl = tree.lengths()
    while need > 0:
        m = int(need * 1.15) + 8
        src = rng.integers(0, tree.n, m) if leaf_ids is None else rng.choice(leaf_ids, m)
        n = tl[src]
        w = np.rint(rng.normal(window_frac * n, window_sd * n)).astype(np.int64)
        w = np.clip(w, np.minimum(min_window, n), n if max_window is None else np.minimum(n, max_window))
        st = (rng.random(m) * (n - w + 1)).astype(np.int64)
        fwd = rng.random(m) < fwd_ratio
        # window bases in read orientation
        woff = np.zeros(m + 1, np.int64)
        woff[1:] = np.cumsum(w)
        T = int(woff[-1])
        row = np.repeat(np.arange(m), w)
        i_in = np.arange(T, dtype=np.int64) - woff[row]
        src_start = tree.base_off[:-1].astype(np.int64)[src] + st
        pos = np.where(fwd[row], src_start[row] + i_in, src_start[row] + (w[row] - 1 - i_in))
        b = tree.codes[pos]
        b = np.where(fwd[row], b, 3 - b).astype(np.uint8)
        # block-wise Q (flat stream of blocks over the concatenated windows)
        nblk = T // 2 + 16
        bs = _sample_cdf(rng, _QBLOCK_CDF, nblk)
        while bs.sum() < T:
            bs = np.concatenate([bs, _sample_cdf(rng, _QBLOCK_CDF, nblk)])
        central = _sample_cdf(rng, _Q_CDF, len(bs))
        blk = np.repeat(np.arange(len(bs)), bs)[:T]
        inner = np.searchsorted(_QINNER_CDF, rng.random(T), side="right").astype(np.int64)
        q = np.clip(central[blk] + inner - 3 - q_shift, 1, 40).astype(np.int64)
        # per-base errors
        p_err = np.power(10.0, -q / 10.0)
        err = rng.random(T) < p_err
        et = rng.random(T)
        is_del = err & (et < 0.3)
        is_ins = err & (et >= 0.3) & (et < 0.6)
        is_snp = err & (et >= 0.6)
        snp_shift = rng.integers(1, 4, T, dtype=np.uint8)
        b2 = np.where(is_snp, (b + snp_shift) & 3, b).astype(np.uint8)
        q2 = np.where(is_snp, np.maximum(q - 2, 0), q)
        cnt = np.where(is_del, 0, np.where(is_ins, 2, 1)).astype(np.int64)
        ob = np.repeat(b2, cnt)
        oq = np.repeat(q2, cnt).astype(np.uint8)
        # first element of each inserted pair is a random base
        out_pos = np.cumsum(cnt) - cnt
        ins_first = out_pos[is_ins]
        ob[ins_first] = rng.integers(0, 4, len(ins_first), dtype=np.uint8)
        new_len = np.bincount(row, weights=cnt, minlength=m).astype(np.int64)
        noff = np.zeros(m + 1, np.int64)
        noff[1:] = np.cumsum(new_len)
        qsum = np.bincount(np.repeat(np.arange(m), new_len), weights=oq, minlength=m)
        ok = (new_len >= min_length) & (qsum / np.maximum(new_len, 1) >= min_mean_qual)
        keep = np.nonzero(ok)[0][:need]
        p2 = _ragged_positions(noff[:-1][keep], new_len[keep])
        out_codes.append(ob[p2])
        out_qual.append(oq[p2])
        out_len.append(new_len[keep])
        out_src.append(src[keep])
        need -= len(keep)
    lens = np.concatenate(out_len)
    base_off = np.zeros(len(lens) + 1, np.uint64)
    base_off[1:] = np.cumsum(lens)
    return SeqSet(np.concatenate(out_codes), base_off, np.concatenate(out_qual),
                  {"kind": "reads", "seed": seed, "src_leaf": np.concatenate(out_src)})


def taxonomy(tree: SeqSet) -> dict:
    """The taxonomy of a make_tree() tree as the reference builds it from FASTA headers
    (tax/store.rs:127-247): root rank Kingdom (`k:Plantae`), branches p > c > o > f > g, and the sequences as
    LEAVES named by their species (several leaves may share a species name; a leaf's payload_id is its FASTA
    order = its index in `tree`).  Returned as the flat PRE-ORDER node list of the C ABI (myrio_store_new):
    kind (0 branch, 1 leaf), arg (number of children | payload_id), names; plus rank (Rank discriminant of every
    node: branches 7..2, leaves 1) and root_rank."""
    parents = tree.meta["parents"]  # per level p, c, o, f, g, s, leaf: index of the parent one level up
    letters = ["p", "c", "o", "f", "g", "s"]
    n_levels = len(parents)
    # children ranges: parents are monotone, so the children of node j at level l are contiguous at level l + 1
    starts = []
    for li in range(1, n_levels):
        par = parents[li]
        n_par = len(parents[li - 1])
        st = np.searchsorted(par, np.arange(n_par + 1), side="left")
        starts.append(st)
    kind, arg, names, rank = [0], [len(parents[0])], ["k:Plantae"], [7]
    # iterative DFS over (level, index); species nodes (level 5) are not branches: their leaves hang under the genus
    def leaves_of_genus(g):
        out = []
        s0, s1 = starts[4][g], starts[4][g + 1]          # species of genus g
        for sp in range(s0, s1):
            l0, l1 = starts[5][sp], starts[5][sp + 1]    # leaves of species sp
            out.extend((int(leaf), f"s:Species{sp}") for leaf in range(l0, l1))
        return out
    stack = [(0, j) for j in range(len(parents[0]) - 1, -1, -1)]
    while stack:
        li, j = stack.pop()
        if li == 4:  # genus: children are leaves
            lv = leaves_of_genus(j)
            kind.append(0)
            arg.append(len(lv))
            names.append(f"g:Genus{j}")
            rank.append(2)
            for pid, nm in lv:
                kind.append(1)
                arg.append(pid)
                names.append(nm)
                rank.append(1)
            continue
        c0, c1 = int(starts[li][j]), int(starts[li][j + 1])
        kind.append(0)
        arg.append(c1 - c0)
        names.append(f"{letters[li]}:{letters[li].upper()}{j}")
        rank.append(6 - li)
        for c in range(c1 - 1, c0 - 1, -1):
            stack.append((li + 1, c))
    return {"kind": np.array(kind, np.uint8), "arg": np.array(arg, np.uint64), "names": names,
            "rank": np.array(rank, np.uint8), "root_rank": 7}


def concat(parts: list) -> SeqSet:
    """Concatenation of ragged batches (reads of several genes in one FASTQ)."""
    lens = np.concatenate([p.lengths() for p in parts])
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    qual = None if parts[0].qual is None else np.concatenate([p.qual for p in parts])
    return SeqSet(np.concatenate([p.codes for p in parts]), off, qual, {"kind": parts[0].meta.get("kind")})


# BASELINE.json configs[2]: four gene trees: name, mean amplicon length, tree seed
FOUR_GENES = [("matK", 850, 1003), ("rbcL", 1400, 1004), ("ITS", 650, 1005), ("trnH-psbA", 500, 1006)]


def make_gene_tree(n_leaves: int, mean_len: int, seed: int) -> SeqSet:
    return make_tree(n_leaves, seed, mean_len=mean_len, sd_len=int(0.06 * mean_len), min_len=int(0.7 * mean_len),
                     max_len=int(1.35 * mean_len))


def make_mixed_reads(trees: list, n_reads: int, seed: int, *, q_shift: int = -3, shuffle: bool = True) -> SeqSet:
    """configs[2] reads: an equal mix over the gene trees, near-full-length amplicon reads (window ~0.95 n,
    rbcL windowed to ~1 kb), Q shifted UP by 3 (`q_shift` = -3; the observed nanopore Q distribution alone
    gives 9.6 %) so that the realised per-base error is ~5 % (4.7 %)
    (meta['mean_phred_error']); meta['gene'] holds the tree of origin."""
    per = n_reads // len(trees)
    parts = [make_reads(t, per + (n_reads - per * len(trees) if i == 0 else 0), seed + i, q_shift=q_shift,
                        window_frac=0.95, window_sd=0.04, max_window=1050) for i, t in enumerate(trees)]
    gene = np.concatenate([np.full(p.n, i, np.int64) for i, p in enumerate(parts)])
    out = concat(parts)
    if shuffle:
        order = np.random.default_rng([seed, 0x5EED]).permutation(out.n)
        out, gene = out.subset(order), gene[order]
    out.meta = {"kind": "reads", "seed": seed, "gene": gene,
                "mean_phred_error": float(np.mean(np.power(10.0, -out.qual.astype(np.float64) / 10.0)))}
    return out


# ------------------------------------------------------------------ packing

def pack_dna2(s: SeqSet):
    """bio-seq Seq<Dna> layout (BitVec<usize, Lsb0>): base i at bits 2*(i%32) of
    word i/32.  Each sequence starts on a 16-byte boundary (even word index).
    Returns (words uint64, word_off uint64[n+1])."""
    lens = s.lengths()
    nw = ((lens + 31) // 32 + 1) // 2 * 2
    word_off = np.zeros(s.n + 1, np.uint64)
    word_off[1:] = np.cumsum(nw)
    total_w = int(word_off[-1])
    words = np.zeros(total_w, np.uint64)
    if total_w == 0:
        return words, word_off
    row = np.repeat(np.arange(s.n), lens)
    i_in = np.arange(len(s.codes), dtype=np.int64) - s.base_off[:-1].astype(np.int64)[row]
    padded = np.zeros(total_w * 32, np.uint8)
    padded[word_off[:-1].astype(np.int64)[row] * 32 + i_in] = s.codes & 3
    shifts = (2 * np.arange(32, dtype=np.uint64))
    step = 1 << 20
    for b in range(0, total_w, step):
        e = min(total_w, b + step)
        blk = padded[b * 32:e * 32].reshape(-1, 32).astype(np.uint64)
        words[b:e] = np.bitwise_or.reduce(blk << shifts, axis=1)
    return words, word_off


DNA_TO_IUPAC = np.array([8, 4, 2, 1], np.uint8)  # bio-seq Iupac: A=0b1000 C=0b0100 G=0b0010 T=0b0001


def pack_iupac4(codes4: np.ndarray, base_off: np.ndarray):
    """bio-seq Seq<Iupac> layout: symbol i at bits 4*(i%16) of word i/16; each
    sequence starts on a 16-byte boundary.  codes4: one 4-bit code per byte."""
    base_off = np.asarray(base_off, np.uint64)
    n = len(base_off) - 1
    lens = np.diff(base_off.astype(np.int64))
    nw = ((lens + 15) // 16 + 1) // 2 * 2
    word_off = np.zeros(n + 1, np.uint64)
    word_off[1:] = np.cumsum(nw)
    total_w = int(word_off[-1])
    words = np.zeros(total_w, np.uint64)
    if total_w == 0:
        return words, word_off
    row = np.repeat(np.arange(n), lens)
    i_in = np.arange(len(codes4), dtype=np.int64) - base_off[:-1].astype(np.int64)[row]
    padded = np.zeros(total_w * 16, np.uint8)
    padded[word_off[:-1].astype(np.int64)[row] * 16 + i_in] = codes4 & 15
    shifts = (4 * np.arange(16, dtype=np.uint64))
    step = 1 << 20
    for b in range(0, total_w, step):
        e = min(total_w, b + step)
        blk = padded[b * 16:e * 16].reshape(-1, 16).astype(np.uint64)
        words[b:e] = np.bitwise_or.reduce(blk << shifts, axis=1)
    return words, word_off


def tree_to_iupac(tree: SeqSet) -> np.ndarray:
    """One 4-bit IUPAC code per byte for an ACGT tree."""
    return DNA_TO_IUPAC[tree.codes & 3]


ASCII = np.frombuffer(b"ACGT", np.uint8)


def to_ascii(s: SeqSet, i: int) -> str:
    return ASCII[s.seq(i)].tobytes().decode()


IUPAC_ASCII = np.frombuffer(b"-TGKCYSBAWRDMHVN", np.uint8)  # index = bio-seq's 4-bit Iupac code


def iupac_to_ascii(codes4: np.ndarray) -> np.ndarray:
    """FASTA text (one byte per symbol) of 4-bit IUPAC codes."""
    return IUPAC_ASCII[np.asarray(codes4, np.uint8) & 15]
