"""Generates tests/golden/*.json from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

  python tests/golden/make_golden.py [/root/reference]

phred_correct_f32.json / phred_error_f32.json: the 128 decimal literals of
myrio-core/src/constants.rs:11,28 rounded to f32 with exact rational arithmetic
(what rustc does for an f32 literal), stored as uint32 bit patterns.
"""
import json
import os
import re
import struct
import sys
from fractions import Fraction


def f32_round_exact(lit: str) -> int:
    """Nearest-even f32 of a decimal literal, exactly; returns the bit pattern."""
    x = Fraction(lit)
    if x == 0:
        return 0
    # candidate from double rounding, then fix up by comparing exact neighbours
    def bits(f):
        return struct.unpack("<I", struct.pack("<f", f))[0]

    def val(b):
        return Fraction(struct.unpack("<f", struct.pack("<I", b))[0])

    b = bits(float(x))
    best = b
    for cand in (b - 1, b, b + 1):
        if cand < 0:
            continue
        d_c, d_b = abs(val(cand) - x), abs(val(best) - x)
        if d_c < d_b or (d_c == d_b and cand % 2 == 0 and best % 2 == 1):
            best = cand
    return best


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = open(os.path.join(ref, "myrio-core/src/constants.rs")).read()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, out in (("Q_TO_BP_CALL_CORRECT_PROB_MAP", "phred_correct_f32.json"),
                      ("Q_TO_BP_CALL_ERROR_PROB_MAP", "phred_error_f32.json")):
        m = re.search(r"pub const " + name + r": \[Float; 128\] = \[(.*?)\];", src, re.S)
        lits = [t.strip() for t in m.group(1).split(",") if t.strip()]
        assert len(lits) == 128
        bits = [f32_round_exact(t) for t in lits]
        json.dump({"source": "myrio-core/src/constants.rs " + name, "f32_bits": bits},
                  open(os.path.join(here, out), "w"))
        print(out, "ok")


def make_tiny_myrtree():
    """tests/golden/tiny.myrtree: a three-leaf store WRITTEN BY THE LIBRARY's writer (myrio_store_to_bytes, zstd
    level 6, one chunk), with the cached counts of k = 6 and 18 taken from the CPU oracle (32 resamples, seed 0:
    leaf 1 holds ambiguity codes, so its counts depend on the documented counter-based resampler).  No GPU and no
    reference checkout needed; tests/test_myrtree.py re-reads it and compares its sections with hand-assembled
    bincode bytes."""
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))
    sys.path.insert(0, root)
    import numpy as np
    import oracle
    from myrio_b200 import api
    iupac = {c: i for i, c in enumerate("-TGKCYSBAWRDMHVN")}
    seqs = ["ACGTACGTTTGACCA", "ACGTACGTTTGACCANNRY-ACGT", "GGGG"]
    codes = np.array([iupac[c] for s in seqs for c in s], np.uint8)
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    f = api.MyrTreeFile.new("trnH-psbA", api.RANKS["Family"], [0, 0, 1, 1, 0, 1], [2, 2, 0, 1, 1, 2],
                            ["f:Pinaceae", "g:Abies", "s:Abies alba", "s:Abies alba", "g:Picea", "s:Picea abies"],
                            codes, off)
    for k in (6, 18):
        c = oracle.count_leaves(codes, off, k, 32)
        f.append_counts_host(k, c.row_off, c.keys, c.vals, c.aux)
    f.save_to_file(os.path.join(here, "tiny.myrtree"), 6, threads=1)
    print("tiny.myrtree ok", os.path.getsize(os.path.join(here, "tiny.myrtree")), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "myrtree":
        make_tiny_myrtree()
    else:
        main()
        make_tiny_myrtree()
