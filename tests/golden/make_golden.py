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


if __name__ == "__main__":
    main()
