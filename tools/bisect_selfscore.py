"""Diagnosis: leaves scored against their own tree (top-1 must be the leaf itself, score ~1), top-x path vs oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from myrio_b200 import api, synth  # noqa: E402


def main():
    ctx = api.Context(0)
    tree = synth.make_tree(2500, 91)
    st = api.TaxTreeStore(ctx, tree)
    ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, keep_leaf_vecs=True)
    q = ttc.leaf_vecs
    res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 5, q_first=1000, q_count=200)
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
    bad = 0
    for j in range(200):
        k, v = oln.row(1000 + j)
        s = oracle.score_all(oln, k, v)
        ts, ti = oracle.topx(s, 5)
        ok = (ti == res.top_ids[j]).all() and (ts.view(np.uint32) == res.top_scores[j].view(np.uint32)).all()
        if not ok:
            bad += 1
            if bad <= 3:
                print("query", j, "oracle", ti, ts, "gpu", res.top_ids[j], res.top_scores[j])
    print("ENV", {k: v for k, v in os.environ.items() if k.startswith("MYRIO_")}, "mismatching queries:", bad)


if __name__ == "__main__":
    main()
