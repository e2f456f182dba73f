import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from myrio_b200 import api, synth
ctx = api.Context(0)
tree = synth.make_tree(2500, 91)
st = api.TaxTreeStore(ctx, tree)
ttc = api.TaxTreeCompute.from_store_tree(st, 18, 32, keep_leaf_vecs=True)
q = ttc.leaf_vecs
res = api.TaxTreeResults.from_compute_tree(q, ttc, api.SIM_COSINE, 5, q_first=1111, q_count=1, full_scores=True)
oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
k, v = oln.row(1111)
s = oracle.score_all(oln, k, v)
g = res.scores[0]
bad = np.nonzero(s.view(np.uint32) != g.view(np.uint32))[0]
print("rows mismatching leaves:", len(bad), bad[:40])
# deficit in units of p = v1*unit
ro, keys, vals, _ = oln.row_off, oln.keys, oln.vals, None
v1 = v.min()
for l in bad[:12]:
    kk, vv = oln.row(int(l))
    p = np.float32(v1) * np.float32(vv.min())
    print(int(l), float(s[l]), float(g[l]), "deficit/p = %.2f" % ((s[l] - g[l]) / p), "shared", len(np.intersect1d(k, kk)))
print("top", res.top_ids[0], res.top_scores[0])
