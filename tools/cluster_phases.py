"""Kernel-time breakdown of clustering::cluster on the bench read batch (one GPU): CUDA-event time per
kernel name prefix (myrio_prof_query) next to the wall clock of the call, with and without silhouette
trimming.  usage: python tools/cluster_phases.py [n_reads]   (MYRIO_CLUSTER_TIMING=1 adds the host marks)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from myrio_b200 import api, synth  # noqa: E402

PREFIXES = ["k1_", "compact_rows", "scan_", "row_norm", "row_scale", "k6_slab_count", "k6_slab_fill",
            "k6_slab_schedule", "k6_pair_store", "k6_pair_argmax", "k6_argmax_rows", "k6_recenter_accumulate",
            "k6_recenter_finalize", "k6_pair_silhouette", "k6_silsum_rows", ""]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ctx = api.Context(0)
    tree = synth.make_tree(100_000, 1002, root_seed=1002, shard=0)
    reads = synth.make_reads(tree, n, 2002)
    m = api.MyrSeqs(ctx, reads)
    for sil in (None, 1.4):
        prm = api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=4, silhouette_trimming=sil)
        api.cluster(m, prm)
        ctx.prof_reset()
        ctx.prof_enable(True)
        t0 = time.perf_counter()
        res = api.cluster(m, prm)
        wall = (time.perf_counter() - t0) * 1e3
        out = {"silhouette": sil, "reads": n, "call_ms": round(wall, 2), "iters": res.n_iters}
        for p in PREFIXES:
            ms, cnt = ctx.prof_query(p)
            out[p or "all_kernels"] = [round(ms, 3), cnt]
        ctx.prof_enable(False)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
