"""Kernel-time breakdown of the index build (K2 + normalise + K3) on the bench tree (one GPU)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from myrio_b200 import api, synth  # noqa: E402

NAMES = ["k2_", "compact_rows", "row_norm", "row_scale", "k3_make_pairs", "k3_radix", "k3_mark", "k3_localize", "scan_",
         "k3_gather", "k3_emit_unique", "k3_count_odd", "k3_classify_dense", "k3_dense_index", "k3_fill_bitmaps",
         "k3_iota", "k3_make_entries", "k3_gdict_insert", "k3_flag_odd", "k3_flag_ordered", ""]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ctx = api.Context(0)
    tree = synth.make_tree(n, 1002, root_seed=1002, shard=0)
    store = api.TaxTreeStore(ctx, tree)
    store.compute_and_append_kmer_counts(18, 32)
    ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32)
    del ttc
    ctx.prof_reset()
    ctx.prof_enable(True)
    t0 = time.perf_counter()
    ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32)
    ctx.sync()
    wall = (time.perf_counter() - t0) * 1e3
    out = {"leaves": n, "from_store_tree_ms": round(wall, 2)}
    for p in NAMES:
        ms, cnt = ctx.prof_query(p)
        if cnt:
            out[p or "all_kernels"] = [round(ms, 3), cnt]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
