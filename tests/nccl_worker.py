"""Worker of tests/test_gpu_nccl.py: launched by torch.distributed.run with one rank per GPU.
The collectives under test run INSIDE libmyrio_b200.so (myrio_comm_init / myrio_score_topx_sharded /
myrio_cluster_sharded without a callback); torch.distributed only does the rendezvous."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import oracle
    from myrio_b200 import api, synth
    from myrio_b200 import dist as mdist

    ctx = api.Context(local)
    info = mdist.init_library_comm(ctx)
    assert info["nranks"] == world and info["rank"] == rank and info["nccl_version"] > 20000, info

    # ---- sharded scoring: contiguous payload-id shards, replicated queries
    n_leaves, n_reads, x = 3000, 400, 50
    tree = synth.make_tree(n_leaves, 501)
    reads = synth.make_reads(tree, n_reads, 502)
    lo, hi = mdist.shard_bounds(n_leaves, rank, world)
    store = api.TaxTreeStore(ctx, tree.subset(np.arange(lo, hi)), leaf_id_base=lo)
    ttc = api.TaxTreeCompute.from_store_tree(store, 18, 32, tile_leaves=256)
    qn = api.MyrSeqs(ctx, reads).compute_kmer_counts(18, 0.1).into_normalized_l2()
    oln = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
    oqn = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
    for max_batch in (0, 150):  # one batch / three batches (the ranks must cut at the same places)
        s, i = mdist.sharded_topx(ctx, ttc, qn, x, api.SIM_COSINE, max_batch=max_batch)
        s, i = s.cpu().numpy(), i.cpu().numpy().view(np.uint32)
        for q in range(0, n_reads, 3):
            k, v = oqn.row(q)
            es, ei = oracle.topx(oracle.score_all(oln, k, v), x)
            assert (ei == i[q]).all() and (bits(es) == bits(s[q])).all(), (rank, max_batch, q)
    # the plain exchange through torch.distributed gives the same lists
    s2, i2 = mdist.sharded_topx_allgather(ctx, ttc, qn, x, api.SIM_COSINE)
    assert (i2.cpu().numpy().view(np.uint32) == i).all() and (bits(s2.cpu().numpy()) == bits(s)).all()

    # ---- row-sharded clustering with the device-side exchange (no callback)
    t1, t2 = synth.make_tree(80, 61), synth.make_tree(80, 62)
    mixed = synth.concat([synth.make_reads(t1, 300, 63), synth.make_reads(t2, 300, 64)])
    m = api.MyrSeqs(ctx, mixed)
    oc = oracle.count_reads(mixed.codes, mixed.qual, mixed.base_off, 6, 0.2)
    for sil in (None, 1.4):
        prm = api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=3, silhouette_trimming=sil)
        out = mdist.cluster_sharded(m, prm)
        assert "ncclAllGather" in out.exchange["path"]
        oassign, oiters, osil = oracle.cluster(oc, 20, None, 3, silhouette=sil)
        assert (out.assign == oassign).all() and out.n_iters == oiters, (rank, sil)
        if sil is not None:
            both = ~np.isnan(osil)
            assert (bits(out.silhouette[both]) == bits(osil[both])).all()
    dist.barrier()
    if rank == 0:
        print(f"NCCL_WORKER_OK world={world} nccl={info['nccl_version']}", flush=True)
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
