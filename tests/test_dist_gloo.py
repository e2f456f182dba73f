"""CPU-only, world_size 2 over gloo: the host logic of the sharded path
(myrio_b200/dist.py): contiguous leaf shards, replicated (all-gathered) queries,
shard-local top-x, one all-gather, merge with the (score desc, payload_id asc)
rule.  The per-shard scores come from the oracle here (no GPU in this container);
the merge rule is checked against the unsharded oracle result, which is what the
CUDA merge kernel (K7) is tested against on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
X = 12


def merge_rule(gs: np.ndarray, gi: np.ndarray, x: int):
    """[parts, Q, x] -> [Q, x]: composite order used by topx_merge_kernel (score.cu)."""
    parts, Q, _ = gs.shape
    out_s = np.zeros((Q, x), np.float32)
    out_i = np.zeros((Q, x), np.uint32)
    for q in range(Q):
        s = gs[:, q, :].reshape(-1)
        i = gi[:, q, :].reshape(-1).astype(np.uint32)
        order = np.lexsort((i, -s.astype(np.float64)))
        out_s[q], out_i[q] = s[order][:x], i[order][:x]
    return out_s, out_i


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from myrio_b200 import dist as mdist
    from myrio_b200 import synth

    n_leaves, n_reads = 240, 18
    tree = synth.make_tree(n_leaves, 77)                      # every rank derives the same tree
    lo, hi = mdist.shard_bounds(n_leaves, rank, world)
    shard = tree.subset(np.arange(lo, hi))
    rlo, rhi = mdist.shard_bounds(n_reads, rank, world)
    all_reads = synth.make_reads(tree, n_reads, 78)
    mine = all_reads.subset(np.arange(rlo, rhi))
    reads = mdist.allgather_reads(mine)                       # queries are broadcast
    assert reads.n == n_reads
    assert (reads.codes == all_reads.codes).all() and (reads.qual == all_reads.qual).all()
    assert (reads.base_off == all_reads.base_off).all()

    oq = oracle.normalize(oracle.count_reads(reads.codes, reads.qual, reads.base_off, 18, 0.1))
    ol = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(shard), shard.base_off, 18, 32,
                                              leaf_id_base=lo))
    ls = np.zeros((n_reads, X), np.float32)
    li = np.zeros((n_reads, X), np.uint32)
    for q in range(n_reads):
        k, v = oq.row(q)
        ls[q], li[q] = oracle.topx(oracle.score_all(ol, k, v), X, id_base=lo)
    gs, gi = mdist.allgather_topx(torch.from_numpy(ls), torch.from_numpy(li.view(np.int32)))
    assert tuple(gs.shape) == (world, n_reads, X)
    ms, mi = merge_rule(gs.numpy(), gi.numpy().view(np.uint32), X)

    full = oracle.normalize(oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, 18, 32))
    for q in range(n_reads):
        k, v = oq.row(q)
        es, ei = oracle.topx(oracle.score_all(full, k, v), X)
        assert (ei == mi[q]).all() and (es.view(np.uint32) == ms[q].view(np.uint32)).all()

    # the two-phase exchange (dist.sharded_topx_two_phase): every rank's bound = its own X-th best score,
    # all-reduce(MAX), shard lists pruned to the global bound (padding = score 0 / id 0xFFFFFFFF) -- the
    # merged list is unchanged although the shard lists got shorter
    bound = torch.from_numpy(ls[:, X - 1].copy().view(np.int32))
    dist.all_reduce(bound, op=dist.ReduceOp.MAX)
    gb = bound.numpy().view(np.float32)
    ps, pi = ls.copy(), li.copy()
    drop = ps < gb[:, None]
    ps[drop], pi[drop] = 0.0, 0xFFFFFFFF
    gs2, gi2 = mdist.allgather_topx(torch.from_numpy(ps), torch.from_numpy(pi.view(np.int32)))
    ms2, mi2 = merge_rule(gs2.numpy(), gi2.numpy().view(np.uint32), X)
    assert (mi2 == mi).all() and (ms2.view(np.uint32) == ms.view(np.uint32)).all()
    assert drop.any()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_sharded_topx_host_logic_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_cover_everything():
    from myrio_b200.dist import shard_bounds
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _cb_worker(rank, world, port, tmp):
    """The exchange of the row-sharded clustering (myrio_shard.allgather) over gloo: called through the
    C function-pointer type exactly as libmyrio_b200.so calls it."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myrio_b200 import _lib
    from myrio_b200 import dist as mdist
    cb = _lib.ALLGATHER_FN(mdist.make_allgather_callback())
    for chunk in (1, 7, 4096):
        # float payload like (best similarity | silhouette sums), chunk elements per rank
        buf = np.full(world * chunk, -1.0, np.float32)
        buf[rank * chunk:(rank + 1) * chunk] = np.arange(chunk, dtype=np.float32) + 1000 * rank
        rc = cb(None, buf.ctypes.data_as(C.c_void_p), chunk * 4)
        assert rc == 0
        for r in range(world):
            assert (buf[r * chunk:(r + 1) * chunk] == np.arange(chunk, dtype=np.float32) + 1000 * r).all()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"cb{rank}"), "w").write("ok")


def test_cluster_shard_allgather_callback_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_cb_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "cb0").exists() and (tmp_path / "cb1").exists()
