"""Multi-GPU plumbing: reference leaves are sharded across ranks in
contiguous payload-id blocks, queries are replicated (all-gathered once), every
rank produces a shard-local top-x and ONE exchange step -- an all-gather of
Q x x (score, global payload id) pairs -- feeds the merge kernel (K7).

torch.distributed is used for the rendezvous and the collectives only
(NCCL on GPUs; gloo in the CPU tests of this module's host logic)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .synth import SeqSet


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of n units (leaves / reads) over `world` ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for_backend() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allgather_ragged(local: np.ndarray, group=None) -> list[np.ndarray]:
    """All-gather 1-D arrays of different lengths (one per rank), same dtype."""
    world = dist.get_world_size(group)
    dev = _device_for_backend()
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    raw = np.zeros(m * local.dtype.itemsize, np.uint8)
    raw[:local.nbytes] = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    mine = torch.from_numpy(raw).to(dev)
    bufs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    out = []
    for s, b in zip(sizes, bufs):
        out.append(b.cpu().numpy()[:s * local.dtype.itemsize].view(local.dtype).copy())
    return out


def allgather_reads(local: SeqSet, group=None) -> SeqSet:
    """Queries are broadcast: every rank ends up with the reads of all ranks, in rank order."""
    codes = allgather_ragged(local.codes, group)
    quals = allgather_ragged(local.qual, group)
    lens = allgather_ragged(local.lengths().astype(np.int64), group)
    all_lens = np.concatenate(lens)
    off = np.zeros(len(all_lens) + 1, np.uint64)
    off[1:] = np.cumsum(all_lens)
    return SeqSet(np.concatenate(codes), off, np.concatenate(quals), {"kind": "reads"})


def allgather_topx(scores: torch.Tensor, ids: torch.Tensor, group=None):
    """The one exchange step: [Q, x] per rank -> [world, Q, x] on every rank."""
    world = dist.get_world_size(group)
    gs = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    gi = torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(gs.view(-1), scores.contiguous().view(-1), group=group) \
        if dist.get_backend() == "nccl" else dist.all_gather(list(gs.unbind(0)), scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi.view(-1), ids.contiguous().view(-1), group=group) \
        if dist.get_backend() == "nccl" else dist.all_gather(list(gi.unbind(0)), ids.contiguous(), group=group)
    return gs, gi


def init_library_comm(ctx, group=None) -> dict:
    """Gives the library its own NCCL communicator over the ranks of `group`: rank 0 draws the unique id
    (myrio_comm_unique_id), torch.distributed only carries the 128 bytes, every rank calls myrio_comm_init.
    From then on the collectives of the hot path (bound all-reduce, top-x all-gather, clustering's per-row
    exchange) run inside libmyrio_b200.so on the context's stream."""
    from .api import Context
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    info = ctx.comm_info()
    if info["nranks"] == world and info["nccl_version"]:
        return info
    box = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(world, rank, box[0])
    return ctx.comm_info()


def sharded_topx_two_phase(ctx, ttcompute, queries, x: int, similarity: int = 0, group=None, max_batch: int = 0,
                           out=None):
    """The multi-GPU step with the bound exchange -- one library call (myrio_score_topx_sharded): per batch
    of queries the seed pass on this rank's leaves, ncclAllReduce(MAX) of the per-query bounds, the sweep with
    the global bound, ONE ncclAllGather of the shard lists as 8-byte composites, the merge (K7); all on the
    library's stream with the library's communicator (init_library_comm)."""
    from . import _lib
    Q = queries.n_rows
    dev = torch.device("cuda", ctx.device)
    init_library_comm(ctx, group)
    if out is not None:  # caller-owned result tensors (a steady-state loop then never touches the allocator)
        out_s, out_i = out
    else:
        with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream, device=dev)):
            out_s = torch.empty((Q, x), dtype=torch.float32, device=dev)
            out_i = torch.empty((Q, x), dtype=torch.int32, device=dev)
    _lib.check(ctx._L.myrio_score_topx_sharded(ctx.h, ttcompute.h, queries.h, 0, Q, x, similarity,
                                               C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()),
                                               max_batch))
    ctx.sync()
    return out_s, out_i


def sharded_topx(ctx, ttcompute, queries, x: int, similarity: int = 0, group=None, max_batch: int = 0, out=None):
    """The multi-GPU scoring step: shard-local top-x on this rank's leaves, exchange, merge (K7).
    Returns device tensors (scores [Q,x] f32, ids [Q,x] i32 holding u32 payload ids), identical on every
    rank and identical to the single-GPU result: two-phase with the bound exchange over NCCL inside the
    library (`sharded_topx_two_phase`); `sharded_topx_allgather` is the plain exchange of exact shard lists
    through torch.distributed (cross-check, and the gloo path of the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl":
        return sharded_topx_two_phase(ctx, ttcompute, queries, x, similarity, group, max_batch, out)
    return sharded_topx_allgather(ctx, ttcompute, queries, x, similarity, group, out)


def sharded_topx_allgather(ctx, ttcompute, queries, x: int, similarity: int = 0, group=None, out=None):
    """Exact shard-local top-x (K4+K5), all-gather of Q x x (score, id) pairs, merge (K7)."""
    from . import _lib
    Q = queries.n_rows
    dev = torch.device("cuda", ctx.device)
    single = not dist.is_initialized() or dist.get_world_size(group) == 1
    if out is not None and single:
        ls, li = out
    else:
        ls = torch.empty((Q, x), dtype=torch.float32, device=dev)
        li = torch.empty((Q, x), dtype=torch.int32, device=dev)
        torch.cuda.current_stream(dev).synchronize()
    _lib.check(ctx._L.myrio_score_topx_async(ctx.h, ttcompute.h, queries.h, 0, Q, x, similarity,
                                             C.c_void_p(ls.data_ptr()), C.c_void_p(li.data_ptr())))
    ctx.sync()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ls, li
    gs, gi = allgather_topx(ls, li, group)
    torch.cuda.current_stream(dev).synchronize()
    out_s, out_i = torch.empty_like(ls), torch.empty_like(li)
    _lib.check(ctx._L.myrio_topx_merge_async(ctx.h, C.c_void_p(gs.data_ptr()), C.c_void_p(gi.data_ptr()),
                                             gs.shape[0], Q, x, C.c_void_p(out_s.data_ptr()),
                                             C.c_void_p(out_i.data_ptr())))
    ctx.sync()
    return out_s, out_i


def make_allgather_callback(group=None):
    """The `myrio_allgather_fn` of a row-sharded clustering call over torch.distributed: host_buf
    holds `world` chunks of nbytes, this rank's chunk is filled; all chunks are filled on return.
    (NCCL: staged through a device tensor; the exchanged arrays are a few bytes per read.)"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _device_for_backend()

    import time
    stats = {"calls": 0, "seconds": 0.0, "bytes": 0}

    def allgather(user, host_buf, nbytes):
        t0 = time.perf_counter()
        try:
            nbytes = int(nbytes)
            stats["calls"] += 1
            stats["bytes"] += nbytes * world
            arr = np.ctypeslib.as_array(C.cast(host_buf, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
            mine = torch.from_numpy(arr[rank * nbytes:(rank + 1) * nbytes].copy()).to(dev)
            if dist.get_backend(group) == "nccl":
                out = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(out, mine, group=group)
                arr[:] = out.cpu().numpy()
            else:
                bufs = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(bufs, mine, group=group)
                for r, b in enumerate(bufs):
                    arr[r * nbytes:(r + 1) * nbytes] = b.numpy()
            return 0
        except Exception:  # never unwind through the C ABI
            import traceback
            traceback.print_exc()
            return 1
        finally:
            stats["seconds"] += time.perf_counter() - t0

    allgather.stats = stats
    return allgather


def cluster_sharded(myrseqs, params, group=None):
    """clustering::cluster over all ranks of `group`: the reads are replicated, the rows of
    every similarity tile are split across the ranks, per-row results are all-gathered.  Same result on
    every rank as api.cluster on one GPU."""
    from . import api
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return api.cluster(myrseqs, params)
    if dist.get_backend(group) == "nccl":
        # device-side exchange: ncclAllGather on the library's own communicator, no host staging
        init_library_comm(myrseqs.ctx, group)
        out = api.cluster(myrseqs, params, shard=(dist.get_rank(group), dist.get_world_size(group), None))
        out.exchange = {"path": "ncclAllGather inside libmyrio_b200.so"}
        return out
    cb = make_allgather_callback(group)  # any other backend (gloo in the CPU tests): the callback form
    out = api.cluster(myrseqs, params, shard=(dist.get_rank(group), dist.get_world_size(group), cb))
    out.exchange = dict(cb.stats)  # calls / seconds / bytes spent in the all-gathers (includes waiting for peers)
    out.exchange["path"] = "callback (torch.distributed)"
    return out
