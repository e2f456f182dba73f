// comm.cu -- the multi-GPU entry points of the library: one process per GPU, the
// reference leaves sharded over the ranks in contiguous payload-id blocks, queries replicated.
//
//   myrio_comm_init            ncclCommInitRank on the context's device
//   myrio_score_topx_sharded   per batch of queries: seed pass on this rank's leaves ->
//                              ncclAllReduce(MAX) of the per-query bounds (4 B each) -> sweep with the
//                              exchanged bound -> ONE ncclAllGather of the shard lists (8-byte composites)
//                              -> K7 merge.  Everything is enqueued on the context's stream.
//   comm_allgather_inplace     used by the row-sharded cluster() (cluster.cu) for its per-row results.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the SONAME both the system package and the wheel
// bundled with PyTorch carry -- inside a process that has already loaded PyTorch the loader returns that
// same copy), so the library itself has no link-time dependency on NCCL and single-GPU users never load
// it.  The reference has nothing to compare with: it is a single-process rayon program.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "comm.cuh"
#include "score.cuh"

namespace myrio {

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("MYRIO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *s) { return dlsym(api.handle, s); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce ||
        !api.AllGather || !api.GetErrorString)
        fail(MYRIO_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2: %s)", dlerror() ? dlerror() : "missing symbols");
    return api;
}

#define MYRIO_NCCL(expr)                                                                              \
    do {                                                                                              \
        ncclResult_t r__ = (expr);                                                                    \
        if (r__ != ncclSuccess)                                                                       \
            ::myrio::fail(MYRIO_ERR_COMM, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(r__)); \
    } while (0)

__global__ void set_u32_kernel(uint32_t *p, uint32_t v) { *p = v; }
}  // namespace

void comm_unique_id(uint8_t *id128) {
    static_assert(sizeof(ncclUniqueId) == MYRIO_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    MYRIO_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
}

void comm_init(myrio_ctx *ctx, uint32_t nranks, uint32_t rank, const uint8_t *id128) {
    if (nranks == 0 || rank >= nranks) fail(MYRIO_ERR_BAD_ARG, "bad rank %u of %u", rank, nranks);
    if (ctx->comm) fail(MYRIO_ERR_BAD_ARG, "this context already has a communicator");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    MYRIO_NCCL(nccl().CommInitRank(&comm, (int)nranks, id, (int)rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = nranks;
}

void comm_destroy(myrio_ctx *ctx) {
    if (!ctx->comm) return;
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
}

int comm_nccl_version() {
    int v = 0;
    if (nccl().GetVersion) nccl().GetVersion(&v);
    return v;
}

// buf holds `world` chunks of bytes_per_rank; chunk `rank` is filled on entry, all chunks on return
void comm_allgather_inplace(myrio_ctx *ctx, void *d_buf, uint64_t bytes_per_rank) {
    if (!ctx->comm) fail(MYRIO_ERR_BAD_ARG, "no communicator: call myrio_comm_init first");
    if (bytes_per_rank == 0) return;
    const uint8_t *mine = static_cast<const uint8_t *>(d_buf) + (size_t)ctx->comm_rank * bytes_per_rank;
    MYRIO_NCCL(nccl().AllGather(mine, d_buf, bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
}

void comm_allreduce_max_u32(myrio_ctx *ctx, uint32_t *d_buf, uint64_t n) {
    if (!ctx->comm) fail(MYRIO_ERR_BAD_ARG, "no communicator: call myrio_comm_init first");
    if (n == 0) return;
    MYRIO_NCCL(nccl().AllReduce(d_buf, d_buf, n, ncclUint32, ncclMax, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
}

// The multi-GPU scoring step.  Identical result on every rank, identical to the single-GPU top-x over the
// union of the shards (the bound is a lower bound of the global x-th best score; ties are resolved by the
// composite key in K7).
//
// No host synchronisation between the collectives of a call: the whole sequence of every batch is queued on the
// context's stream and the host waits once, at the end.  (A first version read an error flag back after every
// all-reduce; with the queue drained twice per step any host-side hiccup -- an NVML query holding the driver lock
// was enough -- left all N GPUs idle for its duration: profiles/r02_smi_hiccups.md.)
// Error agreement without a mid-call sync: a rank whose seed pass fails joins the batch's collectives with zero
// bounds, an empty list and flag = 1 in the extra element of the bound array; the MAX all-reduce spreads the
// flag, a kernel ORs it into a per-call word, and every rank reads that word after its last batch: the failing
// rank rethrows its own error, its peers return MYRIO_ERR_COMM.  Nobody waits for a rank that is not coming.
// The batch size (the ranks must cut the queries at the same places) is agreed once per (index, x) with an
// all-reduce and kept with the index.
__global__ void or_flag_kernel(const uint32_t *__restrict__ flag, uint32_t *__restrict__ acc) {
    if (*flag) *acc = 1u;
}

void score_topx_sharded(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                        uint64_t q_count, uint32_t x, int32_t sim, float *d_top_scores, uint32_t *d_top_ids,
                        uint64_t max_batch) {
    if (!ctx->comm || ctx->comm_world == 1) {
        score_topx_device(ctx, ix, queries, q_first, q_count, x, sim, d_top_scores, d_top_ids, false);
        return;
    }
    const uint32_t W = ctx->comm_world;
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    ctx->xchg_bound.reserve(std::min<uint64_t>(std::max<uint64_t>(q_count, 1), 0x7FFFFFFFull) + 4);
    if (ix->agreed_x != x || ix->agreed_world != W) {
        // smallest capacity over the ranks (the tile count of a shard decides how many queries' candidates fit):
        // ~0 - cap so that MAX gives the minimum.  One synchronising exchange per (index, x), not per call.
        const uint64_t cap_local = std::min<uint64_t>(score_topx_batch_capacity(ix, x), 0x7FFFFFFFull);
        uint32_t h_cap = 0xFFFFFFFFu - (uint32_t)cap_local;
        h2d(ctx, ctx->xchg_bound.p, &h_cap, 1);
        MYRIO_NCCL(nccl().AllReduce(ctx->xchg_bound.p, ctx->xchg_bound.p, 1, ncclUint32, ncclMax, comm, ctx->stream));
        d2h(ctx, &h_cap, ctx->xchg_bound.p, 1);
        sync(ctx);
        ix->agreed_cap = 0xFFFFFFFFu - h_cap;
        ix->agreed_x = x;
        ix->agreed_world = W;
    }
    uint64_t batch = ix->agreed_cap;
    if (max_batch) batch = std::min(batch, max_batch);
    batch = std::max<uint64_t>(batch, 1);

    DevBuf<uint32_t> &bound = ctx->xchg_bound;
    DevBuf<uint64_t> &top = ctx->xchg_top, &all = ctx->xchg_all;
    const uint64_t bq = std::min(batch, std::max<uint64_t>(q_count, 1));
    bound.reserve(bq + 4);  // [0, qc) bounds, [qc] this batch's error flag; the per-call flag word sits at the end
    top.reserve(bq * x);
    all.reserve((uint64_t)W * bq * x);
    uint32_t *d_call_flag = bound.p + bound.n - 1;
    MYRIO_CK(cudaMemsetAsync(d_call_flag, 0, sizeof(uint32_t), ctx->stream));
    Failure local{MYRIO_OK, ""};
    for (uint64_t q0 = 0; q0 < q_count; q0 += batch) {
        const uint64_t qc = std::min(batch, q_count - q0);
        bool ok = local.code == MYRIO_OK;
        if (ok) {
            try {
                score_topx_seed_device(ctx, ix, queries, q_first + q0, qc, x, sim, bound.p);
            } catch (const Failure &e) {
                local = e;
                ok = false;
            }
        }
        if (!ok) MYRIO_CK(cudaMemsetAsync(bound.p, 0, qc * sizeof(uint32_t), ctx->stream));
        MYRIO_LAUNCH(ctx, "k7_set_flag", set_u32_kernel, 1, 1, 0, bound.p + qc, ok ? 0u : 1u);
        MYRIO_NCCL(nccl().AllReduce(bound.p, bound.p, qc + 1, ncclUint32, ncclMax, comm, ctx->stream));
        MYRIO_LAUNCH(ctx, "k7_or_flag", or_flag_kernel, 1, 1, 0, bound.p + qc, d_call_flag);
        if (ok) {
            try {
                score_topx_sweep_device(ctx, ix, qc, x, sim, bound.p, top.p);
            } catch (const Failure &e) {
                local = e;
                ok = false;
            }
        }
        if (!ok) MYRIO_CK(cudaMemsetAsync(top.p, 0, qc * x * sizeof(uint64_t), ctx->stream));
        MYRIO_NCCL(nccl().AllGather(top.p, all.p, qc * x, ncclUint64, comm, ctx->stream));
        topx_merge_composites_device(ctx, all.p, W, qc, x, d_top_scores + q0 * x, d_top_ids + q0 * x);
    }
    uint32_t any_err = 0;
    d2h(ctx, &any_err, d_call_flag, 1);
    sync(ctx);
    if (local.code != MYRIO_OK) throw local;
    if (any_err) fail(MYRIO_ERR_COMM, "the scoring pass failed on another rank");
}

}  // namespace myrio
