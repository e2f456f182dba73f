// capi.cu -- the extern "C" boundary declared in include/myrio_b200.h.
#include <memory>
#include <mutex>

#include "cluster.cuh"
#include "comm.cuh"
#include "index.cuh"
#include "kmer_count.cuh"
#include "score.cuh"

namespace myrio {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
cudaStream_t &alloc_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}
}  // namespace myrio

using namespace myrio;

extern "C" {

const char *myrio_last_error(void) { return g_last_error.c_str(); }
int32_t myrio_version(void) { return 100; }

int32_t myrio_ctx_create(int32_t device, myrio_ctx **out) {
    return guarded([&] {
        if (!out) fail(MYRIO_ERR_BAD_ARG, "out is NULL");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            fail(MYRIO_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                 cudaGetErrorString(e));
        if (device < 0 || device >= count) fail(MYRIO_ERR_BAD_ARG, "device %d out of range (%d)", device, count);
        MYRIO_CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        MYRIO_CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            fail(MYRIO_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
        auto ctx = std::make_unique<myrio_ctx>();
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        MYRIO_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        {  // Keep freed scratch cached in the device's default stream-ordered pool instead of returning it to
           // the driver at every synchronisation (a steady-state step then never calls the allocator).  This is a
           // PROCESS-WIDE setting of that pool: MYRIO_POOL_RELEASE_THRESHOLD_MB bounds what stays cached (0 = the
           // CUDA default, release everything at the next synchronisation; unset = keep everything).
            cudaMemPool_t pool;
            MYRIO_CK(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t threshold = UINT64_MAX;
            if (const char *e = getenv("MYRIO_POOL_RELEASE_THRESHOLD_MB")) threshold = strtoull(e, nullptr, 10) << 20;
            MYRIO_CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        }
        ScopedCtx sc(ctx.get());
        upload_phred_table(ctx.get());
        sync(ctx.get());
        *out = ctx.release();
    });
}

void myrio_ctx_destroy(myrio_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    alloc_stream() = ctx->stream;
    for (auto &e : ctx->prof) {
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    ctx->score_rows.release();
    ctx->score_counters.release();
    ctx->topx_cand.release();
    ctx->topx_cand_cnt.release();
    ctx->topx_gthr.release();
    ctx->xchg_bound.release();
    ctx->xchg_top.release();
    ctx->xchg_all.release();
    ctx->clock_probes.release();
    comm_destroy(ctx);
    if (ctx->plan) plan_scratch_free(ctx->plan);
    ctx->plan = nullptr;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    alloc_stream() = nullptr;
    delete ctx;
}

int32_t myrio_ctx_sync(myrio_ctx *ctx) {
    return guarded([&] {
        ScopedCtx sc(ctx);
        sync(ctx);
    });
}

uint64_t myrio_ctx_launch_count(const myrio_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t myrio_ctx_set_kmer_bit_order(myrio_ctx *ctx, int32_t msb_first) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        ctx->kmer_msb_first = msb_first ? 1u : 0u;
    });
}

int32_t myrio_ctx_set_topx_bound_pruning(myrio_ctx *ctx, int32_t on) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        ctx->topx_bound_pruning = on ? 1 : 0;
    });
}

void *myrio_ctx_stream(const myrio_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t myrio_prof_enable(myrio_ctx *ctx, int32_t on) {
    return guarded([&] { ctx->prof_on = on != 0; });
}

int32_t myrio_prof_reset(myrio_ctx *ctx) {
    return guarded([&] {
        ScopedCtx sc(ctx);
        sync(ctx);
        for (auto &e : ctx->prof) {
            cudaEventDestroy(e.start);
            cudaEventDestroy(e.stop);
        }
        ctx->prof.clear();
    });
}

int32_t myrio_prof_query(myrio_ctx *ctx, const char *prefix, double *ms_total, uint64_t *launches) {
    return guarded([&] {
        ScopedCtx sc(ctx);
        sync(ctx);
        double ms = 0;
        uint64_t n = 0;
        const size_t pl = prefix ? strlen(prefix) : 0;
        for (auto &e : ctx->prof) {
            if (pl && e.name.compare(0, pl, prefix) != 0) continue;
            float t = 0;
            MYRIO_CK(cudaEventElapsedTime(&t, e.start, e.stop));
            ms += t;
            ++n;
        }
        if (ms_total) *ms_total = ms;
        if (launches) *launches = n;
    });
}

int32_t myrio_bench_smem_bandwidth(myrio_ctx *ctx, double *gbps, double *ms) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        ScopedCtx sc(ctx);
        bench_smem_bandwidth(ctx, gbps, ms);
    });
}

int32_t myrio_bench_clock_probe(myrio_ctx *ctx, uint32_t slot) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        ScopedCtx sc(ctx);
        bench_clock_probe(ctx, slot);
    });
}

int32_t myrio_bench_clock_read(myrio_ctx *ctx, uint32_t n, double *mhz) {
    return guarded([&] {
        if (!ctx || (n && !mhz)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        bench_clock_read(ctx, n, mhz);
    });
}

// ------------------------------------------------------------- sequences

static myrio_seqs *upload_seqs(myrio_ctx *ctx, int kind, const uint64_t *words, const uint64_t *word_off,
                               const uint64_t *base_off, const uint8_t *qual, uint64_t n) {
    if (n && (!words || !word_off || !base_off)) fail(MYRIO_ERR_BAD_ARG, "NULL input array");
    if (kind == 0 && n && !qual) fail(MYRIO_ERR_BAD_ARG, "NULL quality array");
    auto s = std::make_unique<myrio_seqs>();
    s->kind = kind;
    s->n = n;
    s->h_base_off.assign(n + 1, 0);
    if (n) memcpy(s->h_base_off.data(), base_off, (n + 1) * sizeof(uint64_t));
    const uint64_t total_w = n ? word_off[n] : 0, total_b = n ? base_off[n] : 0;
    const uint64_t per_word = kind == 0 ? 32 : 16;
    for (uint64_t r = 0; r < n; ++r) {
        if (word_off[r] & 1ull) fail(MYRIO_ERR_BAD_ARG, "sequence %llu does not start on an even word", (unsigned long long)r);
        const uint64_t len = base_off[r + 1] - base_off[r];
        if ((word_off[r + 1] - word_off[r]) * per_word < len)
            fail(MYRIO_ERR_BAD_ARG, "sequence %llu: word range too short for %llu symbols", (unsigned long long)r,
                 (unsigned long long)len);
    }
    s->words.alloc(total_w + 4);  // kernels may read one 16-byte vector past a sequence
    s->word_off.alloc(n + 1);
    s->base_off.alloc(n + 1);
    MYRIO_CK(cudaMemsetAsync(s->words.p, 0, (total_w + 4) * sizeof(uint64_t), ctx->stream));
    h2d(ctx, s->words.p, words, total_w);
    if (n) {
        h2d(ctx, s->word_off.p, word_off, n + 1);
        h2d(ctx, s->base_off.p, base_off, n + 1);
    } else {
        MYRIO_CK(cudaMemsetAsync(s->word_off.p, 0, sizeof(uint64_t), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(s->base_off.p, 0, sizeof(uint64_t), ctx->stream));
    }
    if (kind == 0) {
        s->qual.alloc(total_b + 32);
        MYRIO_CK(cudaMemsetAsync(s->qual.p, 0, total_b + 32, ctx->stream));
        h2d(ctx, s->qual.p, qual, total_b);
    } else {
        // ambiguity codes per leaf (a nibble with more than one bit set) -> scratch sizing
        s->h_aux.assign(n, 0);
        for (uint64_t r = 0; r < n; ++r) {
            const uint64_t len = base_off[r + 1] - base_off[r];
            const uint64_t nw = (len + 15) / 16;
            uint32_t amb = 0;
            for (uint64_t w = 0; w < nw; ++w) {
                const uint64_t a = words[word_off[r] + w];
                uint64_t t = (a & (a >> 1) & 0x7777777777777777ull) | (a & (a >> 2) & 0x3333333333333333ull) |
                             (a & (a >> 3) & 0x1111111111111111ull);
                t |= t >> 1;
                t |= t >> 2;
                amb += (uint32_t)__builtin_popcountll(t & 0x1111111111111111ull);
            }
            s->h_aux[r] = amb;
        }
    }
    sync(ctx);
    return s.release();
}

int32_t myrio_reads_upload(myrio_ctx *ctx, const uint64_t *words, const uint64_t *word_off,
                           const uint64_t *base_off, const uint8_t *qual, uint64_t n_reads,
                           myrio_seqs **out) {
    return guarded([&] {
        if (!ctx || !out) fail(MYRIO_ERR_BAD_ARG, "NULL ctx/out");
        ScopedCtx sc(ctx);
        *out = upload_seqs(ctx, 0, words, word_off, base_off, qual, n_reads);
    });
}

int32_t myrio_reads_from_fastq(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint8_t *qual_ascii,
                               const uint64_t *base_off, uint64_t n_records, uint64_t min_length,
                               float min_mean_qual, uint8_t max_qual, uint8_t *keep, myrio_seqs **out) {
    return guarded([&] {
        if (!ctx || !out) fail(MYRIO_ERR_BAD_ARG, "NULL ctx/out");
        ScopedCtx sc(ctx);
        *out = reads_from_fastq(ctx, seq_ascii, qual_ascii, base_off, n_records, min_length, min_mean_qual,
                                max_qual, keep);
    });
}

int32_t myrio_leaves_upload(myrio_ctx *ctx, const uint64_t *words, const uint64_t *word_off,
                            const uint64_t *base_off, uint64_t n_leaves, myrio_seqs **out) {
    return guarded([&] {
        if (!ctx || !out) fail(MYRIO_ERR_BAD_ARG, "NULL ctx/out");
        ScopedCtx sc(ctx);
        *out = upload_seqs(ctx, 1, words, word_off, base_off, nullptr, n_leaves);
    });
}

int32_t myrio_leaves_from_fasta(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint64_t *base_off,
                                uint64_t n_records, uint8_t *keep, myrio_seqs **out) {
    return guarded([&] {
        if (!ctx || !out) fail(MYRIO_ERR_BAD_ARG, "NULL ctx/out");
        ScopedCtx sc(ctx);
        *out = leaves_from_fasta(ctx, seq_ascii, base_off, n_records, keep);
    });
}

int32_t myrio_seqs_count(const myrio_seqs *s, uint64_t *n_seqs) {
    return guarded([&] {
        if (!s || !n_seqs) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        *n_seqs = s->n;
    });
}

void myrio_seqs_free(myrio_seqs *s) { delete s; }

// ------------------------------------------------------------- counting

int32_t myrio_count_reads(myrio_ctx *ctx, const myrio_seqs *reads, uint32_t k, float cutoff,
                          myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !reads || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = count_reads(ctx, reads, k, cutoff);
    });
}

int32_t myrio_count_leaves(myrio_ctx *ctx, const myrio_seqs *leaves, uint32_t k, uint32_t nb_resamples,
                           uint64_t seed, uint64_t leaf_id_base, myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !leaves || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = count_leaves(ctx, leaves, k, nb_resamples, seed, leaf_id_base);
    });
}

// --------------------------------------------------------------- svecs

int32_t myrio_svecs_info(const myrio_svecs *v, uint64_t *n_rows, uint64_t *nnz, int32_t *vtype) {
    return guarded([&] {
        if (!v) fail(MYRIO_ERR_BAD_ARG, "NULL svecs");
        if (n_rows) *n_rows = v->n_rows;
        if (nnz) *nnz = v->nnz;
        if (vtype) *vtype = v->vtype;
    });
}

int32_t myrio_svecs_download(myrio_ctx *ctx, const myrio_svecs *v, uint64_t *row_off, uint64_t *keys,
                             void *vals, void *aux) {
    return guarded([&] {
        if (!ctx || !v) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        if (row_off) d2h(ctx, row_off, v->row_off.p, v->n_rows + 1);
        if (keys) d2h(ctx, keys, v->keys.p, v->nnz);
        if (vals) {
            if (v->vtype == MYRIO_VAL_F32)
                d2h(ctx, (float *)vals, v->vals_f32.p, v->nnz);
            else
                d2h(ctx, (uint16_t *)vals, v->vals_u16.p, v->nnz);
        }
        if (aux) {
            if (v->aux_kind == 1)
                d2h(ctx, (uint64_t *)aux, v->aux_u64.p, v->n_rows);
            else if (v->aux_kind == 2)
                d2h(ctx, (float *)aux, v->aux_f32.p, v->n_rows);
            else
                fail(MYRIO_ERR_BAD_ARG, "this batch has no aux column");
        }
        sync(ctx);
    });
}

int32_t myrio_svecs_upload(myrio_ctx *ctx, uint64_t n_rows, const uint64_t *row_off, const uint64_t *keys,
                           const void *vals, int32_t vtype, myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !out || !row_off) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        if (vtype != MYRIO_VAL_F32 && vtype != MYRIO_VAL_U16) fail(MYRIO_ERR_BAD_ARG, "bad vtype");
        ScopedCtx sc(ctx);
        const uint64_t nnz = row_off[n_rows];
        if (nnz && (!keys || !vals)) fail(MYRIO_ERR_BAD_ARG, "NULL keys/vals");
        auto v = std::make_unique<myrio_svecs>();
        v->n_rows = n_rows;
        v->nnz = nnz;
        v->vtype = vtype;
        v->row_off.alloc(n_rows + 1);
        v->keys.alloc(std::max<uint64_t>(nnz, 1));
        h2d(ctx, v->row_off.p, row_off, n_rows + 1);
        h2d(ctx, v->keys.p, keys, nnz);
        if (vtype == MYRIO_VAL_F32) {
            v->vals_f32.alloc(std::max<uint64_t>(nnz, 1));
            h2d(ctx, v->vals_f32.p, (const float *)vals, nnz);
        } else {
            v->vals_u16.alloc(std::max<uint64_t>(nnz, 1));
            h2d(ctx, v->vals_u16.p, (const uint16_t *)vals, nnz);
        }
        sync(ctx);
        *out = v.release();
    });
}

int32_t myrio_svecs_normalize(myrio_ctx *ctx, const myrio_svecs *in, myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !in || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = svecs_normalize(ctx, in);
    });
}

void myrio_svecs_free(myrio_svecs *v) { delete v; }

int32_t myrio_svecs_device_ptrs(const myrio_svecs *v, const uint64_t **row_off, const uint64_t **keys,
                                const void **vals, const void **aux) {
    return guarded([&] {
        if (!v) fail(MYRIO_ERR_BAD_ARG, "NULL svecs");
        if (row_off) *row_off = v->row_off.p;
        if (keys) *keys = v->keys.p;
        if (vals) *vals = v->vtype == MYRIO_VAL_F32 ? (const void *)v->vals_f32.p : (const void *)v->vals_u16.p;
        if (aux) *aux = v->aux_kind == 1 ? (const void *)v->aux_u64.p : v->aux_kind == 2 ? (const void *)v->aux_f32.p : nullptr;
    });
}

// --------------------------------------------------------------- index

int32_t myrio_index_build(myrio_ctx *ctx, const myrio_svecs *leaf_vecs_normalized, uint32_t leaf_id_base,
                          uint32_t tile_leaves, myrio_index **out) {
    return guarded([&] {
        if (!ctx || !leaf_vecs_normalized || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = index_build(ctx, leaf_vecs_normalized, leaf_id_base, tile_leaves);
    });
}

int32_t myrio_index_info(const myrio_index *ix, uint64_t *n_leaves, uint64_t *n_tiles, uint64_t *n_postings,
                         uint64_t *n_unique_total, uint32_t *tile_leaves) {
    return guarded([&] {
        if (!ix) fail(MYRIO_ERR_BAD_ARG, "NULL index");
        if (n_leaves) *n_leaves = ix->n_leaves;
        if (n_tiles) *n_tiles = ix->tiles.size();
        if (n_postings) *n_postings = ix->n_postings;
        if (n_unique_total) *n_unique_total = ix->n_unique_total;
        if (tile_leaves) *tile_leaves = ix->tile_leaves;
    });
}

int32_t myrio_index_export_tile(myrio_ctx *ctx, const myrio_index *ix, uint64_t tile, uint64_t *n_unique,
                                uint64_t *n_postings, uint64_t *ukeys, uint64_t *offs, uint32_t *post_leaf,
                                float *post_val) {
    return guarded([&] {
        if (!ctx || !ix) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        index_export_tile(ctx, ix, tile, n_unique, n_postings, ukeys, offs, post_leaf, post_val);
    });
}

void myrio_index_free(myrio_index *ix) { delete ix; }

// ------------------------------------------------------------- scoring

int32_t myrio_score_all(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                        uint64_t q_count, int32_t similarity, float *scores) {
    return guarded([&] {
        if (!ctx || !ix || !queries || (!scores && q_count)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_all(ctx, ix, queries, q_first, q_count, similarity, scores);
    });
}

int32_t myrio_score_topx_async(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                               uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                               float *d_top_scores, uint32_t *d_top_ids) {
    return guarded([&] {
        if (!ctx || !ix || !queries || !d_top_scores || !d_top_ids) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_topx_device(ctx, ix, queries, q_first, q_count, x, similarity, d_top_scores, d_top_ids, false);
    });
}

int32_t myrio_score_topx(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                         uint64_t q_count, uint32_t x, int32_t similarity, float *top_scores,
                         uint32_t *top_ids) {
    return guarded([&] {
        if (!ctx || !ix || !queries || !top_scores || !top_ids) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        DevBuf<float> ds(std::max<uint64_t>(q_count * x, 1));
        DevBuf<uint32_t> di(std::max<uint64_t>(q_count * x, 1));
        score_topx_device(ctx, ix, queries, q_first, q_count, x, similarity, ds.p, di.p, true);
        d2h(ctx, top_scores, ds.p, q_count * x);
        d2h(ctx, top_ids, di.p, q_count * x);
        sync(ctx);
    });
}

int32_t myrio_score_all_direct(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, const myrio_svecs *queries,
                               uint64_t q_first, uint64_t q_count, int32_t similarity, float *scores) {
    return guarded([&] {
        if (!ctx || !leaf_vecs || !queries || (!scores && q_count)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_all_direct(ctx, leaf_vecs, queries, q_first, q_count, similarity, scores);
    });
}

int32_t myrio_score_topx_direct_async(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, uint32_t leaf_id_base,
                                      const myrio_svecs *queries, uint64_t q_first, uint64_t q_count,
                                      uint32_t x, int32_t similarity, float *d_top_scores,
                                      uint32_t *d_top_ids) {
    return guarded([&] {
        if (!ctx || !leaf_vecs || !queries || !d_top_scores || !d_top_ids) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_topx_direct_device(ctx, leaf_vecs, leaf_id_base, queries, q_first, q_count, x, similarity,
                                 d_top_scores, d_top_ids);
    });
}

int32_t myrio_score_topx_direct(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, uint32_t leaf_id_base,
                                const myrio_svecs *queries, uint64_t q_first, uint64_t q_count, uint32_t x,
                                int32_t similarity, float *top_scores, uint32_t *top_ids) {
    return guarded([&] {
        if (!ctx || !leaf_vecs || !queries || !top_scores || !top_ids) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        DevBuf<float> ds(std::max<uint64_t>(q_count * x, 1));
        DevBuf<uint32_t> di(std::max<uint64_t>(q_count * x, 1));
        score_topx_direct_device(ctx, leaf_vecs, leaf_id_base, queries, q_first, q_count, x, similarity, ds.p,
                                 di.p);
        d2h(ctx, top_scores, ds.p, q_count * x);
        d2h(ctx, top_ids, di.p, q_count * x);
        sync(ctx);
    });
}

int32_t myrio_topx_merge_async(myrio_ctx *ctx, const float *d_scores_in, const uint32_t *d_ids_in,
                               uint32_t n_parts, uint64_t q_count, uint32_t x, float *d_scores_out,
                               uint32_t *d_ids_out) {
    return guarded([&] {
        if (!ctx || !d_scores_in || !d_ids_in || !d_scores_out || !d_ids_out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        topx_merge_device(ctx, d_scores_in, d_ids_in, n_parts, q_count, x, d_scores_out, d_ids_out);
    });
}

int32_t myrio_score_topx_seed_async(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                                    uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                                    uint32_t *d_bound) {
    return guarded([&] {
        if (!ctx || !ix || !queries || !d_bound) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_topx_seed_device(ctx, ix, queries, q_first, q_count, x, similarity, d_bound);
    });
}

int32_t myrio_score_topx_sweep_async(myrio_ctx *ctx, const myrio_index *ix, uint64_t q_count, uint32_t x,
                                     int32_t similarity, const uint32_t *d_bound, uint64_t *d_top) {
    return guarded([&] {
        if (!ctx || !ix || !d_top) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_topx_sweep_device(ctx, ix, q_count, x, similarity, d_bound, d_top);
    });
}

int32_t myrio_topx_merge_composites_async(myrio_ctx *ctx, const uint64_t *d_in, uint32_t n_parts,
                                          uint64_t q_count, uint32_t x, float *d_scores_out,
                                          uint32_t *d_ids_out) {
    return guarded([&] {
        if (!ctx || !d_in || !d_scores_out || !d_ids_out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        topx_merge_composites_device(ctx, d_in, n_parts, q_count, x, d_scores_out, d_ids_out);
    });
}

// ------------------------------------------------------------- multi-GPU

int32_t myrio_comm_unique_id(uint8_t id[MYRIO_UNIQUE_ID_BYTES]) {
    return guarded([&] {
        if (!id) fail(MYRIO_ERR_BAD_ARG, "NULL id");
        comm_unique_id(id);
    });
}

int32_t myrio_comm_init(myrio_ctx *ctx, uint32_t nranks, uint32_t rank, const uint8_t id[MYRIO_UNIQUE_ID_BYTES]) {
    return guarded([&] {
        if (!ctx || !id) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        comm_init(ctx, nranks, rank, id);
    });
}

int32_t myrio_comm_destroy(myrio_ctx *ctx) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        ScopedCtx sc(ctx);
        comm_destroy(ctx);
    });
}

int32_t myrio_comm_info(const myrio_ctx *ctx, uint32_t *rank, uint32_t *nranks, int32_t *nccl_version) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        if (rank) *rank = ctx->comm_rank;
        if (nranks) *nranks = ctx->comm_world;
        if (nccl_version) *nccl_version = ctx->comm ? comm_nccl_version() : 0;
    });
}

int32_t myrio_score_topx_sharded(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                                 uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                                 float *d_top_scores, uint32_t *d_top_ids, uint64_t max_batch) {
    return guarded([&] {
        if (!ctx || !ix || !queries || !d_top_scores || !d_top_ids) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        score_topx_sharded(ctx, ix, queries, q_first, q_count, x, similarity, d_top_scores, d_top_ids, max_batch);
    });
}

int32_t myrio_score_stats(myrio_ctx *ctx, uint64_t *n_query_keys, uint64_t *n_key_hits,
                          uint64_t *n_postings_touched) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        if (n_query_keys) *n_query_keys = ctx->last_stats[0];
        if (n_key_hits) *n_key_hits = ctx->last_stats[1];
        if (n_postings_touched) *n_postings_touched = ctx->last_stats[2];
    });
}

// ---------------------------------------------------------- clustering

int32_t myrio_cluster(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *params,
                      const myrio_svecs *initial_centroids, int32_t *assign, uint32_t *n_iters,
                      float *silhouette) {
    return guarded([&] {
        if (!ctx || !reads || !params || (!assign && reads->n)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        cluster_reads(ctx, reads, params, initial_centroids, nullptr, assign, n_iters, silhouette);
    });
}

int32_t myrio_cluster_sharded(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *params,
                              const myrio_svecs *initial_centroids, const myrio_shard *shard, int32_t *assign,
                              uint32_t *n_iters, float *silhouette) {
    return guarded([&] {
        if (!ctx || !reads || !params || !shard || (!assign && reads->n)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        cluster_reads(ctx, reads, params, initial_centroids, shard, assign, n_iters, silhouette);
    });
}

int32_t myrio_compute_cluster_centroid(myrio_ctx *ctx, const myrio_svecs *counts, const uint64_t *member_ids,
                                       uint64_t m, myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !counts || !out || (!member_ids && m)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = cluster_centroid(ctx, counts, member_ids, m);
    });
}

int32_t myrio_reweight_fingerprints(myrio_ctx *ctx, const myrio_svecs *local_fingerprints,
                                    const myrio_svecs *naive_query, int32_t similarity, float alpha,
                                    myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !local_fingerprints || !naive_query || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = reweight_fingerprints(ctx, local_fingerprints, naive_query, similarity, alpha);
    });
}

int32_t myrio_pairwise(myrio_ctx *ctx, const myrio_svecs *rows, const myrio_svecs *cols, int32_t similarity,
                       float *out) {
    return guarded([&] {
        if (!ctx || !rows || !cols || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        pairwise(ctx, rows, cols, similarity, out);
    });
}

int32_t myrio_cluster_stats(myrio_ctx *ctx, uint64_t *n_pair_similarities, uint64_t *n_terms) {
    return guarded([&] {
        if (!ctx) fail(MYRIO_ERR_BAD_ARG, "NULL ctx");
        if (n_pair_similarities) *n_pair_similarities = ctx->cluster_stats[0];
        if (n_terms) *n_terms = ctx->cluster_stats[1];
    });
}

}  // extern "C"
