// score.cu -- K4 (query x leaf similarity over the inverted index), K5 (top-x
// select) and K7 (merge of shard-local top-x lists).
//
// Reference semantics (myrio-core/src/tax/results.rs:82-93 with
// similarity.rs:87-92 and data/sparse.rs:317-360): score[leaf] = sum over the
// keys shared by query and leaf, in ASCENDING KEY ORDER, of q_val * leaf_val,
// accumulated in f32 with separate multiply and add (Rust never fuses).  The
// rounding of that sequential sum is part of the result, so the
// kernel keeps the order: one warp owns the accumulators of one (query, tile)
// pair and walks the query's keys in ascending order; the lanes fan out over
// the posting list of ONE key at a time (a leaf occurs at most once per list),
// so no two lanes ever update the same accumulator and no atomics are needed.
#include <algorithm>

#include "index.cuh"
#include "kmer_count.cuh"
#include "score.cuh"

namespace myrio {

struct ScoreArgs {
    const TileInfo *tiles;
    const uint64_t *postings;
    const uint64_t *q_row_off;
    const uint64_t *q_keys;
    const float *q_vals;
    uint64_t q_first;
    uint32_t q_count;
    uint32_t tile_leaves;
    unsigned long long n_tasks;
    float *scores;  // [q_count][ld]            (full-row mode)
    uint64_t ld;
    unsigned long long *counters;  // [0] task counter, [1..3] stats
    // fused top-x mode: every (query, tile) task leaves its candidates for the global top-x
    uint32_t x;
    uint32_t n_tiles;
    uint32_t leaf_id_base;
    uint64_t *cand;      // composites [q_count][n_tiles][x]
    uint32_t *cand_cnt;  // [q_count][n_tiles]
    uint32_t *gthr;      // [q_count] running lower bound (score bits) of the query's x-th best score
};

__device__ __forceinline__ float finite_or_zero(float x) {
    return (fabsf(x) <= 3.402823466e+38f) ? x : 0.0f;  // false for NaN and inf
}

template <int SIM>
__device__ __forceinline__ float sim_transform(float dot) {
    if (SIM == MYRIO_SIM_JACARD_TANIMOTO)  // similarity.rs:143-149 (pre-normalised)
        return finite_or_zero(__fdiv_rn(dot, __fsub_rn(2.0f, dot)));
    return dot;  // similarity.rs:87-92
}

// larger composite = better rank: score bits (scores are >= +0) then smaller id first
__device__ __forceinline__ uint64_t topx_composite(uint32_t score_bits, uint32_t id) {
    return ((uint64_t)score_bits << 32) | (uint64_t)(0xFFFFFFFFu - id);
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ uint32_t warp_incl_scan_u32s(uint32_t v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// K5, fused into K4: with one tile's scores still in shared memory the warp keeps the
// tile's candidates for the query's global top-x (tax/results.rs:310-329).  A running
// per-query lower bound on the x-th best score (max over finished tiles of the tile's
// own x-th best) prunes later tiles: most tasks take the "few survivors" path (one
// pass); only the first tiles of a query run the 4 x 8-bit warp radix select.
// Pruning never drops a member of the global top-x: an element below the bound has
// at least x better elements in the tile that produced the bound.
template <int SIM>
__device__ __forceinline__ void tile_topx(float *acc, uint32_t *hist, uint32_t nl, uint32_t x,
                                          uint32_t gid_base, uint32_t *gthr_q, uint64_t *cand,
                                          uint32_t *cand_cnt, unsigned lane) {
    uint32_t thr = 0;
    if (lane == 0) thr = *reinterpret_cast<volatile uint32_t *>(gthr_q);
    thr = __shfl_sync(0xffffffffu, thr, 0);
    uint32_t cnt = 0;
    for (uint32_t i = lane; i < nl; i += 32) {
        const float s = sim_transform<SIM>(acc[i]);
        acc[i] = s;
        cnt += (__float_as_uint(s) >= thr) ? 1u : 0u;
    }
    cnt = warp_sum_u32(cnt);
    __syncwarp();
    const bool select = cnt > x;
    uint32_t T = thr, need_eq = 0;
    if (select) {
        // x-th largest score among the survivors: radix select on the f32 bit pattern
        uint32_t prefix = 0, need = x;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
#pragma unroll
            for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0;
            __syncwarp();
            for (uint32_t base = 0; base < nl; base += 32) {
                const uint32_t i = base + lane;
                uint32_t bin = 0xFFFFFFFFu;
                if (i < nl) {
                    const uint32_t b = __float_as_uint(acc[i]);
                    if (b >= thr && (pass == 0 || (b >> (shift + 8)) == prefix)) bin = (b >> shift) & 255u;
                }
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (bin != 0xFFFFFFFFu && (int)lane == __ffs(peers) - 1) hist[bin] += __popc(peers);
                __syncwarp();
            }
            // lane l owns bins 255-8l .. 248-8l (descending); find where the running count reaches need
            uint32_t local[8], lsum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                local[j] = hist[255 - 8 * lane - j];
                lsum += local[j];
            }
            const uint32_t incl = warp_incl_scan_u32s(lsum, lane);
            const unsigned reach = __ballot_sync(0xffffffffu, incl >= need);
            const int owner = __ffs(reach) - 1;
            uint32_t bin = 0, rest = 0;
            if ((int)lane == owner) {
                uint32_t run = incl - lsum;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (run + local[j] >= need) {
                        bin = 255 - 8 * lane - j;
                        rest = need - run;
                        break;
                    }
                    run += local[j];
                }
            }
            bin = __shfl_sync(0xffffffffu, bin, owner);
            need = __shfl_sync(0xffffffffu, rest, owner);
            prefix = (prefix << 8) | bin;
            __syncwarp();
        }
        T = prefix;
        need_eq = need;  // how many elements equal to T are taken (smallest ids first)
        if (lane == 0 && T > thr) atomicMax(gthr_q, T);
    }
    uint32_t out = 0, eq_taken = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < nl; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < nl;
        const uint32_t b = valid ? __float_as_uint(acc[i]) : 0u;
        bool take;
        if (select) {
            const bool eq = valid && b == T;
            const unsigned m_eq = __ballot_sync(0xffffffffu, eq);
            take = (valid && b > T) || (eq && eq_taken + __popc(m_eq & lt) < need_eq);
            eq_taken += __popc(m_eq);
        } else {
            take = valid && b >= thr;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (take) cand[out + __popc(m & lt)] = topx_composite(b, gid_base + i);
        out += __popc(m);
    }
    if (lane == 0) *cand_cnt = out;
}

// Postings fetched per lane and chunk: a chunk is up to 32*SC_U postings of ONE key
// (distinct leaves => SC_U*32 independent accumulator updates => full ILP), and two
// chunks are kept in flight (registers) so that the L2/HBM latency of chunk n+1
// hides behind the shared-memory updates of chunk n, across key boundaries too.
static constexpr int SC_U = 8;

template <int SIM, bool FUSED>
__global__ void __launch_bounds__(512) score_tiles_kernel(ScoreArgs a) {
    extern __shared__ float acc_all[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: tile accumulators + dummy slot for masked-off lanes (+ 256-bin histogram)
    const uint32_t acc_stride = a.tile_leaves + 32 + (FUSED ? 256 : 0);
    float *acc = acc_all + (size_t)warp * acc_stride;
    uint32_t *hist = reinterpret_cast<uint32_t *>(acc + a.tile_leaves + 32);
    const uint64_t dummy_posting = (uint64_t)a.tile_leaves << 32;  // leaf = dummy slot, val = +0.0
    unsigned long long st_keys = 0, st_hits = 0, st_post = 0;

    for (;;) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(a.counters, 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const uint32_t t = (uint32_t)(task / a.q_count);  // tile-major: a tile's postings stay hot in L2
        const uint32_t ql = (uint32_t)(task % a.q_count);
        const TileInfo *ti = a.tiles + t;
        const uint32_t nl = ti->n_leaves;
        const uint32_t mask = ti->dict_mask;
        const uint4 *dict = reinterpret_cast<const uint4 *>(ti->dict);
        const uint64_t *post = a.postings + ti->post_begin;

        for (uint32_t i = lane; i < nl; i += 32) acc[i] = 0.0f;
        if (lane == 0) acc[a.tile_leaves] = 0.0f;
        __syncwarp();

        const uint64_t qb = a.q_row_off[a.q_first + ql], qe = a.q_row_off[a.q_first + ql + 1];
        st_keys += (lane == 0) ? (qe - qb) : 0;

        // software-pipelined dictionary probe: the first slot of batch n+1 is in flight
        // while the posting lists of batch n are processed
        uint64_t key = 0, nkey = 0;
        float qv = 0.0f, nqv = 0.0f;
        uint32_t h = 0, nh = 0;
        uint4 slot = make_uint4(0, 0, 0, 0), nslot = make_uint4(0, 0, 0, 0);
        if (qb + lane < qe) {
            key = a.q_keys[qb + lane];
            qv = a.q_vals[qb + lane];
            h = dict_hash(key) & mask;
            slot = __ldg(dict + h);
        }
        for (uint64_t base = qb; base < qe; base += 32) {
            const bool valid = base + lane < qe;
            {  // prefetch the next batch of query keys and their first dictionary slot
                const uint64_t nidx = base + 32 + lane;
                nslot = make_uint4(0, 0, 0, 0);
                if (nidx < qe) {
                    nkey = a.q_keys[nidx];
                    nqv = a.q_vals[nidx];
                    nh = dict_hash(nkey) & mask;
                    nslot = __ldg(dict + nh);
                }
            }
            uint32_t begin = 0, len = 0;
            if (valid) {
                for (;;) {
                    if (slot.w == 0u) break;  // empty slot: key absent from this tile
                    if ((((uint64_t)slot.y << 32) | slot.x) == key) {
                        begin = slot.z;
                        len = slot.w;
                        break;
                    }
                    h = (h + 1) & mask;
                    slot = __ldg(dict + h);
                }
            }
            unsigned hits = __ballot_sync(0xffffffffu, len != 0u);
            st_hits += (lane == 0) ? __popc(hits) : 0;

            // cursor over the hit lists, in lane (= ascending key) order
            int src = -1;
            uint32_t cb = 0, cl = 0, coff = 0;
            float cv = 0.0f;
            auto next_hit = [&]() {
                if (!hits) {
                    src = -1;
                    return;
                }
                src = __ffs(hits) - 1;
                hits &= hits - 1;
                cb = __shfl_sync(0xffffffffu, begin, src);
                cl = __shfl_sync(0xffffffffu, len, src);
                cv = __shfl_sync(0xffffffffu, qv, src);
                coff = 0;
                st_post += (lane == 0) ? cl : 0;
            };
            auto fetch = [&](uint64_t(&p)[SC_U], uint32_t &n, float &v) -> bool {
                if (src < 0) return false;
                n = min(cl - coff, 32u * SC_U);
                v = cv;
                const uint64_t *ptr = post + cb + coff + lane;
                if (n == 32u * SC_U) {  // full chunk: unpredicated loads
#pragma unroll
                    for (int u = 0; u < SC_U; ++u) p[u] = __ldg(ptr + u * 32);
                } else {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        p[u] = (u * 32u + lane < n) ? __ldg(ptr + u * 32) : dummy_posting;
                }
                coff += n;
                if (coff >= cl) next_hit();
                return true;
            };
            auto process = [&](const uint64_t(&p)[SC_U], uint32_t n, float v) {
                float tacc[SC_U];
                if (n == 32u * SC_U) {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u) tacc[u] = acc[(uint32_t)(p[u] >> 32)];
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        acc[(uint32_t)(p[u] >> 32)] =
                            __fadd_rn(tacc[u], __fmul_rn(v, __uint_as_float((uint32_t)p[u])));
                } else {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        if (u * 32u < n) tacc[u] = acc[(uint32_t)(p[u] >> 32)];
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        if (u * 32u < n)
                            acc[(uint32_t)(p[u] >> 32)] =
                                __fadd_rn(tacc[u], __fmul_rn(v, __uint_as_float((uint32_t)p[u])));
                }
                __syncwarp();  // the next chunk may belong to another key and hit the same leaves
            };
            next_hit();
            uint64_t pa[SC_U], pb[SC_U];
            uint32_t na = 0, nb = 0;
            float va = 0.0f, vb = 0.0f;
            bool have_a = fetch(pa, na, va);
            while (have_a) {
                const bool have_b = fetch(pb, nb, vb);
                process(pa, na, va);
                if (!have_b) break;
                have_a = fetch(pa, na, va);
                process(pb, nb, vb);
            }
            key = nkey;
            qv = nqv;
            h = nh;
            slot = nslot;
        }
        if (FUSED) {
            const size_t slot = (size_t)ql * a.n_tiles + t;
            tile_topx<SIM>(acc, hist, nl, a.x, a.leaf_id_base + ti->leaf_begin, a.gthr + ql,
                           a.cand + slot * a.x, a.cand_cnt + slot, lane);
        } else {
            float *out = a.scores + (size_t)ql * a.ld + ti->leaf_begin;
            for (uint32_t i = lane; i < nl; i += 32) out[i] = sim_transform<SIM>(acc[i]);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (st_keys) atomicAdd(a.counters + 1, st_keys);
        if (st_hits) atomicAdd(a.counters + 2, st_hits);
        if (st_post) atomicAdd(a.counters + 3, st_post);
    }
}

// ------------------------------------------------------------------ K5: top-x

static constexpr int TX_THREADS = 256;
static constexpr int TX_MAX_X = 1024;
static constexpr int TX_SORT_CAP = 4096;  // composites sorted at once per query (32 KB)

// descending bitonic sort of S (pow2) u64 in shared memory
__device__ __forceinline__ void bitonic_desc(uint64_t *k, uint32_t S) {
    for (uint32_t size = 2; size <= S; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (S >> 1); t += blockDim.x) {
                uint32_t lo = 2u * t - (t & (stride - 1u)), hi = lo + stride;
                bool desc = (lo & size) == 0u;
                uint64_t x = k[lo], y = k[hi];
                if ((x < y) == desc) {
                    k[lo] = y;
                    k[hi] = x;
                }
            }
        }
    __syncthreads();
}

// One CTA per query: merges the per-tile candidate lists into the final top-x,
// order (score desc, payload_id asc) == descending composite.
__global__ void __launch_bounds__(TX_THREADS) topx_cand_merge_kernel(const uint64_t *__restrict__ cand,
                                                                     const uint32_t *__restrict__ cand_cnt,
                                                                     uint32_t n_tiles, uint32_t x,
                                                                     float *__restrict__ out_s,
                                                                     uint32_t *__restrict__ out_i) {
    __shared__ uint64_t buf[TX_SORT_CAP];
    const size_t q = blockIdx.x;
    uint32_t filled = 0;  // buf[0..filled) = best so far (sorted) + appended candidates
    for (uint32_t t = 0; t <= n_tiles; ++t) {
        const uint32_t cnt = t < n_tiles ? cand_cnt[q * n_tiles + t] : 0u;
        if (t == n_tiles || filled + cnt > TX_SORT_CAP) {
            uint32_t S = 2;
            while (S < filled) S <<= 1;
            for (uint32_t j = filled + threadIdx.x; j < S; j += TX_THREADS) buf[j] = 0;
            bitonic_desc(buf, S);
            filled = min(filled, x);
        }
        if (t < n_tiles) {
            const uint64_t *src = cand + (q * n_tiles + t) * x;
            for (uint32_t j = threadIdx.x; j < cnt; j += TX_THREADS) buf[filled + j] = src[j];
            filled += cnt;
            __syncthreads();
        }
    }
    for (uint32_t j = threadIdx.x; j < x; j += TX_THREADS) {
        if (j < filled) {
            const uint64_t c = buf[j];
            out_s[q * x + j] = __uint_as_float((uint32_t)(c >> 32));
            out_i[q * x + j] = 0xFFFFFFFFu - (uint32_t)c;
        } else {
            out_s[q * x + j] = 0.0f;
            out_i[q * x + j] = 0xFFFFFFFFu;
        }
    }
}

// ------------------------------------------------------- K7: merge of shards

// in: [n_parts][q][x] ; ids are global payload ids, padding = (0.0f, 0xFFFFFFFF)
__global__ void __launch_bounds__(256) topx_merge_kernel(const float *__restrict__ in_s,
                                                         const uint32_t *__restrict__ in_i,
                                                         uint32_t n_parts, uint64_t q_count, uint32_t x,
                                                         float *__restrict__ out_s,
                                                         uint32_t *__restrict__ out_i) {
    extern __shared__ uint64_t mk[];
    const uint64_t q = blockIdx.x;
    const uint32_t total = n_parts * x;
    uint32_t S = 2;
    while (S < total) S <<= 1;
    for (uint32_t j = threadIdx.x; j < S; j += blockDim.x) {
        uint64_t c = 0;
        if (j < total) {
            const uint32_t p = j / x, e = j % x;
            const size_t src = ((size_t)p * q_count + q) * x + e;
            const uint32_t id = in_i[src];
            // padding (id 0xFFFFFFFF, score 0) maps to composite 0 = worst
            c = ((uint64_t)__float_as_uint(in_s[src]) << 32) | (uint64_t)(0xFFFFFFFFu - id);
        }
        mk[j] = c;
    }
    bitonic_desc(mk, S);
    for (uint32_t j = threadIdx.x; j < x; j += blockDim.x) {
        const uint64_t c = mk[j];
        out_s[q * x + j] = __uint_as_float((uint32_t)(c >> 32));
        out_i[q * x + j] = 0xFFFFFFFFu - (uint32_t)c;
    }
}

// ------------------------------------------------------------------ host side

static void check_queries(const myrio_svecs *q, uint64_t q_first, uint64_t q_count) {
    if (q->vtype != MYRIO_VAL_F32) fail(MYRIO_ERR_BAD_ARG, "queries must be f32 (pre-normalised) vectors");
    if (q_first + q_count > q->n_rows) fail(MYRIO_ERR_BAD_ARG, "query range out of bounds");
}

static void check_sim(int32_t sim) {
    if (sim != MYRIO_SIM_COSINE && sim != MYRIO_SIM_JACARD_TANIMOTO)
        fail(MYRIO_ERR_UNSUPPORTED, "similarity %d is not implemented on the GPU path (Cosine, JacardTanimoto are)", sim);
}

template <bool FUSED>
static void launch_score(myrio_ctx *ctx, const myrio_index *ix, ScoreArgs &a, int32_t sim) {
    const size_t per_warp = ((size_t)ix->tile_leaves + 32 + (FUSED ? 256 : 0)) * sizeof(float);
    int warps = (int)std::min<size_t>(16, (ctx->smem_optin - 1024) / per_warp);
    if (warps < 1) fail(MYRIO_ERR_BAD_ARG, "tile_leaves=%u does not fit in shared memory", ix->tile_leaves);
    const size_t smem = per_warp * warps;
    const unsigned grid = (unsigned)std::min<unsigned long long>(a.n_tasks, (unsigned long long)ctx->sm_count);
    const char *name = FUSED ? "k4_score_tiles_topx" : "k4_score_tiles_rows";
    if (sim == MYRIO_SIM_COSINE) {
        auto kern = score_tiles_kernel<MYRIO_SIM_COSINE, FUSED>;
        MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MYRIO_LAUNCH(ctx, name, kern, grid, warps * 32, smem, a);
    } else {
        auto kern = score_tiles_kernel<MYRIO_SIM_JACARD_TANIMOTO, FUSED>;
        MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MYRIO_LAUNCH(ctx, name, kern, grid, warps * 32, smem, a);
    }
}

static ScoreArgs base_args(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q0,
                           uint64_t qn) {
    ScoreArgs a{};
    a.tiles = ix->d_tiles.p;
    a.postings = ix->postings.p;
    a.q_row_off = queries->row_off.p;
    a.q_keys = queries->keys.p;
    a.q_vals = queries->vals_f32.p;
    a.q_first = q0;
    a.q_count = (uint32_t)qn;
    a.tile_leaves = ix->tile_leaves;
    a.n_tiles = (uint32_t)ix->tiles.size();
    a.n_tasks = (unsigned long long)qn * ix->tiles.size();
    a.ld = ix->n_leaves;
    a.leaf_id_base = ix->leaf_id_base;
    a.counters = ctx->score_counters.p;
    return a;
}

static void reset_stats(myrio_ctx *ctx) {
    ctx->last_stats[0] = ctx->last_stats[1] = ctx->last_stats[2] = 0;
    ctx->score_counters.reserve(4);
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
}

static void accumulate_stats(myrio_ctx *ctx) {
    unsigned long long h[4] = {0, 0, 0, 0};
    d2h(ctx, h, ctx->score_counters.p, 4);
    sync(ctx);
    ctx->last_stats[0] += h[1];
    ctx->last_stats[1] += h[2];
    ctx->last_stats[2] += h[3];
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
}

void score_all(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
               uint64_t q_count, int32_t sim, float *scores_host) {
    check_queries(queries, q_first, q_count);
    check_sim(sim);
    const uint64_t L = ix->n_leaves;
    reset_stats(ctx);
    if (q_count == 0 || L == 0) return;
    const uint64_t budget = 1ull << 31;  // bytes of score rows kept on the device at once
    const uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, budget / (L * 4)), q_count);
    ctx->score_rows.reserve(per * L);
    for (uint64_t q0 = 0; q0 < q_count; q0 += per) {
        const uint64_t qn = std::min(per, q_count - q0);
        MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        ScoreArgs a = base_args(ctx, ix, queries, q_first + q0, qn);
        a.scores = ctx->score_rows.p;
        launch_score<false>(ctx, ix, a, sim);
        d2h(ctx, scores_host + q0 * L, ctx->score_rows.p, qn * L);
        accumulate_stats(ctx);
    }
    sync(ctx);
}

void score_topx_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                       uint64_t q_first, uint64_t q_count, uint32_t x, int32_t sim, float *d_top_scores,
                       uint32_t *d_top_ids, bool collect_stats) {
    check_queries(queries, q_first, q_count);
    check_sim(sim);
    if (x == 0 || x > TX_MAX_X) fail(MYRIO_ERR_BAD_ARG, "x must be in 1..%d", TX_MAX_X);
    reset_stats(ctx);
    if (q_count == 0) return;
    const uint64_t n_tiles = std::max<uint64_t>(ix->tiles.size(), 1);
    const uint64_t budget = 4ull << 30;  // bytes of per-tile candidates kept at once
    const uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, budget / (n_tiles * x * 8)), q_count);
    DevBuf<uint64_t> cand(per * n_tiles * x);
    DevBuf<uint32_t> cand_cnt(per * n_tiles), gthr(per);
    for (uint64_t q0 = 0; q0 < q_count; q0 += per) {
        const uint64_t qn = std::min(per, q_count - q0);
        MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(gthr.p, 0, qn * sizeof(uint32_t), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(cand_cnt.p, 0, qn * n_tiles * sizeof(uint32_t), ctx->stream));
        if (!ix->tiles.empty()) {
            ScoreArgs a = base_args(ctx, ix, queries, q_first + q0, qn);
            a.x = x;
            a.cand = cand.p;
            a.cand_cnt = cand_cnt.p;
            a.gthr = gthr.p;
            launch_score<true>(ctx, ix, a, sim);
        }
        MYRIO_LAUNCH(ctx, "k5_topx_merge_tiles", topx_cand_merge_kernel, (unsigned)qn, TX_THREADS, 0, cand.p,
                     cand_cnt.p, (uint32_t)ix->tiles.size(), x, d_top_scores + q0 * x, d_top_ids + q0 * x);
        if (collect_stats) accumulate_stats(ctx);
    }
}

void topx_merge_device(myrio_ctx *ctx, const float *d_scores_in, const uint32_t *d_ids_in, uint32_t n_parts,
                       uint64_t q_count, uint32_t x, float *d_scores_out, uint32_t *d_ids_out) {
    if (x == 0 || n_parts == 0) fail(MYRIO_ERR_BAD_ARG, "x and n_parts must be > 0");
    uint64_t total = (uint64_t)n_parts * x, S = 2;
    while (S < total) S <<= 1;
    const size_t smem = S * sizeof(uint64_t);
    if (smem > ctx->smem_optin) fail(MYRIO_ERR_UNSUPPORTED, "n_parts*x=%llu too large to merge", (unsigned long long)total);
    if (q_count == 0) return;
    MYRIO_CK(cudaFuncSetAttribute(topx_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, "k7_topx_merge", topx_merge_kernel, (unsigned)q_count, 256, smem, d_scores_in, d_ids_in,
                 n_parts, q_count, x, d_scores_out, d_ids_out);
}

}  // namespace myrio
