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
// That is score_tiles_kernel (round 1).  The default since round 2 is COUNT MODE
// (score_tiles_count_kernel): nearly all terms of a (query, leaf) pair are the same
// f32 product, equal addends commute, so they are counted (bitmaps into bit-sliced
// counters, sparse postings into u16 counters) and added n times at the end -- only
// for the leaves whose upper bound can still reach the query's top-x -- while the
// rare odd terms keep their place in key order.  File map: tile_topx (K5 fused),
// score_tiles_kernel, count mode, K5 / K7 merges, the plan kernels (K4a), host side.
#include <algorithm>

#include "index.cuh"
#include "kmer_count.cuh"
#include "merge.cuh"
#include "score.cuh"

namespace myrio {

struct __align__(4) HitRec {  // one key of the query that occurs in the tile, in key order
    uint32_t begin;
    uint32_t len;  // DENSE_FLAG | ORDERED_FLAG | df
    float qv;
};

struct ScoreArgs {
    const TileInfo *tiles;
    const uint64_t *postings;
    const HitRec *plan;         // hit lists of all (query, tile) tasks of the batch
    const uint64_t *plan_off;   // [q_count * plan_stride (+ 1)]: a task's hit list is [off[s], off[s + 1])
    uint32_t plan_stride;       // entries per query: n_tiles + 1 (plan_build_kernel) or n_tiles (contiguous lists)
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
    // seeding of the bound: pass 1 scores every query against its SEED tile only (the tile holding most of
    // its keys, i.e. its closest relatives), pass 2 sweeps all other tasks tile-major with the bound already high
    const float *qv1;      // [q_count] the query's common value (count mode): its smallest value
    const uint32_t *seed;  // [q_count] seed tile (pass 1); its plan_off entry carries PLAN_SEED_FLAG (pass 2 skips it)
    uint32_t seed_mode;    // 0 = no seeding, 1 = seed pass (task == query), 2 = sweep
    const uint32_t *seed_order;  // seed pass: task i scores query seed_order[i] (queries grouped by seed tile) or i
    uint32_t ub_prune;     // count mode, fused top-x: finish only the scores whose upper bound reaches the bound
    // count mode: a query's common dense keys per tile as bits (bit = bitmap index in the tile) instead of hit
    // records: row ql of mask_stride words, the tile's words at TileInfo::mask_off (0 = everything is a record)
    const uint32_t *dmask;
    uint32_t mask_stride;
};
static constexpr uint64_t PLAN_SEED_FLAG = 1ull << 63;

__device__ __forceinline__ uint32_t post_leaf(uint64_t p) { return (uint32_t)(p >> 32) & POSTING_LEAF_MASK; }

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
// second half: acc[] holds the tile's final (transformed) scores, `cnt` of them are >= thr
__device__ __forceinline__ void tile_topx_emit(float *acc, uint32_t *hist, uint32_t nl, uint32_t x,
                                               uint32_t gid_base, uint32_t *gthr_q, uint64_t *cand,
                                               uint32_t *cand_cnt, unsigned lane, uint32_t thr, uint32_t cnt);

__device__ __forceinline__ uint32_t read_bound(const uint32_t *gthr_q, unsigned lane) {
    uint32_t thr = 0;
    if (lane == 0) thr = *reinterpret_cast<const volatile uint32_t *>(gthr_q);
    return __shfl_sync(0xffffffffu, thr, 0);
}

template <int SIM>
__device__ __forceinline__ void tile_topx(float *acc, uint32_t *hist, uint32_t nl, uint32_t x,
                                          uint32_t gid_base, uint32_t *gthr_q, uint64_t *cand,
                                          uint32_t *cand_cnt, unsigned lane) {
    const uint32_t thr = read_bound(gthr_q, lane);
    uint32_t cnt = 0;
    for (uint32_t i = lane; i < nl; i += 32) {
        const float s = sim_transform<SIM>(acc[i]);
        acc[i] = s;
        cnt += (__float_as_uint(s) >= thr) ? 1u : 0u;
    }
    cnt = warp_sum_u32(cnt);
    __syncwarp();
    tile_topx_emit(acc, hist, nl, x, gid_base, gthr_q, cand, cand_cnt, lane, thr, cnt);
}

__device__ __forceinline__ void tile_topx_emit(float *acc, uint32_t *hist, uint32_t nl, uint32_t x,
                                               uint32_t gid_base, uint32_t *gthr_q, uint64_t *cand,
                                               uint32_t *cand_cnt, unsigned lane, uint32_t thr, uint32_t cnt) {
    if (cnt == 0) {  // nothing in this tile can enter the query's top-x
        if (lane == 0) *cand_cnt = 0;
        return;
    }
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

// K4.  One warp owns one (query, tile) task.
//
// Phase A -- the task's hit list (the keys of the query that occur in the tile, IN KEY ORDER,
// with where their postings are) was laid out by the plan kernels; it is staged in shared
// memory.  Phase B -- accumulate: the hit list is walked in order.  A sparse key is streamed in
// chunks of up to 32*SC_U postings (distinct leaves => independent updates, full ILP);
// a dense key is one 128-byte bitmap applied to lane-owned accumulators that stay in
// registers while dense keys follow one another.  Two chunks are kept in flight so the
// load of chunk n+1 (possibly of the next key) hides behind the updates of chunk n.
static constexpr int SC_U = 2;
static constexpr int SC_MAX_WARPS = 20;  // 96 registers per thread (racc[32] in registers), ~11 KB of smem per warp
static constexpr uint32_t SC_HITS = 160;  // hit-list capacity per warp (drained when nearly full)

__host__ __device__ inline size_t score_warp_smem_bytes(uint32_t tile_leaves, bool fused) {
    return ((size_t)tile_leaves + 32 + (fused ? 256 : 0) + DENSE_MAX_TILE) * sizeof(float) +
           SC_HITS * sizeof(HitRec);
}

// STATS: count the postings touched (bench / tests; the timed path leaves it out).
// UREG: the unit values of the lane-owned leaves live in registers (3 instead of 4 instructions per
// dense slot, 128 registers -> 16 warps) instead of shared memory (96 registers -> 20 warps).
static constexpr int SC_UREG_WARPS = 16;
template <int SIM, bool FUSED, bool STATS, bool UREG>
__global__ void __launch_bounds__((UREG ? SC_UREG_WARPS : SC_MAX_WARPS) * 32) score_tiles_kernel(ScoreArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: tile accumulators + dummy slot for masked-off lanes (+ 256-bin histogram) + hit list
    uint8_t *wbase = smem_raw + (size_t)warp * score_warp_smem_bytes(a.tile_leaves, FUSED);
    float *acc = reinterpret_cast<float *>(wbase);
    uint32_t *hist = reinterpret_cast<uint32_t *>(acc + a.tile_leaves + 32);
    float *usm = acc + a.tile_leaves + 32 + (FUSED ? 256 : 0);  // unit values of the lane-owned leaves
    HitRec *hits = reinterpret_cast<HitRec *>(usm + DENSE_MAX_TILE);
    const uint64_t dummy_posting = (uint64_t)a.tile_leaves << 32;  // leaf = dummy slot, val = +0.0
    unsigned long long st_post = 0;
    uint32_t cached_tile = 0xFFFFFFFFu;  // tile whose unit values sit in usm (and whose TileInfo is in registers)
    uint32_t nl = 0;
    const uint64_t *post = nullptr;
    const uint32_t *bitmaps = nullptr;
    const TileInfo *ti = a.tiles;
    bool dense_tile = false;

    // task ids are handed out by an atomic counter; the id of the NEXT task is requested at the top of
    // the current one, so its round trip to L2 is off the critical path (a batch has < 2^32 tasks)
    uint32_t next_task = 0;
    if (lane == 0) next_task = (uint32_t)atomicAdd(a.counters, 1ull);
    for (;;) {
        const uint32_t task = __shfl_sync(0xffffffffu, next_task, 0);
        if (task >= a.n_tasks) break;
        if (lane == 0) next_task = (uint32_t)atomicAdd(a.counters, 1ull);
        uint32_t t, ql;
        if (a.seed_mode == 1) {
            ql = a.seed_order ? __ldg(a.seed_order + task) : task;
            t = __ldg(a.seed + ql);
        } else {
            t = task / a.q_count;  // tile-major: a tile's index stays hot in L2
            ql = task - t * a.q_count;
        }
        // the task's slice of the plan (requested first: everything below overlaps its latency)
        const size_t slot = (size_t)ql * a.n_tiles + t, pslot = (size_t)ql * a.plan_stride + t;
        const uint64_t hb_raw = a.plan_off[pslot];
        const uint64_t hb = hb_raw & ~PLAN_SEED_FLAG, he = a.plan_off[pslot + 1] & ~PLAN_SEED_FLAG;
        if (a.seed_mode == 2 && (hb_raw & PLAN_SEED_FLAG)) continue;  // scored by the seed pass
        if (t != cached_tile) {
            ti = a.tiles + t;
            nl = ti->n_leaves;
            post = a.postings + ti->post_begin;
            bitmaps = ti->bitmaps;
            dense_tile = ti->n_dense != 0;  // implies tile_leaves <= DENSE_MAX_TILE
        }

        // Dense keys (tiles of <= 1024 leaves): this lane owns leaves 32*r + lane.  Their
        // accumulators live in registers while dense keys follow one another (flushed to /
        // reloaded from shared memory around sparse chunks); their unit values sit in shared
        // memory (bank == lane, conflict-free) -- keeping them in registers too costs a quarter
        // of the resident warps and measured slower (profiles/r01_k4_variants.md).
        float ureg[UREG ? 32 : 1];
        // tile-major task order: a warp's consecutive tasks are mostly in the same tile, whose unit
        // values are still in its shared-memory slice
        const bool unit_cached = !UREG && t == cached_tile;
        cached_tile = t;
        if (dense_tile && !unit_cached) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float u = (32u * r + lane < nl) ? __ldg(ti->unit + 32 * r + lane) : 0.0f;
                if (UREG)
                    ureg[UREG ? r : 0] = u;
                else
                    usm[32 * r + lane] = u;
            }
        } else if (UREG) {
#pragma unroll
            for (int r = 0; r < 32; ++r) ureg[UREG ? r : 0] = 0.0f;
        }
        for (uint32_t i = lane; i < a.tile_leaves + 32; i += 32) acc[i] = 0.0f;  // incl. the dummy slots
        __syncwarp();

        // ---- phase B: walk hits[0..n_hits) in order.  The register accumulators are local to
        // the walk (dead during the probe phase, which then has the registers for deep MLP);
        // between walks the accumulators live in shared memory.
        auto process_hits = [&](uint32_t n_hits) {
            float racc[32];
            bool in_regs = false;
            auto flush_regs = [&]() {  // registers -> shared memory (slot 32*r + lane, bank == lane)
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    if (32u * r < a.tile_leaves) acc[32 * r + lane] = racc[r];
                in_regs = false;
                __syncwarp();
            };
            auto reload_regs = [&]() {
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 32; ++r) racc[r] = (32u * r < a.tile_leaves) ? acc[32 * r + lane] : 0.0f;
                in_regs = true;
            };
            uint32_t hi = 0;  // next hit to open
            uint32_t cb = 0, cl = 0, coff = 0;
            float cv = 0.0f;
            bool cdense = false, open = false;
            auto next_hit = [&]() {
                if (hi >= n_hits) {
                    open = false;
                    return;
                }
                const HitRec h = hits[hi++];  // same address for all lanes: broadcast
                cb = h.begin;
                cdense = (h.len & DENSE_FLAG) != 0u;
                cl = h.len & LEN_MASK;
                cv = h.qv;
                coff = 0;
                open = true;
                if (STATS) st_post += (lane == 0) ? cl : 0;
            };
            constexpr uint32_t DENSE_CHUNK = 0xFFFFFFFFu;  // chunk-size marker of a dense key
            auto fetch = [&](uint64_t(&p)[SC_U], uint32_t &n, float &v) -> bool {
                if (!open) return false;
                v = cv;
                if (cdense) {  // the whole key is one 128-byte bitmap: word `lane`
                    n = DENSE_CHUNK;
                    p[0] = __ldg(bitmaps + (size_t)cb * 32 + lane);
                    next_hit();
                    return true;
                }
                n = min(cl - coff, 32u * SC_U);
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
                if (n == DENSE_CHUNK) {
                    // lane-owned leaves in registers: no shared-memory traffic, no ordering hazard
                    // with other dense keys; branch-free select per slot
                    if (!in_regs) reload_regs();
                    const uint32_t w = (uint32_t)p[0];
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float s2 = __fadd_rn(racc[r], __fmul_rn(v, UREG ? ureg[UREG ? r : 0] : usm[32 * r + lane]));
                        racc[r] = (w & (1u << r)) ? s2 : racc[r];
                    }
                    return;
                }
                if (in_regs) flush_regs();
                float tacc[SC_U];
                if (n == 32u * SC_U) {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u) tacc[u] = acc[post_leaf(p[u])];
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        acc[post_leaf(p[u])] =
                            __fadd_rn(tacc[u], __fmul_rn(v, __uint_as_float((uint32_t)p[u])));
                } else {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        if (u * 32u < n) tacc[u] = acc[post_leaf(p[u])];
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        if (u * 32u < n)
                            acc[post_leaf(p[u])] =
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
            if (in_regs) flush_regs();
        };

        // ---- phase A: the task's hit list comes from the plan (plan_* kernels below): stage it in
        // shared memory SC_HITS records at a time (coalesced) and walk it.  ONE textual call of
        // process_hits keeps the lambda inlined (racc stays in registers).
        {
            for (uint64_t h0 = hb; h0 < he; h0 += SC_HITS) {
                const uint32_t n_hits = (uint32_t)min((uint64_t)SC_HITS, he - h0);
                for (uint32_t i = lane; i < n_hits; i += 32) hits[i] = a.plan[h0 + i];
                __syncwarp();
                process_hits(n_hits);
                __syncwarp();
            }
        }

        if (FUSED) {
            tile_topx<SIM>(acc, hist, nl, a.x, a.leaf_id_base + ti->leaf_begin, a.gthr + ql,
                           a.cand + slot * a.x, a.cand_cnt + slot, lane);
        } else {
            float *out = a.scores + (size_t)ql * a.ld + ti->leaf_begin;
            for (uint32_t i = lane; i < nl; i += 32) out[i] = sim_transform<SIM>(acc[i]);
        }
        __syncwarp();
    }
    if (STATS && lane == 0 && st_post) atomicAdd(a.counters + 3, st_post);
}


// ------------------------------------------------------------------ K4, count mode
//
// The same scores with about half the instructions.  For one (query, leaf) pair almost every shared key
// contributes THE SAME f32 product p = fl(v1 * unit[leaf]): the query's k-mers nearly all have count 1 (value
// v1 after normalisation) and the leaf's nearly all the common count (value unit[leaf]).  Equal addends commute,
// so the order of the reference's sequential sum only matters relative to the few ODD terms (a k-mer counted
// twice in the read or in the leaf).  The kernel therefore COUNTS the common terms instead of adding them --
// dense keys: one bitmap word per lane added into bit-sliced counters held in registers (12 logic instructions
// for 32 leaves instead of 128); sparse postings: a u16 increment in shared memory -- and replays them as
// n sequential additions of p per leaf at the end of the task:  score = fl(fl(fl(0 + p) + p) + ...).  An odd
// term (query value != v1, or a posting flagged POSTING_ODD by the index build) first replays the leaf's pending
// common terms, which are exactly the ones that precede it in key order, then is added itself: bit-identical to
// the walk of score_tiles_kernel and to the reference's merge-join, whatever the data.
static_assert(SC_U == 2, "scc_odd_sparse_chunk takes two postings");
static constexpr int SCC_PLANES = 6;        // bit-sliced counters: up to 63 dense keys between two flushes (7 planes: measured no faster)
static constexpr int SCC_MAX_WARPS = 24;

__host__ __device__ inline size_t score_count_warp_smem_bytes(uint32_t tile_leaves, bool fused) {
    (void)fused;  // the histogram's space doubles as the dense-key list of a split hit list: always there
    return ((size_t)tile_leaves + 32 + 256) * sizeof(float) +                            // acc + histogram
           (((size_t)tile_leaves + 32) * sizeof(uint16_t) + 15) / 16 * 16 +             // shared-key counters
           SC_HITS * sizeof(HitRec);
}

// The odd terms are rare: their code lives out of line so that the hot loop of the kernel stays small (the
// instruction cache is the first thing a 24-warp persistent kernel of this size runs out of: ncu `no_instruction`).
// An odd term for `leaf`: replay its pending common terms, then add the term (reference order).
__device__ __forceinline__ void scc_odd_update(float *acc, uint16_t *cnt, const float *unit, uint32_t nl, float v1,
                                               uint32_t leaf, float term) {
    const float p = __fmul_rn(v1, __ldg(unit + min(leaf, nl - 1)));
    float s = acc[leaf];
    for (uint32_t n = cnt[leaf]; n; --n) s = __fadd_rn(s, p);
    cnt[leaf] = 0;
    acc[leaf] = __fadd_rn(s, term);
}
// a sparse chunk that holds at least one odd posting (or belongs to a query key whose value is not v1)
__device__ __noinline__ void scc_odd_sparse_chunk(float *acc, uint16_t *cnt, const float *unit, uint32_t nl,
                                                  uint32_t tile_leaves, float v1, float v, uint64_t p0, uint64_t p1,
                                                  uint32_t n, bool odd_key) {
    const uint64_t p[SC_U] = {p0, p1};
#pragma unroll
    for (int u = 0; u < SC_U; ++u)
        if (u * 32u < n) {
            const uint32_t leaf = post_leaf(p[u]);
            const bool odd = (odd_key || (p[u] & POSTING_ODD) != 0ull) && leaf < tile_leaves;
            if (odd)
                scc_odd_update(acc, cnt, unit, nl, v1, leaf, __fmul_rn(v, __uint_as_float((uint32_t)p[u])));
            else
                cnt[leaf] = (uint16_t)(cnt[leaf] + 1u);
        }
}
// a dense key of a query value other than v1: every posting carries unit[leaf], the query value is the odd one
__device__ __noinline__ void scc_odd_dense_key(float *acc, uint16_t *cnt, const float *unit, uint32_t nl, float v1,
                                               float v, uint32_t w, uint32_t lane) {
    for (int r = 0; r < 32; ++r)
        if ((w >> r) & 1u) {
            const uint32_t leaf = 32u * r + lane;
            scc_odd_update(acc, cnt, unit, nl, v1, leaf, __fmul_rn(v, __ldg(unit + leaf)));
        }
}

template <int SIM, bool FUSED, bool STATS>
__global__ void __launch_bounds__(SCC_MAX_WARPS * 32) score_tiles_count_kernel(ScoreArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = smem_raw + (size_t)warp * score_count_warp_smem_bytes(a.tile_leaves, FUSED);
    float *acc = reinterpret_cast<float *>(wbase);
    uint32_t *hist = reinterpret_cast<uint32_t *>(acc + a.tile_leaves + 32);
    uint16_t *cnt = reinterpret_cast<uint16_t *>(acc + a.tile_leaves + 32 + 256);
    HitRec *hits = reinterpret_cast<HitRec *>(reinterpret_cast<uint8_t *>(cnt) +
                                              (((size_t)a.tile_leaves + 32) * sizeof(uint16_t) + 15) / 16 * 16);
    const uint64_t dummy_posting = (uint64_t)a.tile_leaves << 32;  // leaf = dummy slot, not odd
    unsigned long long st_post = 0;
    uint32_t cached_tile = 0xFFFFFFFFu;
    uint32_t nl = 0;
    const uint64_t *post = nullptr;
    const uint32_t *bitmaps = nullptr;
    const float *unit = nullptr;
    float unit_max = 0.0f;
    uint32_t t_mo = 0, t_mw = 0;  // the tile's words in a query's dense-key mask row
    bool acc_dirty = true;        // acc[] may hold something else than +0
    const TileInfo *ti = a.tiles;

    // ---- task pipeline.  Task ids come from an atomic counter.  The id of the task after next is requested in
    // the middle of the current task, where the next task's plan slice (two offsets) and common value are loaded
    // as well: all three round trips to L2 are off the critical path.
    const unsigned lt = (1u << lane) - 1u;
    constexpr uint32_t TASK_BATCH = 4;  // consecutive tasks (same tile, neighbouring queries) per atomic
    uint32_t next_task = 0;
    if (lane == 0) next_task = (uint32_t)atomicAdd(a.counters, (unsigned long long)TASK_BATCH);
    auto load_desc = [&](uint32_t tk, uint32_t &tt, uint32_t &qq, uint64_t &b, uint64_t &e, float &v) {
        if (tk >= a.n_tasks) return;
        if (a.seed_mode == 1) {
            qq = a.seed_order ? __ldg(a.seed_order + tk) : tk;  // grouped by seed tile: neighbours share the tile
            tt = __ldg(a.seed + qq);
        } else {
            tt = tk / a.q_count;  // tile-major: a tile's index stays hot in L2
            qq = tk - tt * a.q_count;
        }
        const size_t sl = (size_t)qq * a.plan_stride + tt;
        b = a.plan_off[sl];
        e = a.plan_off[sl + 1];
        v = __ldg(a.qv1 + qq);  // the query's common value
    };
    uint32_t task = __shfl_sync(0xffffffffu, next_task, 0);
    uint32_t task_end = task + TASK_BATCH;  // end of the batch `task` belongs to
    uint32_t t = 0, ql = 0;
    uint64_t hb_raw = 0, he_raw = 0;
    float v1 = 0.0f;
    load_desc(task, t, ql, hb_raw, he_raw, v1);
    if (lane == 0) next_task = (uint32_t)atomicAdd(a.counters, (unsigned long long)TASK_BATCH);
    while (task < a.n_tasks) {
        const size_t slot = (size_t)ql * a.n_tiles + t;
        const uint64_t hb = hb_raw & ~PLAN_SEED_FLAG, he = he_raw & ~PLAN_SEED_FLAG;
        const bool skip = a.seed_mode == 2 && (hb_raw & PLAN_SEED_FLAG);  // scored by the seed pass
        // bit-sliced counters of the lane-owned leaves 32*r + lane (bit r of plane j = bit j of the count)
        uint32_t pl[SCC_PLANES];
#pragma unroll
        for (int j = 0; j < SCC_PLANES; ++j) pl[j] = 0;
        uint32_t pending = 0;  // dense keys counted since the last flush (warp-uniform)
        bool had_ordered = false;  // some batch of the task took the ordered walk (partial sums may be non-zero)
        uint32_t mask0 = 0;
        bool any_hits = he > hb;  // hit records, or (below) bits in the dense-key mask
        if (!skip) {
        if (t != cached_tile) {
            ti = a.tiles + t;
            nl = ti->n_leaves;
            post = a.postings + ti->post_begin;
            bitmaps = ti->bitmaps;
            unit = ti->unit;
            unit_max = ti->unit_max;
            t_mo = ti->mask_off;
            t_mw = a.mask_stride ? ti->mask_words : 0u;
            cached_tile = t;
        }
        // first words of the query's dense-key mask for this tile (requested before the accumulators are cleared)
        if (t_mw && lane < 8 && lane < t_mw) mask0 = __ldg(a.dmask + (size_t)ql * a.mask_stride + t_mo + lane);
        // the counters are cleared for every task; the partial sums only if the previous task touched them (the
        // usual task -- no ordered batch, no survivor of the bound pass -- never does)
        if ((a.tile_leaves & 7u) == 0u) {  // 16-byte stores (the per-warp slices are 16-byte aligned)
            const uint32_t n4 = (a.tile_leaves + 32) >> 2, n8 = (a.tile_leaves + 32) >> 3;
            if (acc_dirty)
                for (uint32_t i = lane; i < n4; i += 32) reinterpret_cast<float4 *>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t i = lane; i < n8; i += 32) reinterpret_cast<uint4 *>(cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            for (uint32_t i = lane; i < a.tile_leaves + 32; i += 32) {
                if (acc_dirty) acc[i] = 0.0f;
                cnt[i] = 0;
            }
        }
        acc_dirty = true;  // until the task proves otherwise
        __syncwarp();
        }

        auto flush_planes = [&]() {
            if (pending == 0) return;
            // rolled on purpose: the flush is inlined at several (mostly cold) sites and the kernel is
            // instruction-cache sensitive; with the bound pass of the fused top-x it rarely runs at all
#pragma unroll 2
            for (int r = 0; r < 32; ++r) {
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < SCC_PLANES; ++j) c |= ((pl[j] >> r) & 1u) << j;
                if (32u * r < a.tile_leaves && c) cnt[32 * r + lane] = (uint16_t)(cnt[32 * r + lane] + c);
            }
#pragma unroll
            for (int j = 0; j < SCC_PLANES; ++j) pl[j] = 0;
            pending = 0;
            __syncwarp();
        };
        auto process_hits = [&](uint32_t n_hits) {
            uint32_t hi = 0;
            uint32_t cb = 0, cl = 0, coff = 0;
            float cv = 0.0f;
            bool cdense = false, open = false;
            auto next_hit = [&]() {
                if (hi >= n_hits) {
                    open = false;
                    return;
                }
                const HitRec h = hits[hi++];
                cb = h.begin;
                cdense = (h.len & DENSE_FLAG) != 0u;
                cl = h.len & LEN_MASK;
                cv = h.qv;
                coff = 0;
                open = true;
                if (STATS) st_post += (lane == 0) ? cl : 0;
            };
            constexpr uint32_t DENSE_CHUNK = 0xFFFFFFFFu;
            auto fetch = [&](uint64_t(&p)[SC_U], uint32_t &n, float &v) -> bool {
                if (!open) return false;
                v = cv;
                if (cdense) {
                    n = DENSE_CHUNK;
                    p[0] = __ldg(bitmaps + (size_t)cb * 32 + lane);
                    next_hit();
                    return true;
                }
                n = min(cl - coff, 32u * SC_U);
                const uint64_t *ptr = post + cb + coff + lane;
                if (n == 32u * SC_U) {
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
                const bool odd_key = v != v1;  // warp-uniform
                if (n == DENSE_CHUNK) {
                    const uint32_t w = (uint32_t)p[0];
                    if (!odd_key) {
                        // ripple-carry add of one bit per leaf into the bit-sliced counters (an earlier batch of
                        // the task may have left them full: the dense sweep stops at 63 as well)
                        if (pending >= (1u << SCC_PLANES) - 1u) flush_planes();
                        uint32_t carry = w;
#pragma unroll
                        for (int j = 0; j < SCC_PLANES; ++j) {
                            const uint32_t t2 = pl[j] & carry;
                            pl[j] ^= carry;
                            carry = t2;
                        }
                        ++pending;
                    } else {
                        // every posting of a dense key carries unit[leaf]; the query value is the odd one
                        flush_planes();
                        scc_odd_dense_key(acc, cnt, unit, nl, v1, v, w, lane);
                        __syncwarp();
                    }
                    return;
                }
                bool any_odd = odd_key && n != 0u;
#pragma unroll
                for (int u = 0; u < SC_U; ++u) any_odd |= (p[u] & POSTING_ODD) != 0ull;  // dummy postings carry no flag
                if (!__any_sync(0xffffffffu, any_odd)) {
#pragma unroll
                    for (int u = 0; u < SC_U; ++u)
                        if (u * 32u < n) {
                            const uint32_t leaf = post_leaf(p[u]);
                            cnt[leaf] = (uint16_t)(cnt[leaf] + 1u);
                        }
                } else {
                    flush_planes();
                    scc_odd_sparse_chunk(acc, cnt, unit, nl, a.tile_leaves, v1, v, p[0], p[1], n, odd_key);
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
        };

        uint32_t *dlist = hist;  // dense-key list of a split hit list / of a mask chunk (the radix-select
                                  // histogram's space, idle until the top-x)
        // dense keys dlist[0, nd): eight bitmap words at a time go through a carry-save adder tree into the
        // bit-sliced counters (a full adder is two LOP3: 20 logic instructions per eight words instead of 96 for
        // eight ripple-carry additions); the next eight words are requested before the current ones are added
        auto dense_sweep = [&](uint32_t nd) {
            uint32_t w[8], wn[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const uint32_t *bm_lane = bitmaps + lane;
            auto load8 = [&](uint32_t(&x)[8], uint32_t b0) {
                uint32_t idx[8];
                if ((a.tile_leaves & 3u) == 0u) {  // the list is 16-byte aligned, b0 a multiple of 8
                    const uint4 i0 = *reinterpret_cast<const uint4 *>(dlist + b0);
                    const uint4 i1 = *reinterpret_cast<const uint4 *>(dlist + b0 + 4);
                    idx[0] = i0.x, idx[1] = i0.y, idx[2] = i0.z, idx[3] = i0.w;
                    idx[4] = i1.x, idx[5] = i1.y, idx[6] = i1.z, idx[7] = i1.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) idx[u] = dlist[b0 + u];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = b0 + u < nd ? __ldg(bm_lane + idx[u] * 32u) : 0u;
            };
            // full adder on 32 independent bit positions
            auto csa = [](uint32_t x, uint32_t y, uint32_t z, uint32_t &sum, uint32_t &car) {
                sum = x ^ y ^ z;
                car = (x & y) | (z & (x | y));
            };
            load8(w, 0);
            for (uint32_t d0 = 0; d0 < nd; d0 += 8) {
                if (pending > (1u << SCC_PLANES) - 9u) flush_planes();
                if (d0 + 8 < nd) load8(wn, d0 + 8);
                if (STATS) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) st_post += __popc(w[u]);  // summed over the lanes at the end: df
                }
                uint32_t s1, c1, s2, c2, s3, c3, c4, c5, c6, c7, t0;
                csa(w[0], w[1], w[2], s1, c1);          // weight 1
                csa(w[3], w[4], w[5], s2, c2);
                csa(w[6], w[7], pl[0], s3, c3);
                csa(s1, s2, s3, pl[0], c4);
                csa(c1, c2, c3, t0, c5);                // weight 2
                csa(t0, c4, pl[1], pl[1], c6);
                csa(c5, c6, pl[2], pl[2], c7);          // weight 4
#pragma unroll
                for (int j = 3; j < SCC_PLANES; ++j) {  // weight 8 and up: ripple the single carry
                    const uint32_t t2 = pl[j] & c7;
                    pl[j] ^= c7;
                    c7 = t2;
                }
                pending += min(8u, nd - d0);
#pragma unroll
                for (int u = 0; u < 8; ++u) w[u] = wn[u];
            }
        };

        if (!skip && t_mw) {
            // the query's common dense keys in this tile: bits of its mask row (plan_build_kernel), 8 words
            // (256 keys) at a time expanded to bitmap indexes
            const uint32_t *mrow = a.dmask + (size_t)ql * a.mask_stride + t_mo;
            for (uint32_t w0 = 0; w0 < t_mw; w0 += 8) {
                uint32_t m = w0 == 0 ? mask0 : ((lane < 8 && w0 + lane < t_mw) ? __ldg(mrow + w0 + lane) : 0u);
                const uint32_t nb = __popc(m);
                const uint32_t incl = warp_incl_scan_u32s(nb, lane);
                const uint32_t nd = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t pos = incl - nb;
                any_hits |= nd != 0u;
                while (m) {
                    const uint32_t b = __ffs(m) - 1;
                    m &= m - 1;
                    dlist[pos++] = (w0 + lane) * 32 + b;
                }
                __syncwarp();
                dense_sweep(nd);
                __syncwarp();
            }
        }

        if (!skip)
        for (uint64_t h0 = hb; h0 < he; h0 += SC_HITS) {
            const uint32_t n_hits = (uint32_t)min((uint64_t)SC_HITS, he - h0);
            // Staging.  Common terms commute, so a hit list without an ORDERED hit (nearly all of them) is split on
            // the way in: the bitmap indexes of its dense keys go to dlist[] (the radix-select histogram's space,
            // idle until the top-x), its sparse keys are compacted to the front of the hit buffer as (begin, len).
            uint2 *shits = reinterpret_cast<uint2 *>(hits);
            uint32_t nd = 0, ns = 0;
            bool ordered = false;
            unsigned long long tsum = 0;
            for (uint32_t base = 0; base < n_hits; base += 32) {
                const uint32_t i = base + lane;
                HitRec h;
                h.begin = 0;
                h.len = 0;
                if (i < n_hits) h = a.plan[h0 + i];
                const bool is_d = (h.len & DENSE_FLAG) != 0u, is_s = !is_d && i < n_hits;
                ordered |= (h.len & ORDERED_FLAG) != 0u;
                const unsigned md = __ballot_sync(0xffffffffu, is_d), ms = __ballot_sync(0xffffffffu, is_s);
                if (is_d) dlist[nd + __popc(md & lt)] = h.begin;
                if (is_s) shits[ns + __popc(ms & lt)] = make_uint2(h.begin, h.len & LEN_MASK);
                nd += __popc(md);
                ns += __popc(ms);
                if (STATS && is_s) tsum += h.len & LEN_MASK;  // (a dense key's df is counted from its bitmap)
            }
            ordered = __any_sync(0xffffffffu, ordered);
            __syncwarp();
            if (ordered) {  // rare: the list is walked in key order
                had_ordered = true;
                for (uint32_t i = lane; i < n_hits; i += 32) hits[i] = a.plan[h0 + i];
                __syncwarp();
                process_hits(n_hits);
                __syncwarp();
                continue;
            }
            if (STATS) st_post += tsum;
            dense_sweep(nd);
            // sparse keys: the lanes fan out over a key's postings; the first chunks of four keys are requested
            // together, and the counters are bumped with shared-memory atomics on the u32 that holds two of them
            // (two keys may hit the same leaf: no barrier between keys is needed this way)
            {
                uint32_t *cnt32 = reinterpret_cast<uint32_t *>(cnt);
                auto bump = [&](uint64_t p) {
                    const uint32_t leaf = (uint32_t)(p >> 32);  // no flag bits in an unordered list
                    atomicAdd(cnt32 + (leaf >> 1), 1u << ((leaf & 1u) * 16u));
                };
                for (uint32_t s0 = 0; s0 < ns; s0 += 4) {
                    uint2 rec[4];
                    uint64_t p[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        rec[u] = s0 + u < ns ? shits[s0 + u] : make_uint2(0u, 0u);
                        p[u] = lane < rec[u].y ? __ldg(post + rec[u].x + lane) : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (lane < rec[u].y) bump(p[u]);
                        for (uint32_t off = 32 + lane; off < rec[u].y; off += 32) bump(__ldg(post + rec[u].x + off));
                    }
                }
            }
            __syncwarp();
        }

        // ---- the next task's descriptor (its id was requested half a task ago), and the id after that
        uint32_t n_task = task + 1, n_end = task_end;
        if (n_task == task_end) {  // the batch is used up: take the one requested a batch ago, request another
            n_task = __shfl_sync(0xffffffffu, next_task, 0);
            n_end = n_task + TASK_BATCH;
            if (lane == 0) next_task = (uint32_t)atomicAdd(a.counters, (unsigned long long)TASK_BATCH);
        }
        uint32_t n_t = 0, n_ql = 0;
        uint64_t n_hb = 0, n_he = 0;
        float n_v1 = 0.0f;
        load_desc(n_task, n_t, n_ql, n_hb, n_he, n_v1);
        auto advance = [&]() {
            task = n_task;
            task_end = n_end;
            t = n_t;
            ql = n_ql;
            hb_raw = n_hb;
            he_raw = n_he;
            v1 = n_v1;
        };
        if (skip) {
            advance();
            continue;
        }

        // ---- bound pass (fused top-x, once the query has a bound).  The score of a leaf is the f32 sequence
        // s <- fl(s + p), n times, from its partial sum s0: every rounding adds at most 2^-24 relative, so
        // score <= (s0 + n p)(1 + 2^-24)^n <= 1.004 (s0 + n p) for n < 65535.  The bound is formed with the largest
        // unit value of the tile (p <= pmax = fl(v1 unit_max)) and with the dense keys still pending in the
        // bit-sliced counters taken as ALL present (n <= cnt + pending); a leaf whose bound stays below the query's
        // running bound cannot enter its top-x: its shared keys were counted, its score is not finished (left at
        // +0, below any non-zero bound).  Only the survivors -- the query's relatives -- get their unit value
        // loaded, their dense count unsliced and their additions replayed, exactly as in the full replay below.
        bool emit_ready = false;  // acc[] holds the final scores that matter, e_cnt of them reach the bound e_thr
        uint32_t e_thr = 0, e_cnt = 0;
        if (FUSED && a.ub_prune) {
            const uint32_t thr = read_bound(a.gthr + ql, lane);
            if (thr != 0u) {
                const float thr_f = __uint_as_float(thr);
                const float pmax = __fmul_rn(v1, unit_max);
                // Integer form of the same test for the usual task (no ordered batch: every partial sum is +0;
                // Cosine): ub(n) = 1.01 n pmax is monotone in n, so "ub < bound" is "n < n_need" with
                // n_need = floor(bound / (1.0101 pmax)) - 1 (the margin of 1e-4 and of one count covers the roundings
                // of the division).  Eight counters per 16-byte load, one add + one OR per pair of counters: a
                // counter c >= T sets bit 15 of (c + 0x8000 - T).  Survivors (rare) are listed, the pending dense
                // counts flushed, and each survivor is finished by its own lane.
                if (SIM == MYRIO_SIM_COSINE && !had_ordered && (a.tile_leaves & 7u) == 0u && any_hits) {
                    const float q = __fdiv_rn(thr_f, __fmul_rn(pmax, 1.0101f));  // +inf for pmax == 0
                    uint32_t n_need = q < 1.0e9f ? (uint32_t)q : 1000000000u;
                    n_need = n_need > 1u ? n_need - 1u : 0u;
                    const uint32_t T = n_need > pending ? n_need - pending : 0u;
                    if (T >= 1u && T <= 0x8000u) {
                        // (counters stay below 0x8000 + T: a count is at most the query's length, and count mode
                        // is off for leaf vectors of 65 535+ keys; the general pass below needs no such assumption)
                        const uint32_t KK = (0x8000u - T) * 0x00010001u;
                        // survivor leaves [0, SL_CAP) and their number [SL_CAP] live in the hit buffer (idle now)
                        constexpr uint32_t SL_CAP = SC_HITS * sizeof(HitRec) / sizeof(uint32_t) - 1;
                        uint32_t *slist = reinterpret_cast<uint32_t *>(hits);
                        if (lane == 0) slist[SL_CAP] = 0;
                        __syncwarp();
                        const uint4 *cnt4 = reinterpret_cast<const uint4 *>(cnt);
                        for (uint32_t i = lane; i * 8u < nl; i += 32) {
                            const uint4 v = cnt4[i];
                            if ((((v.x + KK) | (v.y + KK) | (v.z + KK) | (v.w + KK)) & 0x80008000u) == 0u) continue;
                            const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const uint32_t c = (ww[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu, leaf = i * 8u + e;
                                if (c >= T && leaf < nl) {
                                    const uint32_t pos = atomicAdd(&slist[SL_CAP], 1u);
                                    if (pos < SL_CAP) slist[pos] = leaf;
                                }
                            }
                        }
                        __syncwarp();
                        const uint32_t n_surv = slist[SL_CAP];
                        if (n_surv <= SL_CAP) {
                            uint32_t c_cnt = 0;
                            if (n_surv) {
                                flush_planes();  // the survivors' counts become complete
                                for (uint32_t j = lane; j < n_surv; j += 32) {
                                    const uint32_t leaf = slist[j];
                                    const float p = __fmul_rn(v1, __ldg(unit + leaf));
                                    float sc = 0.0f;
                                    for (uint32_t n = cnt[leaf]; n; --n) sc = __fadd_rn(sc, p);
                                    acc[leaf] = sc;  // Cosine: the dot product is the score
                                    c_cnt += (__float_as_uint(sc) >= thr) ? 1u : 0u;
                                }
                                c_cnt = warp_sum_u32(c_cnt);
                                __syncwarp();
                            }
                            acc_dirty = n_surv != 0u;
                            e_thr = thr;
                            e_cnt = c_cnt;
                            emit_ready = true;
                        }
                    }
                }
                const uint32_t magic = 0x4B000000u + pending;  // float(2^23 + n) for n < 2^23
                uint32_t c_cnt = 0;
                if (!emit_ready && any_hits) {
#pragma unroll 4
                    for (uint32_t r = 0; 32u * r < nl; ++r) {
                        const uint32_t leaf = 32u * r + lane;  // leaf < tile_leaves + 32: inside the padded arrays
                        const float s0 = acc[leaf];
                        const uint32_t n0 = cnt[leaf];
                        const float nf = __fadd_rn(__uint_as_float(magic + n0), -8388608.0f);
                        float ub = __fmul_rn(__fadd_rn(s0, __fmul_rn(nf, pmax)), 1.01f);
                        if (SIM == MYRIO_SIM_JACARD_TANIMOTO)  // monotone in the dot product while it is < 2
                            ub = ub < 1.5f ? sim_transform<SIM>(ub) : 3.0e38f;
                        const bool need = leaf < nl && !(ub < thr_f);
                        if (need || s0 != 0.0f) {
                            float sc = 0.0f;
                            if (need) {
                                const float p = __fmul_rn(v1, __ldg(unit + leaf));
                                uint32_t c = 0;
#pragma unroll
                                for (int j = 0; j < SCC_PLANES; ++j) c |= ((pl[j] >> r) & 1u) << j;
                                sc = s0;
                                for (uint32_t n = n0 + c; n; --n) sc = __fadd_rn(sc, p);
                                sc = sim_transform<SIM>(sc);
                                c_cnt += (__float_as_uint(sc) >= thr) ? 1u : 0u;
                            }
                            acc[leaf] = sc;
                        }
                    }
                }
                if (!emit_ready) {
                    e_cnt = warp_sum_u32(c_cnt);
                    e_thr = thr;
                    emit_ready = true;
                    __syncwarp();
                }
            }
        }
        if (!emit_ready) {
        flush_planes();

        // replay: n sequential additions of p per leaf; four leaves of a lane in flight (rows r0 .. r0+3 of its
        // slots: conflict-free shared-memory accesses, 32 consecutive leaves across the warp), in lock step up to
        // the shortest of the four, then predicated.  (Variants measured in profiles/r02_k4_variants.md: giving a
        // lane four CONSECUTIVE leaves or finishing the remainders one chain at a time is slower -- the chains are
        // latency-bound, four independent ones per lane are needed.)
        for (uint32_t r0 = 0; any_hits && 32u * r0 < nl; r0 += 4) {  // a task without hits leaves all scores at +0
            float s[4], p[4];
            uint32_t n[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t leaf = 32u * (r0 + j) + lane;
                const bool ok = leaf < nl;
                n[j] = ok ? cnt[leaf] : 0u;
                s[j] = ok ? acc[leaf] : 0.0f;
                p[j] = ok ? __fmul_rn(v1, __ldg(unit + leaf)) : 0.0f;
            }
            const uint32_t nmin = min(min(n[0], n[1]), min(n[2], n[3]));
            const uint32_t nmax = max(max(n[0], n[1]), max(n[2], n[3]));
            uint32_t i = 0;
            for (; i < nmin; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = __fadd_rn(s[j], p[j]);
            }
            for (; i < nmax; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = i < n[j] ? __fadd_rn(s[j], p[j]) : s[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t leaf = 32u * (r0 + j) + lane;
                if (leaf < nl) acc[leaf] = s[j];
            }
        }
        __syncwarp();

        if (FUSED) {  // all scores are finished: how many reach the bound?
            e_thr = read_bound(a.gthr + ql, lane);
            uint32_t c = 0;
            for (uint32_t i = lane; i < nl; i += 32) {
                const float sc = sim_transform<SIM>(acc[i]);
                acc[i] = sc;
                c += (__float_as_uint(sc) >= e_thr) ? 1u : 0u;
            }
            e_cnt = warp_sum_u32(c);
            __syncwarp();
        } else {
            float *out = a.scores + (size_t)ql * a.ld + ti->leaf_begin;
            for (uint32_t i = lane; i < nl; i += 32) out[i] = sim_transform<SIM>(acc[i]);
        }
        }
        // one top-x tail for the three ways a task ends (integer bound pass, general bound pass, full replay)
        if (FUSED)
            tile_topx_emit(acc, hist, nl, a.x, a.leaf_id_base + ti->leaf_begin, a.gthr + ql, a.cand + slot * a.x,
                           a.cand_cnt + slot, lane, e_thr, e_cnt);
        __syncwarp();
        advance();
    }
    if (STATS) {  // lane 0 counted the walked lists, every lane its share of the split ones
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) st_post += __shfl_xor_sync(0xffffffffu, st_post, d);
        if (lane == 0 && st_post) atomicAdd(a.counters + 3, st_post);
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
// order (score desc, payload_id asc) == descending composite.  The candidate counts
// are prefix-summed so that the (few hundred) survivors are gathered in one go and
// sorted once; only a query with more than TX_SORT_CAP candidates takes the chunked path.
__global__ void __launch_bounds__(TX_THREADS) topx_cand_merge_kernel(const uint64_t *__restrict__ cand,
                                                                     const uint32_t *__restrict__ cand_cnt,
                                                                     const uint32_t *__restrict__ gthr,
                                                                     uint32_t n_tiles, uint32_t x,
                                                                     float *__restrict__ out_s,
                                                                     uint32_t *__restrict__ out_i,
                                                                     uint64_t *__restrict__ out_c) {
    __shared__ uint64_t buf[TX_SORT_CAP];
    __shared__ uint32_t s_warp[TX_THREADS / 32 + 1];
    __shared__ uint32_t s_total;
    const size_t q = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *cnts = cand_cnt + q * n_tiles;
    // pass 1: total number of candidates
    uint32_t local = 0;
    for (uint32_t t = tid; t < n_tiles; t += TX_THREADS) local += cnts[t];
    local = warp_sum_u32(local);
    if (lane == 0) s_warp[warp] = local;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < TX_THREADS / 32; ++w) s += s_warp[w];
        s_total = s;
    }
    __syncthreads();
    const uint32_t total = s_total;
    __syncthreads();  // everyone has read the total before the cursor below reuses s_total
    // The FINAL bound of the query (max over its tiles of the tile's x-th best score) is known
    // now: the early tiles wrote their candidates against a looser bound, most of which fall
    // below it and cannot be in the top-x (at least x scores are >= the bound).
    const uint32_t thr = gthr[q];
    uint32_t filled = 0;
    if (total <= TX_SORT_CAP) {
        // gather: one warp per tile, positions from an atomic cursor (order is fixed by the sort)
        if (tid == 0) s_total = 0;
        __syncthreads();
        for (uint32_t t = warp; t < n_tiles; t += TX_THREADS / 32) {
            const uint32_t cnt = cnts[t];
            const uint64_t *src = cand + (q * n_tiles + t) * x;
            for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                const uint32_t j = j0 + lane;
                const uint64_t c = j < cnt ? src[j] : 0ull;
                const bool keep = j < cnt && (uint32_t)(c >> 32) >= thr;
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m == 0) continue;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_total, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) buf[base + __popc(m & ((1u << lane) - 1u))] = c;
            }
        }
        __syncthreads();
        filled = s_total;
        uint32_t S = 2;
        while (S < filled) S <<= 1;
        for (uint32_t j = filled + tid; j < S; j += TX_THREADS) buf[j] = 0;
        bitonic_desc(buf, S);
        filled = min(filled, x);
    } else {
        for (uint32_t t = 0; t <= n_tiles; ++t) {
            const uint32_t cnt = t < n_tiles ? cnts[t] : 0u;
            if (t == n_tiles || filled + cnt > TX_SORT_CAP) {
                uint32_t S = 2;
                while (S < filled) S <<= 1;
                for (uint32_t j = filled + tid; j < S; j += TX_THREADS) buf[j] = 0;
                bitonic_desc(buf, S);
                filled = min(filled, x);
            }
            if (t < n_tiles) {
                const uint64_t *src = cand + (q * n_tiles + t) * x;
                for (uint32_t j = tid; j < cnt; j += TX_THREADS) buf[filled + j] = src[j];
                filled += cnt;
                __syncthreads();
            }
        }
    }
    for (uint32_t j = tid; j < x; j += TX_THREADS) {
        const uint64_t c = j < filled ? buf[j] : 0ull;  // padding = (score 0, id 0xFFFFFFFF) = composite 0
        if (out_c) {
            out_c[q * x + j] = c;
        } else {
            out_s[q * x + j] = __uint_as_float((uint32_t)(c >> 32));
            out_i[q * x + j] = 0xFFFFFFFFu - (uint32_t)c;
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

// K7 on composites: in_c = [n_parts][q][x] of (score bits << 32 | ~payload id), padding = 0
__global__ void __launch_bounds__(256) topx_merge_composites_kernel(const uint64_t *__restrict__ in_c,
                                                                    uint32_t n_parts, uint64_t q_count, uint32_t x,
                                                                    float *__restrict__ out_s,
                                                                    uint32_t *__restrict__ out_i) {
    // Every part is sorted (best first, 0 = padding) and the composites are distinct (payload ids), so the place
    // of an element in the merged list is its place in its own part + the number of greater elements in every
    // other part (a binary search each): no sort, one barrier.
    extern __shared__ uint64_t mk[];  // [n_parts][x]
    const uint64_t q = blockIdx.x;
    const uint32_t total = n_parts * x;
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        const uint32_t p = j / x, e = j % x;
        mk[j] = in_c[((size_t)p * q_count + q) * x + e];
    }
    for (uint32_t j = threadIdx.x; j < x; j += blockDim.x) {  // places nobody reaches stay padding
        out_s[q * x + j] = 0.0f;
        out_i[q * x + j] = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        const uint64_t c = mk[j];
        if (c == 0ull) continue;
        const uint32_t p = j / x;
        uint32_t rank = j - p * x;
        for (uint32_t o = 0; o < n_parts && rank < x; ++o) {
            if (o == p) continue;
            const uint64_t *lst = mk + (size_t)o * x;
            uint32_t lo = 0, hi = x;  // first position whose element is not greater than c
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lst[mid] > c) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < x) {
            out_s[q * x + rank] = __uint_as_float((uint32_t)(c >> 32));
            out_i[q * x + rank] = 0xFFFFFFFFu - (uint32_t)c;
        }
    }
}

// ------------------------------------------------------------------ plan

// A query key is looked up ONCE per query in the index-wide dictionary (not once per tile);
// the result is a short list of (tile, begin, len) entries.  plan_probe counts the hits of
// every (query, tile) task, a prefix sum lays the hit lists out, plan_fill writes them in
// ascending key order.  The scoring kernel then streams its hit list and never probes.
static constexpr int PL_WARPS = 8;

// one warp per query: dictionary lookups (lanes over keys) + per-tile hit counts
__global__ void __launch_bounds__(PL_WARPS * 32) plan_probe_kernel(
    const uint64_t *__restrict__ q_row_off, const uint64_t *__restrict__ q_keys, uint64_t q_first,
    uint32_t q_count, const DictSlot *__restrict__ gdict, uint32_t gmask,
    const EntryRec *__restrict__ entries, uint32_t n_tiles, uint2 *__restrict__ kref,
    uint32_t *__restrict__ hitcnt, const float *__restrict__ q_vals, float *__restrict__ qv1) {
    extern __shared__ uint32_t pl_count[];  // [PL_WARPS][n_tiles]: the warp's query owns one row of hitcnt
    const uint32_t ql = blockIdx.x * PL_WARPS + (threadIdx.x >> 5);
    if (ql >= q_count) return;
    const unsigned lane = threadIdx.x & 31;
    uint32_t *cnt = pl_count + (size_t)(threadIdx.x >> 5) * n_tiles;
    for (uint32_t t = lane; t < n_tiles; t += 32) cnt[t] = 0;
    __syncwarp();
    const uint64_t kbase = q_row_off[q_first];
    const uint64_t qb = q_row_off[q_first + ql], qe = q_row_off[q_first + ql + 1];
    const uint4 *dict = reinterpret_cast<const uint4 *>(gdict);
    float vmin = 3.402823466e+38f;
    for (uint64_t i = qb + lane; i < qe; i += 32) {
        vmin = fminf(vmin, q_vals[i]);
        const uint64_t key = q_keys[i];
        uint32_t h = dict_hash(key) & gmask;
        uint32_t begin = 0, len = 0;
        for (;;) {
            const uint4 s = __ldg(dict + h);
            if (s.w == 0u) break;  // empty slot: the key occurs nowhere in this index
            if ((((uint64_t)s.y << 32) | s.x) == key) {
                begin = s.z;
                len = s.w;
                break;
            }
            h = (h + 1) & gmask;
        }
        kref[i - kbase] = make_uint2(begin, len);
        for (uint32_t e = 0; e < len; ++e) atomicAdd(&cnt[entries[begin + e].tile], 1u);
    }
    __syncwarp();
    uint32_t *out = hitcnt + (size_t)ql * n_tiles;
    for (uint32_t t = lane; t < n_tiles; t += 32) out[t] = cnt[t];
    // the query's common value for the count-mode kernel: its smallest value (a read's k-mers of count 1)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
    if (lane == 0) qv1[ql] = qe > qb ? vmin : 0.0f;
}

// one warp per query: keys in ascending order, the lanes fan out over ONE key's entries
// (distinct tiles => no two lanes bump the same cursor)
__global__ void __launch_bounds__(PL_WARPS * 32) plan_fill_kernel(
    const uint64_t *__restrict__ q_row_off, const float *__restrict__ q_vals, uint64_t q_first,
    uint32_t q_count, const EntryRec *__restrict__ entries, uint32_t n_tiles,
    const uint2 *__restrict__ kref, const uint64_t *__restrict__ plan_off, HitRec *__restrict__ plan,
    uint32_t skip_words, const float *__restrict__ qv1s) {
    extern __shared__ uint32_t pl_cursor[];  // [PL_WARPS][n_tiles], relative to the query's first hit
    const uint32_t ql = blockIdx.x * PL_WARPS + (threadIdx.x >> 5);
    if (ql >= q_count) return;
    // queries of at most 32*skip_words keys were placed by plan_fill_ranked_kernel
    if ((q_row_off[q_first + ql + 1] - q_row_off[q_first + ql] + 31) / 32 <= skip_words) return;
    const unsigned lane = threadIdx.x & 31;
    const float qv1 = qv1s[ql];  // the query's common value (plan_probe_kernel)
    uint32_t *cur = pl_cursor + (size_t)(threadIdx.x >> 5) * n_tiles;
    const uint64_t *off = plan_off + (size_t)ql * n_tiles;
    const uint64_t base = off[0];
    for (uint32_t t = lane; t < n_tiles; t += 32) cur[t] = (uint32_t)(off[t] - base);
    __syncwarp();
    const uint64_t kbase = q_row_off[q_first];
    const uint64_t qb = q_row_off[q_first + ql], qe = q_row_off[q_first + ql + 1];
    for (uint64_t r0 = qb; r0 < qe; r0 += 32) {
        uint2 kr = make_uint2(0, 0);
        float qv = 0.0f;
        if (r0 + lane < qe) {
            kr = kref[r0 + lane - kbase];
            qv = q_vals[r0 + lane];
        }
        unsigned have = __ballot_sync(0xffffffffu, kr.y != 0u);
        while (have) {
            const int src = __ffs(have) - 1;
            have &= have - 1;
            const uint32_t eb = __shfl_sync(0xffffffffu, kr.x, src);
            const uint32_t en = __shfl_sync(0xffffffffu, kr.y, src);
            const float v = __shfl_sync(0xffffffffu, qv, src);
            for (uint32_t e = lane; e < en; e += 32) {
                const EntryRec er = entries[eb + e];
                HitRec h;
                h.begin = er.begin;
                h.len = er.len | (v != qv1 ? ORDERED_FLAG : 0u);
                h.qv = v;
                plan[base + cur[er.tile]++] = h;
            }
            __syncwarp();
        }
    }
}

// Fast variant of plan_fill: one CTA per query.  present[tile][key/32] (shared memory) is the
// bit-matrix "key i of the query occurs in tile t"; the place of a hit inside its (query, tile)
// list is the number of earlier keys present in that tile = a prefix popcount, so all hits are
// placed in parallel (no per-key serialisation).  Needs n_tiles * ceil(nq/32) words of shared
// memory; plan_fill_kernel is the fallback for queries/tilings that do not fit.
static constexpr int PF_THREADS = 256;
__global__ void __launch_bounds__(PF_THREADS) plan_fill_ranked_kernel(
    const uint64_t *__restrict__ q_row_off, const float *__restrict__ q_vals, uint64_t q_first,
    const EntryRec *__restrict__ entries, uint32_t n_tiles, uint32_t words_cap,
    const uint2 *__restrict__ kref, const uint64_t *__restrict__ plan_off, HitRec *__restrict__ plan,
    const float *__restrict__ qv1s) {
    // shared: present[n_tiles][words] | hit_prefix[32*words_cap + 1] | ent_begin[32*words_cap]
    // (staging the query's plan region in shared memory for coalesced stores was measured slower:
    // the extra shared memory costs more occupancy than the scattered 12-byte stores cost)
    extern __shared__ uint32_t pf_smem[];
    __shared__ uint32_t s_warp[PF_THREADS / 32 + 1];
    const uint32_t ql = blockIdx.x;
    const float qv1 = qv1s[ql];  // the query's common value (plan_probe_kernel)
    const uint64_t kbase = q_row_off[q_first];
    const uint64_t qb = q_row_off[q_first + ql], qe = q_row_off[q_first + ql + 1];
    const uint32_t nq = (uint32_t)(qe - qb);
    const uint32_t words = (nq + 31) >> 5;
    if (words > words_cap || nq == 0) return;  // long queries: fallback kernel
    uint32_t *present = pf_smem;
    uint32_t *prefix = pf_smem + (size_t)n_tiles * words_cap;
    uint32_t *ebegin = prefix + 32 * words_cap + 1;
    const uint64_t *off = plan_off + (size_t)ql * n_tiles;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < n_tiles * words; i += PF_THREADS) present[i] = 0;
    // exclusive prefix of the entry counts over the query's keys (chunks of PF_THREADS keys)
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nq; base += PF_THREADS) {
        const uint32_t i = base + tid;
        uint2 kr = make_uint2(0, 0);
        if (i < nq) kr = kref[qb + i - kbase];
        uint32_t inc = kr.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < warp; ++w) wbase += s_warp[w];
        if (i < nq) {
            prefix[i] = carry + wbase + inc - kr.y;
            ebegin[i] = kr.x;
        }
        uint32_t total = 0;
        for (uint32_t w = 0; w < PF_THREADS / 32; ++w) total += s_warp[w];
        carry += total;
        __syncthreads();
    }
    if (tid == 0) prefix[nq] = carry;
    __syncthreads();
    const uint32_t H = carry;  // hits of this query over all tiles
    auto key_of_hit = [&](uint32_t j) {  // last i with prefix[i] <= j
        uint32_t lo = 0, hi = nq;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (prefix[mid] <= j) lo = mid; else hi = mid;
        }
        return lo;
    };
    // pass 1: one bit per (key, tile) hit -- one thread per hit
    for (uint32_t j = tid; j < H; j += PF_THREADS) {
        const uint32_t i = key_of_hit(j);
        const uint32_t tile = entries[ebegin[i] + (j - prefix[i])].tile;
        atomicOr(&present[tile * words + (i >> 5)], 1u << (i & 31));
    }
    __syncthreads();
    // pass 2: rank of key i in tile t = popcount of present[t] below bit i
    for (uint32_t j = tid; j < H; j += PF_THREADS) {
        const uint32_t i = key_of_hit(j);
        const EntryRec er = entries[ebegin[i] + (j - prefix[i])];
        const uint32_t *row = present + er.tile * words;
        uint32_t rank = __popc(row[i >> 5] & ((1u << (i & 31)) - 1u));
        for (uint32_t w = 0; w < (i >> 5); ++w) rank += __popc(row[w]);
        HitRec h;
        h.begin = er.begin;
        h.qv = q_vals[qb + i];
        h.len = er.len | (h.qv != qv1 ? ORDERED_FLAG : 0u);
        plan[off[er.tile] + rank] = h;
    }
}

// ---- plan, second generation: two kernels and a scan over the QUERIES (not over the (query, tile) tasks).
// plan_count_kernel: one warp per query -- dictionary lookups only; a key's slot already says in how many tiles
// it occurs, so the query's total number of hits is known without touching the entry lists.
__global__ void __launch_bounds__(PL_WARPS * 32) plan_count_kernel(
    const uint64_t *__restrict__ q_row_off, const uint64_t *__restrict__ q_keys, uint64_t q_first,
    uint32_t q_count, const DictSlot *__restrict__ gdict, uint32_t gmask, uint2 *__restrict__ kref,
    uint32_t *__restrict__ qhits, const float *__restrict__ q_vals, float *__restrict__ qv1) {
    const uint32_t ql = blockIdx.x * PL_WARPS + (threadIdx.x >> 5);
    if (ql >= q_count) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t kbase = q_row_off[q_first];
    const uint64_t qb = q_row_off[q_first + ql], qe = q_row_off[q_first + ql + 1];
    const uint4 *dict = reinterpret_cast<const uint4 *>(gdict);
    float vmin = 3.402823466e+38f;
    uint32_t total = 0;
    for (uint64_t i = qb + lane; i < qe; i += 32) {
        vmin = fminf(vmin, q_vals[i]);
        const uint64_t key = q_keys[i];
        uint32_t h = dict_hash(key) & gmask;
        uint32_t begin = 0, len = 0;
        for (;;) {
            const uint4 s = __ldg(dict + h);
            if (s.w == 0u) break;  // empty slot: the key occurs nowhere in this index
            if ((((uint64_t)s.y << 32) | s.x) == key) {
                begin = s.z;
                len = s.w;
                break;
            }
            h = (h + 1) & gmask;
        }
        kref[i - kbase] = make_uint2(begin, len);
        total += len;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
        total += __shfl_xor_sync(0xffffffffu, total, d);
    }
    if (lane == 0) {
        qv1[ql] = qe > qb ? vmin : 0.0f;  // the query's common value for the count-mode kernel
        qhits[ql] = total;
    }
}

// plan_build_kernel: one CTA per query.  Sweep 1 sets present[tile][key] for every hit (each warp flattens the
// entry lists of 32 keys over its lanes); the tiles' hit counts and the prefix popcounts of the bit rows give the
// place of every hit -- rank of key i in tile t = keys below i present in t -- so sweep 2 writes all hit records
// in ascending key order without any serialisation.  The query's tile offsets (plan_off row of n_tiles + 1
// entries) start at qoff[q]; its seed tile (most hits, lowest tile on ties) is chosen on the way.
//
// Dense-key masks (mask_stride > 0; count-mode scoring).  A COMMON dense hit -- dense key, query value == the
// query's common value -- needs no record at all: in a hit list without ordered hits its whole content is "bitmap
// b of the tile", so it is one bit (bit b of the tile's words in the query's mask row).  On barcode-shaped data
// three hits out of four are of this kind.  A tile whose list holds an ordered hit keeps records for everything
// (the walk in key order needs them) and gets an empty mask.
static constexpr int PB_THREADS = 256;
static constexpr int PB_WARP_WORDS = 100;  // per warp: prefix[33] | entry begins[32] | query values[32]
__host__ __device__ inline size_t plan_build_smem_bytes(uint32_t n_tiles, uint32_t words, uint32_t mask_stride) {
    const size_t per_bit = mask_stride ? 10 : 6;  // present (+ dense-common) bit rows + u16 prefix popcounts
    return ((size_t)n_tiles * words * per_bit + 3) / 4 * 4 + ((size_t)n_tiles + 1) * sizeof(uint32_t) +
           (mask_stride ? ((size_t)(n_tiles + 31) / 32 + mask_stride + words + 1) * sizeof(uint32_t) : 0) +
           (PB_THREADS / 32) * PB_WARP_WORDS * sizeof(uint32_t);
}
__global__ void __launch_bounds__(PB_THREADS) plan_build_kernel(
    const uint64_t *__restrict__ q_row_off, const float *__restrict__ q_vals, uint64_t q_first,
    const EntryRec *__restrict__ entries, uint32_t n_tiles, uint32_t words_cap, const uint2 *__restrict__ kref,
    const uint64_t *__restrict__ qoff, const float *__restrict__ qv1s, uint64_t *__restrict__ plan_off,
    HitRec *__restrict__ plan, uint32_t *__restrict__ seed, const uint32_t *__restrict__ mask_off,
    uint32_t mask_stride, uint32_t *__restrict__ dmask_out) {
    extern __shared__ uint32_t pb_smem[];
    __shared__ uint32_t s_scan[PB_THREADS / 32 + 1];
    __shared__ unsigned long long s_best;
    const uint32_t ql = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t kbase = q_row_off[q_first];
    const uint64_t qb = q_row_off[q_first + ql], qe = q_row_off[q_first + ql + 1];
    const uint32_t nq = (uint32_t)(qe - qb);
    const uint32_t words = max((nq + 31) >> 5, 1u);  // <= words_cap (the host checked the longest query)
    const uint64_t base = qoff[ql];
    uint64_t *off_row = plan_off + (size_t)ql * (n_tiles + 1);
    const bool masks = mask_stride != 0;
    uint32_t *present = pb_smem;                                                      // [n_tiles][words]
    uint32_t *pdc = present + (masks ? (size_t)n_tiles * words : 0);                  // [n_tiles][words] dense-common
    uint16_t *pfx = reinterpret_cast<uint16_t *>(pdc + (size_t)n_tiles * words);      // [n_tiles][words]
    uint32_t *toff = pb_smem + ((size_t)n_tiles * words_cap * (masks ? 10 : 6) + 3) / 4;  // [n_tiles + 1]
    uint32_t *tord = toff + n_tiles + 1;                                              // [(n_tiles+31)/32] ordered tiles
    uint32_t *dm = tord + (n_tiles + 31) / 32;                                        // [mask_stride]
    uint32_t *krec = dm + mask_stride;  // [words_cap] keys with a hit that is NOT a mask bit | [1] any ordered tile
    uint32_t *wpre = toff + n_tiles + 1 + (masks ? (n_tiles + 31) / 32 + mask_stride + words_cap + 1 : 0) +
                     (size_t)warp * PB_WARP_WORDS;
    uint32_t *wbeg = wpre + 33;
    float *wval = reinterpret_cast<float *>(wbeg + 32);
    const float qv1 = qv1s[ql];
    for (uint32_t i = tid; i < n_tiles * words * (masks ? 2u : 1u); i += PB_THREADS) present[i] = 0;
    if (masks)
        for (uint32_t i = tid; i < (n_tiles + 31) / 32 + mask_stride + words_cap + 1; i += PB_THREADS) tord[i] = 0;
    if (tid == 0) s_best = 0;
    __syncthreads();

    // the hits of 32 consecutive keys flattened over the lanes of a warp, FOUR hits of a lane at a time so that
    // their (dependent, L2-latency) entry loads overlap: f(key index[4], slot[4], entry index[4], n)
    auto sweep = [&](bool only_record_keys, auto f) {
        for (uint32_t k0 = warp * 32; k0 < nq; k0 += PB_THREADS) {
            const uint32_t i = k0 + lane;
            uint2 kr = make_uint2(0u, 0u);
            float v = 0.0f;
            if (i < nq) {
                kr = kref[qb + i - kbase];
                v = q_vals[qb + i];
                // second sweep: a key all of whose hits became mask bits has nothing to write
                if (only_record_keys && !((krec[i >> 5] >> (i & 31)) & 1u)) kr.y = 0;
            }
            uint32_t inc = kr.y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (unsigned)d) inc += t;
            }
            __syncwarp();  // the previous round's readers are done
            wpre[lane + 1] = inc;
            if (lane == 0) wpre[0] = 0;
            wbeg[lane] = kr.x;
            wval[lane] = v;
            __syncwarp();
            const uint32_t H = wpre[32];
            for (uint32_t j0 = lane; j0 < H; j0 += 128) {
                uint32_t ki[4], sl[4], en[4];
                uint32_t n = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + 32u * u;
                    uint32_t lo = 0;  // last slot with wpre[slot] <= j (empty lists are skipped)
                    if (j < H) {
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1)
                            if (wpre[lo + step] <= j) lo += step;
                        n = u + 1;
                    }
                    ki[u] = k0 + lo;
                    sl[u] = lo;
                    en[u] = wbeg[lo] + (min(j, H - 1) - wpre[lo]);
                }
                f(ki, sl, en, n);
            }
        }
    };
    if (!masks) {
        sweep(false, [&](const uint32_t(&ki)[4], const uint32_t(&)[4], const uint32_t(&en)[4], uint32_t n) {
            uint32_t tl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((uint32_t)u < n) tl[u] = entries[en[u]].tile;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((uint32_t)u < n) atomicOr(&present[tl[u] * words + (ki[u] >> 5)], 1u << (ki[u] & 31));
        });
    } else {
        sweep(false, [&](const uint32_t(&ki)[4], const uint32_t(&sl)[4], const uint32_t(&en)[4], uint32_t n) {
            EntryRec er[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((uint32_t)u < n) er[u] = entries[en[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((uint32_t)u < n) {
                    const uint32_t t = er[u].tile, w = ki[u] >> 5, bit = 1u << (ki[u] & 31);
                    atomicOr(&present[t * words + w], bit);
                    const bool ord = (er[u].len & ORDERED_FLAG) != 0u || wval[sl[u]] != qv1;
                    bool masked = false;
                    if (ord) {
                        atomicOr(&tord[t >> 5], 1u << (t & 31));
                        krec[words_cap] = 1u;  // some tile keeps records for everything: no key may be skipped
                    } else if (er[u].len & DENSE_FLAG) {
                        // the tile's mask covers its HOT dense keys (bitmap indexes below 32 * mask words); a cold
                        // dense key stays a record (its DENSE_FLAG sends it through the dense sweep all the same)
                        const uint32_t mo = __ldg(mask_off + t), mw = __ldg(mask_off + t + 1) - mo;
                        if ((er[u].begin >> 5) < mw) {
                            atomicOr(&pdc[t * words + w], bit);
                            atomicOr(&dm[mo + (er[u].begin >> 5)], 1u << (er[u].begin & 31));
                            masked = true;
                        }
                    }
                    if (!masked) atomicOr(&krec[w], bit);
                }
        });
    }
    __syncthreads();
    // hit count of every tile + prefix popcounts of its bit row
    const uint32_t C = (n_tiles + PB_THREADS - 1) / PB_THREADS;  // consecutive tiles per thread
    uint32_t mine = 0;
    unsigned long long best = 0;
    for (uint32_t t = tid * C; t < min(n_tiles, (tid + 1) * C); ++t) {
        uint32_t all = 0, run = 0;
        const bool keep_all = masks && ((tord[t >> 5] >> (t & 31)) & 1u);
        if (keep_all)  // an ordered list: records for every hit, no mask
            for (uint32_t w = __ldg(mask_off + t); w < __ldg(mask_off + t + 1); ++w) dm[w] = 0;
        for (uint32_t w = 0; w < words; ++w) {
            uint32_t bits = present[t * words + w];
            all += __popc(bits);
            if (masks && !keep_all) {
                bits &= ~pdc[t * words + w];  // the common dense hits live in the mask
                present[t * words + w] = bits;
            }
            pfx[t * words + w] = (uint16_t)run;
            run += __popc(bits);
        }
        toff[t] = run;
        mine += run;
        best = max(best, ((unsigned long long)all << 32) | (unsigned long long)(0xFFFFFFFFu - t));
    }
    // exclusive scan of the tile counts (block scan over the per-thread chunk sums)
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 31) s_scan[warp] = inc;
    if (lane == 0) atomicMax(&s_best, best);
    __syncthreads();
    uint32_t run = inc - mine;
    for (uint32_t w = 0; w < warp; ++w) run += s_scan[w];
    for (uint32_t t = tid * C; t < min(n_tiles, (tid + 1) * C); ++t) {
        const uint32_t c = toff[t];
        toff[t] = run;
        run += c;
    }
    if (tid == PB_THREADS - 1) toff[n_tiles] = run;
    __syncthreads();
    const uint32_t seed_tile = 0xFFFFFFFFu - (uint32_t)s_best;
    for (uint32_t t = tid; t <= n_tiles; t += PB_THREADS)
        off_row[t] = (base + toff[t]) | ((seed && t == seed_tile) ? PLAN_SEED_FLAG : 0ull);
    if (seed && tid == 0) seed[ql] = seed_tile;
    if (masks)
        for (uint32_t i = tid; i < mask_stride; i += PB_THREADS) dmask_out[(size_t)ql * mask_stride + i] = dm[i];
    sweep(masks && krec[words_cap] == 0u,
          [&](const uint32_t(&ki)[4], const uint32_t(&sl)[4], const uint32_t(&en)[4], uint32_t n) {
        EntryRec er[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if ((uint32_t)u < n) er[u] = entries[en[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if ((uint32_t)u < n) {
                const uint32_t i = ki[u], w = i >> 5;
                const uint32_t bits = present[er[u].tile * words + w];
                if (!((bits >> (i & 31)) & 1u)) continue;  // a common dense hit: in the mask
                const uint32_t rank = pfx[er[u].tile * words + w] + __popc(bits & ((1u << (i & 31)) - 1u));
                HitRec h;
                h.begin = er[u].begin;
                h.qv = wval[sl[u]];
                h.len = er[u].len | (h.qv != qv1 ? ORDERED_FLAG : 0u);
                plan[base + toff[er[u].tile] + rank] = h;
            }
    });
}

// one warp per query: its seed tile = the tile that holds most of its keys (ties: the lowest tile); the
// tile's plan_off entry is flagged so that the sweep skips the (query, seed tile) task
__global__ void __launch_bounds__(256) seed_tile_kernel(const uint32_t *__restrict__ hitcnt, uint32_t q_count,
                                                        uint32_t n_tiles, uint64_t *__restrict__ plan_off,
                                                        uint32_t *__restrict__ seed) {
    const uint32_t ql = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ql >= q_count) return;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t *row = hitcnt + (size_t)ql * n_tiles;
    // composite (count, ~tile): max = most hits, lowest tile on ties
    unsigned long long best = 0;
    for (uint32_t t = lane; t < n_tiles; t += 32)
        best = max(best, ((unsigned long long)row[t] << 32) | (unsigned long long)(0xFFFFFFFFu - t));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0) {
        const uint32_t t = 0xFFFFFFFFu - (uint32_t)best;
        seed[ql] = t;
        plan_off[(size_t)ql * n_tiles + t] |= PLAN_SEED_FLAG;
    }
}

__global__ void __launch_bounds__(256) max_row_len_kernel(const uint64_t *__restrict__ row_off, uint64_t n,
                                                          unsigned long long *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? row_off[i + 1] - row_off[i] : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

static uint64_t queries_max_keys(myrio_ctx *ctx, const myrio_svecs *q) {
    if (q->n_rows == 0) return 0;
    DevBuf<unsigned long long> d(1);
    MYRIO_CK(cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), ctx->stream));
    MYRIO_LAUNCH(ctx, "max_row_len", max_row_len_kernel, (unsigned)((q->n_rows + 255) / 256), 256, 0, q->row_off.p,
                 q->n_rows, d.p);
    unsigned long long h = 0;
    d2h(ctx, &h, d.p, 1);
    sync(ctx);
    return h;
}

struct Plan {
    DevBuf<uint2> kref;
    DevBuf<uint32_t> hitcnt;
    DevBuf<uint64_t> off;
    DevBuf<HitRec> hits;
    DevBuf<uint32_t> seed;
    DevBuf<float> qv1;
    DevBuf<uint32_t> qhits;  // hits per query (plan_count_kernel)
    DevBuf<uint64_t> qoff;   // their exclusive scan: where a query's hit lists start
    uint64_t n_hits = 0;
    uint32_t stride = 0;     // plan_off entries per query: n_tiles + 1 (plan_build_kernel) or n_tiles (fallback)
    bool seeded = false;     // plan_build_kernel chose the seed tiles and flagged them
    DevBuf<uint32_t> seed_order;  // queries grouped by seed tile (seed pass)
    bool has_seed_order = false;
    DevBuf<uint32_t> dmask;  // [queries][mask_stride] dense-key masks (count mode)
    uint32_t mask_stride = 0;
};
static bool k4_count_mode(const myrio_index *ix);

static bool plan_second_generation() {  // MYRIO_K4A_PLAN=1 keeps the probe / scan / ranked-fill kernels of round 1
    static const bool v = [] {
        const char *e = getenv("MYRIO_K4A_PLAN");
        return !(e && e[0] == '1');
    }();
    return v;
}

// plan for queries [q0, q0+qn); returns false (nothing built) when the hit lists would not fit `budget`.
// want_seed: also choose every query's seed tile (fused top-x with seeding).
static bool build_plan(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q0, uint64_t qn,
                       uint64_t budget_bytes, Plan &pl, bool want_seed = false) {
    const uint64_t n_tiles = ix->tiles.size();
    const uint64_t n_tasks = qn * n_tiles;
    pl.seeded = false;
    pl.kref.reserve(std::max<uint64_t>(queries->nnz, 1));
    pl.qv1.reserve(std::max<uint64_t>(qn, 1));
    const unsigned grid = (unsigned)((qn + PL_WARPS - 1) / PL_WARPS);
    const uint64_t max_q_keys = std::max<uint64_t>(queries_max_keys(ctx, queries), 1);
    const uint32_t words_all = (uint32_t)((max_q_keys + 31) / 32);
    pl.mask_stride = 0;
    // dense-key masks: count-mode scoring only (the add-as-you-go kernel needs a record with the query value for
    // every hit), and only while the CTA's bit matrices stay small enough for several CTAs per SM
    static const bool masks_on = [] {
        const char *e = getenv("MYRIO_K4A_MASKS");
        return !(e && e[0] == '0');
    }();
    uint32_t mask_stride = masks_on && k4_count_mode(ix) ? ix->mask_words_total : 0;
    if (mask_stride && plan_build_smem_bytes((uint32_t)n_tiles, words_all, mask_stride) > 100 * 1024) mask_stride = 0;
    if (plan_second_generation() && max_q_keys < 65536 &&
        plan_build_smem_bytes((uint32_t)n_tiles, words_all, mask_stride) + 4096 <= ctx->smem_optin) {
        pl.qhits.reserve(qn);
        pl.qoff.reserve(qn + 1);
        MYRIO_LAUNCH(ctx, "k4a_plan_count", plan_count_kernel, grid, PL_WARPS * 32, 0, queries->row_off.p,
                     queries->keys.p, q0, (uint32_t)qn, ix->gdict.p, ix->gdict_mask, pl.kref.p, pl.qhits.p,
                     queries->vals_f32.p, pl.qv1.p);
        exclusive_scan_u32_to_u64(ctx, pl.qhits.p, pl.qoff.p, qn);
        uint64_t kb[2];
        d2h(ctx, &kb[0], queries->row_off.p + q0, 1);
        d2h(ctx, &kb[1], queries->row_off.p + q0 + qn, 1);
        d2h(ctx, &pl.n_hits, pl.qoff.p + qn, 1);
        sync(ctx);
        if (pl.n_hits * sizeof(HitRec) > budget_bytes && qn > 1) return false;
        ctx->last_stats[0] += kb[1] - kb[0];
        ctx->last_stats[1] += pl.n_hits;
        pl.hits.reserve(std::max<uint64_t>(pl.n_hits, 1));
        pl.stride = (uint32_t)n_tiles + 1;
        pl.off.reserve(qn * pl.stride + 1);
        const bool seed = want_seed && n_tiles > 1;
        if (seed) pl.seed.reserve(qn);
        if (mask_stride) pl.dmask.reserve(qn * mask_stride);
        const size_t psm = plan_build_smem_bytes((uint32_t)n_tiles, words_all, mask_stride);
        MYRIO_CK(cudaFuncSetAttribute(plan_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        MYRIO_LAUNCH(ctx, "k4a_plan_build", plan_build_kernel, (unsigned)qn, PB_THREADS, psm, queries->row_off.p,
                     queries->vals_f32.p, q0, ix->gentries.p, (uint32_t)n_tiles, words_all, pl.kref.p, pl.qoff.p,
                     pl.qv1.p, pl.off.p, pl.hits.p, seed ? pl.seed.p : (uint32_t *)nullptr, ix->mask_off.p,
                     mask_stride, mask_stride ? pl.dmask.p : (uint32_t *)nullptr);
        pl.seeded = seed;
        pl.mask_stride = mask_stride;
        return true;
    }
    // ---- round-1 path (very long queries x very many tiles): per-task hit counts, a scan over all tasks
    uint64_t kb[2];
    d2h(ctx, &kb[0], queries->row_off.p + q0, 1);
    d2h(ctx, &kb[1], queries->row_off.p + q0 + qn, 1);
    sync(ctx);
    const uint64_t nnz = kb[1] - kb[0];
    pl.hitcnt.reserve(std::max<uint64_t>(n_tasks, 1));
    pl.off.reserve(n_tasks + 1);
    pl.stride = (uint32_t)n_tiles;
    const size_t smem = (size_t)PL_WARPS * n_tiles * sizeof(uint32_t);
    if (smem > ctx->smem_optin) fail(MYRIO_ERR_UNSUPPORTED, "too many tiles (%llu) for the plan kernels", (unsigned long long)n_tiles);
    MYRIO_CK(cudaFuncSetAttribute(plan_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, "k4a_plan_probe", plan_probe_kernel, grid, PL_WARPS * 32, smem, queries->row_off.p,
                 queries->keys.p, q0, (uint32_t)qn, ix->gdict.p, ix->gdict_mask, ix->gentries.p, (uint32_t)n_tiles,
                 pl.kref.p, pl.hitcnt.p, queries->vals_f32.p, pl.qv1.p);
    exclusive_scan_u32_to_u64(ctx, pl.hitcnt.p, pl.off.p, n_tasks);
    d2h(ctx, &pl.n_hits, pl.off.p + n_tasks, 1);
    sync(ctx);
    if (pl.n_hits * sizeof(HitRec) > budget_bytes && qn > 1) return false;
    ctx->last_stats[0] += nnz;
    ctx->last_stats[1] += pl.n_hits;
    pl.hits.reserve(std::max<uint64_t>(pl.n_hits, 1));
    // queries whose presence bit-matrix fits in shared memory take the parallel kernel, the rest
    // (very long queries on indexes with very many tiles) the per-key sequential one
    const size_t smem_budget = std::min<size_t>(ctx->smem_optin, 96 * 1024);
    // shared memory per word of query keys: n_tiles presence words + 64 prefix/begin words
    uint32_t words_cap = (uint32_t)std::min<uint64_t>((max_q_keys + 31) / 32,
                                                      (smem_budget - 64) / ((n_tiles + 64) * sizeof(uint32_t)));
    if (words_cap) {
        const size_t psm = ((size_t)n_tiles * words_cap + 64 * words_cap + 1) * sizeof(uint32_t);
        MYRIO_CK(cudaFuncSetAttribute(plan_fill_ranked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        MYRIO_LAUNCH(ctx, "k4a_plan_fill_ranked", plan_fill_ranked_kernel, (unsigned)qn, PF_THREADS, psm,
                     queries->row_off.p, queries->vals_f32.p, q0, ix->gentries.p, (uint32_t)n_tiles, words_cap,
                     pl.kref.p, pl.off.p, pl.hits.p, pl.qv1.p);
    }
    if ((max_q_keys + 31) / 32 > words_cap) {
        MYRIO_CK(cudaFuncSetAttribute(plan_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MYRIO_LAUNCH(ctx, "k4a_plan_fill", plan_fill_kernel, grid, PL_WARPS * 32, smem, queries->row_off.p,
                     queries->vals_f32.p, q0, (uint32_t)qn, ix->gentries.p, (uint32_t)n_tiles, pl.kref.p, pl.off.p,
                     pl.hits.p, words_cap, pl.qv1.p);
    }
    return true;
}

// Queries in the order of their seed tiles (a counting sort by one CTA: the seed pass then works tile by tile
// like the sweep -- a warp's consecutive tasks share the tile's unit values, bitmaps and postings -- instead of
// jumping to a random tile per task).
static constexpr uint32_t SO_MAX_TILES = 8192;
__global__ void __launch_bounds__(1024) seed_order_kernel(const uint32_t *__restrict__ seed, uint32_t q_count,
                                                          uint32_t n_tiles, uint32_t *__restrict__ order) {
    extern __shared__ uint32_t so_cnt[];  // [n_tiles]
    __shared__ uint32_t s_warp[33];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t t = tid; t < n_tiles; t += blockDim.x) so_cnt[t] = 0;
    __syncthreads();
    for (uint32_t q = tid; q < q_count; q += blockDim.x) atomicAdd(&so_cnt[seed[q]], 1u);
    __syncthreads();
    // exclusive scan of the counters (consecutive tiles per thread)
    const uint32_t C = (n_tiles + blockDim.x - 1) / blockDim.x;
    uint32_t mine = 0;
    for (uint32_t t = tid * C; t < min(n_tiles, (tid + 1) * C); ++t) mine += so_cnt[t];
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t run = inc - mine;
    for (uint32_t w = 0; w < warp; ++w) run += s_warp[w];
    for (uint32_t t = tid * C; t < min(n_tiles, (tid + 1) * C); ++t) {
        const uint32_t c = so_cnt[t];
        so_cnt[t] = run;
        run += c;
    }
    __syncthreads();
    for (uint32_t q = tid; q < q_count; q += blockDim.x) order[atomicAdd(&so_cnt[seed[q]], 1u)] = q;
}

// seed tiles of a fused top-x step (unless plan_build_kernel chose them already) and the queries grouped by them
static void choose_seed_tiles(myrio_ctx *ctx, const myrio_index *ix, Plan &pl, uint64_t qn) {
    if (!pl.seeded) {
        pl.seed.reserve(qn);
        MYRIO_LAUNCH(ctx, "k4a_seed_tiles", seed_tile_kernel, (unsigned)((qn + 7) / 8), 256, 0, pl.hitcnt.p,
                     (uint32_t)qn, (uint32_t)ix->tiles.size(), pl.off.p, pl.seed.p);
    }
    static const bool order_on = [] {  // MYRIO_K4_SEED_ORDER=0: seed tasks in query order (comparison runs)
        const char *e = getenv("MYRIO_K4_SEED_ORDER");
        return !(e && e[0] == '0');
    }();
    pl.has_seed_order = order_on && ix->tiles.size() <= SO_MAX_TILES && qn < 0xFFFFFFFFull;
    if (pl.has_seed_order) {
        pl.seed_order.reserve(qn);
        const size_t smem = ix->tiles.size() * sizeof(uint32_t);
        MYRIO_LAUNCH(ctx, "k4a_seed_order", seed_order_kernel, 1, 1024, smem, pl.seed.p, (uint32_t)qn,
                     (uint32_t)ix->tiles.size(), pl.seed_order.p);
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

static bool k4_unit_in_registers() {  // MYRIO_K4_UREG=1/0 (tuning)
    static const bool v = [] {
        const char *e = getenv("MYRIO_K4_UREG");
        return e ? e[0] == '1' : false;
    }();
    return v;
}

template <int SIM, bool FUSED, bool STATS, bool UREG>
static void launch_score_variant(myrio_ctx *ctx, const myrio_index *ix, ScoreArgs &a) {
    const size_t per_warp = score_warp_smem_bytes(ix->tile_leaves, FUSED);
    int warps = (int)std::min<size_t>(UREG ? SC_UREG_WARPS : SC_MAX_WARPS, (ctx->smem_optin - 1024) / per_warp);
    if (warps < 1) fail(MYRIO_ERR_BAD_ARG, "tile_leaves=%u does not fit in shared memory", ix->tile_leaves);
    const size_t smem = per_warp * warps;
    if (a.n_tasks >= 0xFFF00000ull) fail(MYRIO_ERR_BAD_ARG, "too many (query, tile) tasks in one batch");
    const unsigned grid = (unsigned)std::min<unsigned long long>(a.n_tasks, (unsigned long long)ctx->sm_count);
    const char *name = FUSED ? "k4_score_tiles_topx" : "k4_score_tiles_rows";
    auto kern = score_tiles_kernel<SIM, FUSED, STATS, UREG>;
    MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, name, kern, grid, warps * 32, smem, a);
}

// MYRIO_K4_COUNT=0 selects the add-as-you-go kernel (score_tiles_kernel); count mode is the default where it
// applies: u16 shared-key counters (no leaf vector of 65535+ keys) and an index whose postings are mostly
// unit-valued (odd terms are exact but slow: they flush the register counters)
static bool k4_count_mode(const myrio_index *ix) {
    static const int v = [] {
        const char *e = getenv("MYRIO_K4_COUNT");
        return e ? atoi(e) : -1;
    }();
    if (v == 0) return false;
    if (ix->max_leaf_nnz >= 65535) return false;  // u16 shared-key counters
    if (v > 0) return true;
    return ix->n_odd_postings * 50 <= ix->n_postings;  // <= 2 % odd postings
}

template <int SIM, bool FUSED, bool STATS>
static void launch_score_count_variant(myrio_ctx *ctx, const myrio_index *ix, ScoreArgs &a) {
    const size_t per_warp = score_count_warp_smem_bytes(ix->tile_leaves, FUSED);
    int warps = (int)std::min<size_t>(SCC_MAX_WARPS, (ctx->smem_optin - 1024) / per_warp);
    if (warps < 1) fail(MYRIO_ERR_BAD_ARG, "tile_leaves=%u does not fit in shared memory", ix->tile_leaves);
    const size_t smem = per_warp * warps;
    if (a.n_tasks >= 0xFFF00000ull) fail(MYRIO_ERR_BAD_ARG, "too many (query, tile) tasks in one batch");
    const unsigned grid = (unsigned)std::min<unsigned long long>(a.n_tasks, (unsigned long long)ctx->sm_count);
    const char *name = FUSED ? "k4_score_tiles_topx" : "k4_score_tiles_rows";
    auto kern = score_tiles_count_kernel<SIM, FUSED, STATS>;
    MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, name, kern, grid, warps * 32, smem, a);
}

template <bool FUSED>
static void launch_score(myrio_ctx *ctx, const myrio_index *ix, ScoreArgs &a, int32_t sim, bool stats) {
    if (k4_count_mode(ix)) {
        if (sim == MYRIO_SIM_COSINE) {
            if (stats) launch_score_count_variant<MYRIO_SIM_COSINE, FUSED, true>(ctx, ix, a);
            else launch_score_count_variant<MYRIO_SIM_COSINE, FUSED, false>(ctx, ix, a);
        } else {
            if (stats) launch_score_count_variant<MYRIO_SIM_JACARD_TANIMOTO, FUSED, true>(ctx, ix, a);
            else launch_score_count_variant<MYRIO_SIM_JACARD_TANIMOTO, FUSED, false>(ctx, ix, a);
        }
        return;
    }
    const bool ureg = k4_unit_in_registers();
#define MYRIO_K4_DISPATCH(SIMV)                                                         \
    do {                                                                                \
        if (stats) {                                                                    \
            if (ureg) launch_score_variant<SIMV, FUSED, true, true>(ctx, ix, a);        \
            else launch_score_variant<SIMV, FUSED, true, false>(ctx, ix, a);            \
        } else {                                                                        \
            if (ureg) launch_score_variant<SIMV, FUSED, false, true>(ctx, ix, a);       \
            else launch_score_variant<SIMV, FUSED, false, false>(ctx, ix, a);           \
        }                                                                               \
    } while (0)
    if (sim == MYRIO_SIM_COSINE)
        MYRIO_K4_DISPATCH(MYRIO_SIM_COSINE);
    else
        MYRIO_K4_DISPATCH(MYRIO_SIM_JACARD_TANIMOTO);
#undef MYRIO_K4_DISPATCH
}

// MYRIO_K4_UB=0: the fused top-x finishes EVERY score of every task (the full replay) instead of only those
// whose upper bound reaches the query's running bound (comparison runs; the result is the same)
static bool k4_bound_pruning() {
    static const bool v = [] {
        const char *e = getenv("MYRIO_K4_UB");
        return !(e && e[0] == '0');
    }();
    return v;
}

static ScoreArgs base_args(myrio_ctx *ctx, const myrio_index *ix, const Plan &pl, uint64_t qn) {
    ScoreArgs a{};
    a.tiles = ix->d_tiles.p;
    a.postings = ix->postings.p;
    a.plan = pl.hits.p;
    a.plan_off = pl.off.p;
    a.plan_stride = pl.stride;
    a.dmask = pl.mask_stride ? pl.dmask.p : nullptr;
    a.mask_stride = pl.mask_stride;
    a.qv1 = pl.qv1.p;
    a.q_count = (uint32_t)qn;
    a.tile_leaves = ix->tile_leaves;
    a.n_tiles = (uint32_t)ix->tiles.size();
    a.n_tasks = (unsigned long long)qn * ix->tiles.size();
    a.ld = ix->n_leaves;
    a.leaf_id_base = ix->leaf_id_base;
    a.counters = ctx->score_counters.p;
    a.ub_prune = (ctx->topx_bound_pruning < 0 ? k4_bound_pruning() : ctx->topx_bound_pruning != 0) ? 1u : 0u;
    return a;
}

static void reset_stats(myrio_ctx *ctx) {
    ctx->twophase_ready = false;  // every scoring entry point rebuilds the plan / candidate scratch
    ctx->last_stats[0] = ctx->last_stats[1] = ctx->last_stats[2] = 0;
    ctx->score_counters.reserve(4);
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
}

static void accumulate_stats(myrio_ctx *ctx) {
    unsigned long long h[4] = {0, 0, 0, 0};
    d2h(ctx, h, ctx->score_counters.p, 4);
    sync(ctx);
    ctx->last_stats[2] += h[3];
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
}

static constexpr uint64_t PLAN_BUDGET = 12ull << 30;  // bytes of hit lists kept at once

static bool topx_seeding() {  // MYRIO_K4_SEED=0 disables the seed pass (comparison runs)
    static const bool v = [] {
        const char *e = getenv("MYRIO_K4_SEED");
        return !(e && e[0] == '0');
    }();
    return v;
}

void score_all(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
               uint64_t q_count, int32_t sim, float *scores_host) {
    check_queries(queries, q_first, q_count);
    check_sim(sim);
    const uint64_t L = ix->n_leaves;
    reset_stats(ctx);
    if (q_count == 0 || L == 0) return;
    const uint64_t budget = 1ull << 31;  // bytes of score rows kept on the device at once
    uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, budget / (L * 4)), q_count);
    if (!ctx->plan) ctx->plan = plan_scratch_new();
    Plan &pl = *static_cast<Plan *>(ctx->plan);
    for (uint64_t q0 = 0; q0 < q_count;) {
        uint64_t qn = std::min(per, q_count - q0);
        while (!build_plan(ctx, ix, queries, q_first + q0, qn, PLAN_BUDGET, pl)) qn = (qn + 1) / 2;
        ctx->score_rows.reserve(qn * L);
        MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        ScoreArgs a = base_args(ctx, ix, pl, qn);
        a.scores = ctx->score_rows.p;
        launch_score<false>(ctx, ix, a, sim, true);
        d2h(ctx, scores_host + q0 * L, ctx->score_rows.p, qn * L);
        accumulate_stats(ctx);
        q0 += qn;
    }
    sync(ctx);
}

// Score rows of up to `q_max` queries starting at q_first, left ON THE DEVICE in ctx->score_rows ([qn][n_leaves],
// payload order) for the results pass (results.cu).  Returns the number of queries done (bounded by the 2 GB
// row budget and by what the hit lists allow).
uint64_t score_rows_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                           uint64_t q_max, int32_t sim) {
    check_queries(queries, q_first, q_max);
    check_sim(sim);
    const uint64_t L = ix->n_leaves;
    reset_stats(ctx);
    if (q_max == 0 || L == 0) return q_max;
    const uint64_t budget = 1ull << 31;
    uint64_t qn = std::min<uint64_t>(std::max<uint64_t>(1, budget / (L * 4)), q_max);
    if (!ctx->plan) ctx->plan = plan_scratch_new();
    Plan &pl = *static_cast<Plan *>(ctx->plan);
    while (!build_plan(ctx, ix, queries, q_first, qn, PLAN_BUDGET, pl)) qn = (qn + 1) / 2;
    ctx->score_rows.reserve(qn * L);
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
    ScoreArgs a = base_args(ctx, ix, pl, qn);
    a.scores = ctx->score_rows.p;
    launch_score<false>(ctx, ix, a, sim, false);
    return qn;
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
    const uint64_t budget = 8ull << 30;  // bytes of per-tile candidates kept at once
    const uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, budget / (n_tiles * x * 8)), q_count);
    // grow-only scratch of the context: a steady-state step never goes back to the allocator
    DevBuf<uint64_t> &cand = ctx->topx_cand;
    DevBuf<uint32_t> &cand_cnt = ctx->topx_cand_cnt, &gthr = ctx->topx_gthr;
    if (!ctx->plan) ctx->plan = plan_scratch_new();
    Plan &pl = *static_cast<Plan *>(ctx->plan);
    for (uint64_t q0 = 0; q0 < q_count;) {
        uint64_t qn = std::min(per, q_count - q0);
        const bool seeding = ix->tiles.size() > 1 && topx_seeding();
        while (!build_plan(ctx, ix, queries, q_first + q0, qn, PLAN_BUDGET, pl, seeding)) qn = (qn + 1) / 2;
        cand.reserve(qn * n_tiles * x);
        cand_cnt.reserve(qn * n_tiles);
        gthr.reserve(qn);
        MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(gthr.p, 0, qn * sizeof(uint32_t), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(cand_cnt.p, 0, qn * n_tiles * sizeof(uint32_t), ctx->stream));
        if (!ix->tiles.empty()) {
            ScoreArgs a = base_args(ctx, ix, pl, qn);
            a.x = x;
            a.cand = cand.p;
            a.cand_cnt = cand_cnt.p;
            a.gthr = gthr.p;
            if (seeding) {
                // pass 1: every query against its seed tile; pass 2: everything else, tile-major
                choose_seed_tiles(ctx, ix, pl, qn);
                ScoreArgs s1 = a;
                s1.seed = pl.seed.p;
                s1.seed_order = pl.has_seed_order ? pl.seed_order.p : nullptr;
                s1.seed_mode = 1;
                s1.n_tasks = qn;
                launch_score<true>(ctx, ix, s1, sim, collect_stats);
                MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
                a.seed = pl.seed.p;
                a.seed_mode = 2;
            }
            launch_score<true>(ctx, ix, a, sim, collect_stats);
        }
        MYRIO_LAUNCH(ctx, "k5_topx_merge_tiles", topx_cand_merge_kernel, (unsigned)qn, TX_THREADS, 0, cand.p,
                     cand_cnt.p, gthr.p, (uint32_t)ix->tiles.size(), x, d_top_scores + q0 * x,
                     d_top_ids + q0 * x, (uint64_t *)nullptr);
        if (collect_stats) accumulate_stats(ctx);
        q0 += qn;
    }
}

// ---- two-phase sharded top-x: seed pass | exchange of the bounds between ranks | sweep + merge
//
// A shard that does not hold a query's clade seeds a weak bound from its own tiles; the maximum of the
// ranks' bounds is still a lower bound of the query's GLOBAL x-th best score (x leaves of some tile of some
// rank reach it), so after an all-reduce(MAX) of Q x 4 bytes every rank sweeps with the strongest bound.
// The shard-local result may then hold fewer than x entries (padded): only the merged list is exact.
uint64_t score_topx_batch_capacity(const myrio_index *ix, uint32_t x) {
    const uint64_t n_tiles = std::max<uint64_t>(ix->tiles.size(), 1);
    return std::max<uint64_t>(1, (8ull << 30) / (n_tiles * x * 8));
}

void score_topx_seed_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                            uint64_t q_count, uint32_t x, int32_t sim, uint32_t *d_bound) {
    check_queries(queries, q_first, q_count);
    check_sim(sim);
    if (x == 0 || x > TX_MAX_X) fail(MYRIO_ERR_BAD_ARG, "x must be in 1..%d", TX_MAX_X);
    if (q_count > score_topx_batch_capacity(ix, x))
        fail(MYRIO_ERR_UNSUPPORTED, "two-phase top-x needs the queries in one batch (%llu > %llu)",
             (unsigned long long)q_count, (unsigned long long)score_topx_batch_capacity(ix, x));
    reset_stats(ctx);
    ctx->twophase_ready = false;
    if (q_count == 0) return;
    const uint64_t n_tiles = std::max<uint64_t>(ix->tiles.size(), 1);
    if (!ctx->plan) ctx->plan = plan_scratch_new();
    Plan &pl = *static_cast<Plan *>(ctx->plan);
    if (!build_plan(ctx, ix, queries, q_first, q_count, ~0ull, pl, ix->tiles.size() > 1))
        fail(MYRIO_ERR_UNSUPPORTED, "two-phase top-x: the hit lists do not fit");
    ctx->topx_cand.reserve(q_count * n_tiles * x);
    ctx->topx_cand_cnt.reserve(q_count * n_tiles);
    ctx->topx_gthr.reserve(q_count);
    MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
    MYRIO_CK(cudaMemsetAsync(ctx->topx_gthr.p, 0, q_count * sizeof(uint32_t), ctx->stream));
    MYRIO_CK(cudaMemsetAsync(ctx->topx_cand_cnt.p, 0, q_count * n_tiles * sizeof(uint32_t), ctx->stream));
    ctx->twophase_seeded = false;
    if (ix->tiles.size() > 1) {
        choose_seed_tiles(ctx, ix, pl, q_count);
        ScoreArgs s1 = base_args(ctx, ix, pl, q_count);
        s1.x = x;
        s1.cand = ctx->topx_cand.p;
        s1.cand_cnt = ctx->topx_cand_cnt.p;
        s1.gthr = ctx->topx_gthr.p;
        s1.seed = pl.seed.p;
        s1.seed_order = pl.has_seed_order ? pl.seed_order.p : nullptr;
        s1.seed_mode = 1;
        s1.n_tasks = q_count;
        launch_score<true>(ctx, ix, s1, sim, false);
        ctx->twophase_seeded = true;
    }
    MYRIO_CK(cudaMemcpyAsync(d_bound, ctx->topx_gthr.p, q_count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->twophase_ready = true;
    ctx->twophase_q = q_count;
    ctx->twophase_x = x;
    ctx->twophase_ix = ix;
    ctx->twophase_sim = sim;
}

void score_topx_sweep_device(myrio_ctx *ctx, const myrio_index *ix, uint64_t q_count, uint32_t x, int32_t sim,
                             const uint32_t *d_bound, uint64_t *d_top) {
    check_sim(sim);
    // the plan, the candidate buffers and the seed flags of the seed pass live in the context: any other
    // scoring call in between (it clears twophase_ready) or a different index / similarity invalidates them
    if (!ctx->twophase_ready || ctx->twophase_q != q_count || ctx->twophase_x != x || ctx->twophase_ix != ix ||
        ctx->twophase_sim != sim)
        fail(MYRIO_ERR_BAD_ARG, "myrio_score_topx_sweep_async must directly follow myrio_score_topx_seed_async with the same "
                                "index, similarity, x and query count");
    ctx->twophase_ready = false;
    if (q_count == 0) return;
    Plan &pl = *static_cast<Plan *>(ctx->plan);
    if (d_bound)  // the exchanged (raised) bounds; also the final filter of the merge
        MYRIO_CK(cudaMemcpyAsync(ctx->topx_gthr.p, d_bound, q_count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (!ix->tiles.empty()) {
        MYRIO_CK(cudaMemsetAsync(ctx->score_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        ScoreArgs a = base_args(ctx, ix, pl, q_count);
        a.x = x;
        a.cand = ctx->topx_cand.p;
        a.cand_cnt = ctx->topx_cand_cnt.p;
        a.gthr = ctx->topx_gthr.p;
        if (ctx->twophase_seeded) {
            a.seed = pl.seed.p;
            a.seed_mode = 2;
        }
        launch_score<true>(ctx, ix, a, sim, false);
    }
    MYRIO_LAUNCH(ctx, "k5_topx_merge_tiles", topx_cand_merge_kernel, (unsigned)q_count, TX_THREADS, 0,
                 ctx->topx_cand.p, ctx->topx_cand_cnt.p, ctx->topx_gthr.p, (uint32_t)ix->tiles.size(), x,
                 (float *)nullptr, (uint32_t *)nullptr, d_top);
}

void topx_merge_composites_device(myrio_ctx *ctx, const uint64_t *d_in, uint32_t n_parts, uint64_t q_count,
                                  uint32_t x, float *d_scores_out, uint32_t *d_ids_out) {
    if (x == 0 || n_parts == 0) fail(MYRIO_ERR_BAD_ARG, "x and n_parts must be > 0");
    uint64_t total = (uint64_t)n_parts * x, S = 2;
    while (S < total) S <<= 1;
    const size_t smem = S * sizeof(uint64_t);
    if (smem > ctx->smem_optin) fail(MYRIO_ERR_UNSUPPORTED, "n_parts*x=%llu too large to merge", (unsigned long long)total);
    if (q_count == 0) return;
    MYRIO_CK(cudaFuncSetAttribute(topx_merge_composites_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, "k7_topx_merge", topx_merge_composites_kernel, (unsigned)q_count, 256, smem, d_in, n_parts,
                 q_count, x, d_scores_out, d_ids_out);
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

// ------------------------------------------ direct (index-free) scoring

// Exact top-x of one score row in global memory, one CTA per query: 4 x 8-bit radix select on the
// f32 bit pattern (scores are >= +0), then the elements above the x-th best plus the first ties
// by ascending id are gathered and sorted by the composite (score desc, payload_id asc) --
// tax/results.rs:310-329.
__global__ void __launch_bounds__(TX_THREADS) topx_rows_kernel(const float *__restrict__ rows, uint64_t L,
                                                               uint32_t x, uint32_t id_base,
                                                               float *__restrict__ out_s,
                                                               uint32_t *__restrict__ out_i) {
    __shared__ uint64_t buf[TX_MAX_X];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_warp[TX_THREADS / 32];
    __shared__ uint32_t s_prefix, s_need, s_cursor, s_eq_base;
    const size_t q = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *row = reinterpret_cast<const uint32_t *>(rows + q * L);
    const uint32_t take = (uint32_t)min((uint64_t)x, L);
    uint32_t prefix = 0, need = take;
    if (L > x) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            hist[tid] = 0;  // TX_THREADS == 256
            __syncthreads();
            for (uint64_t i = tid; i < L; i += TX_THREADS) {
                const uint32_t b = row[i];
                if (pass == 0 || (b >> (shift + 8)) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t run = 0, bin = 255;
                for (;; --bin) {
                    if (run + hist[bin] >= need || bin == 0) break;
                    run += hist[bin];
                }
                s_prefix = (prefix << 8) | bin;
                s_need = need - run;
            }
            __syncthreads();
            prefix = s_prefix;
            need = s_need;
            __syncthreads();
        }
    }
    const uint32_t T = prefix, need_eq = (L > x) ? need : 0xFFFFFFFFu;  // L <= x: everything is taken
    if (tid == 0) {
        s_cursor = 0;
        s_eq_base = 0;
    }
    __syncthreads();
    for (uint64_t base = 0; base < L; base += TX_THREADS) {
        const uint64_t i = base + tid;
        const bool valid = i < L;
        const uint32_t b = valid ? row[i] : 0u;
        const bool gt = valid && (L <= x || b > T);
        const bool eq = valid && L > x && b == T;
        // ties are taken in ascending id order: ordered block-wide count of the eq flags
        const unsigned m_eq = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp[warp] = __popc(m_eq);
        __syncthreads();
        uint32_t before = s_eq_base;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        const bool take_eq = eq && (before + __popc(m_eq & ((1u << lane) - 1u)) < need_eq);
        if (gt || take_eq) buf[atomicAdd(&s_cursor, 1u)] = topx_composite(b, id_base + (uint32_t)i);
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (uint32_t w = 0; w < TX_THREADS / 32; ++w) tot += s_warp[w];
            s_eq_base += tot;
        }
        __syncthreads();
    }
    const uint32_t filled = s_cursor;  // == take
    uint32_t S = 2;
    while (S < filled) S <<= 1;
    for (uint32_t j = filled + tid; j < S; j += TX_THREADS) buf[j] = 0;
    bitonic_desc(buf, S);
    for (uint32_t j = tid; j < x; j += TX_THREADS) {
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

static void check_direct(const myrio_svecs *leaves, const myrio_svecs *queries, uint64_t q_first, uint64_t q_count,
                         int32_t sim) {
    check_queries(queries, q_first, q_count);
    if (leaves->vtype != MYRIO_VAL_F32) fail(MYRIO_ERR_BAD_ARG, "leaf vectors must be f32 (pre-normalised)");
    if (sim != MYRIO_SIM_COSINE && sim != MYRIO_SIM_JACARD_TANIMOTO && sim != MYRIO_SIM_OVERLAP)
        fail(MYRIO_ERR_BAD_ARG, "unknown similarity %d", sim);
    if (leaves->n_rows > 0xFFFFFFFFull) fail(MYRIO_ERR_BAD_ARG, "too many leaves");
}

// score rows of queries [q0, q0+qn) against all leaves into d_rows [qn][L]
static void direct_rows(myrio_ctx *ctx, const myrio_svecs *leaves, const myrio_svecs *queries, uint64_t q0,
                        uint64_t qn, int32_t sim, float *d_rows) {
    // rows of the kernel = leaves (consecutive threads walk consecutive leaves against one query)
    MergeArgs a{};
    a.r_off = leaves->row_off.p;
    a.r_keys = leaves->keys.p;
    a.r_vals = leaves->vals_f32.p;
    a.n_rows = (uint32_t)leaves->n_rows;
    a.c_off = queries->row_off.p;
    a.c_keys = queries->keys.p;
    a.c_vals = queries->vals_f32.p;
    a.c_first = (uint32_t)q0;
    a.n_cols = (uint32_t)qn;
    a.sim = sim;
    a.out = d_rows;
    a.ld = leaves->n_rows;
    a.transposed = true;  // out[query][leaf]
    launch_merge_pairs(ctx, a, "k4d_score_direct");
}

// the same through the merge-join path (every Similarity, incl. Overlap); rows left in ctx->score_rows
uint64_t score_rows_direct_device(myrio_ctx *ctx, const myrio_svecs *leaves, const myrio_svecs *queries,
                                  uint64_t q_first, uint64_t q_max, int32_t sim) {
    check_direct(leaves, queries, q_first, q_max, sim);
    const uint64_t L = leaves->n_rows;
    if (q_max == 0 || L == 0) return q_max;
    const uint64_t qn = std::min<uint64_t>(std::max<uint64_t>(1, (1ull << 31) / (L * 4)), q_max);
    ctx->score_rows.reserve(qn * L);
    direct_rows(ctx, leaves, queries, q_first, qn, sim, ctx->score_rows.p);
    return qn;
}

void score_all_direct(myrio_ctx *ctx, const myrio_svecs *leaves, const myrio_svecs *queries, uint64_t q_first,
                      uint64_t q_count, int32_t sim, float *scores_host) {
    check_direct(leaves, queries, q_first, q_count, sim);
    const uint64_t L = leaves->n_rows;
    if (q_count == 0 || L == 0) return;
    const uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, (1ull << 31) / (L * 4)), q_count);
    for (uint64_t q0 = 0; q0 < q_count; q0 += per) {
        const uint64_t qn = std::min(per, q_count - q0);
        ctx->score_rows.reserve(qn * L);
        direct_rows(ctx, leaves, queries, q_first + q0, qn, sim, ctx->score_rows.p);
        d2h(ctx, scores_host + q0 * L, ctx->score_rows.p, qn * L);
        sync(ctx);
    }
}

void score_topx_direct_device(myrio_ctx *ctx, const myrio_svecs *leaves, uint32_t leaf_id_base,
                              const myrio_svecs *queries, uint64_t q_first, uint64_t q_count, uint32_t x,
                              int32_t sim, float *d_top_scores, uint32_t *d_top_ids) {
    check_direct(leaves, queries, q_first, q_count, sim);
    if (x == 0 || x > TX_MAX_X) fail(MYRIO_ERR_BAD_ARG, "x must be in 1..%d", TX_MAX_X);
    const uint64_t L = leaves->n_rows;
    if (q_count == 0) return;
    const uint64_t per = std::min<uint64_t>(std::max<uint64_t>(1, (1ull << 31) / (std::max<uint64_t>(L, 1) * 4)), q_count);
    for (uint64_t q0 = 0; q0 < q_count; q0 += per) {
        const uint64_t qn = std::min(per, q_count - q0);
        ctx->score_rows.reserve(std::max<uint64_t>(qn * L, 1));
        if (L) direct_rows(ctx, leaves, queries, q_first + q0, qn, sim, ctx->score_rows.p);
        MYRIO_LAUNCH(ctx, "k5_topx_rows", topx_rows_kernel, (unsigned)qn, TX_THREADS, 0, ctx->score_rows.p, L, x,
                     leaf_id_base, d_top_scores + q0 * x, d_top_ids + q0 * x);
    }
}

// the plan scratch of a context (allocated lazily, grow-only)
void plan_scratch_free(void *p) { delete static_cast<Plan *>(p); }
void *plan_scratch_new() { return new Plan(); }

}  // namespace myrio
