// index.cu -- K3: build the tiled inverted index (k-mer key -> posting list of
// (leaf, normalised count)) from the per-leaf sorted sparse vectors.
//
// Replaces the `Box<[SFVec]>` of L2-normalised leaf vectors that
// TaxTreeCompute::from_store_tree builds (myrio-core/src/tax/compute.rs:100-118)
// as the structure the score loop (tax/results.rs:82-93) runs against.
//
// Leaves are cut into tiles of `tile_leaves` consecutive payload ids (the scoring
// kernel keeps one tile's accumulators on chip).  The whole index is built at once,
// whatever the number of tiles:
//   1. (key, (leaf, value)) pairs in leaf order;
//   2. ONE stable LSD radix sort: by key, then by tile -> postings grouped by tile and,
//      inside a tile, ordered by (key, leaf);
//   3. run heads -> unique (tile, key) entries + absolute posting offsets;
//   4. unit[leaf] = smallest value of the leaf; a key that occurs in >= 1/4 of a tile's
//      leaves (HOT) -- or, on an index that will be scored in count mode, in >= 32 leaves (cold) --
//      always with value == unit[leaf], is DENSE: one 128-byte bitmap; hot keys are numbered first
//      inside a tile (their bitmap indexes are the bits of the plan's dense-key masks);
//   5. the (tile, key) entries re-sorted by key + ONE open-addressing dictionary
//      key -> [EntryRec{tile, begin, len | DENSE_FLAG | ORDERED_FLAG}] (tiles ascending): a query key is
//      looked up once per query, not once per tile;
//   6. postings whose value is not unit[leaf] get POSTING_ODD, the sparse entries that hold one ORDERED_FLAG.
#include <algorithm>

#include "index.cuh"
#include "kmer_count.cuh"

namespace myrio {

// payload = (shard-local leaf << 32) | value bits; key_or collects the key bits in use
__global__ void __launch_bounds__(256) make_pairs_kernel(const uint64_t *__restrict__ row_off,
                                                         const uint64_t *__restrict__ keys,
                                                         const float *__restrict__ vals,
                                                         uint64_t n_rows, uint64_t *__restrict__ pkeys,
                                                         uint64_t *__restrict__ pvals,
                                                         unsigned long long *__restrict__ key_or,
                                                         float *__restrict__ unit) {
    // key_or[0]: OR of all keys; key_or[1]: longest row
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = row_off[row], e = row_off[row + 1];
    uint64_t acc = 0;
    float m = 3.402823466e+38f;
    for (uint64_t i = b + lane; i < e; i += 32) {
        const uint64_t k = keys[i];
        const float v = vals[i];
        acc |= k;
        m = fminf(m, v);
        pkeys[i] = k;
        pvals[i] = (row << 32) | (uint64_t)__float_as_uint(v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc |= __shfl_xor_sync(0xffffffffu, acc, d);
        m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
    }
    if (lane == 0) {
        if (acc) atomicOr(key_or, (unsigned long long)acc);
        atomicMax(key_or + 1, (unsigned long long)(e - b));
        unit[row] = e > b ? m : 0.0f;  // the leaf's smallest normalised value (count == common count)
    }
}

__global__ void __launch_bounds__(256) gather_u64_kernel(const uint64_t *__restrict__ src,
                                                         const uint64_t *__restrict__ idx, uint64_t n,
                                                         uint64_t *__restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// tile of a posting position / unique-entry index: last t with bounds[t] <= pos
__device__ __forceinline__ uint32_t tile_of(const uint64_t *__restrict__ bounds, uint32_t n_tiles, uint64_t pos) {
    uint32_t lo = 0, hi = n_tiles;  // invariant: bounds[lo] <= pos < bounds[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (bounds[mid] <= pos)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// heads of the key runs (tile starts are added by mark_tile_starts_kernel)
__global__ void __launch_bounds__(256) mark_heads_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                         uint32_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) mark_tile_starts_kernel(const uint64_t *__restrict__ bounds,
                                                               uint32_t n_tiles, uint64_t n,
                                                               uint32_t *__restrict__ flags) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_tiles && bounds[t] < n) flags[bounds[t]] = 1u;
}

__global__ void __launch_bounds__(256) localize_leaves_kernel(uint64_t *__restrict__ payload, uint64_t n,
                                                              uint32_t tile_leaves) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t p = payload[i];
    payload[i] = ((uint64_t)((uint32_t)(p >> 32) % tile_leaves) << 32) | (p & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(256) emit_unique_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                          const uint32_t *__restrict__ flags,
                                                          const uint64_t *__restrict__ uidx,
                                                          uint64_t *__restrict__ ukeys,
                                                          uint32_t *__restrict__ ubegin) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) {
        const uint64_t u = uidx[i];
        ukeys[u] = keys[i];
        ubegin[u] = (uint32_t)i;
    }
    if (i == 0) ubegin[uidx[n]] = (uint32_t)n;
}

// A unique (tile, key) entry is DENSE when the key occurs in at least min_df leaves of the tile and every one of
// its postings carries exactly unit[leaf]: it is then also stored as a 128-byte bitmap.  flags: 2 = HOT (in at
// least max(32, n_leaves/4) leaves: the plan keeps such keys as mask bits), 1 = cold dense, 0 = sparse.
// min_df = 0 means "hot only" (the add-as-you-go kernel pays 32 slots per dense key whatever its df; the
// count-mode kernel pays three logic instructions: measured best from 32 postings on).
__global__ void __launch_bounds__(256) classify_dense_kernel(const uint64_t *__restrict__ postings,
                                                             const uint32_t *__restrict__ ubegin,
                                                             uint64_t n_unique,
                                                             const uint64_t *__restrict__ ustart,
                                                             uint32_t n_tiles, uint64_t n_leaves,
                                                             uint32_t tile_leaves,
                                                             const float *__restrict__ unit, uint32_t min_df,
                                                             uint32_t *__restrict__ hot, uint32_t *__restrict__ cold) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const uint32_t b = ubegin[u], e = ubegin[u + 1];
    uint32_t cls = 0;
    if (e - b >= (min_df ? min(min_df, 32u) : 32u)) {
        const uint32_t t = tile_of(ustart, n_tiles, u);
        const uint64_t leaf0 = (uint64_t)t * tile_leaves;
        const uint32_t nl = (uint32_t)min((uint64_t)tile_leaves, n_leaves - leaf0);
        const uint32_t hot_df = max(32u, nl / 4u);
        if (e - b >= hot_df)
            cls = 2;
        else if (min_df && e - b >= min_df)
            cls = 1;
        if (cls) {
            const float *un = unit + leaf0;
            for (uint32_t i = b; cls && i < e; ++i) {
                const uint64_t p = postings[i];
                if ((uint32_t)p != __float_as_uint(un[(uint32_t)(p >> 32)])) cls = 0u;
            }
        }
    }
    hot[u] = cls == 2u;
    cold[u] = cls == 1u;
}

// bitmap index of every dense entry: inside a tile the hot keys come first (their indexes are the bit positions
// of the plan's dense-key masks), then the cold ones; tiles follow one another.  dflags[u] = dense at all.
__global__ void __launch_bounds__(256) dense_index_kernel(const uint32_t *__restrict__ hot,
                                                          const uint32_t *__restrict__ cold,
                                                          const uint64_t *__restrict__ hidx,
                                                          const uint64_t *__restrict__ cidx, uint64_t n_unique,
                                                          const uint64_t *__restrict__ ustart, uint32_t n_tiles,
                                                          uint32_t *__restrict__ dflags, uint64_t *__restrict__ didx) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const uint32_t t = tile_of(ustart, n_tiles, u);
    const uint64_t us = ustart[t], ue = ustart[t + 1];
    const uint64_t hot0 = hidx[us], cold0 = cidx[us], n_hot = hidx[ue] - hot0;
    dflags[u] = hot[u] | cold[u];
    didx[u] = hot0 + cold0 + (hot[u] ? hidx[u] - hot0 : n_hot + (cidx[u] - cold0));
}

// number of postings whose value is not unit[leaf] (what flag_odd_postings_kernel will flag), counted before the
// dense classification: it decides whether the index is scored in count mode, hence how low the dense bar is
__global__ void __launch_bounds__(256) count_odd_postings_kernel(const uint64_t *__restrict__ postings, uint64_t n,
                                                                 const uint64_t *__restrict__ bounds, uint32_t n_tiles,
                                                                 uint32_t tile_leaves, const float *__restrict__ unit,
                                                                 unsigned long long *__restrict__ n_odd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool odd = false;
    if (i < n) {
        const uint64_t p = postings[i];
        const uint32_t t = tile_of(bounds, n_tiles, i);
        odd = (uint32_t)p != __float_as_uint(unit[(uint64_t)t * tile_leaves + (uint32_t)(p >> 32)]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, odd);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_odd, (unsigned long long)__popc(m));
}

// postings whose value is not unit[leaf] get POSTING_ODD (see index.cuh); run last, on the final posting array
__global__ void __launch_bounds__(256) flag_odd_postings_kernel(uint64_t *__restrict__ postings, uint64_t n,
                                                                const uint64_t *__restrict__ bounds, uint32_t n_tiles,
                                                                uint32_t tile_leaves, const float *__restrict__ unit,
                                                                unsigned long long *__restrict__ n_odd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool odd = false;
    if (i < n) {
        const uint64_t p = postings[i];
        const uint32_t t = tile_of(bounds, n_tiles, i);
        const uint64_t leaf = (uint64_t)t * tile_leaves + (uint32_t)(p >> 32);
        odd = (uint32_t)p != __float_as_uint(unit[leaf]);
        if (odd) postings[i] = p | POSTING_ODD;
    }
    const unsigned m = __ballot_sync(0xffffffffu, odd);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_odd, (unsigned long long)__popc(m));
}

// a sparse (tile, key) entry whose posting list holds an odd posting gets ORDERED_FLAG (index.cuh); run after
// flag_odd_postings_kernel.  One thread per entry (lists are ~12 postings long on average).
__global__ void __launch_bounds__(256) flag_ordered_entries_kernel(EntryRec *__restrict__ entries, uint64_t n_unique,
                                                                   const uint64_t *__restrict__ postings,
                                                                   const uint64_t *__restrict__ bounds) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_unique) return;
    const EntryRec e = entries[j];
    if (e.len & DENSE_FLAG) return;  // every posting of a dense key carries unit[leaf]
    const uint64_t *p = postings + bounds[e.tile] + e.begin;
    uint64_t any = 0;
    for (uint32_t i = 0; i < (e.len & LEN_MASK); ++i) any |= p[i];
    if (any & POSTING_ODD) entries[j].len = e.len | ORDERED_FLAG;
}

// one warp per dense entry: lane l builds word l (leaves with leaf & 31 == l, bit = leaf >> 5)
__global__ void __launch_bounds__(256) fill_bitmaps_kernel(const uint64_t *__restrict__ postings,
                                                           const uint32_t *__restrict__ ubegin,
                                                           uint64_t n_unique, const uint32_t *__restrict__ flags,
                                                           const uint64_t *__restrict__ didx,
                                                           uint32_t *__restrict__ bitmaps) {
    const uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n_unique || !flags[u]) return;
    const unsigned lane = threadIdx.x & 31;
    uint32_t w = 0;
    for (uint32_t i = ubegin[u]; i < ubegin[u + 1]; ++i) {
        const uint32_t leaf = (uint32_t)(postings[i] >> 32);
        if ((leaf & 31u) == lane) w |= 1u << (leaf >> 5);
    }
    bitmaps[didx[u] * 32 + lane] = w;
}

// entry index u (tile-major, key ascending inside a tile) -> EntryRec
__global__ void __launch_bounds__(256) make_entries_kernel(const uint64_t *__restrict__ ent_u, uint64_t n_unique,
                                                           const uint32_t *__restrict__ ubegin,
                                                           const uint32_t *__restrict__ dflags,
                                                           const uint64_t *__restrict__ didx,
                                                           const uint64_t *__restrict__ ustart,
                                                           const uint64_t *__restrict__ bounds,
                                                           const uint64_t *__restrict__ dstart, uint32_t n_tiles,
                                                           EntryRec *__restrict__ entries) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_unique) return;
    const uint64_t u = ent_u[j];
    const uint32_t t = tile_of(ustart, n_tiles, u);
    EntryRec e;
    e.tile = t;
    e.len = ubegin[u + 1] - ubegin[u];
    if (dflags[u]) {
        e.begin = (uint32_t)(didx[u] - dstart[t]);
        e.len |= DENSE_FLAG;
    } else {
        e.begin = (uint32_t)(ubegin[u] - bounds[t]);
    }
    entries[j] = e;
}

// heads of the key-sorted entry list -> dictionary slots key -> (first entry, number of tiles)
__global__ void __launch_bounds__(256) gdict_insert_kernel(const uint64_t *__restrict__ ent_keys, uint64_t n_unique,
                                                           const uint32_t *__restrict__ flags,
                                                           const uint64_t *__restrict__ hidx,
                                                           const uint32_t *__restrict__ hbegin, DictSlot *dict,
                                                           uint32_t mask) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_unique || !flags[j]) return;
    const uint64_t key = ent_keys[j];
    const uint64_t h0 = hidx[j];
    const uint32_t len = hbegin[h0 + 1] - hbegin[h0];
    uint32_t h = dict_hash(key) & mask;
    for (;;) {
        // claim an empty slot through its `len` word (len != 0 for every real entry)
        if (atomicCAS(&dict[h].len, 0u, len) == 0u) {
            dict[h].key = key;
            dict[h].begin = (uint32_t)j;
            return;
        }
        h = (h + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) iota_kernel(uint64_t *__restrict__ v, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

myrio_index *index_build(myrio_ctx *ctx, const myrio_svecs *leaves, uint32_t leaf_id_base,
                         uint32_t tile_leaves) {
    if (leaves->vtype != MYRIO_VAL_F32)
        fail(MYRIO_ERR_BAD_ARG, "index_build needs normalised f32 leaf vectors (myrio_svecs_normalize)");
    if (tile_leaves == 0) tile_leaves = 1024;
    if (tile_leaves > 49152) fail(MYRIO_ERR_BAD_ARG, "tile_leaves must be <= 49152");
    if (leaves->n_rows >= 0xFFFFFFFFull) fail(MYRIO_ERR_BAD_ARG, "too many leaves for u32 payload ids");
    if (leaves->nnz >= 0xFFFFFFFFull)
        fail(MYRIO_ERR_UNSUPPORTED, "an index shard holds at most 2^32-1 postings (got %llu): shard the leaves",
             (unsigned long long)leaves->nnz);
    auto ix = std::make_unique<myrio_index>();
    const uint64_t L = leaves->n_rows, P = leaves->nnz;
    ix->n_leaves = L;
    ix->leaf_id_base = leaf_id_base;
    ix->tile_leaves = tile_leaves;
    ix->n_postings = P;
    const uint64_t n_tiles = (L + tile_leaves - 1) / tile_leaves;
    ix->tiles.resize(n_tiles);
    if (L == 0) return ix.release();

    // tile boundaries in posting space (the sort keeps the number of postings per tile)
    std::vector<uint64_t> rows(n_tiles + 1), bounds(n_tiles + 1);
    for (uint64_t t = 0; t <= n_tiles; ++t) rows[t] = std::min<uint64_t>(t * tile_leaves, L);
    DevBuf<uint64_t> d_rows(n_tiles + 1), d_bounds(n_tiles + 1);
    h2d(ctx, d_rows.p, rows.data(), n_tiles + 1);
    MYRIO_LAUNCH(ctx, "k3_gather", gather_u64_kernel, (unsigned)((n_tiles + 256) / 256), 256, 0,
                 leaves->row_off.p, d_rows.p, n_tiles + 1, d_bounds.p);
    d2h(ctx, bounds.data(), d_bounds.p, n_tiles + 1);

    // 1. pairs + unit values
    DevBuf<uint64_t> pk(std::max<uint64_t>(P, 1)), pv(std::max<uint64_t>(P, 1)),
        pk2(std::max<uint64_t>(P, 1)), pv2(std::max<uint64_t>(P, 1));
    DevBuf<unsigned long long> key_or(3);  // [0] OR of the keys, [1] longest leaf vector, [2] odd postings
    ix->unit.alloc(L);
    MYRIO_CK(cudaMemsetAsync(key_or.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
    MYRIO_LAUNCH(ctx, "k3_make_pairs", make_pairs_kernel, (unsigned)((L + 7) / 8), 256, 0,
                 leaves->row_off.p, leaves->keys.p, leaves->vals_f32.p, L, pk.p, pv.p, key_or.p,
                 ix->unit.p);
    unsigned long long h_or2[2] = {0, 0};
    d2h(ctx, h_or2, key_or.p, 2);
    std::vector<float> h_unit(L);  // per-tile maxima (TileInfo::unit_max)
    d2h(ctx, h_unit.data(), ix->unit.p, L);
    sync(ctx);
    const unsigned long long h_or = h_or2[0];
    ix->max_leaf_nnz = h_or2[1];
    int key_bits = 0;
    while (key_bits < 64 && (h_or >> key_bits) != 0) ++key_bits;
    if (key_bits == 0) key_bits = 1;

    // 2. one stable sort: by key, then by tile
    DevBuf<uint32_t> hist;
    DevBuf<uint64_t> offs;
    const bool in_tmp = radix_sort_pairs_raw(ctx, pk.p, pv.p, pk2.p, pv2.p, P, key_bits, hist, offs,
                                             tile_leaves, n_tiles);
    uint64_t *sk = in_tmp ? pk2.p : pk.p, *sv = in_tmp ? pv2.p : pv.p;
    hist.release();
    offs.release();
    (in_tmp ? pk : pk2).release();  // the other key buffer is no longer needed

    uint64_t U = 0, n_dense = 0;
    std::vector<uint64_t> ustart(n_tiles + 1, 0), dstart(n_tiles + 1, 0);
    std::vector<uint32_t> n_hot_of(n_tiles, 0);  // hot dense keys per tile (the plan's mask bits)
    DevBuf<uint64_t> d_ustart(n_tiles + 1);
    DevBuf<uint32_t> dflags;
    DevBuf<uint64_t> didx;
    if (P) {
        // 3. run heads -> unique entries
        const unsigned g = (unsigned)((P + 255) / 256);
        DevBuf<uint32_t> flags(P);
        DevBuf<uint64_t> uidx(P + 1);
        MYRIO_LAUNCH(ctx, "k3_mark_heads", mark_heads_kernel, g, 256, 0, sk, P, flags.p);
        MYRIO_LAUNCH(ctx, "k3_mark_tile_starts", mark_tile_starts_kernel, (unsigned)((n_tiles + 255) / 256), 256,
                     0, d_bounds.p, (uint32_t)n_tiles, P, flags.p);
        MYRIO_LAUNCH(ctx, "k3_localize_leaves", localize_leaves_kernel, g, 256, 0, sv, P, tile_leaves);
        exclusive_scan_u32_to_u64(ctx, flags.p, uidx.p, P);
        MYRIO_LAUNCH(ctx, "k3_gather", gather_u64_kernel, (unsigned)((n_tiles + 256) / 256), 256, 0, uidx.p,
                     d_bounds.p, n_tiles + 1, d_ustart.p);
        d2h(ctx, ustart.data(), d_ustart.p, n_tiles + 1);
        sync(ctx);
        U = ustart[n_tiles];
        ix->ukeys.alloc(U);
        ix->ubegin.alloc(U + 1);
        MYRIO_LAUNCH(ctx, "k3_emit_unique", emit_unique_kernel, g, 256, 0, sk, P, flags.p, uidx.p, ix->ukeys.p,
                     ix->ubegin.p);
        // 4. dense entries -> bitmaps
        dflags.alloc(U);
        didx.alloc(U + 1);
        DevBuf<uint64_t> d_dstart(n_tiles + 1);
        std::vector<uint64_t> hstart(n_tiles + 1, 0);
        if (tile_leaves <= DENSE_MAX_TILE) {
            // will the index be scored in count mode (score.cu: <= 2 % odd postings, short leaf vectors)?
            unsigned long long h_odd0 = 0;
            MYRIO_CK(cudaMemsetAsync(key_or.p + 2, 0, sizeof(unsigned long long), ctx->stream));
            MYRIO_LAUNCH(ctx, "k3_count_odd", count_odd_postings_kernel, g, 256, 0, sv, P, d_bounds.p,
                         (uint32_t)n_tiles, tile_leaves, ix->unit.p, key_or.p + 2);
            d2h(ctx, &h_odd0, key_or.p + 2, 1);
            sync(ctx);
            MYRIO_CK(cudaMemsetAsync(key_or.p + 2, 0, sizeof(unsigned long long), ctx->stream));
            static const uint32_t cold_df = [] {  // MYRIO_DENSE_MIN_DF: 0 = hot keys only (round-1 behaviour)
                const char *e = getenv("MYRIO_DENSE_MIN_DF");
                return e ? (uint32_t)atoi(e) : 32u;
            }();
            const bool count_mode = h_odd0 * 50 <= P && ix->max_leaf_nnz < 65535;
            DevBuf<uint32_t> hot(U), cold(U);
            DevBuf<uint64_t> hidx(U + 1), cidx(U + 1), d_hstart(n_tiles + 1);
            MYRIO_LAUNCH(ctx, "k3_classify_dense", classify_dense_kernel, (unsigned)((U + 255) / 256), 256, 0, sv,
                         ix->ubegin.p, U, d_ustart.p, (uint32_t)n_tiles, L, tile_leaves, ix->unit.p,
                         count_mode ? cold_df : 0u, hot.p, cold.p);
            exclusive_scan_u32_to_u64(ctx, hot.p, hidx.p, U);
            exclusive_scan_u32_to_u64(ctx, cold.p, cidx.p, U);
            MYRIO_LAUNCH(ctx, "k3_dense_index", dense_index_kernel, (unsigned)((U + 255) / 256), 256, 0, hot.p, cold.p,
                         hidx.p, cidx.p, U, d_ustart.p, (uint32_t)n_tiles, dflags.p, didx.p);
            // per tile: first bitmap (hot + cold before the tile) and number of hot keys
            MYRIO_LAUNCH(ctx, "k3_gather", gather_u64_kernel, (unsigned)((n_tiles + 256) / 256), 256, 0, hidx.p,
                         d_ustart.p, n_tiles + 1, d_hstart.p);
            MYRIO_LAUNCH(ctx, "k3_gather", gather_u64_kernel, (unsigned)((n_tiles + 256) / 256), 256, 0, cidx.p,
                         d_ustart.p, n_tiles + 1, d_dstart.p);
            std::vector<uint64_t> cstart(n_tiles + 1, 0);
            d2h(ctx, hstart.data(), d_hstart.p, n_tiles + 1);
            d2h(ctx, cstart.data(), d_dstart.p, n_tiles + 1);
            sync(ctx);
            for (uint64_t t = 0; t <= n_tiles; ++t) dstart[t] = hstart[t] + cstart[t];
        } else {
            MYRIO_CK(cudaMemsetAsync(dflags.p, 0, U * sizeof(uint32_t), ctx->stream));
            MYRIO_CK(cudaMemsetAsync(didx.p, 0, (U + 1) * sizeof(uint64_t), ctx->stream));
            sync(ctx);
        }
        n_dense = dstart[n_tiles];
        n_hot_of.assign(n_tiles, 0);
        for (uint64_t t = 0; t < n_tiles; ++t) n_hot_of[t] = (uint32_t)(hstart[t + 1] - hstart[t]);
        if (n_dense) {
            ix->bitmaps.alloc(n_dense * 32);
            MYRIO_LAUNCH(ctx, "k3_fill_bitmaps", fill_bitmaps_kernel, (unsigned)((U + 7) / 8), 256, 0, sv,
                         ix->ubegin.p, U, dflags.p, didx.p, ix->bitmaps.p);
        }
    } else {
        ix->ubegin.alloc(1);
        MYRIO_CK(cudaMemsetAsync(ix->ubegin.p, 0, sizeof(uint32_t), ctx->stream));
    }
    ix->n_unique_total = U;
    ix->n_dense_total = n_dense;

    // 5. index-wide key map: entries re-sorted by key (stable: tiles stay ascending inside a key),
    //    one EntryRec per (tile, key), and a dictionary key -> entry range
    if (U) {
        DevBuf<uint64_t> ek(U), eu(U), ek2(U), eu2(U), d_dstart(n_tiles + 1);
        h2d(ctx, d_dstart.p, dstart.data(), n_tiles + 1);
        MYRIO_CK(cudaMemcpyAsync(ek.p, ix->ukeys.p, U * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        MYRIO_LAUNCH(ctx, "k3_iota", iota_kernel, (unsigned)((U + 255) / 256), 256, 0, eu.p, U);
        DevBuf<uint32_t> hist2;
        DevBuf<uint64_t> offs2;
        const bool tmp2 = radix_sort_pairs_raw(ctx, ek.p, eu.p, ek2.p, eu2.p, U, key_bits, hist2, offs2);
        const uint64_t *sk2 = tmp2 ? ek2.p : ek.p, *su2 = tmp2 ? eu2.p : eu.p;
        ix->gentries.alloc(U);
        MYRIO_LAUNCH(ctx, "k3_make_entries", make_entries_kernel, (unsigned)((U + 255) / 256), 256, 0, su2, U,
                     ix->ubegin.p, dflags.p, didx.p, d_ustart.p, d_bounds.p, d_dstart.p, (uint32_t)n_tiles,
                     ix->gentries.p);
        DevBuf<uint32_t> hflags(U), hbegin;
        DevBuf<uint64_t> hidx(U + 1), hkeys;
        MYRIO_LAUNCH(ctx, "k3_mark_heads", mark_heads_kernel, (unsigned)((U + 255) / 256), 256, 0, sk2, U, hflags.p);
        exclusive_scan_u32_to_u64(ctx, hflags.p, hidx.p, U);
        uint64_t Ug = 0;
        d2h(ctx, &Ug, hidx.p + U, 1);
        sync(ctx);
        hkeys.alloc(Ug);
        hbegin.alloc(Ug + 1);
        MYRIO_LAUNCH(ctx, "k3_emit_unique", emit_unique_kernel, (unsigned)((U + 255) / 256), 256, 0, sk2, U, hflags.p,
                     hidx.p, hkeys.p, hbegin.p);
        uint64_t cap = 16;
        while (cap < 4 * Ug) cap <<= 1;
        if (cap > 0x80000000ull) fail(MYRIO_ERR_UNSUPPORTED, "too many distinct keys for the index dictionary");
        ix->gdict.alloc(cap);
        ix->gdict_mask = (uint32_t)(cap - 1);
        ix->n_unique_keys = Ug;
        MYRIO_CK(cudaMemsetAsync(ix->gdict.p, 0, cap * sizeof(DictSlot), ctx->stream));
        MYRIO_LAUNCH(ctx, "k3_gdict_insert", gdict_insert_kernel, (unsigned)((U + 255) / 256), 256, 0, sk2, U,
                     hflags.p, hidx.p, hbegin.p, ix->gdict.p, ix->gdict_mask);
        sync(ctx);
    } else {
        ix->gdict.alloc(16);
        ix->gdict_mask = 15;
        MYRIO_CK(cudaMemsetAsync(ix->gdict.p, 0, 16 * sizeof(DictSlot), ctx->stream));
    }
    for (uint64_t t = 0; t < n_tiles; ++t) {
        TileInfo &ti = ix->tiles[t];
        ti.post_begin = bounds[t];
        ti.n_post = bounds[t + 1] - bounds[t];
        ti.n_unique = ustart[t + 1] - ustart[t];
        ti.leaf_begin = (uint32_t)(t * tile_leaves);
        ti.n_leaves = (uint32_t)(std::min<uint64_t>(L, (t + 1) * tile_leaves) - t * tile_leaves);
        ti.ukeys = ix->ukeys.p + ustart[t];
        ti.ubegin = ix->ubegin.p + ustart[t];
        ti.bitmaps = ix->bitmaps.p + dstart[t] * 32;
        ti.unit = ix->unit.p + ti.leaf_begin;
        ti.n_dense = dstart[t + 1] - dstart[t];
        ti.unit_max = *std::max_element(h_unit.begin() + ti.leaf_begin, h_unit.begin() + ti.leaf_begin + ti.n_leaves);
    }
    {
        std::vector<uint32_t> moff(n_tiles + 1, 0);
        for (uint64_t t = 0; t < n_tiles; ++t) {
            ix->tiles[t].mask_off = moff[t];
            ix->tiles[t].mask_words = (n_hot_of[t] + 31) / 32;  // hot keys first: their bitmap indexes are the bits
            moff[t + 1] = moff[t] + ix->tiles[t].mask_words;
        }
        ix->mask_words_total = moff[n_tiles];
        ix->mask_off.alloc(n_tiles + 1);
        h2d(ctx, ix->mask_off.p, moff.data(), n_tiles + 1);
        sync(ctx);  // moff is a local
    }
    if (P) {
        MYRIO_LAUNCH(ctx, "k3_flag_odd", flag_odd_postings_kernel, (unsigned)((P + 255) / 256), 256, 0, sv, P,
                     d_bounds.p, (uint32_t)n_tiles, tile_leaves, ix->unit.p, key_or.p + 2);
        unsigned long long h_odd = 0;
        d2h(ctx, &h_odd, key_or.p + 2, 1);
        sync(ctx);
        ix->n_odd_postings = h_odd;
        if (h_odd && U)
            MYRIO_LAUNCH(ctx, "k3_flag_ordered", flag_ordered_entries_kernel, (unsigned)((U + 255) / 256), 256, 0,
                         ix->gentries.p, U, sv, d_bounds.p);
    }
    ix->postings = std::move(in_tmp ? pv2 : pv);
    ix->d_tiles.alloc(n_tiles);
    h2d(ctx, ix->d_tiles.p, ix->tiles.data(), n_tiles);
    sync(ctx);
    return ix.release();
}

__global__ void __launch_bounds__(256) export_postings_kernel(const uint64_t *__restrict__ postings,
                                                              uint64_t n, uint32_t *__restrict__ leaf,
                                                              float *__restrict__ val) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t p = postings[i];
        leaf[i] = (uint32_t)(p >> 32) & POSTING_LEAF_MASK;
        val[i] = __uint_as_float((uint32_t)p);
    }
}

void index_export_tile(myrio_ctx *ctx, const myrio_index *ix, uint64_t tile, uint64_t *n_unique,
                       uint64_t *n_postings, uint64_t *ukeys, uint64_t *offs, uint32_t *post_leaf,
                       float *post_val) {
    if (tile >= ix->tiles.size()) fail(MYRIO_ERR_BAD_ARG, "tile %llu out of range", (unsigned long long)tile);
    const TileInfo &ti = ix->tiles[tile];
    if (n_unique) *n_unique = ti.n_unique;
    if (n_postings) *n_postings = ti.n_post;
    if (ukeys) d2h(ctx, ukeys, ti.ukeys, ti.n_unique);
    if (offs) {
        std::vector<uint32_t> tmp(ti.n_unique + 1);
        d2h(ctx, tmp.data(), ti.ubegin, ti.n_unique + 1);
        sync(ctx);
        for (uint64_t i = 0; i <= ti.n_unique; ++i) offs[i] = tmp[i] - ti.post_begin;  // tile-relative
        if (ti.n_unique == 0) offs[0] = 0;
    }
    if ((post_leaf || post_val) && ti.n_post) {
        DevBuf<uint32_t> dl(ti.n_post);
        DevBuf<float> dv(ti.n_post);
        MYRIO_LAUNCH(ctx, "export_postings", export_postings_kernel, (unsigned)((ti.n_post + 255) / 256),
                     256, 0, ix->postings.p + ti.post_begin, ti.n_post, dl.p, dv.p);
        if (post_leaf) d2h(ctx, post_leaf, dl.p, ti.n_post);
        if (post_val) d2h(ctx, post_val, dv.p, ti.n_post);
        sync(ctx);
    }
    sync(ctx);
}

}  // namespace myrio
