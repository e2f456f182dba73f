// index.cu -- K3: build the tiled CSR inverted index (k-mer key -> posting list of
// (leaf, normalised count)) from the per-leaf sorted sparse vectors.
//
// Replaces the `Box<[SFVec]>` of L2-normalised leaf vectors that
// TaxTreeCompute::from_store_tree builds (myrio-core/src/tax/compute.rs:100-118)
// as the structure the score loop (tax/results.rs:82-93) runs against.
//
// Leaves are cut into tiles of `tile_leaves` consecutive payload ids (the scoring
// kernel keeps one tile's accumulators in shared memory).  Per tile: stable LSD
// radix sort of (key, (leaf,value)) pairs -> posting lists ordered by leaf inside
// each key; run heads -> unique keys + offsets; open-addressing dictionary
// key -> (begin, len).
#include <algorithm>

#include "index.cuh"
#include "kmer_count.cuh"

namespace myrio {

__global__ void __launch_bounds__(256) make_pairs_kernel(const uint64_t *__restrict__ row_off,
                                                         const uint64_t *__restrict__ keys,
                                                         const float *__restrict__ vals,
                                                         uint64_t n_rows, uint32_t tile_leaves,
                                                         uint64_t *__restrict__ pkeys,
                                                         uint64_t *__restrict__ pvals,
                                                         unsigned long long *__restrict__ key_or) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = row_off[row], e = row_off[row + 1];
    const uint64_t leaf_local = row % tile_leaves;
    uint64_t acc = 0;
    for (uint64_t i = b + lane; i < e; i += 32) {
        const uint64_t k = keys[i];
        acc |= k;
        pkeys[i] = k;
        pvals[i] = (leaf_local << 32) | (uint64_t)__float_as_uint(vals[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0 && acc) atomicOr(key_or, (unsigned long long)acc);
}

__global__ void __launch_bounds__(256) mark_heads_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                         uint32_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
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

// unit[leaf] = smallest normalised value of the leaf's row (count == common count)
__global__ void __launch_bounds__(256) row_min_kernel(const uint64_t *__restrict__ row_off,
                                                      const float *__restrict__ vals, uint64_t n_rows,
                                                      float *__restrict__ unit) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned lane = threadIdx.x & 31;
    float m = 3.402823466e+38f;
    for (uint64_t i = row_off[row] + lane; i < row_off[row + 1]; i += 32) m = fminf(m, vals[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) unit[row] = (row_off[row + 1] > row_off[row]) ? m : 0.0f;
}

// A unique key is DENSE when it occurs in at least dense_min leaves of the tile and every
// one of its postings carries exactly unit[leaf]; flags[u] = 1 then.
__global__ void __launch_bounds__(256) classify_dense_kernel(const uint64_t *__restrict__ postings,
                                                             const uint32_t *__restrict__ ubegin,
                                                             uint64_t n_unique, const float *__restrict__ unit,
                                                             uint32_t dense_min, uint32_t *__restrict__ flags) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const uint32_t b = ubegin[u], e = ubegin[u + 1];
    uint32_t dense = (e - b) >= dense_min ? 1u : 0u;
    for (uint32_t i = b; dense && i < e; ++i) {
        const uint64_t p = postings[i];
        if ((uint32_t)p != __float_as_uint(unit[(uint32_t)(p >> 32)])) dense = 0u;
    }
    flags[u] = dense;
}

// one warp per dense key: fill its 32-word bitmap (word = leaf & 31, bit = leaf >> 5)
__global__ void __launch_bounds__(256) fill_bitmaps_kernel(const uint64_t *__restrict__ postings,
                                                           const uint32_t *__restrict__ ubegin,
                                                           uint64_t n_unique, const uint32_t *__restrict__ flags,
                                                           const uint64_t *__restrict__ didx,
                                                           uint32_t *__restrict__ bitmaps) {
    const uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n_unique || !flags[u]) return;
    const unsigned lane = threadIdx.x & 31;
    uint32_t *bm = bitmaps + didx[u] * 32;
    uint32_t w = 0;
    // lane l owns word l: scan the key's postings and keep the leaves with leaf & 31 == l
    for (uint32_t i = ubegin[u]; i < ubegin[u + 1]; ++i) {
        const uint32_t leaf = (uint32_t)(postings[i] >> 32);
        if ((leaf & 31u) == lane) w |= 1u << (leaf >> 5);
    }
    bm[lane] = w;
}

__global__ void __launch_bounds__(256) dict_insert_kernel(const uint64_t *__restrict__ ukeys,
                                                          const uint32_t *__restrict__ ubegin,
                                                          uint64_t n_unique, const uint32_t *__restrict__ flags,
                                                          const uint64_t *__restrict__ didx, DictSlot *dict,
                                                          uint32_t mask) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    const uint64_t key = ukeys[u];
    uint32_t begin = ubegin[u], len = ubegin[u + 1] - begin;
    if (flags && flags[u]) {
        begin = (uint32_t)didx[u];
        len |= DENSE_FLAG;
    }
    uint32_t h = dict_hash(key) & mask;
    for (;;) {
        // claim an empty slot through its `len` word (len != 0 for every real entry)
        if (atomicCAS(&dict[h].len, 0u, len) == 0u) {
            dict[h].key = key;
            dict[h].begin = begin;
            return;
        }
        h = (h + 1) & mask;
    }
}

myrio_index *index_build(myrio_ctx *ctx, const myrio_svecs *leaves, uint32_t leaf_id_base,
                         uint32_t tile_leaves) {
    if (leaves->vtype != MYRIO_VAL_F32)
        fail(MYRIO_ERR_BAD_ARG, "index_build needs normalised f32 leaf vectors (myrio_svecs_normalize)");
    if (tile_leaves == 0) tile_leaves = 1024;
    if (tile_leaves > 49152) fail(MYRIO_ERR_BAD_ARG, "tile_leaves must be <= 49152");
    if (leaves->n_rows >= 0xFFFFFFFFull) fail(MYRIO_ERR_BAD_ARG, "too many leaves for u32 payload ids");
    auto ix = std::make_unique<myrio_index>();
    const uint64_t L = leaves->n_rows, P = leaves->nnz;
    ix->n_leaves = L;
    ix->leaf_id_base = leaf_id_base;
    ix->tile_leaves = tile_leaves;
    ix->n_postings = P;
    const uint64_t n_tiles = (L + tile_leaves - 1) / tile_leaves;
    ix->tiles.resize(n_tiles);
    ix->dicts.resize(n_tiles);
    ix->ukeys.resize(n_tiles);
    ix->ubegin.resize(n_tiles);
    if (L == 0) return ix.release();

    // tile boundaries in posting space
    std::vector<uint64_t> bounds(n_tiles + 1);
    for (uint64_t t = 0; t <= n_tiles; ++t) {
        const uint64_t row = std::min<uint64_t>(t * tile_leaves, L);
        d2h(ctx, &bounds[t], leaves->row_off.p + row, 1);
    }
    DevBuf<uint64_t> pk(std::max<uint64_t>(P, 1)), pv(std::max<uint64_t>(P, 1)),
        pk2(std::max<uint64_t>(P, 1)), pv2(std::max<uint64_t>(P, 1));
    DevBuf<unsigned long long> key_or(1);
    MYRIO_CK(cudaMemsetAsync(key_or.p, 0, sizeof(unsigned long long), ctx->stream));
    MYRIO_LAUNCH(ctx, "k3_make_pairs", make_pairs_kernel, (unsigned)((L + 7) / 8), 256, 0,
                 leaves->row_off.p, leaves->keys.p, leaves->vals_f32.p, L, tile_leaves, pk.p, pv.p,
                 key_or.p);
    unsigned long long h_or = 0;
    d2h(ctx, &h_or, key_or.p, 1);
    sync(ctx);
    int key_bits = 0;
    while (key_bits < 64 && (h_or >> key_bits) != 0) ++key_bits;
    if (key_bits == 0) key_bits = 1;

    ix->unit.alloc(L);
    MYRIO_LAUNCH(ctx, "k3_row_min", row_min_kernel, (unsigned)((L + 7) / 8), 256, 0, leaves->row_off.p,
                 leaves->vals_f32.p, L, ix->unit.p);
    ix->bitmaps.resize(n_tiles);
    const bool dense_ok = tile_leaves <= DENSE_MAX_TILE;

    DevBuf<uint32_t> hist, flags, dflags;
    DevBuf<uint64_t> offs, uidx, didx;
    bool in_tmp = false;
    uint64_t max_tile_post = 0;
    for (uint64_t t = 0; t < n_tiles; ++t) max_tile_post = std::max(max_tile_post, bounds[t + 1] - bounds[t]);
    flags.alloc(std::max<uint64_t>(max_tile_post, 1));
    uidx.alloc(max_tile_post + 1);

    for (uint64_t t = 0; t < n_tiles; ++t) {
        const uint64_t pb = bounds[t], np = bounds[t + 1] - pb;
        TileInfo &ti = ix->tiles[t];
        ti.post_begin = pb;
        ti.n_post = np;
        ti.leaf_begin = (uint32_t)(t * tile_leaves);
        ti.n_leaves = (uint32_t)(std::min<uint64_t>(L, (t + 1) * tile_leaves) - t * tile_leaves);
        if (np >= 0xFFFFFFFFull) fail(MYRIO_ERR_UNSUPPORTED, "tile with >= 2^32 postings; lower tile_leaves");
        in_tmp = radix_sort_pairs_raw(ctx, pk.p + pb, pv.p + pb, pk2.p + pb, pv2.p + pb, np, key_bits,
                                      hist, offs);
        const uint64_t *sk = (in_tmp ? pk2.p : pk.p) + pb;
        uint64_t U = 0;
        if (np) {
            const unsigned g = (unsigned)((np + 255) / 256);
            MYRIO_LAUNCH(ctx, "k3_mark_heads", mark_heads_kernel, g, 256, 0, sk, np, flags.p);
            exclusive_scan_u32_to_u64(ctx, flags.p, uidx.p, np);
            d2h(ctx, &U, uidx.p + np, 1);
            sync(ctx);
            ix->ukeys[t].alloc(U);
            ix->ubegin[t].alloc(U + 1);
            MYRIO_LAUNCH(ctx, "k3_emit_unique", emit_unique_kernel, g, 256, 0, sk, np, flags.p, uidx.p,
                         ix->ukeys[t].p, ix->ubegin[t].p);
        } else {
            ix->ubegin[t].alloc(1);
            MYRIO_CK(cudaMemsetAsync(ix->ubegin[t].p, 0, sizeof(uint32_t), ctx->stream));
        }
        // dense keys -> bitmaps
        uint64_t n_dense = 0;
        const uint64_t *sv = (in_tmp ? pv2.p : pv.p) + pb;
        if (dense_ok && U) {
            const uint32_t dense_min = std::max<uint32_t>(32u, ti.n_leaves / 4u);
            dflags.reserve(U);
            didx.reserve(U + 1);
            MYRIO_LAUNCH(ctx, "k3_classify_dense", classify_dense_kernel, (unsigned)((U + 255) / 256), 256, 0, sv,
                         ix->ubegin[t].p, U, ix->unit.p + ti.leaf_begin, dense_min, dflags.p);
            exclusive_scan_u32_to_u64(ctx, dflags.p, didx.p, U);
            d2h(ctx, &n_dense, didx.p + U, 1);
            sync(ctx);
            if (n_dense) {
                ix->bitmaps[t].alloc(n_dense * 32);
                MYRIO_LAUNCH(ctx, "k3_fill_bitmaps", fill_bitmaps_kernel, (unsigned)((U + 7) / 8), 256, 0, sv,
                             ix->ubegin[t].p, U, dflags.p, didx.p, ix->bitmaps[t].p);
            }
        }
        uint32_t cap = 16;
        while (cap < 2 * U) cap <<= 1;
        ix->dicts[t].alloc(cap);
        MYRIO_CK(cudaMemsetAsync(ix->dicts[t].p, 0, (size_t)cap * sizeof(DictSlot), ctx->stream));
        if (U)
            MYRIO_LAUNCH(ctx, "k3_dict_insert", dict_insert_kernel, (unsigned)((U + 255) / 256), 256, 0,
                         ix->ukeys[t].p, ix->ubegin[t].p, U, n_dense ? dflags.p : nullptr, didx.p,
                         ix->dicts[t].p, cap - 1);
        ti.n_unique = U;
        ti.dict = ix->dicts[t].p;
        ti.dict_mask = cap - 1;
        ti.ukeys = ix->ukeys[t].p;
        ti.ubegin = ix->ubegin[t].p;
        ti.bitmaps = ix->bitmaps[t].p;
        ti.unit = ix->unit.p + ti.leaf_begin;
        ti.n_dense = n_dense;
        ix->n_dense_total += n_dense;
        ix->n_unique_total += U;
    }
    sync(ctx);
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
        leaf[i] = (uint32_t)(p >> 32);
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
        for (uint64_t i = 0; i <= ti.n_unique; ++i) offs[i] = tmp[i];
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
