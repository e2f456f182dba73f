// index.cuh -- the tiled inverted index over the reference leaves (K3 output).
#pragma once
#include "common.cuh"

namespace myrio {

// Slot of the index-wide key dictionary: open addressing, linear probing, len == 0 <=> empty.
// key -> entries[begin .. begin+len): one EntryRec per tile that holds the key, tiles ascending.
struct __align__(16) DictSlot {
    uint64_t key;
    uint32_t begin;
    uint32_t len;
};
// One (tile, key) entry.  Sparse key: begin = first posting (relative to the tile's posting
// range), len = df.  Dense key (DENSE_FLAG set in len): begin = index of the key's 32-word
// bitmap in the tile's bitmap array, len & ~DENSE_FLAG = df.
struct EntryRec {
    uint32_t tile;
    uint32_t begin;
    uint32_t len;
};
static constexpr uint32_t DENSE_FLAG = 0x80000000u;
// bit 30 of EntryRec::len / HitRec::len: the term needs its place in key order -- a sparse posting list that
// holds a POSTING_ODD posting (EntryRec, set by the index build) or, in a hit list, also a query value that is not
// the query's common value (HitRec, set by the plan kernels).  A hit list without this bit consists of common
// terms only, which commute: the count-mode scoring kernel then takes its dense and sparse keys in two sweeps.
static constexpr uint32_t ORDERED_FLAG = 0x40000000u;
static constexpr uint32_t LEN_MASK = 0x3FFFFFFFu;  // df <= tile_leaves <= 49152
// dense keys need lane-owned accumulators: leaf = 32*r + lane with r < 32
static constexpr uint32_t DENSE_MAX_TILE = 1024;

struct TileInfo {
    uint64_t post_begin;  // offset of the tile's postings in the global posting array
    uint64_t n_post;
    uint64_t n_unique;
    uint32_t leaf_begin;  // first leaf (shard-local payload index) of the tile
    uint32_t n_leaves;
    const uint64_t *ukeys;    // sorted unique keys of the tile        (export / tests)
    const uint32_t *ubegin;   // ABSOLUTE first posting of each unique key, +1 sentinel
    // dense keys: a posting whose value equals unit[leaf] (the leaf's smallest normalised
    // value, i.e. count == the common count) is one bit: word `leaf & 31`, bit `leaf >> 5`
    const uint32_t *bitmaps;  // [n_dense][32]
    const float *unit;        // [n_leaves] (slice of the index-wide array)
    uint64_t n_dense;
    float unit_max;           // largest unit value of the tile (upper bounds of the count-mode top-x)
    // dense-key masks of the plan: a query's common dense keys in this tile are bits of mask_words words
    // (bit = the key's bitmap index) at word mask_off of the query's mask row (index-wide: mask_words_total)
    uint32_t mask_off, mask_words;
};

__device__ __forceinline__ uint32_t dict_hash(uint64_t key) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    return (uint32_t)key;
}

// posting = (value bits, tile-local leaf) packed as one 8-byte word
struct __align__(8) Posting {
    float val;
    uint32_t leaf;
};
// bit 63 of a posting word: its value differs from unit[leaf] (a k-mer that occurs more than once in the
// leaf, or a resampled ambiguity window).  The count-mode scoring kernel treats such postings one by one, in
// key order; all others contribute the same product per (query, leaf) and are merely COUNTED.
static constexpr uint64_t POSTING_ODD = 1ull << 63;
static constexpr uint32_t POSTING_LEAF_MASK = 0x7FFFFFFFu;

}  // namespace myrio

struct myrio_index {
    uint64_t n_leaves = 0;
    uint32_t leaf_id_base = 0;
    uint32_t tile_leaves = 0;
    uint64_t n_postings = 0, n_unique_total = 0;
    myrio::DevBuf<uint64_t> postings;  // Posting[n_postings]
    std::vector<myrio::TileInfo> tiles;
    myrio::DevBuf<myrio::TileInfo> d_tiles;
    // index-wide arrays; every TileInfo points at its slice
    myrio::DevBuf<myrio::DictSlot> gdict;     // key -> entry range
    uint32_t gdict_mask = 0;
    myrio::DevBuf<myrio::EntryRec> gentries;  // [n_unique_total], grouped by key, tiles ascending
    uint64_t n_unique_keys = 0;
    myrio::DevBuf<uint64_t> ukeys;    // [n_unique_total]
    myrio::DevBuf<uint32_t> ubegin;   // [n_unique_total + 1], absolute posting offsets
    myrio::DevBuf<uint32_t> bitmaps;  // [n_dense_total][32]
    myrio::DevBuf<float> unit;        // [n_leaves]
    uint64_t n_dense_total = 0;
    uint32_t mask_words_total = 0;    // words of one query's dense-key mask row (sum of the tiles' mask_words)
    myrio::DevBuf<uint32_t> mask_off; // [n_tiles + 1] (device copy for the plan kernels)
    // myrio_score_topx_sharded: queries per exchange batch agreed by all ranks for (x, world) -- see comm.cu
    mutable uint32_t agreed_x = 0, agreed_world = 0;
    mutable uint64_t agreed_cap = 0;
    uint64_t n_odd_postings = 0;   // postings flagged POSTING_ODD
    uint64_t max_leaf_nnz = 0;     // longest leaf vector (the u16 shared-key counters of count mode need < 65535)
};

namespace myrio {
myrio_index *index_build(myrio_ctx *ctx, const myrio_svecs *leaves, uint32_t leaf_id_base,
                         uint32_t tile_leaves);
void index_export_tile(myrio_ctx *ctx, const myrio_index *ix, uint64_t tile, uint64_t *n_unique,
                       uint64_t *n_postings, uint64_t *ukeys, uint64_t *offs, uint32_t *post_leaf,
                       float *post_val);
bool radix_sort_pairs_raw(myrio_ctx *ctx, uint64_t *keys, uint64_t *vals, uint64_t *keys_tmp,
                          uint64_t *vals_tmp, size_t n, int key_bits, DevBuf<uint32_t> &hist,
                          DevBuf<uint64_t> &offs, uint32_t group_div = 0, uint64_t n_groups = 0);
}  // namespace myrio
