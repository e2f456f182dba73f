// myrtree.cuh -- host-side content of a .myrtree file (myrtree.cu)
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "common.cuh"

namespace myrio {
struct StoreNode {   // one node of the taxonomy in PRE-ORDER (the order bincode nests them, tax/core.rs:355-372)
    uint8_t kind;    // 0 = Node::Branch, 1 = Node::Leaf
    std::string name;
    uint64_t arg;    // branch: number of children; leaf: payload_id
};
struct StoreCounts {  // kmer_store_counts_vec[i] of every payload, as one CSR batch (payload order)
    std::vector<uint64_t> row_off, keys;
    std::vector<uint16_t> vals;
    std::vector<float> rescale;
};
}  // namespace myrio

struct myrio_store {  // TaxTreeStore (tax/store.rs:43-48) minus the file path
    std::string gene;
    uint32_t root_rank = 1;  // Rank discriminant (tax/clade.rs:93-104): Domain = 8 ... Species = 1
    std::vector<myrio::StoreNode> nodes;
    uint64_t n_payloads = 0;
    // StorePayload::seq of every payload: 4-bit packed, every sequence on an even word (16-byte aligned)
    std::vector<uint64_t> seq_words, seq_word_off, seq_base_off;
    std::vector<uint64_t> k_precomputed;
    std::vector<myrio::StoreCounts> counts;  // one per k_precomputed entry
};

namespace myrio {
unsigned store_default_threads();
std::vector<uint8_t> store_to_bytes(const myrio_store &st, int zstd_level, unsigned threads);
myrio_store *store_from_bytes(const uint8_t *data, uint64_t len, unsigned threads);
myrio_seqs *store_leaves_to_device(myrio_ctx *ctx, const myrio_store &st);
myrio_svecs *store_counts_to_device(myrio_ctx *ctx, const myrio_store &st, uint64_t k_index);
void store_put_counts(myrio_ctx *ctx, myrio_store &st, uint64_t k, const myrio_svecs *counts, bool overwrite);
}  // namespace myrio
