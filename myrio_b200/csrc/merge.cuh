// merge.cuh -- the reference's own pairwise algorithm on the device: a two-pointer merge-join
// over two sorted sparse vectors (data/sparse.rs:317-360 dot, :143-221 merge_and_apply), for
// every Similarity (similarity.rs:40-51).  It is the only form in which Overlap can be computed
// bit-exactly (its sum of maxima runs over the UNION of the keys in key order), and an
// index-free cross-check of Cosine / JacardTanimoto.
#pragma once
#include "common.cuh"

namespace myrio {

struct MergeArgs {
    // rows / columns: CSR batches with optional id lists (NULL = identity)
    const uint64_t *r_off, *r_keys;
    const float *r_vals;
    const uint32_t *row_ids;
    uint32_t n_rows;
    const uint64_t *c_off, *c_keys;
    const float *c_vals;
    const uint64_t *c_end;  // optional: end of every column (NULL = c_off[c + 1])
    const uint32_t *col_ids;
    uint32_t c_first, n_cols;
    int sim;
    float *out;  // transposed: out[c * ld + r]; row-major: out[r * ld + c]   (c relative to c_first)
    uint64_t ld;
    bool transposed;
};

// out = simfunc(rows_i, cols_j) for all pairs; one thread per pair
void launch_merge_pairs(myrio_ctx *ctx, const MergeArgs &a, const char *name);

}  // namespace myrio
