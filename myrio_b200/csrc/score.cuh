// score.cuh -- scoring entry points (score.cu)
#pragma once
#include "index.cuh"

namespace myrio {
void score_all(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
               uint64_t q_count, int32_t sim, float *scores_host);
uint64_t score_rows_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                           uint64_t q_max, int32_t sim);
uint64_t score_rows_direct_device(myrio_ctx *ctx, const myrio_svecs *leaves, const myrio_svecs *queries,
                                  uint64_t q_first, uint64_t q_max, int32_t sim);
void score_topx_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                       uint64_t q_first, uint64_t q_count, uint32_t x, int32_t sim, float *d_top_scores,
                       uint32_t *d_top_ids, bool collect_stats);
void topx_merge_device(myrio_ctx *ctx, const float *d_scores_in, const uint32_t *d_ids_in, uint32_t n_parts,
                       uint64_t q_count, uint32_t x, float *d_scores_out, uint32_t *d_ids_out);
// index-free scoring by merge-join over the normalised leaf vectors (all similarities, incl. Overlap)
void score_all_direct(myrio_ctx *ctx, const myrio_svecs *leaves, const myrio_svecs *queries, uint64_t q_first,
                      uint64_t q_count, int32_t sim, float *scores_host);
void score_topx_direct_device(myrio_ctx *ctx, const myrio_svecs *leaves, uint32_t leaf_id_base,
                              const myrio_svecs *queries, uint64_t q_first, uint64_t q_count, uint32_t x,
                              int32_t sim, float *d_top_scores, uint32_t *d_top_ids);
// two-phase sharded top-x (seed pass | exchange of bounds | sweep) and the merge of composite lists
void score_topx_seed_device(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                            uint64_t q_count, uint32_t x, int32_t sim, uint32_t *d_bound);
void score_topx_sweep_device(myrio_ctx *ctx, const myrio_index *ix, uint64_t q_count, uint32_t x, int32_t sim,
                             const uint32_t *d_bound, uint64_t *d_top);
void topx_merge_composites_device(myrio_ctx *ctx, const uint64_t *d_in, uint32_t n_parts, uint64_t q_count,
                                  uint32_t x, float *d_scores_out, uint32_t *d_ids_out);
// queries per two-phase batch on this shard (the per-tile candidate buffer is capped at 8 GB)
uint64_t score_topx_batch_capacity(const myrio_index *ix, uint32_t x);
void *plan_scratch_new();
void plan_scratch_free(void *p);
}  // namespace myrio
