// comm.cuh -- multi-GPU plumbing inside the library (comm.cu): NCCL bound at run time
#pragma once
#include "index.cuh"

namespace myrio {
void comm_unique_id(uint8_t *id128);
void comm_init(myrio_ctx *ctx, uint32_t nranks, uint32_t rank, const uint8_t *id128);
void comm_destroy(myrio_ctx *ctx);
int comm_nccl_version();
// d_buf = `world` chunks of bytes_per_rank (device memory); chunk `rank` filled on entry, all on return
void comm_allgather_inplace(myrio_ctx *ctx, void *d_buf, uint64_t bytes_per_rank);
void comm_allreduce_max_u32(myrio_ctx *ctx, uint32_t *d_buf, uint64_t n);
void score_topx_sharded(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries, uint64_t q_first,
                        uint64_t q_count, uint32_t x, int32_t sim, float *d_top_scores, uint32_t *d_top_ids,
                        uint64_t max_batch);
}  // namespace myrio
