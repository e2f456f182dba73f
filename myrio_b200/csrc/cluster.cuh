// cluster.cuh -- clustering entry points (cluster.cu)
#pragma once
#include "common.cuh"

namespace myrio {
void cluster_reads(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *p,
                   const myrio_svecs *init, const myrio_shard *shard, int32_t *assign, uint32_t *n_iters,
                   float *silhouette);
myrio_svecs *cluster_centroid(myrio_ctx *ctx, const myrio_svecs *counts, const uint64_t *member_ids,
                              uint64_t m, const float *d_row_scale = nullptr);
myrio_svecs *reweight_fingerprints(myrio_ctx *ctx, const myrio_svecs *local, const myrio_svecs *naive_query,
                                   int32_t sim, float alpha);
void pairwise(myrio_ctx *ctx, const myrio_svecs *rows, const myrio_svecs *cols, int32_t sim, float *out);
}  // namespace myrio
