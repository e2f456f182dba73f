// internal entry points of the library (one per translation unit)
#pragma once
#include <memory>

#include "common.cuh"

namespace myrio {

// kmer_count.cu
void upload_phred_table(myrio_ctx *ctx);
myrio_svecs *count_reads(myrio_ctx *ctx, const myrio_seqs *reads, uint32_t k, float cutoff);
myrio_svecs *count_leaves(myrio_ctx *ctx, const myrio_seqs *leaves, uint32_t k, uint32_t nb_resamples,
                          uint64_t seed, uint64_t leaf_id_base);

// ingest.cu
myrio_seqs *reads_from_fastq(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint8_t *qual_ascii,
                             const uint64_t *base_off, uint64_t n, uint64_t min_length, float min_mean_qual,
                             uint8_t max_qual, uint8_t *keep_host);
myrio_seqs *leaves_from_fasta(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint64_t *base_off, uint64_t n,
                              uint8_t *keep_host);

// microbench.cu
void bench_smem_bandwidth(myrio_ctx *ctx, double *gbps, double *ms);
void bench_clock_probe(myrio_ctx *ctx, uint32_t slot);
void bench_clock_read(myrio_ctx *ctx, uint32_t n, double *mhz);

// svecs.cu
myrio_svecs *svecs_normalize(myrio_ctx *ctx, const myrio_svecs *in);

// index.cu / score.cu / cluster.cu: see their own headers

}  // namespace myrio
