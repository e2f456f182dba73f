/*
 * myrio_b200.h -- C ABI of the B200-native (sm_100a) implementation of Myrio's
 * data-parallel hot path: k-mer counting of reads and reference leaves, the
 * inverted index over the leaves, query x leaf similarity scoring with top-x,
 * and the read-clustering similarity kernels.
 *
 * The reference (anesthetice/Myrio v0.3.0) has no FFI; its boundary is the Rust
 * API of `myrio-core` (myrio-core/src/prelude.rs:1-12).  Each entry point below
 * names the reference function whose body it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the Rust `extern "C"` block and the
 * call-site changes; rust/myrio-b200-sys holds that crate as source.
 *
 * Conventions
 *   - plain C types; every function returns a myrio_status (0 = ok);
 *     myrio_last_error() gives a thread-local message.  No unwinding crosses
 *     the ABI.
 *   - all *host* pointers are caller-owned for the duration of the call;
 *     results with a size known up front are caller-allocated, results with
 *     a data-dependent size live in opaque device-resident handles
 *     (myrio_svecs, myrio_index) and are fetched with *_info + *_download.
 *   - a myrio_ctx is bound to one GPU and one stream and is NOT thread-safe;
 *     distinct contexts may be used concurrently.  Calls return after the
 *     work has completed unless the name ends in `_async`: those leave their
 *     kernels queued on the context's stream (results are valid after
 *     myrio_ctx_sync or after later work on that stream); the scoring ones may
 *     still synchronise the stream internally while they size their hit lists.
 *   - scratch comes from the device's default stream-ordered memory pool, whose
 *     release threshold myrio_ctx_create raises so that freed scratch stays cached
 *     (process-wide for that device; MYRIO_POOL_RELEASE_THRESHOLD_MB bounds it).
 *   - there is no CPU fallback: without a CUDA device myrio_ctx_create fails.
 *
 * Encodings (bio-seq 0.14.2, the reference's sequence crate):
 *   Seq<Dna>   2 bits/base  A=0 C=1 G=2 T=3, base i at bits 2*(i%32) of u64 word i/32
 *   Seq<Iupac> 4 bits/base  A=8 C=4 G=2 T=1, ambiguity = OR, gap 'X' = 0,
 *                           symbol i at bits 4*(i%16) of u64 word i/16
 *   k-mer key  usize::from(&Kmer<Dna,K>): base j of the window at bits 2j..2j+1
 *   In the packed inputs below every sequence starts at an EVEN word index
 *   (16-byte aligned) so the kernels can use 128-bit loads.
 */
#ifndef MYRIO_B200_H
#define MYRIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only these entry points are exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum {
    MYRIO_OK = 0,
    MYRIO_ERR_INVALID_K = 1,   /* k outside 2..32: the reference panics (myrio-proc/src/lib.rs:70) */
    MYRIO_ERR_BAD_ARG = 2,
    MYRIO_ERR_CUDA = 3,
    MYRIO_ERR_OOM = 4,
    MYRIO_ERR_BAD_QUALITY = 5, /* Phred >= 128: table index out of bounds in the reference */
    MYRIO_ERR_UNSUPPORTED = 6,
    MYRIO_ERR_EMPTY = 7,       /* cluster(): no read survives t2 / no centroid (reference panics) */
    MYRIO_ERR_NONFINITE = 8,   /* SimScore::try_new(..).unwrap() would panic (clustering.rs:125) */
    MYRIO_ERR_BAD_BASE = 9,    /* a non-ACGT base: read_fastq fails the whole file (io/mod.rs:46-47) */
    MYRIO_ERR_COMM = 10,       /* NCCL failure, or another rank of a sharded call failed */
    MYRIO_ERR_FORMAT = 11      /* malformed .myrtree bytes (tax/store.rs:408-454 returns Error) */
} myrio_status;

/* myrio-core/src/similarity.rs:26-33 */
typedef enum { MYRIO_SIM_COSINE = 0, MYRIO_SIM_JACARD_TANIMOTO = 1, MYRIO_SIM_OVERLAP = 2 } myrio_similarity;

typedef enum { MYRIO_VAL_F32 = 0, MYRIO_VAL_U16 = 1 } myrio_vtype;

typedef struct myrio_ctx myrio_ctx;       /* one GPU + stream + scratch                     */
typedef struct myrio_seqs myrio_seqs;     /* device-resident packed sequences (+ quality)   */
typedef struct myrio_svecs myrio_svecs;   /* device-resident batch of SparseVec<T> (CSR)    */
typedef struct myrio_index myrio_index;   /* device-resident tiled inverted index of leaves */
typedef struct myrio_store myrio_store;   /* host-side content of a .myrtree file (TaxTreeStore)  */

const char *myrio_last_error(void);
int32_t myrio_version(void);

int32_t myrio_ctx_create(int32_t device, myrio_ctx **out);
void myrio_ctx_destroy(myrio_ctx *ctx);
int32_t myrio_ctx_sync(myrio_ctx *ctx);
/* the context's cudaStream_t (as void*), so callers can record their own events on it */
void *myrio_ctx_stream(const myrio_ctx *ctx);
/* number of kernels this library has launched on the context so far */
uint64_t myrio_ctx_launch_count(const myrio_ctx *ctx);
/* Integer value of a k-mer key (`usize::from(&Kmer<Dna, K>)` of the pinned bio-seq fork, Cargo.toml:23):
 * 0 (default) = base j of the window at bits 2j..2j+1 (Seq is a BitVec<usize, Lsb0>); 1 = base j at bits
 * 2(k-1-j).  The dependency is not vendored in the reference tree, so the order is an option rather than a
 * constant; it decides the stored .myrtree keys and, through the sort order, the f32 summation order.
 * Applies to the myrio_count_* calls made on this context afterwards. */
int32_t myrio_ctx_set_kmer_bit_order(myrio_ctx *ctx, int32_t msb_first);
/* Fused top-x (myrio_score_topx*): 1 (default) = a (query, leaf) score is FINISHED only if its upper bound
 * reaches the query's running x-th best score (every pair's shared keys are still counted; the top-x is
 * bit-identical either way); 0 = every score of every pair is finished, as myrio_score_all does.  A
 * measurement / comparison switch; the environment variable MYRIO_K4_UB=0 sets the default to 0. */
int32_t myrio_ctx_set_topx_bound_pruning(myrio_ctx *ctx, int32_t on);
/* optional per-kernel CUDA-event timing (events on the context's stream) */
int32_t myrio_prof_enable(myrio_ctx *ctx, int32_t on);
int32_t myrio_prof_reset(myrio_ctx *ctx);
/* accumulated device time and launch count of the kernels whose name starts with `prefix` */
int32_t myrio_prof_query(myrio_ctx *ctx, const char *prefix, double *ms_total, uint64_t *launches);

/* measured aggregate shared-memory read bandwidth (GB/s; conflict-free 16-byte loads, one 1024-thread CTA
 * per SM) -- the denominator of the roofline fraction of the clustering tile kernels (bench.py) */
int32_t myrio_bench_smem_bandwidth(myrio_ctx *ctx, double *gbps, double *ms);
/* SM clock under load without a driver call: myrio_bench_clock_probe enqueues a ~0.5 ms one-warp kernel on the
 * context's stream that counts SM cycles (%clock64) against %globaltimer into `slot` (< 64);
 * myrio_bench_clock_read synchronises and returns the first n slots in MHz (0 = never written). */
int32_t myrio_bench_clock_probe(myrio_ctx *ctx, uint32_t slot);
int32_t myrio_bench_clock_read(myrio_ctx *ctx, uint32_t n, double *mhz);

/* ---- sequences ---------------------------------------------------------- */

/* A batch of MyrSeq (myrio-core/src/data/myrseq.rs:21-33): `sequence` as packed
 * Seq<Dna> words, `quality` as raw Phred bytes.  base_off[n+1] / word_off[n+1]
 * are prefix sums; read r has base_off[r+1]-base_off[r] bases. */
int32_t myrio_reads_upload(myrio_ctx *ctx, const uint64_t *words, const uint64_t *word_off,
                           const uint64_t *base_off, const uint8_t *qual, uint64_t n_reads,
                           myrio_seqs **out);
/* FASTQ ingestion on the device: the per-record conversion of io::read_fastq
 * (myrio-core/src/io/mod.rs:46-48: Seq::<Dna>::from_str, quality = byte saturating_sub 33)
 * followed by MyrSeq::pre_process (data/myrseq.rs:113-131).  seq_ascii / qual_ascii are the
 * concatenated sequence and quality lines (no newlines), base_off[n+1] their prefix sums.
 * A record is kept iff it is non-empty, len >= min_length, mean quality >= min_mean_qual and
 * max quality <= max_qual; keep[n] (optional) receives 0/1 and *out holds the kept reads in
 * input order.  Any base other than upper-case ACGT -> MYRIO_ERR_BAD_BASE. */
int32_t myrio_reads_from_fastq(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint8_t *qual_ascii,
                               const uint64_t *base_off, uint64_t n_records, uint64_t min_length,
                               float min_mean_qual, uint8_t max_qual, uint8_t *keep, myrio_seqs **out);
/* The `seq: Seq<Iupac>` of every StorePayload (myrio-core/src/tax/store.rs:28-32)
 * in payload_id order. */
int32_t myrio_leaves_upload(myrio_ctx *ctx, const uint64_t *words, const uint64_t *word_off,
                            const uint64_t *base_off, uint64_t n_leaves, myrio_seqs **out);
/* FASTA ingestion on the device: the `Seq::<Iupac>::from_str(&string_seq)` of every record in
 * TaxTreeStore::load_from_fasta_string (myrio-core/src/tax/store.rs:166-195).  seq_ascii = the
 * concatenated sequence lines of all records (headers and newlines removed), base_off[n+1] their
 * prefix sums.  Upper-case IUPAC nucleotide codes and '-' (gap, Iupac::X) are valid; a record holding
 * any other byte is skipped like the reference does (store.rs:181-187: reported, no payload), so
 * keep[n] (optional) receives 0/1 and *out holds the kept records in input order = payload_id order. */
int32_t myrio_leaves_from_fasta(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint64_t *base_off,
                                uint64_t n_records, uint8_t *keep, myrio_seqs **out);
/* number of sequences held by a batch */
int32_t myrio_seqs_count(const myrio_seqs *s, uint64_t *n_seqs);
void myrio_seqs_free(myrio_seqs *s);

/* ---- k-mer counting ----------------------------------------------------- */

/* MyrSeq::compute_kmer_counts(k, cutoff) for every read
 * (myrio-core/src/data/myrseq.rs:76-111; callers clustering.rs:78-90,
 * myrio-cli/src/app/process/run.rs:233-241,301-309).
 * out: rows = sorted (usize key, f32 count); aux = nb_hck as u64 per read. */
int32_t myrio_count_reads(myrio_ctx *ctx, const myrio_seqs *reads, uint32_t k, float cutoff,
                          myrio_svecs **out);

/* compute_kmer_store_counts_for_fasta_seq for every leaf = the body of
 * TaxTreeStore::compute_and_append_kmer_counts / _overwrite_
 * (myrio-core/src/tax/mod.rs:70-148, tax/store.rs:265-308).
 * out: rows = sorted (usize key, u16 count); aux = rescale factor (f32) per leaf.
 * Ambiguity codes are resampled nb_resamples times with the counter-based
 * generator documented in DESIGN.md (the reference seeds from the OS). */
int32_t myrio_count_leaves(myrio_ctx *ctx, const myrio_seqs *leaves, uint32_t k,
                           uint32_t nb_resamples, uint64_t seed, uint64_t leaf_id_base,
                           myrio_svecs **out);

/* ---- sparse-vector batches ---------------------------------------------- */

int32_t myrio_svecs_info(const myrio_svecs *v, uint64_t *n_rows, uint64_t *nnz, int32_t *vtype);
/* row_off[n_rows+1], keys[nnz], vals[nnz] (f32 or u16), aux[n_rows] (u64 for f32
 * batches made by myrio_count_reads, f32 for u16 batches); any pointer may be NULL. */
int32_t myrio_svecs_download(myrio_ctx *ctx, const myrio_svecs *v, uint64_t *row_off,
                             uint64_t *keys, void *vals, void *aux);
/* e.g. the cached `kmer_store_counts_vec[i]` of a loaded .myrtree, or centroids */
int32_t myrio_svecs_upload(myrio_ctx *ctx, uint64_t n_rows, const uint64_t *row_off,
                           const uint64_t *keys, const void *vals, int32_t vtype,
                           myrio_svecs **out);
/* SparseVec::into_normalized_l2 per row (data/sparse.rs:366-388); u16 input is
 * first converted with Float::from (tax/compute.rs:100-108). */
int32_t myrio_svecs_normalize(myrio_ctx *ctx, const myrio_svecs *in, myrio_svecs **out);
void myrio_svecs_free(myrio_svecs *v);

/* ---- .myrtree files ----------------------------------------------------- *
 * TaxTreeStore (myrio-core/src/tax/store.rs:43-48) and its byte format, to_bytes / from_bytes
 * (store.rs:342-454): magic "MYRIO-\xCE\xA8", u64 header length, header {core chunk info, payload chunk
 * infos, k_precomputed}, zstd(core without payloads), zstd(chunk of payloads) x C; bincode 2 fixed-int
 * little endian inside (data/sparse.rs:413-436, tax/core.rs:332-410).  Two encodings are decided by code
 * outside the reference tree and are ASSUMED (DESIGN.md lists them): Seq<Iupac> = u64 symbol count +
 * Vec<usize> raw words; Rank = u32 discriminant.
 *
 * The taxonomy crosses the ABI as a flat PRE-ORDER node list: kind[i] (0 = branch, 1 = leaf),
 * arg[i] (branch: number of children, leaf: payload_id), names concatenated with name_off[n+1].
 * Leaf sequences use the packed Seq<Iupac> layout of myrio_leaves_upload. */
int32_t myrio_store_new(const char *gene, uint32_t root_rank, uint64_t n_nodes, const uint8_t *node_kind,
                        const uint64_t *node_arg, const uint64_t *name_off, const char *names,
                        uint64_t n_payloads, const uint64_t *seq_words, const uint64_t *seq_word_off,
                        const uint64_t *seq_base_off, myrio_store **out);
/* TaxTreeStore::from_bytes (store.rs:408-454).  threads = 0 -> hardware concurrency.
 * Malformed input -> MYRIO_ERR_FORMAT (the reference returns tax::Error). */
int32_t myrio_store_from_bytes(const uint8_t *data, uint64_t len, uint32_t threads, myrio_store **out);
/* TaxTreeStore::to_bytes (store.rs:342-405).  *out is library-allocated (myrio_bytes_free).  The payloads are
 * cut into chunks of len < threads ? len : len / threads + 1 (store.rs:363-367: threads plays
 * rayon::current_num_threads(); 0 -> hardware concurrency).  A store without payloads cannot be
 * serialised (the reference panics in par_chunks(0)) -> MYRIO_ERR_EMPTY. */
int32_t myrio_store_to_bytes(const myrio_store *st, int32_t zstd_level, uint32_t threads, uint8_t **out,
                             uint64_t *len);
void myrio_bytes_free(uint8_t *p);
/* save_to_file / load_from_file (store.rs:53-93): the bytes above, written to / read from `path` */
int32_t myrio_store_save(const myrio_store *st, const char *path, int32_t zstd_level, uint32_t threads);
int32_t myrio_store_load(const char *path, uint32_t threads, myrio_store **out);
int32_t myrio_store_info(const myrio_store *st, uint64_t *n_nodes, uint64_t *n_payloads, uint64_t *n_k,
                         uint32_t *root_rank, uint64_t *names_bytes, uint64_t *gene_bytes,
                         uint64_t *seq_words_total);
/* any pointer may be NULL; sizes from myrio_store_info (gene: gene_bytes + 1 with the terminating 0) */
int32_t myrio_store_tree(const myrio_store *st, char *gene, uint8_t *node_kind, uint64_t *node_arg,
                         uint64_t *name_off, char *names, uint64_t *k_precomputed);
int32_t myrio_store_seqs(const myrio_store *st, uint64_t *seq_words, uint64_t *seq_word_off,
                         uint64_t *seq_base_off);
/* size of / host copy of kmer_store_counts_vec[k_index] over all payloads (CSR, u16 counts, f32 rescale) */
int32_t myrio_store_counts_info(const myrio_store *st, uint64_t k_index, uint64_t *nnz);
int32_t myrio_store_counts_host(const myrio_store *st, uint64_t k_index, uint64_t *row_off, uint64_t *keys,
                                uint16_t *counts, float *rescale);
/* the payload sequences as a device batch for myrio_count_leaves */
int32_t myrio_store_leaves(myrio_ctx *ctx, const myrio_store *st, myrio_seqs **out);
/* the cached counts of k_precomputed[k_index] as a device batch (u16 values + rescale), ready for
 * myrio_svecs_normalize + myrio_index_build: TaxTreeCompute::from_store_tree's cache hit (compute.rs:80-85) */
int32_t myrio_store_counts(myrio_ctx *ctx, const myrio_store *st, uint64_t k_index, myrio_svecs **out);
/* compute_and_append_kmer_counts / compute_and_overwrite_kmer_counts (store.rs:265-308), second half: the
 * batch myrio_count_leaves produced for k is written back into the store (append: pushed after the existing
 * entries; overwrite: replaces them all) -- the GPU-built counts reach the cache file without a host hop
 * through the caller. */
int32_t myrio_store_append_counts(myrio_ctx *ctx, myrio_store *st, uint64_t k, const myrio_svecs *counts);
int32_t myrio_store_overwrite_counts(myrio_ctx *ctx, myrio_store *st, uint64_t k, const myrio_svecs *counts);
/* the same from host arrays (CSR over the payloads; e.g. counts that never were on this GPU) */
int32_t myrio_store_append_counts_host(myrio_store *st, uint64_t k, const uint64_t *row_off, const uint64_t *keys,
                                       const uint16_t *counts, const float *rescale);
/* TaxTreeStore::shrink (store.rs:259-262) */
int32_t myrio_store_shrink(myrio_store *st);
void myrio_store_free(myrio_store *st);

/* ---- reference index ---------------------------------------------------- */

/* Replaces the `Box<[SFVec]>` payloads built in TaxTreeCompute::from_store_tree
 * (myrio-core/src/tax/compute.rs:100-118): the normalised leaf vectors are
 * inverted into per-tile CSR posting lists (key -> (leaf, value)), leaves in
 * payload_id order.  leaf_id_base is added to reported ids (multi-GPU shards).
 * tile_leaves = 0 picks the default. */
int32_t myrio_index_build(myrio_ctx *ctx, const myrio_svecs *leaf_vecs_normalized,
                          uint32_t leaf_id_base, uint32_t tile_leaves, myrio_index **out);
int32_t myrio_index_info(const myrio_index *ix, uint64_t *n_leaves, uint64_t *n_tiles,
                         uint64_t *n_postings, uint64_t *n_unique_total, uint32_t *tile_leaves);
/* one tile's CSR: ukeys[U], offs[U+1] (relative to the tile), post_leaf[P] (tile-local
 * leaf index), post_val[P].  Call with NULL arrays to get U and P. */
int32_t myrio_index_export_tile(myrio_ctx *ctx, const myrio_index *ix, uint64_t tile,
                                uint64_t *n_unique, uint64_t *n_postings, uint64_t *ukeys,
                                uint64_t *offs, uint32_t *post_leaf, float *post_val);
void myrio_index_free(myrio_index *ix);

/* ---- scoring ------------------------------------------------------------ */

/* The loop of TaxTreeResults::from_compute_tree (myrio-core/src/tax/results.rs:82-93):
 * scores[q*n_leaves + payload_id] = simfunc(leaf, query) for queries
 * [q_first, q_first+q_count) of `queries` (pre-normalised rows). */
int32_t myrio_score_all(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                        uint64_t q_first, uint64_t q_count, int32_t similarity, float *scores);
/* Same scores, reduced on the device to the global top-x of cut()
 * (tax/results.rs:310-329), order (score desc, payload_id asc); entries past
 * min(x, n_leaves) are (0.0f, 0xFFFFFFFF).  top_*[q*x + j]. */
int32_t myrio_score_topx(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                         uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                         float *top_scores, uint32_t *top_ids);
/* device-output variant (for the NCCL all-gather of shard-local results);
 * d_* are device pointers of q_count*x elements; asynchronous on the ctx stream */
int32_t myrio_score_topx_async(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                               uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                               float *d_top_scores, uint32_t *d_top_ids);
/* merge n_parts shard-local top-x lists laid out [part][q][x] into [q][x] (device pointers) */
int32_t myrio_topx_merge_async(myrio_ctx *ctx, const float *d_scores_in, const uint32_t *d_ids_in,
                               uint32_t n_parts, uint64_t q_count, uint32_t x, float *d_scores_out,
                               uint32_t *d_ids_out);
/* Index-free scoring: the reference's own loop (tax/results.rs:82-93), one merge-join
 * (data/sparse.rs:317-360, :143-221) per (leaf, query) pair over the normalised leaf vectors
 * `leaf_vecs` (payload order).  Every Similarity is available here, including Overlap
 * (similarity.rs:172-182), whose sum of maxima runs over the union of the keys and cannot be
 * formed from an inverted index; Cosine / JacardTanimoto give the same bits as the index path.
 * Outputs as myrio_score_all / myrio_score_topx(_async); ids are leaf_id_base + payload id. */
int32_t myrio_score_all_direct(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, const myrio_svecs *queries,
                               uint64_t q_first, uint64_t q_count, int32_t similarity, float *scores);
int32_t myrio_score_topx_direct(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, uint32_t leaf_id_base,
                                const myrio_svecs *queries, uint64_t q_first, uint64_t q_count, uint32_t x,
                                int32_t similarity, float *top_scores, uint32_t *top_ids);
int32_t myrio_score_topx_direct_async(myrio_ctx *ctx, const myrio_svecs *leaf_vecs, uint32_t leaf_id_base,
                                      const myrio_svecs *queries, uint64_t q_first, uint64_t q_count,
                                      uint32_t x, int32_t similarity, float *d_top_scores,
                                      uint32_t *d_top_ids);
/* Two-phase shard-local top-x for the multi-GPU path (one process per GPU, leaves sharded):
 *   1. myrio_score_topx_seed_async : plan + every query against its seed tile on this shard;
 *      d_bound[q] = running lower bound of the query's x-th best score (f32 bits, scores >= +0).
 *   2. the caller all-reduces d_bound with MAX over the ranks (Q x 4 bytes; the maximum is still a
 *      lower bound of the GLOBAL x-th best score: x leaves of one tile of one rank reach it).
 *   3. myrio_score_topx_sweep_async: all remaining (query, tile) tasks with the exchanged bounds, then
 *      the shard's list as composites d_top[q][x] = score bits << 32 | (0xFFFFFFFF - payload id), best
 *      first, 0 = padding.  Entries below the global bound are dropped, so the shard-local list may be
 *      shorter than x: only the merge over all ranks (myrio_topx_merge_composites_async after ONE
 *      all-gather of Q*x*8 bytes per rank) is the exact global top-x.
 * The queries must fit one batch (MYRIO_ERR_UNSUPPORTED otherwise: use myrio_score_topx_async). */
int32_t myrio_score_topx_seed_async(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                                    uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                                    uint32_t *d_bound);
int32_t myrio_score_topx_sweep_async(myrio_ctx *ctx, const myrio_index *ix, uint64_t q_count, uint32_t x,
                                     int32_t similarity, const uint32_t *d_bound, uint64_t *d_top);
int32_t myrio_topx_merge_composites_async(myrio_ctx *ctx, const uint64_t *d_in, uint32_t n_parts,
                                          uint64_t q_count, uint32_t x, float *d_scores_out,
                                          uint32_t *d_ids_out);
/* ---- results: everything TaxTreeResults::from_compute_tree does after the score loop -------------
 * (myrio-core/src/tax/results.rs:96-252): per-branch nb_leaves / mean / max over ALL leaves (:96-143),
 * trim(max_amount_of_leaves_per_branch) (:257-307), cut(nb_best_analysis) (:310-381), softmax pooling per
 * rank and confidence (:160-248).  Scores, branch statistics and the per-branch leaf selection run on the GPU
 * (the score rows never leave it); the tree rebuild, pooling and confidence run in the library's host code
 * in the reference's f32 order.  Ties (itertools' k_largest_by_key leaves them unspecified, and the
 * reference's child order comes from a HashMap): equal scores keep the input order of the taxonomy. */
typedef struct myrio_taxonomy myrio_taxonomy; /* the taxonomy of one tree, prepared for the results pass */
typedef struct myrio_results myrio_results;   /* TaxTreeResults of one query */
/* search parameters (myrio-cli/src/app/config.rs:158-175; defaults 15, 100, 2, 20, 1.1, 2.5, 0.9, 0.55) */
typedef struct {
    uint64_t max_amount_of_leaves_per_branch;
    uint64_t nb_best_analysis;
    float lambda_leaf, lambda_branch, mu, gamma, delta, epsilon;
} myrio_results_params;
/* flat PRE-ORDER taxonomy as in myrio_store_new (kind 0 = branch with arg children, 1 = leaf with arg =
 * payload_id); every payload id at most once; root_rank = Rank discriminant of the root (tax/clade.rs:93-104) */
int32_t myrio_taxonomy_create(myrio_ctx *ctx, uint32_t root_rank, uint64_t n_nodes, const uint8_t *node_kind,
                              const uint64_t *node_arg, uint64_t n_payloads, myrio_taxonomy **out);
void myrio_taxonomy_free(myrio_taxonomy *t);
/* TaxTreeResults::from_compute_tree(query, ttcompute, similarity, ...) for queries [q_first, q_first+q_count):
 * out[q] receives one handle per query.  `ix` must index the whole tree (leaf_id_base 0).  leaf_vecs = NULL
 * scores over the index (Cosine / JacardTanimoto); with the normalised leaf vectors the scores come from the
 * merge-join path (every Similarity, incl. Overlap; ix may then be NULL). */
int32_t myrio_results_from_compute_tree(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *leaf_vecs,
                                        myrio_taxonomy *taxonomy, const myrio_svecs *queries, uint64_t q_first,
                                        uint64_t q_count, int32_t similarity, const myrio_results_params *params,
                                        myrio_results **out);
/* TaxTreeResults::cut(nb_best) on an analysed tree (run.rs:392: nb_best_display); force_keep is honoured */
int32_t myrio_results_cut(const myrio_results *r, uint64_t nb_best, myrio_results **out);
int32_t myrio_results_info(const myrio_results *r, uint64_t *n_nodes, uint64_t *n_payloads, uint32_t *root_rank);
/* the tree as flat PRE-ORDER arrays (any pointer may be NULL): node_kind, node_arg (children | payload index),
 * src_node (index of the node in the taxonomy: its name), BRes per node (branches: nb_leaves, mean, max,
 * pool_score, confidence state 0 None / 1 Some(None) / 2 Some(Some(conf)), conf, force_keep), then per payload
 * its score and its payload_id in the scored tree */
int32_t myrio_results_tree(const myrio_results *r, uint8_t *node_kind, uint64_t *node_arg, uint64_t *src_node,
                           uint64_t *nb_leaves, float *mean, float *max, float *pool_score, uint8_t *conf_state,
                           float *conf, uint8_t *force_keep, float *payload_score, uint64_t *payload_src_id);
void myrio_results_free(myrio_results *r);
/* the "in-database confidence" of a run over several genes (myrio-cli/src/app/process/run.rs:405-466).  Per gene
 * g the genus branches of its results tree, genus_off[g] .. genus_off[g+1]: name ids (equal id = equal genus name),
 * pool scores, confidence state / value. */
int32_t myrio_in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                                     const float *pool_score, const uint8_t *conf_state, const float *conf,
                                     float *out);

/* ---- multi-GPU: one process per GPU, reference leaves sharded, queries replicated -------------
 * The reference is a single-process rayon program; these calls are what a multi-GPU binding of
 * TaxTreeResults::from_compute_tree (tax/results.rs:82-93 + cut(), :310-329) uses.
 * NCCL is bound at run time (dlopen libnccl.so.2); single-GPU users never load it.
 *   rank 0: myrio_comm_unique_id(id) -> ship the 128 bytes to every rank (any channel) ->
 *   every rank: myrio_comm_init(ctx, nranks, rank, id) (collective: ncclCommInitRank).
 * The communicator belongs to the context and runs its collectives on the context's stream. */
#define MYRIO_UNIQUE_ID_BYTES 128
int32_t myrio_comm_unique_id(uint8_t id[MYRIO_UNIQUE_ID_BYTES]);
int32_t myrio_comm_init(myrio_ctx *ctx, uint32_t nranks, uint32_t rank, const uint8_t id[MYRIO_UNIQUE_ID_BYTES]);
int32_t myrio_comm_destroy(myrio_ctx *ctx);
/* rank / world of the context's communicator (0 / 1 without one) and the NCCL version bound */
int32_t myrio_comm_info(const myrio_ctx *ctx, uint32_t *rank, uint32_t *nranks, int32_t *nccl_version);
/* The sharded scoring step (collective: every rank calls it with the same queries and x, each with the
 * index of ITS leaf shard, built with leaf_id_base = first payload id of the shard).  Per batch of queries:
 * seed pass -> ncclAllReduce(MAX) of the per-query bounds -> sweep with the exchanged bound ->
 * ONE ncclAllGather of the shard lists (8-byte composites) -> merge (K7), all on the context's stream.
 * d_top_*[q*x + j] (device pointers, every rank receives the full merged result) equal what
 * myrio_score_topx gives on one GPU holding all leaves.  max_batch = 0 lets the library pick.
 * Without a communicator (or nranks == 1) it is myrio_score_topx_async. */
int32_t myrio_score_topx_sharded(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *queries,
                                 uint64_t q_first, uint64_t q_count, uint32_t x, int32_t similarity,
                                 float *d_top_scores, uint32_t *d_top_ids, uint64_t max_batch);

/* exact algorithmic work of the last myrio_score_* call: sum over queries of
 * query keys, of (query key, tile) hits, and of postings touched (T_q) */
int32_t myrio_score_stats(myrio_ctx *ctx, uint64_t *n_query_keys, uint64_t *n_key_hits,
                          uint64_t *n_postings_touched);

/* ---- clustering --------------------------------------------------------- */

/* ClusteringParameters (myrio-core/src/clustering.rs:15-36) */
typedef struct {
    uint32_t k;
    float t1;
    uint64_t t2;
    int32_t similarity;
    uint64_t expected_nb_of_clusters;
    float eta_improvement;
    uint64_t nb_iters_max;
    int32_t has_silhouette_trimming;
    float silhouette_trimming; /* ssdcf */
} myrio_cluster_params;

/* clustering::cluster (myrio-core/src/clustering.rs:55-270).  initial_centroids
 * may be NULL (empty).  assign[i] = cluster index of read i, -1 = rejected by t2,
 * -2 = rejected by silhouette trimming.  silhouette (optional, n_reads floats)
 * receives the silhouette score of each clustered read (NaN where undefined). */
int32_t myrio_cluster(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *params,
                      const myrio_svecs *initial_centroids, int32_t *assign, uint32_t *n_iters,
                      float *silhouette);
/* Row-sharded clustering over `world` GPUs (one process per GPU).  Every rank passes the SAME reads
 * and parameters; the reads x centroids tiles of every iteration and the members x (members +
 * neighbours) silhouette tiles are computed for the rank's contiguous share of the rows and the
 * per-row results are exchanged through `allgather`: host_buf holds `world` chunks of
 * bytes_per_rank, chunk `rank` is filled on entry, all chunks on return (0 = ok).  With
 * allgather == NULL the exchange runs inside the library: ncclAllGather on the context's own
 * communicator (myrio_comm_init with the same rank / world), device to device on the context's
 * stream; the callback remains for callers that bring their own collective.  Results are identical
 * on every rank and identical to myrio_cluster. */
typedef int32_t (*myrio_allgather_fn)(void *user, void *host_buf, uint64_t bytes_per_rank);
typedef struct {
    uint32_t rank, world;
    myrio_allgather_fn allgather;
    void *user;
} myrio_shard;
int32_t myrio_cluster_sharded(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *params,
                              const myrio_svecs *initial_centroids, const myrio_shard *shard,
                              int32_t *assign, uint32_t *n_iters, float *silhouette);
/* compute_cluster_centroid (clustering.rs:44-53) over rows member_ids[0..m) of `counts`
 * (raw, un-normalised f32 counts): out = normalize_l2(sum of the members), a 1-row batch.
 * The per-key sums are sequential f32 additions in member order, which is also
 * `iter.sum::<SFVec>().into_normalized_l2()` -- so the same call builds the FINGERPRINTS of
 * TaxTreeCompute::from_store_tree (tax/compute.rs:51-77): with `counts` = a u16 batch of
 * myrio_count_leaves at cluster_k (converted like kmer_store_counts_to_kmer_counts, tax/mod.rs:55-58:
 * count as f32 * rescale) and member_ids = the leaves sampled from one branch it returns the local
 * fingerprint; over the f32 batch of local fingerprints it returns the global one. */
int32_t myrio_compute_cluster_centroid(myrio_ctx *ctx, const myrio_svecs *counts,
                                       const uint64_t *member_ids, uint64_t m, myrio_svecs **out);
/* The re-weighting of a tree's local fingerprints before the second clustering pass of `myrio run`
 * (myrio-cli/src/app/process/run.rs:243-262): out = normalize_l2(sum_i local_i * exp(1 + alpha *
 * simfunc(local_i, naive_query))), the sum taken in row order.  local_fingerprints: the normalised local
 * fingerprints (rows); naive_query: the centroid of the pass-1 cluster (a 1-row batch); alpha:
 * config.fingerprint.alpha (default 2.5, config.rs:137).  out: a 1-row batch = the cluster's initial centroid. */
int32_t myrio_reweight_fingerprints(myrio_ctx *ctx, const myrio_svecs *local_fingerprints,
                                    const myrio_svecs *naive_query, int32_t similarity, float alpha,
                                    myrio_svecs **out);
/* the reads x reads (or reads x centroids) similarity tile used by the clustering
 * loop: out[i*n_cols + j] = simfunc(rows_i, cols_j), rows/cols pre-normalised */
int32_t myrio_pairwise(myrio_ctx *ctx, const myrio_svecs *rows, const myrio_svecs *cols,
                       int32_t similarity, float *out);

/* exact algorithmic work of the last myrio_cluster / myrio_pairwise call: similarities
 * computed (rows x columns over all K6 tiles) and multiply-add terms (row entries x columns) */
int32_t myrio_cluster_stats(myrio_ctx *ctx, uint64_t *n_pair_similarities, uint64_t *n_terms);

/* device pointers of a batch, for zero-copy interop (torch / NCCL plumbing) */
int32_t myrio_svecs_device_ptrs(const myrio_svecs *v, const uint64_t **row_off,
                                const uint64_t **keys, const void **vals, const void **aux);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* MYRIO_B200_H */
