/*
 * myrio_oracle.h -- CPU restatement of Myrio's data-parallel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under myrio_b200/ (the product) may
 * include, link or call this.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, always as the checker or as the timed CPU baseline, never as the
 * thing shipped.
 *
 * Every function restates one reference function and cites it
 * (paths relative to the reference repository root, Myrio v0.3.0).
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   - pinned by the reference's own known-answer tests (ported verbatim as
 *     inputs/outputs in tests/test_oracle_golden.py):
 *       similarity.rs:217-232, sparse.rs:594-656.
 *   - PARITY UNPINNED (no reference test or golden vector exists, and the
 *     Rust reference cannot be compiled in this image):
 *       the integer value of a k-mer key (bio-seq 0.14.2 fork, un-vendored;
 *       assumed A=0,C=1,G=2,T=3, base i of the window at bits 2i..2i+1),
 *       the IUPAC 4-bit code (assumed A=8,C=4,G=2,T=1, gap X=0),
 *       the quality filter, leaf resampling / cut-off / rescale,
 *       top-x tie order, and every part of cluster().
 *     For these the oracle is a line-by-line restatement of the cited code.
 */
#ifndef MYRIO_ORACLE_H
#define MYRIO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Similarity selector, similarity.rs:26-33 */
enum { MYO_SIM_COSINE = 0, MYO_SIM_JACARD_TANIMOTO = 1, MYO_SIM_OVERLAP = 2 };

/* constants.rs:28  Q_TO_BP_CALL_CORRECT_PROB_MAP as f32 */
const float *myo_phred_correct_table(void);

/* bio-seq Kmer -> usize (assumed, see header): base j at bits 2j */
uint64_t myo_kmer_key(const uint8_t *bases, int k);
void myo_set_kmer_msb_first(int on);

/* data/sparse.rs:229-272 from_unsorted_singles with value=1.0f (f32 "+=").
 * keys_out/vals_out capacity n. returns number of distinct keys. */
size_t myo_sort_reduce_f32(uint64_t *singles, size_t n, uint64_t *keys_out, float *vals_out);
/* same with value=1u16 and wrapping u16 "+=" (release build semantics) */
size_t myo_sort_reduce_u16(uint64_t *singles, size_t n, uint64_t *keys_out, uint16_t *vals_out);

/* data/myrseq.rs:76-111  MyrSeq::compute_kmer_counts
 * bases: one 2-bit code per byte. keys_out/vals_out capacity 2*(n-k+1).
 * returns 0 ok, <0 error (k out of 2..32, quality >= 128). */
int myo_count_read(const uint8_t *bases, const uint8_t *qual, size_t n, int k, float cutoff,
                   uint64_t *keys_out, float *vals_out, size_t *d_out, uint64_t *nb_hck_out);

/* Batch of the above (the reference's callers loop over reads:
 * clustering.rs:78-90, run.rs:233-241,301-309).  Two-phase: pass
 * keys_out==NULL to get row_off only.  threads<=1 -> serial like the
 * reference call sites. */
int myo_count_reads(const uint8_t *bases, const uint8_t *qual, const uint64_t *base_off,
                    size_t n_reads, int k, float cutoff, int threads, uint64_t *row_off,
                    uint64_t *keys_out, float *vals_out, uint64_t *nb_hck_out);

/* Deterministic stand-in for SmallRng in the IUPAC resampler
 * (tax/mod.rs:77-97).  Shared definition with the CUDA path. */
uint32_t myo_resample_pick(uint64_t seed, uint64_t leaf, uint32_t resample, uint32_t pos,
                           uint32_t n_opts);

/* tax/mod.rs:70-148  compute_kmer_store_counts_for_fasta_seq
 * iupac: one 4-bit code per byte.  capacity of outputs: max(1, 2*(n-k+1)*nb_resamples)
 * worst case, but never more than 2*(n-k+1)*min(nb_resamples, ...) distinct;
 * caller passes cap and gets <0 if exceeded. */
int myo_count_leaf(const uint8_t *iupac, size_t n, int k, uint32_t nb_resamples, uint64_t seed,
                   uint64_t leaf_id, uint64_t *keys_out, uint16_t *vals_out, size_t cap,
                   size_t *d_out, float *rescale_out);

/* tax/store.rs:265-285 compute_and_append_kmer_counts: rayon over leaves.
 * two-phase like myo_count_reads (keys_out==NULL -> sizes only). */
int myo_count_leaves(const uint8_t *iupac, const uint64_t *base_off, size_t n_leaves, int k,
                     uint32_t nb_resamples, uint64_t seed, uint64_t leaf_id_base, int threads,
                     uint64_t *row_off, uint64_t *keys_out, uint16_t *vals_out, float *rescale_out);

/* data/sparse.rs:366-383 normalize_l2 (sum of v*v in key order, sqrt, divide) */
void myo_normalize_l2(float *vals, size_t n);
/* tax/compute.rs:100-108  u16 -> f32 -> into_normalized_l2 */
void myo_normalize_u16(const uint16_t *in, size_t n, float *out);
/* row-wise over a CSR batch */
void myo_normalize_rows_f32(const uint64_t *row_off, size_t n_rows, float *vals, int threads);
void myo_normalize_rows_u16(const uint64_t *row_off, size_t n_rows, const uint16_t *in, float *out,
                            int threads);

/* data/sparse.rs:317-360  SparseVec<f32>::dot */
float myo_dot(const uint64_t *ak, const float *av, size_t an, const uint64_t *bk, const float *bv,
              size_t bn);
/* similarity.rs:40-182; pre_normalized selects the *_pre_normalized variants */
float myo_similarity(int sim, int pre_normalized, const uint64_t *ak, const float *av, size_t an,
                     const uint64_t *bk, const float *bv, size_t bn);

/* tax/results.rs:82-93  score every leaf against one query (rayon -> OpenMP) */
void myo_score_all(int sim, const uint64_t *leaf_row_off, const uint64_t *leaf_keys,
                   const float *leaf_vals, size_t n_leaves, const uint64_t *qk, const float *qv,
                   size_t qn, int threads, float *scores_out);

/* tax/results.rs:310-329 cut(): global top-x, order (score desc, payload_id asc).
 * Writes min(x,n) entries, pads the rest with (0.0f, 0xFFFFFFFF). */
void myo_topx(const float *scores, size_t n, uint32_t id_base, size_t x, float *top_scores,
              uint32_t *top_ids);

/* data/sparse.rs:143-221 + 493-519  a + b */
size_t myo_add(const uint64_t *ak, const float *av, size_t an, const uint64_t *bk, const float *bv,
               size_t bn, uint64_t *ck, float *cv);

/* clustering.rs:44-53 compute_cluster_centroid over rows member_ids[0..m) of a CSR batch.
 * out capacity: sum of member nnz. */
size_t myo_centroid(const uint64_t *row_off, const uint64_t *keys, const float *vals,
                    const uint64_t *member_ids, size_t m, uint64_t *ck, float *cv);

/* Reference structure of the inverted index the CUDA path builds (K3).  Not a
 * reference function: the definition of "posting lists" used for parity.
 * For leaves [0,n_leaves): unique keys ascending, offsets, and for each key
 * the (leaf id ascending, value) postings.  Two-phase: ukeys==NULL -> counts. */
void myo_invert(const uint64_t *row_off, const uint64_t *keys, const float *vals, size_t n_leaves,
                size_t *n_unique, uint64_t *ukeys, uint64_t *offs, uint32_t *post_leaf,
                float *post_val);

/* clustering.rs:55-270 cluster(), starting from the per-read counts of step 1.
 * assign_out[i]: cluster index, -1 = rejected by t2 (clustering.rs:82-87),
 * -2 = rejected by silhouette trimming (clustering.rs:254-262).
 * sil_out (optional, n_reads floats): silhouette score of kept reads (NaN if none). */
typedef struct {
    uint64_t t2;
    int similarity;
    size_t n_init;
    const uint64_t *init_row_off;
    const uint64_t *init_keys;
    const float *init_vals;
    size_t expected_nb_of_clusters;
    float eta_improvement;
    size_t nb_iters_max;
    int has_silhouette;
    float ssdcf;
    int threads;
} myo_cluster_params;

int myo_cluster(const uint64_t *row_off, const uint64_t *keys, const float *vals,
                const uint64_t *nb_hck, size_t n_reads, const myo_cluster_params *p,
                int32_t *assign_out, uint32_t *n_iters_out, float *sil_out);

/* data/sparse.rs:274-313 from_unsorted_pairs (f32 values) */
size_t myo_sort_reduce_pairs_f32(const uint64_t *keys, const float *vals, size_t n, uint64_t *keys_out, float *vals_out);
/* myrio-cli/src/app/process/run.rs:243-262 (fingerprint re-weighting); out capacity = local nnz */
size_t myo_reweight_fingerprints(int sim, const uint64_t *row_off, const uint64_t *keys, const float *vals, size_t n,
                                 const uint64_t *qk, const float *qv, size_t qn, float alpha, uint64_t *ck,
                                 float *cv);

/* test hook for full-size checks: restrict the silhouette pass of myo_cluster to the reads with
 * mask[read] != 0 (no trimming then); NULL = the reference behaviour */
void myo_set_silhouette_sample(const uint8_t *mask);

/* pairwise similarity tile: out[i*n_cols+j] = sim(rows_i, cols_j) on pre-normalised rows */
void myo_pairwise(int sim, const uint64_t *r_off, const uint64_t *r_keys, const float *r_vals,
                  size_t n_rows, const uint64_t *c_off, const uint64_t *c_keys, const float *c_vals,
                  size_t n_cols, int threads, float *out);

/* io/mod.rs:46-48 (Seq::<Dna>::from_str + quality byte saturating_sub 33) followed by
 * MyrSeq::pre_process (data/myrseq.rs:113-131).  codes_out/qual_out: capacity = total bases;
 * keep_out[n]; kept_off_out[n+1] = prefix sums of the kept lengths (entries after the kept
 * count are unused).  returns the number of kept records, or -(index+1) of the first record
 * holding a non-ACGT base (the reference fails the whole file). */
long long myo_fastq_preprocess(const uint8_t *seq_ascii, const uint8_t *qual_ascii, const uint64_t *base_off,
                               size_t n, size_t min_length, float min_mean_qual, uint8_t max_qual,
                               uint8_t *codes_out, uint8_t *qual_out, uint8_t *keep_out,
                               uint64_t *kept_off_out);

/* tax/store.rs:178-188 Seq::<Iupac>::from_str per FASTA record: one 4-bit code per byte in codes_out
 * (capacity = total symbols), keep_out[n], kept_off_out[n+1]; returns the number of kept records
 * (a record with an invalid byte is skipped, store.rs:181-187). */
long long myo_fasta_to_iupac(const uint8_t *seq_ascii, const uint64_t *base_off, size_t n, uint8_t *codes_out,
                             uint8_t *keep_out, uint64_t *kept_off_out);

/* ---- .myrtree byte format (myrtree_oracle.c): tax/store.rs:314-454, data/sparse.rs:413-436,
 * tax/core.rs:332-410.  A store is described by flat arrays: taxonomy in PRE-ORDER (kind 0 = branch with
 * arg = number of children, 1 = leaf with arg = payload_id; names concatenated), leaf sequences as one 4-bit
 * IUPAC code per byte, cached counts per k as CSR over the payloads. */
typedef struct {
    const char *gene;
    uint32_t root_rank; /* Rank discriminant, tax/clade.rs:93-104 */
    uint64_t n_nodes;
    const uint8_t *kind;
    const uint64_t *arg;
    const uint64_t *name_off; /* [n_nodes + 1] */
    const char *names;
    uint64_t n_payloads;
    const uint8_t *seq_codes;
    const uint64_t *seq_off; /* [n_payloads + 1] */
    uint64_t n_k;
    const uint64_t *k_precomputed;
    const uint64_t *const *row_off; /* [n_k][n_payloads + 1] */
    const uint64_t *const *keys;
    const uint16_t *const *vals;
    const float *const *rescale; /* [n_k][n_payloads] */
} myo_store_desc;
typedef struct myo_store_owned myo_store_owned;
/* TaxTreeStore::to_bytes; `threads` plays rayon::current_num_threads() in the chunk-size rule (store.rs:363-367).
 * *out is malloc'd (myo_bytes_free). 0 ok; -23 = no payloads (the reference panics) */
int myo_store_encode(const myo_store_desc *d, int zstd_level, unsigned threads, uint8_t **out, uint64_t *out_len);
/* TaxTreeStore::from_bytes; 0 ok, -30 bad magic (NotATaxTreeStore), -31 truncated, -32 zstd, -33 bincode */
int myo_store_decode(const uint8_t *data, uint64_t len, myo_store_owned **out);
/* fills `d` with pointers into `s`; the four arrays of n_k pointers are caller-provided */
void myo_store_view(const myo_store_owned *s, myo_store_desc *d, const uint64_t **row_off, const uint64_t **keys,
                    const uint16_t **vals, const float **rescale);
void myo_store_free(myo_store_owned *s);
/* decompressed section of a file: -1 header, 0 core, 1 + c = payload chunk c */
int myo_store_section(const uint8_t *data, uint64_t len, int which, uint8_t **out, uint64_t *out_len);
void myo_bytes_free(uint8_t *p);

/* ---- host post-processing after the score loop (results_oracle.c): tax/results.rs:96-381,
 * myrio-cli/src/app/process/run.rs:405-466.  Trees are flat PRE-ORDER node lists like the .myrtree ones. */
typedef struct { /* myrio-cli/src/app/config.rs:158-175 (defaults 15, 100, 2, 20, 1.1, 2.5, 0.9, 0.55) */
    uint64_t max_amount_of_leaves_per_branch;
    uint64_t nb_best_analysis;
    float lambda_leaf, lambda_branch, mu, gamma, delta, epsilon;
} myo_results_params;
typedef struct myo_results myo_results;
typedef struct {
    uint64_t n_nodes, n_payloads;
    const uint8_t *kind;      /* 0 branch, 1 leaf */
    const uint64_t *arg;      /* children count | payload id in THIS results tree */
    const uint64_t *src;      /* node index in the input tree (names live there) */
    const uint64_t *nb_leaves;
    const float *mean, *max, *pool_score, *conf;
    const uint8_t *conf_state; /* 0 None, 1 Some(None), 2 Some(Some(conf)) */
    const uint8_t *force_keep;
    const float *payload_score;
    const uint64_t *payload_orig; /* payload id in the input tree */
} myo_results_view_t;
/* TaxTreeResults::from_compute_tree from the scores on (results.rs:96-252): branch statistics over all leaves,
 * trim, cut, pooling, confidence.  0 ok; -40 non-finite pool score (the reference panics), -41 empty result */
int myo_results_from_scores(uint32_t root_rank, uint64_t n_nodes, const uint8_t *kind, const uint64_t *arg,
                            const float *simscores, uint64_t n_payloads, const myo_results_params *p,
                            myo_results **out);
/* TaxTreeResults::cut on an analysed tree (results.rs:310-381; run.rs:392) */
int myo_results_cut(const myo_results *r, uint64_t nb_best, myo_results **out);
void myo_results_view(const myo_results *r, myo_results_view_t *v);
void myo_results_free(myo_results *r);
/* run.rs:405-466 in-database confidence; per gene the genus branches (name ids, pool scores, confidences) */
float myo_in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                                 const float *pool_score, const uint8_t *conf_state, const float *conf);

int myo_max_threads(void);
/* omp_set_num_threads: overrides an inherited OMP_NUM_THREADS (torchrun exports 1) */
void myo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
