//! Raw bindings of `include/myrio_b200.h` (kept in sync by hand; every item cites the header).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

#[repr(C)] pub struct myrio_ctx { _p: [u8; 0] }
#[repr(C)] pub struct myrio_seqs { _p: [u8; 0] }
#[repr(C)] pub struct myrio_svecs { _p: [u8; 0] }
#[repr(C)] pub struct myrio_index { _p: [u8; 0] }
#[repr(C)] pub struct myrio_store { _p: [u8; 0] }
#[repr(C)] pub struct myrio_taxonomy { _p: [u8; 0] }
#[repr(C)] pub struct myrio_results { _p: [u8; 0] }

/// search parameters (myrio-cli/src/app/config.rs:158-175)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct myrio_results_params {
    pub max_amount_of_leaves_per_branch: u64,
    pub nb_best_analysis: u64,
    pub lambda_leaf: f32,
    pub lambda_branch: f32,
    pub mu: f32,
    pub gamma: f32,
    pub delta: f32,
    pub epsilon: f32,
}

pub const MYRIO_OK: i32 = 0;
pub const MYRIO_ERR_INVALID_K: i32 = 1;
pub const MYRIO_ERR_BAD_ARG: i32 = 2;
pub const MYRIO_ERR_CUDA: i32 = 3;
pub const MYRIO_ERR_OOM: i32 = 4;
pub const MYRIO_ERR_BAD_QUALITY: i32 = 5;
pub const MYRIO_ERR_UNSUPPORTED: i32 = 6;
pub const MYRIO_ERR_EMPTY: i32 = 7;
pub const MYRIO_ERR_NONFINITE: i32 = 8;
pub const MYRIO_ERR_BAD_BASE: i32 = 9;

pub const MYRIO_SIM_COSINE: i32 = 0;
pub const MYRIO_SIM_JACARD_TANIMOTO: i32 = 1;
pub const MYRIO_SIM_OVERLAP: i32 = 2;
pub const MYRIO_VAL_F32: i32 = 0;
pub const MYRIO_VAL_U16: i32 = 1;

/// `ClusteringParameters` (myrio-core/src/clustering.rs:15-36) without the centroids.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct myrio_cluster_params {
    pub k: u32,
    pub t1: f32,
    pub t2: u64,
    pub similarity: i32,
    pub expected_nb_of_clusters: u64,
    pub eta_improvement: f32,
    pub nb_iters_max: u64,
    pub has_silhouette_trimming: i32,
    pub silhouette_trimming: f32,
}

/// all-gather of `world` host chunks of `bytes_per_rank` (chunk `rank` filled on entry); 0 = ok
pub type myrio_allgather_fn =
    Option<unsafe extern "C" fn(user: *mut c_void, host_buf: *mut c_void, bytes_per_rank: u64) -> i32>;

/// this rank's place in a row-sharded `myrio_cluster_sharded` call
#[repr(C)]
#[derive(Clone, Copy)]
pub struct myrio_shard {
    pub rank: u32,
    pub world: u32,
    pub allgather: myrio_allgather_fn,
    pub user: *mut c_void,
}

unsafe extern "C" {
    pub fn myrio_last_error() -> *const c_char;
    pub fn myrio_version() -> i32;
    pub fn myrio_ctx_create(device: i32, out: *mut *mut myrio_ctx) -> i32;
    pub fn myrio_ctx_destroy(ctx: *mut myrio_ctx);
    pub fn myrio_ctx_sync(ctx: *mut myrio_ctx) -> i32;
    pub fn myrio_ctx_set_kmer_bit_order(ctx: *mut myrio_ctx, msb_first: i32) -> i32;
    pub fn myrio_ctx_set_topx_bound_pruning(ctx: *mut myrio_ctx, on: i32) -> i32;
    pub fn myrio_ctx_stream(ctx: *const myrio_ctx) -> *mut c_void;
    pub fn myrio_ctx_launch_count(ctx: *const myrio_ctx) -> u64;
    pub fn myrio_prof_enable(ctx: *mut myrio_ctx, on: i32) -> i32;
    pub fn myrio_prof_reset(ctx: *mut myrio_ctx) -> i32;
    pub fn myrio_prof_query(ctx: *mut myrio_ctx, prefix: *const c_char, ms_total: *mut f64, launches: *mut u64) -> i32;

    pub fn myrio_bench_smem_bandwidth(ctx: *mut myrio_ctx, gbps: *mut f64, ms: *mut f64) -> i32;
    pub fn myrio_bench_clock_probe(ctx: *mut myrio_ctx, slot: u32) -> i32;
    pub fn myrio_bench_clock_read(ctx: *mut myrio_ctx, n: u32, mhz: *mut f64) -> i32;

    pub fn myrio_reads_upload(ctx: *mut myrio_ctx, words: *const u64, word_off: *const u64, base_off: *const u64,
                              qual: *const u8, n_reads: u64, out: *mut *mut myrio_seqs) -> i32;
    pub fn myrio_reads_from_fastq(ctx: *mut myrio_ctx, seq_ascii: *const u8, qual_ascii: *const u8, base_off: *const u64,
                                  n_records: u64, min_length: u64, min_mean_qual: f32, max_qual: u8, keep: *mut u8,
                                  out: *mut *mut myrio_seqs) -> i32;
    pub fn myrio_leaves_upload(ctx: *mut myrio_ctx, words: *const u64, word_off: *const u64, base_off: *const u64,
                               n_leaves: u64, out: *mut *mut myrio_seqs) -> i32;
    pub fn myrio_leaves_from_fasta(ctx: *mut myrio_ctx, seq_ascii: *const u8, base_off: *const u64, n_records: u64,
                                   keep: *mut u8, out: *mut *mut myrio_seqs) -> i32;
    pub fn myrio_seqs_count(s: *const myrio_seqs, n_seqs: *mut u64) -> i32;
    pub fn myrio_seqs_free(s: *mut myrio_seqs);

    // .myrtree reader / writer (TaxTreeStore::to_bytes / from_bytes, tax/store.rs:342-454)
    pub fn myrio_store_new(gene: *const c_char, root_rank: u32, n_nodes: u64, node_kind: *const u8, node_arg: *const u64,
                           name_off: *const u64, names: *const c_char, n_payloads: u64, seq_words: *const u64,
                           seq_word_off: *const u64, seq_base_off: *const u64, out: *mut *mut myrio_store) -> i32;
    pub fn myrio_store_from_bytes(data: *const u8, len: u64, threads: u32, out: *mut *mut myrio_store) -> i32;
    pub fn myrio_store_to_bytes(st: *const myrio_store, zstd_level: i32, threads: u32, out: *mut *mut u8, len: *mut u64) -> i32;
    pub fn myrio_bytes_free(p: *mut u8);
    pub fn myrio_store_save(st: *const myrio_store, path: *const c_char, zstd_level: i32, threads: u32) -> i32;
    pub fn myrio_store_load(path: *const c_char, threads: u32, out: *mut *mut myrio_store) -> i32;
    pub fn myrio_store_info(st: *const myrio_store, n_nodes: *mut u64, n_payloads: *mut u64, n_k: *mut u64,
                            root_rank: *mut u32, names_bytes: *mut u64, gene_bytes: *mut u64, seq_words_total: *mut u64) -> i32;
    pub fn myrio_store_tree(st: *const myrio_store, gene: *mut c_char, node_kind: *mut u8, node_arg: *mut u64,
                            name_off: *mut u64, names: *mut c_char, k_precomputed: *mut u64) -> i32;
    pub fn myrio_store_seqs(st: *const myrio_store, seq_words: *mut u64, seq_word_off: *mut u64, seq_base_off: *mut u64) -> i32;
    pub fn myrio_store_counts_info(st: *const myrio_store, k_index: u64, nnz: *mut u64) -> i32;
    pub fn myrio_store_counts_host(st: *const myrio_store, k_index: u64, row_off: *mut u64, keys: *mut u64,
                                   counts: *mut u16, rescale: *mut f32) -> i32;
    pub fn myrio_store_leaves(ctx: *mut myrio_ctx, st: *const myrio_store, out: *mut *mut myrio_seqs) -> i32;
    pub fn myrio_store_counts(ctx: *mut myrio_ctx, st: *const myrio_store, k_index: u64, out: *mut *mut myrio_svecs) -> i32;
    pub fn myrio_store_append_counts(ctx: *mut myrio_ctx, st: *mut myrio_store, k: u64, counts: *const myrio_svecs) -> i32;
    pub fn myrio_store_overwrite_counts(ctx: *mut myrio_ctx, st: *mut myrio_store, k: u64, counts: *const myrio_svecs) -> i32;
    pub fn myrio_store_append_counts_host(st: *mut myrio_store, k: u64, row_off: *const u64, keys: *const u64,
                                          counts: *const u16, rescale: *const f32) -> i32;
    pub fn myrio_store_shrink(st: *mut myrio_store) -> i32;
    pub fn myrio_store_free(st: *mut myrio_store);

    // TaxTreeResults::from_compute_tree after the score loop (tax/results.rs:96-381), run.rs:405-466
    pub fn myrio_taxonomy_create(ctx: *mut myrio_ctx, root_rank: u32, n_nodes: u64, node_kind: *const u8,
                                 node_arg: *const u64, n_payloads: u64, out: *mut *mut myrio_taxonomy) -> i32;
    pub fn myrio_taxonomy_free(t: *mut myrio_taxonomy);
    pub fn myrio_results_from_compute_tree(ctx: *mut myrio_ctx, ix: *const myrio_index, leaf_vecs: *const myrio_svecs,
                                           taxonomy: *mut myrio_taxonomy, queries: *const myrio_svecs, q_first: u64,
                                           q_count: u64, similarity: i32, params: *const myrio_results_params,
                                           out: *mut *mut myrio_results) -> i32;
    pub fn myrio_results_cut(r: *const myrio_results, nb_best: u64, out: *mut *mut myrio_results) -> i32;
    pub fn myrio_results_info(r: *const myrio_results, n_nodes: *mut u64, n_payloads: *mut u64, root_rank: *mut u32) -> i32;
    pub fn myrio_results_tree(r: *const myrio_results, node_kind: *mut u8, node_arg: *mut u64, src_node: *mut u64,
                              nb_leaves: *mut u64, mean: *mut f32, max: *mut f32, pool_score: *mut f32,
                              conf_state: *mut u8, conf: *mut f32, force_keep: *mut u8, payload_score: *mut f32,
                              payload_src_id: *mut u64) -> i32;
    pub fn myrio_results_free(r: *mut myrio_results);
    pub fn myrio_in_database_confidence(n_genes: u64, genus_off: *const u64, name_id: *const u64, pool_score: *const f32,
                                        conf_state: *const u8, conf: *const f32, out: *mut f32) -> i32;

    pub fn myrio_reweight_fingerprints(ctx: *mut myrio_ctx, local_fingerprints: *const myrio_svecs,
                                       naive_query: *const myrio_svecs, similarity: i32, alpha: f32,
                                       out: *mut *mut myrio_svecs) -> i32;

    pub fn myrio_comm_unique_id(id: *mut u8) -> i32;
    pub fn myrio_comm_init(ctx: *mut myrio_ctx, nranks: u32, rank: u32, id: *const u8) -> i32;
    pub fn myrio_comm_destroy(ctx: *mut myrio_ctx) -> i32;
    pub fn myrio_comm_info(ctx: *const myrio_ctx, rank: *mut u32, nranks: *mut u32, nccl_version: *mut i32) -> i32;
    pub fn myrio_score_topx_sharded(ctx: *mut myrio_ctx, ix: *const myrio_index, queries: *const myrio_svecs,
                                    q_first: u64, q_count: u64, x: u32, similarity: i32, d_top_scores: *mut f32,
                                    d_top_ids: *mut u32, max_batch: u64) -> i32;

    pub fn myrio_count_reads(ctx: *mut myrio_ctx, reads: *const myrio_seqs, k: u32, cutoff: f32,
                             out: *mut *mut myrio_svecs) -> i32;
    pub fn myrio_count_leaves(ctx: *mut myrio_ctx, leaves: *const myrio_seqs, k: u32, nb_resamples: u32, seed: u64,
                              leaf_id_base: u64, out: *mut *mut myrio_svecs) -> i32;

    pub fn myrio_svecs_info(v: *const myrio_svecs, n_rows: *mut u64, nnz: *mut u64, vtype: *mut i32) -> i32;
    pub fn myrio_svecs_download(ctx: *mut myrio_ctx, v: *const myrio_svecs, row_off: *mut u64, keys: *mut u64,
                                vals: *mut c_void, aux: *mut c_void) -> i32;
    pub fn myrio_svecs_upload(ctx: *mut myrio_ctx, n_rows: u64, row_off: *const u64, keys: *const u64,
                              vals: *const c_void, vtype: i32, out: *mut *mut myrio_svecs) -> i32;
    pub fn myrio_svecs_normalize(ctx: *mut myrio_ctx, input: *const myrio_svecs, out: *mut *mut myrio_svecs) -> i32;
    pub fn myrio_svecs_free(v: *mut myrio_svecs);
    pub fn myrio_svecs_device_ptrs(v: *const myrio_svecs, row_off: *mut *const u64, keys: *mut *const u64,
                                   vals: *mut *const c_void, aux: *mut *const c_void) -> i32;

    pub fn myrio_index_build(ctx: *mut myrio_ctx, leaf_vecs_normalized: *const myrio_svecs, leaf_id_base: u32,
                             tile_leaves: u32, out: *mut *mut myrio_index) -> i32;
    pub fn myrio_index_info(ix: *const myrio_index, n_leaves: *mut u64, n_tiles: *mut u64, n_postings: *mut u64,
                            n_unique_total: *mut u64, tile_leaves: *mut u32) -> i32;
    pub fn myrio_index_export_tile(ctx: *mut myrio_ctx, ix: *const myrio_index, tile: u64, n_unique: *mut u64,
                                   n_postings: *mut u64, ukeys: *mut u64, offs: *mut u64, post_leaf: *mut u32,
                                   post_val: *mut f32) -> i32;
    pub fn myrio_index_free(ix: *mut myrio_index);

    pub fn myrio_score_all(ctx: *mut myrio_ctx, ix: *const myrio_index, queries: *const myrio_svecs, q_first: u64,
                           q_count: u64, similarity: i32, scores: *mut f32) -> i32;
    pub fn myrio_score_topx(ctx: *mut myrio_ctx, ix: *const myrio_index, queries: *const myrio_svecs, q_first: u64,
                            q_count: u64, x: u32, similarity: i32, top_scores: *mut f32, top_ids: *mut u32) -> i32;
    pub fn myrio_score_topx_async(ctx: *mut myrio_ctx, ix: *const myrio_index, queries: *const myrio_svecs,
                                  q_first: u64, q_count: u64, x: u32, similarity: i32, d_top_scores: *mut f32,
                                  d_top_ids: *mut u32) -> i32;
    pub fn myrio_score_all_direct(ctx: *mut myrio_ctx, leaf_vecs: *const myrio_svecs, queries: *const myrio_svecs,
                                  q_first: u64, q_count: u64, similarity: i32, scores: *mut f32) -> i32;
    pub fn myrio_score_topx_direct(ctx: *mut myrio_ctx, leaf_vecs: *const myrio_svecs, leaf_id_base: u32,
                                   queries: *const myrio_svecs, q_first: u64, q_count: u64, x: u32, similarity: i32,
                                   top_scores: *mut f32, top_ids: *mut u32) -> i32;
    pub fn myrio_score_topx_direct_async(ctx: *mut myrio_ctx, leaf_vecs: *const myrio_svecs, leaf_id_base: u32,
                                         queries: *const myrio_svecs, q_first: u64, q_count: u64, x: u32,
                                         similarity: i32, d_top_scores: *mut f32, d_top_ids: *mut u32) -> i32;
    pub fn myrio_score_topx_seed_async(ctx: *mut myrio_ctx, ix: *const myrio_index, queries: *const myrio_svecs,
                                       q_first: u64, q_count: u64, x: u32, similarity: i32, d_bound: *mut u32) -> i32;
    pub fn myrio_score_topx_sweep_async(ctx: *mut myrio_ctx, ix: *const myrio_index, q_count: u64, x: u32,
                                        similarity: i32, d_bound: *const u32, d_top: *mut u64) -> i32;
    pub fn myrio_topx_merge_composites_async(ctx: *mut myrio_ctx, d_in: *const u64, n_parts: u32, q_count: u64,
                                             x: u32, d_scores_out: *mut f32, d_ids_out: *mut u32) -> i32;
    pub fn myrio_topx_merge_async(ctx: *mut myrio_ctx, d_scores_in: *const f32, d_ids_in: *const u32, n_parts: u32,
                                  q_count: u64, x: u32, d_scores_out: *mut f32, d_ids_out: *mut u32) -> i32;
    pub fn myrio_score_stats(ctx: *mut myrio_ctx, n_query_keys: *mut u64, n_key_hits: *mut u64,
                             n_postings_touched: *mut u64) -> i32;

    pub fn myrio_cluster(ctx: *mut myrio_ctx, reads: *const myrio_seqs, params: *const myrio_cluster_params,
                         initial_centroids: *const myrio_svecs, assign: *mut i32, n_iters: *mut u32,
                         silhouette: *mut f32) -> i32;
    pub fn myrio_cluster_sharded(ctx: *mut myrio_ctx, reads: *const myrio_seqs, params: *const myrio_cluster_params,
                                 initial_centroids: *const myrio_svecs, shard: *const myrio_shard, assign: *mut i32,
                                 n_iters: *mut u32, silhouette: *mut f32) -> i32;
    pub fn myrio_compute_cluster_centroid(ctx: *mut myrio_ctx, counts: *const myrio_svecs, member_ids: *const u64,
                                          m: u64, out: *mut *mut myrio_svecs) -> i32;
    pub fn myrio_pairwise(ctx: *mut myrio_ctx, rows: *const myrio_svecs, cols: *const myrio_svecs, similarity: i32,
                          out: *mut f32) -> i32;
    pub fn myrio_cluster_stats(ctx: *mut myrio_ctx, n_pair_similarities: *mut u64, n_terms: *mut u64) -> i32;
}
