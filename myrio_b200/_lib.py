"""ctypes binding of libmyrio_b200.so (include/myrio_b200.h).

No fallback of any kind: if the CUDA library has not been built (python
__graft_entry__.py build / make -C myrio_b200/csrc) loading raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmyrio_b200.so")

OK = 0
(ERR_INVALID_K, ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_BAD_QUALITY, ERR_UNSUPPORTED, ERR_EMPTY, ERR_NONFINITE, ERR_BAD_BASE,
 ERR_COMM, ERR_FORMAT) = range(1, 12)
UNIQUE_ID_BYTES = 128
SIM_COSINE, SIM_JACARD_TANIMOTO, SIM_OVERLAP = 0, 1, 2
VAL_F32, VAL_U16 = 0, 1

# every symbol include/myrio_b200.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "myrio_last_error", "myrio_version", "myrio_ctx_create", "myrio_ctx_destroy", "myrio_ctx_sync",
    "myrio_ctx_launch_count", "myrio_ctx_set_kmer_bit_order", "myrio_ctx_set_topx_bound_pruning", "myrio_ctx_stream", "myrio_prof_enable", "myrio_prof_reset", "myrio_prof_query", "myrio_bench_smem_bandwidth", "myrio_bench_clock_probe", "myrio_bench_clock_read",
    "myrio_reads_upload", "myrio_reads_from_fastq", "myrio_leaves_upload", "myrio_leaves_from_fasta", "myrio_seqs_count", "myrio_seqs_free", "myrio_count_reads",
    "myrio_count_leaves", "myrio_svecs_info", "myrio_svecs_download", "myrio_svecs_upload",
    "myrio_svecs_normalize", "myrio_svecs_free", "myrio_index_build", "myrio_index_info",
    "myrio_index_export_tile", "myrio_index_free", "myrio_score_all", "myrio_score_topx",
    "myrio_score_topx_async", "myrio_score_all_direct", "myrio_score_topx_direct",
    "myrio_score_topx_direct_async", "myrio_score_topx_seed_async", "myrio_score_topx_sweep_async",
    "myrio_topx_merge_composites_async", "myrio_topx_merge_async", "myrio_score_stats", "myrio_cluster",
    "myrio_cluster_sharded",
    "myrio_compute_cluster_centroid", "myrio_reweight_fingerprints", "myrio_pairwise", "myrio_cluster_stats", "myrio_svecs_device_ptrs",
    "myrio_store_new", "myrio_store_from_bytes", "myrio_store_to_bytes", "myrio_bytes_free", "myrio_store_save",
    "myrio_store_load", "myrio_store_info", "myrio_store_tree", "myrio_store_seqs", "myrio_store_counts_info",
    "myrio_store_counts_host", "myrio_store_leaves", "myrio_store_counts", "myrio_store_append_counts",
    "myrio_store_overwrite_counts", "myrio_store_append_counts_host", "myrio_store_shrink", "myrio_store_free",
    "myrio_taxonomy_create", "myrio_taxonomy_free", "myrio_results_from_compute_tree", "myrio_results_cut",
    "myrio_results_info", "myrio_results_tree", "myrio_results_free", "myrio_in_database_confidence",
    "myrio_comm_unique_id", "myrio_comm_init", "myrio_comm_destroy", "myrio_comm_info", "myrio_score_topx_sharded",
]


# int32_t (*myrio_allgather_fn)(void *user, void *host_buf, uint64_t bytes_per_rank)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64)


class Shard(C.Structure):
    """myrio_shard: this rank's place in a row-sharded myrio_cluster_sharded call"""
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("allgather", ALLGATHER_FN), ("user", C.c_void_p)]


class MyrioError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"myrio_b200 error {code}: {msg}")
        self.code = code


class ClusterParams(C.Structure):
    """myrio_cluster_params (ClusteringParameters, myrio-core/src/clustering.rs:15-36)"""
    _fields_ = [
        ("k", C.c_uint32),
        ("t1", C.c_float),
        ("t2", C.c_uint64),
        ("similarity", C.c_int32),
        ("expected_nb_of_clusters", C.c_uint64),
        ("eta_improvement", C.c_float),
        ("nb_iters_max", C.c_uint64),
        ("has_silhouette_trimming", C.c_int32),
        ("silhouette_trimming", C.c_float),
    ]


class ResultsParams(C.Structure):
    """myrio_results_params (search parameters, myrio-cli/src/app/config.rs:158-175)"""
    _fields_ = [
        ("max_amount_of_leaves_per_branch", C.c_uint64),
        ("nb_best_analysis", C.c_uint64),
        ("lambda_leaf", C.c_float), ("lambda_branch", C.c_float), ("mu", C.c_float),
        ("gamma", C.c_float), ("delta", C.c_float), ("epsilon", C.c_float),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    P = C.POINTER
    L.myrio_last_error.restype = C.c_char_p
    L.myrio_version.restype = i32
    L.myrio_ctx_create.argtypes = [i32, P(vp)]
    L.myrio_ctx_destroy.argtypes = [vp]
    L.myrio_ctx_destroy.restype = None
    L.myrio_ctx_sync.argtypes = [vp]
    L.myrio_ctx_launch_count.argtypes = [vp]
    L.myrio_ctx_launch_count.restype = u64
    L.myrio_ctx_stream.argtypes = [vp]
    L.myrio_ctx_stream.restype = vp
    L.myrio_prof_enable.argtypes = [vp, i32]
    L.myrio_prof_reset.argtypes = [vp]
    L.myrio_prof_query.argtypes = [vp, C.c_char_p, P(C.c_double), P(u64)]
    L.myrio_bench_smem_bandwidth.argtypes = [vp, P(C.c_double), P(C.c_double)]
    L.myrio_bench_clock_probe.argtypes = [vp, u32]
    L.myrio_bench_clock_read.argtypes = [vp, u32, P(C.c_double)]
    L.myrio_reads_upload.argtypes = [vp, vp, vp, vp, vp, u64, P(vp)]
    L.myrio_reads_from_fastq.argtypes = [vp, vp, vp, vp, u64, u64, C.c_float, C.c_uint8, vp, P(vp)]
    L.myrio_leaves_upload.argtypes = [vp, vp, vp, vp, u64, P(vp)]
    L.myrio_leaves_from_fasta.argtypes = [vp, vp, vp, u64, vp, P(vp)]
    L.myrio_seqs_count.argtypes = [vp, P(u64)]
    L.myrio_seqs_free.argtypes = [vp]
    L.myrio_seqs_free.restype = None
    L.myrio_count_reads.argtypes = [vp, vp, u32, C.c_float, P(vp)]
    L.myrio_count_leaves.argtypes = [vp, vp, u32, u32, u64, u64, P(vp)]
    L.myrio_svecs_info.argtypes = [vp, P(u64), P(u64), P(i32)]
    L.myrio_svecs_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.myrio_svecs_upload.argtypes = [vp, u64, vp, vp, vp, i32, P(vp)]
    L.myrio_svecs_normalize.argtypes = [vp, vp, P(vp)]
    L.myrio_svecs_free.argtypes = [vp]
    L.myrio_svecs_free.restype = None
    L.myrio_svecs_device_ptrs.argtypes = [vp, P(vp), P(vp), P(vp), P(vp)]
    L.myrio_index_build.argtypes = [vp, vp, u32, u32, P(vp)]
    L.myrio_index_info.argtypes = [vp, P(u64), P(u64), P(u64), P(u64), P(u32)]
    L.myrio_index_export_tile.argtypes = [vp, vp, u64, P(u64), P(u64), vp, vp, vp, vp]
    L.myrio_index_free.argtypes = [vp]
    L.myrio_index_free.restype = None
    L.myrio_score_all.argtypes = [vp, vp, vp, u64, u64, i32, vp]
    L.myrio_score_topx.argtypes = [vp, vp, vp, u64, u64, u32, i32, vp, vp]
    L.myrio_score_topx_async.argtypes = [vp, vp, vp, u64, u64, u32, i32, vp, vp]
    L.myrio_ctx_set_kmer_bit_order.argtypes = [vp, i32]
    L.myrio_ctx_set_topx_bound_pruning.argtypes = [vp, i32]
    L.myrio_score_all_direct.argtypes = [vp, vp, vp, u64, u64, i32, vp]
    L.myrio_score_topx_direct.argtypes = [vp, vp, u32, vp, u64, u64, u32, i32, vp, vp]
    L.myrio_score_topx_direct_async.argtypes = [vp, vp, u32, vp, u64, u64, u32, i32, vp, vp]
    L.myrio_score_topx_seed_async.argtypes = [vp, vp, vp, u64, u64, u32, i32, vp]
    L.myrio_score_topx_sweep_async.argtypes = [vp, vp, u64, u32, i32, vp, vp]
    L.myrio_topx_merge_composites_async.argtypes = [vp, vp, u32, u64, u32, vp, vp]
    L.myrio_topx_merge_async.argtypes = [vp, vp, vp, u32, u64, u32, vp, vp]
    L.myrio_score_stats.argtypes = [vp, P(u64), P(u64), P(u64)]
    L.myrio_store_new.argtypes = [C.c_char_p, u32, u64, vp, vp, vp, vp, u64, vp, vp, vp, P(vp)]
    L.myrio_store_from_bytes.argtypes = [vp, u64, u32, P(vp)]
    L.myrio_store_to_bytes.argtypes = [vp, i32, u32, P(vp), P(u64)]
    L.myrio_bytes_free.argtypes = [vp]
    L.myrio_bytes_free.restype = None
    L.myrio_store_save.argtypes = [vp, C.c_char_p, i32, u32]
    L.myrio_store_load.argtypes = [C.c_char_p, u32, P(vp)]
    L.myrio_store_info.argtypes = [vp, P(u64), P(u64), P(u64), P(u32), P(u64), P(u64), P(u64)]
    L.myrio_store_tree.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.myrio_store_seqs.argtypes = [vp, vp, vp, vp]
    L.myrio_store_counts_info.argtypes = [vp, u64, P(u64)]
    L.myrio_store_counts_host.argtypes = [vp, u64, vp, vp, vp, vp]
    L.myrio_store_leaves.argtypes = [vp, vp, P(vp)]
    L.myrio_store_counts.argtypes = [vp, vp, u64, P(vp)]
    L.myrio_store_append_counts.argtypes = [vp, vp, u64, vp]
    L.myrio_store_overwrite_counts.argtypes = [vp, vp, u64, vp]
    L.myrio_store_append_counts_host.argtypes = [vp, u64, vp, vp, vp, vp]
    L.myrio_store_shrink.argtypes = [vp]
    L.myrio_store_free.argtypes = [vp]
    L.myrio_store_free.restype = None
    L.myrio_taxonomy_create.argtypes = [vp, u32, u64, vp, vp, u64, P(vp)]
    L.myrio_taxonomy_free.argtypes = [vp]
    L.myrio_taxonomy_free.restype = None
    L.myrio_results_from_compute_tree.argtypes = [vp, vp, vp, vp, vp, u64, u64, i32, P(ResultsParams), P(vp)]
    L.myrio_results_cut.argtypes = [vp, u64, P(vp)]
    L.myrio_results_info.argtypes = [vp, P(u64), P(u64), P(u32)]
    L.myrio_results_tree.argtypes = [vp] + [vp] * 12
    L.myrio_results_free.argtypes = [vp]
    L.myrio_results_free.restype = None
    L.myrio_in_database_confidence.argtypes = [u64, vp, vp, vp, vp, vp, P(C.c_float)]
    L.myrio_comm_unique_id.argtypes = [vp]
    L.myrio_comm_init.argtypes = [vp, u32, u32, vp]
    L.myrio_comm_destroy.argtypes = [vp]
    L.myrio_comm_info.argtypes = [vp, P(u32), P(u32), P(i32)]
    L.myrio_score_topx_sharded.argtypes = [vp, vp, vp, u64, u64, u32, i32, vp, vp, u64]
    L.myrio_cluster.argtypes = [vp, vp, P(ClusterParams), vp, vp, P(u32), vp]
    L.myrio_cluster_sharded.argtypes = [vp, vp, P(ClusterParams), vp, P(Shard), vp, P(u32), vp]
    L.myrio_compute_cluster_centroid.argtypes = [vp, vp, vp, u64, P(vp)]
    L.myrio_reweight_fingerprints.argtypes = [vp, vp, vp, i32, C.c_float, P(vp)]
    L.myrio_pairwise.argtypes = [vp, vp, vp, i32, vp]
    L.myrio_cluster_stats.argtypes = [vp, P(u64), P(u64)]
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ("myrio_last_error", "myrio_ctx_destroy", "myrio_ctx_launch_count", "myrio_seqs_free",
                        "myrio_svecs_free", "myrio_index_free", "myrio_version", "myrio_ctx_stream",
                        "myrio_bytes_free", "myrio_store_free", "myrio_taxonomy_free", "myrio_results_free"):
            f.restype = i32
    _lib = L
    return L


def check(status: int) -> None:
    if status != OK:
        msg = load().myrio_last_error()
        raise MyrioError(status, msg.decode() if msg else "")
