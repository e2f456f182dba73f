"""ctypes binding of the CPU oracle (oracle/myrio_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
(myrio_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmyrio_oracle.so")

SIM_COSINE, SIM_JACARD_TANIMOTO, SIM_OVERLAP = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "myrio_oracle.c")
    hdr = os.path.join(_HERE, "myrio_oracle.h")
    extra = [os.path.join(_HERE, f) for f in ("myrtree_oracle.c", "results_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in [src, hdr] + extra
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


class _ClusterParams(C.Structure):
    _fields_ = [
        ("t2", C.c_uint64),
        ("similarity", C.c_int),
        ("n_init", C.c_size_t),
        ("init_row_off", C.c_void_p),
        ("init_keys", C.c_void_p),
        ("init_vals", C.c_void_p),
        ("expected_nb_of_clusters", C.c_size_t),
        ("eta_improvement", C.c_float),
        ("nb_iters_max", C.c_size_t),
        ("has_silhouette", C.c_int),
        ("ssdcf", C.c_float),
        ("threads", C.c_int),
    ]


class _StoreDesc(C.Structure):
    _fields_ = [
        ("gene", C.c_char_p), ("root_rank", C.c_uint32), ("n_nodes", C.c_uint64), ("kind", C.c_void_p),
        ("arg", C.c_void_p), ("name_off", C.c_void_p), ("names", C.c_void_p), ("n_payloads", C.c_uint64),
        ("seq_codes", C.c_void_p), ("seq_off", C.c_void_p), ("n_k", C.c_uint64), ("k_precomputed", C.c_void_p),
        ("row_off", C.c_void_p), ("keys", C.c_void_p), ("vals", C.c_void_p), ("rescale", C.c_void_p),
    ]


class _ResultsParams(C.Structure):
    _fields_ = [("max_amount_of_leaves_per_branch", C.c_uint64), ("nb_best_analysis", C.c_uint64),
                ("lambda_leaf", C.c_float), ("lambda_branch", C.c_float), ("mu", C.c_float), ("gamma", C.c_float),
                ("delta", C.c_float), ("epsilon", C.c_float)]


class _ResultsView(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_payloads", C.c_uint64), ("kind", C.c_void_p), ("arg", C.c_void_p),
                ("src", C.c_void_p), ("nb_leaves", C.c_void_p), ("mean", C.c_void_p), ("max", C.c_void_p),
                ("pool_score", C.c_void_p), ("conf", C.c_void_p), ("conf_state", C.c_void_p),
                ("force_keep", C.c_void_p), ("payload_score", C.c_void_p), ("payload_orig", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.myo_phred_correct_table.restype = C.POINTER(C.c_float)
        L.myo_kmer_key.restype = C.c_uint64
        L.myo_kmer_key.argtypes = [_u8p, C.c_int]
        L.myo_set_kmer_msb_first.restype = None
        L.myo_set_kmer_msb_first.argtypes = [C.c_int]
        L.myo_sort_reduce_f32.restype = C.c_size_t
        L.myo_sort_reduce_f32.argtypes = [_u64p, C.c_size_t, _u64p, _f32p]
        L.myo_sort_reduce_u16.restype = C.c_size_t
        L.myo_sort_reduce_u16.argtypes = [_u64p, C.c_size_t, _u64p, _u16p]
        L.myo_count_reads.restype = C.c_int
        L.myo_count_reads.argtypes = [_u8p, _u8p, _u64p, C.c_size_t, C.c_int, C.c_float, C.c_int,
                                      _u64p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.myo_resample_pick.restype = C.c_uint32
        L.myo_resample_pick.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.myo_count_leaves.restype = C.c_int
        L.myo_count_leaves.argtypes = [_u8p, _u64p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64,
                                       C.c_uint64, C.c_int, _u64p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.myo_normalize_rows_f32.restype = None
        L.myo_normalize_rows_f32.argtypes = [_u64p, C.c_size_t, _f32p, C.c_int]
        L.myo_normalize_rows_u16.restype = None
        L.myo_normalize_rows_u16.argtypes = [_u64p, C.c_size_t, _u16p, _f32p, C.c_int]
        L.myo_dot.restype = C.c_float
        L.myo_dot.argtypes = [_u64p, _f32p, C.c_size_t, _u64p, _f32p, C.c_size_t]
        L.myo_similarity.restype = C.c_float
        L.myo_similarity.argtypes = [C.c_int, C.c_int, _u64p, _f32p, C.c_size_t, _u64p, _f32p,
                                     C.c_size_t]
        L.myo_score_all.restype = None
        L.myo_score_all.argtypes = [C.c_int, _u64p, _u64p, _f32p, C.c_size_t, _u64p, _f32p,
                                    C.c_size_t, C.c_int, _f32p]
        L.myo_topx.restype = None
        L.myo_topx.argtypes = [_f32p, C.c_size_t, C.c_uint32, C.c_size_t, _f32p, _u32p]
        L.myo_add.restype = C.c_size_t
        L.myo_add.argtypes = [_u64p, _f32p, C.c_size_t, _u64p, _f32p, C.c_size_t, _u64p, _f32p]
        L.myo_centroid.restype = C.c_size_t
        L.myo_centroid.argtypes = [_u64p, _u64p, _f32p, _u64p, C.c_size_t, _u64p, _f32p]
        L.myo_invert.restype = None
        L.myo_invert.argtypes = [_u64p, _u64p, _f32p, C.c_size_t, C.POINTER(C.c_size_t),
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.myo_cluster.restype = C.c_int
        L.myo_cluster.argtypes = [_u64p, _u64p, _f32p, _u64p, C.c_size_t,
                                  C.POINTER(_ClusterParams), _i32p, C.POINTER(C.c_uint32),
                                  C.c_void_p]
        L.myo_sort_reduce_pairs_f32.restype = C.c_size_t
        L.myo_sort_reduce_pairs_f32.argtypes = [_u64p, _f32p, C.c_size_t, _u64p, _f32p]
        L.myo_reweight_fingerprints.restype = C.c_size_t
        L.myo_reweight_fingerprints.argtypes = [C.c_int, _u64p, _u64p, _f32p, C.c_size_t, _u64p, _f32p, C.c_size_t,
                                                C.c_float, _u64p, _f32p]
        L.myo_set_silhouette_sample.restype = None
        L.myo_set_silhouette_sample.argtypes = [C.c_void_p]
        L.myo_pairwise.restype = None
        L.myo_pairwise.argtypes = [C.c_int, _u64p, _u64p, _f32p, C.c_size_t, _u64p, _u64p, _f32p,
                                   C.c_size_t, C.c_int, _f32p]
        L.myo_max_threads.restype = C.c_int
        L.myo_set_num_threads.restype = None
        L.myo_set_num_threads.argtypes = [C.c_int]
        L.myo_fastq_preprocess.restype = C.c_longlong
        L.myo_fastq_preprocess.argtypes = [_u8p, _u8p, _u64p, C.c_size_t, C.c_size_t, C.c_float, C.c_uint8,
                                           _u8p, _u8p, _u8p, _u64p]
        L.myo_fasta_to_iupac.restype = C.c_longlong
        L.myo_fasta_to_iupac.argtypes = [_u8p, _u64p, C.c_size_t, _u8p, _u8p, _u64p]
        L.myo_store_encode.restype = C.c_int
        L.myo_store_encode.argtypes = [C.POINTER(_StoreDesc), C.c_int, C.c_uint, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_uint64)]
        L.myo_store_decode.restype = C.c_int
        L.myo_store_decode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.myo_store_view.restype = None
        L.myo_store_view.argtypes = [C.c_void_p, C.POINTER(_StoreDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.myo_store_free.restype = None
        L.myo_store_free.argtypes = [C.c_void_p]
        L.myo_store_section.restype = C.c_int
        L.myo_store_section.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.myo_bytes_free.restype = None
        L.myo_bytes_free.argtypes = [C.c_void_p]
        L.myo_results_from_scores.restype = C.c_int
        L.myo_results_from_scores.argtypes = [C.c_uint32, C.c_uint64, _u8p, _u64p, _f32p, C.c_uint64,
                                              C.POINTER(_ResultsParams), C.POINTER(C.c_void_p)]
        L.myo_results_cut.restype = C.c_int
        L.myo_results_cut.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.myo_results_view.restype = None
        L.myo_results_view.argtypes = [C.c_void_p, C.POINTER(_ResultsView)]
        L.myo_results_free.restype = None
        L.myo_results_free.argtypes = [C.c_void_p]
        L.myo_in_database_confidence.restype = C.c_float
        L.myo_in_database_confidence.argtypes = [C.c_uint64, _u64p, _u64p, _f32p, _u8p, _f32p]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().myo_max_threads())


def set_num_threads(n: int) -> None:
    """omp_set_num_threads (a launcher may have exported OMP_NUM_THREADS=1)."""
    lib().myo_set_num_threads(int(n))


@dataclass
class Csr:
    """A batch of sparse vectors: row r = keys[row_off[r]:row_off[r+1]] (sorted) + vals."""
    row_off: np.ndarray  # uint64 [n+1]
    keys: np.ndarray     # uint64 [nnz]
    vals: np.ndarray     # float32 or uint16 [nnz]
    aux: np.ndarray | None = None  # nb_hck (uint64) for reads, rescale (float32) for leaves

    @property
    def n_rows(self) -> int:
        return len(self.row_off) - 1

    def row(self, r: int):
        b, e = int(self.row_off[r]), int(self.row_off[r + 1])
        return self.keys[b:e], self.vals[b:e]


def phred_correct_table() -> np.ndarray:
    p = lib().myo_phred_correct_table()
    return np.ctypeslib.as_array(p, shape=(128,)).copy()


def set_kmer_msb_first(on: bool) -> None:
    """The alternative key packing (base j at bits 2(k-1-j)); mirrors Context.set_kmer_bit_order."""
    lib().myo_set_kmer_msb_first(int(bool(on)))


def kmer_key(bases: np.ndarray, k: int) -> int:
    return int(lib().myo_kmer_key(np.ascontiguousarray(bases, np.uint8), k))


def sort_reduce_f32(singles):
    s = np.array(singles, dtype=np.uint64)
    k = np.zeros(max(1, len(s)), np.uint64)
    v = np.zeros(max(1, len(s)), np.float32)
    d = lib().myo_sort_reduce_f32(s, len(s), k, v)
    return k[:d], v[:d]


def sort_reduce_u16(singles):
    s = np.array(singles, dtype=np.uint64)
    k = np.zeros(max(1, len(s)), np.uint64)
    v = np.zeros(max(1, len(s)), np.uint16)
    d = lib().myo_sort_reduce_u16(s, len(s), k, v)
    return k[:d], v[:d]


def fastq_preprocess(seq_ascii: np.ndarray, qual_ascii: np.ndarray, base_off: np.ndarray,
                     min_length: int = 50, min_mean_qual: float = 12.0, max_qual: int = 255):
    """io/mod.rs:46-48 + data/myrseq.rs:113-131. Returns (codes, qual, kept_base_off, keep)."""
    seq_ascii = np.ascontiguousarray(seq_ascii, np.uint8)
    qual_ascii = np.ascontiguousarray(qual_ascii, np.uint8)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n = len(base_off) - 1
    total = max(1, len(seq_ascii))
    codes = np.zeros(total, np.uint8)
    qual = np.zeros(total, np.uint8)
    keep = np.zeros(max(1, n), np.uint8)
    koff = np.zeros(n + 1, np.uint64)
    sa = seq_ascii if len(seq_ascii) else np.zeros(1, np.uint8)
    qa = qual_ascii if len(qual_ascii) else np.zeros(1, np.uint8)
    k = lib().myo_fastq_preprocess(sa, qa, base_off, n, min_length, min_mean_qual, max_qual, codes, qual, keep,
                                   koff)
    if k < 0:
        raise ValueError(f"non-ACGT base in record {-k - 1}")
    koff = koff[:k + 1]
    return codes[:int(koff[-1])], qual[:int(koff[-1])], koff, keep[:n]


def sort_reduce_pairs_f32(pairs):
    """data/sparse.rs:274-313 from_unsorted_pairs."""
    k = np.array([p[0] for p in pairs], np.uint64)
    v = np.array([p[1] for p in pairs], np.float32)
    ko = np.zeros(max(1, len(k)), np.uint64)
    vo = np.zeros(max(1, len(k)), np.float32)
    if len(k) == 0:
        return ko[:0], vo[:0]
    d = lib().myo_sort_reduce_pairs_f32(k, v, len(k), ko, vo)
    return ko[:d], vo[:d]


def reweight_fingerprints(local: "Csr", qk, qv, alpha: float = 2.5, sim: int = SIM_COSINE):
    """myrio-cli/src/app/process/run.rs:243-262"""
    qk, qv, qn = _kv(qk, qv)
    nnz = int(local.row_off[-1])
    ck = np.zeros(nnz + 1, np.uint64)
    cv = np.zeros(nnz + 1, np.float32)
    n = lib().myo_reweight_fingerprints(sim, np.ascontiguousarray(local.row_off, np.uint64),
                                        np.ascontiguousarray(local.keys, np.uint64) if nnz else np.zeros(1, np.uint64),
                                        np.ascontiguousarray(local.vals, np.float32) if nnz else np.zeros(1, np.float32),
                                        local.n_rows, qk, qv, qn, alpha, ck, cv)
    return ck[:n], cv[:n]


def fasta_to_iupac(seq_ascii: np.ndarray, base_off: np.ndarray):
    """tax/store.rs:178-188. Returns (codes4 one per byte, kept_base_off, keep)."""
    seq_ascii = np.ascontiguousarray(seq_ascii, np.uint8)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n = len(base_off) - 1
    codes = np.zeros(max(1, len(seq_ascii)), np.uint8)
    keep = np.zeros(max(1, n), np.uint8)
    koff = np.zeros(n + 1, np.uint64)
    sa = seq_ascii if len(seq_ascii) else np.zeros(1, np.uint8)
    k = lib().myo_fasta_to_iupac(sa, base_off, n, codes, keep, koff)
    koff = koff[:k + 1]
    return codes[:int(koff[-1])], koff, keep[:n].astype(bool)


def count_reads(bases: np.ndarray, qual: np.ndarray, base_off: np.ndarray, k: int, cutoff: float,
                threads: int = 0) -> Csr:
    """data/myrseq.rs:76-111 over a batch. bases: one 2-bit code per byte."""
    bases = np.ascontiguousarray(bases, np.uint8)
    qual = np.ascontiguousarray(qual, np.uint8)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n = len(base_off) - 1
    if len(bases) == 0:
        bases = np.zeros(1, np.uint8)
        qual = np.zeros(1, np.uint8)
    row_off = np.zeros(n + 1, np.uint64)
    hck = np.zeros(max(1, n), np.uint64)
    thr = threads if threads > 0 else max_threads()
    e = lib().myo_count_reads(bases, qual, base_off, n, k, cutoff, thr, row_off, None, None,
                              hck.ctypes.data)
    if e:
        raise ValueError(f"oracle count_reads error {e}")
    nnz = int(row_off[n])
    keys = np.zeros(max(1, nnz), np.uint64)
    vals = np.zeros(max(1, nnz), np.float32)
    e = lib().myo_count_reads(bases, qual, base_off, n, k, cutoff, thr, row_off,
                              keys.ctypes.data, vals.ctypes.data, hck.ctypes.data)
    if e:
        raise ValueError(f"oracle count_reads error {e}")
    return Csr(row_off, keys[:nnz], vals[:nnz], hck[:n])


def count_leaves(iupac: np.ndarray, base_off: np.ndarray, k: int, nb_resamples: int = 32,
                 seed: int = 0, leaf_id_base: int = 0, threads: int = 0) -> Csr:
    """tax/mod.rs:70-148 via tax/store.rs:265-285. iupac: one 4-bit code per byte."""
    iupac = np.ascontiguousarray(iupac, np.uint8)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n = len(base_off) - 1
    if len(iupac) == 0:
        iupac = np.zeros(1, np.uint8)
    row_off = np.zeros(n + 1, np.uint64)
    rescale = np.zeros(max(1, n), np.float32)
    e = lib().myo_count_leaves(iupac, base_off, n, k, nb_resamples, seed, leaf_id_base, threads,
                               row_off, None, None, rescale.ctypes.data)
    if e:
        raise ValueError(f"oracle count_leaves error {e}")
    nnz = int(row_off[n])
    keys = np.zeros(max(1, nnz), np.uint64)
    vals = np.zeros(max(1, nnz), np.uint16)
    e = lib().myo_count_leaves(iupac, base_off, n, k, nb_resamples, seed, leaf_id_base, threads,
                               row_off, keys.ctypes.data, vals.ctypes.data, rescale.ctypes.data)
    if e:
        raise ValueError(f"oracle count_leaves error {e}")
    return Csr(row_off, keys[:nnz], vals[:nnz], rescale[:n])


def normalize(csr: Csr, threads: int = 0) -> Csr:
    """f32 rows: data/sparse.rs:380-383; u16 rows: tax/compute.rs:105-106."""
    n = csr.n_rows
    if csr.vals.dtype == np.uint16:
        out = np.zeros(max(1, len(csr.vals)), np.float32)
        vin = csr.vals if len(csr.vals) else np.zeros(1, np.uint16)
        lib().myo_normalize_rows_u16(csr.row_off, n, vin, out, threads)
        return Csr(csr.row_off, csr.keys, out[:len(csr.vals)], csr.aux)
    vals = np.array(csr.vals, dtype=np.float32, copy=True)
    if len(vals) == 0:
        return Csr(csr.row_off, csr.keys, vals, csr.aux)
    lib().myo_normalize_rows_f32(csr.row_off, n, vals, threads)
    return Csr(csr.row_off, csr.keys, vals, csr.aux)


def _kv(k, v):
    k = np.ascontiguousarray(k, np.uint64)
    v = np.ascontiguousarray(v, np.float32)
    if len(k) == 0:
        return np.zeros(1, np.uint64), np.zeros(1, np.float32), 0
    return k, v, len(k)


def dot(ak, av, bk, bv) -> np.float32:
    ak, av, an = _kv(ak, av)
    bk, bv, bn = _kv(bk, bv)
    return np.float32(lib().myo_dot(ak, av, an, bk, bv, bn))


def similarity(sim: int, pre_normalized: bool, ak, av, bk, bv) -> np.float32:
    ak, av, an = _kv(ak, av)
    bk, bv, bn = _kv(bk, bv)
    return np.float32(lib().myo_similarity(sim, int(pre_normalized), ak, av, an, bk, bv, bn))


def score_all(leaves: Csr, qk, qv, sim: int = SIM_COSINE, threads: int = 0) -> np.ndarray:
    """tax/results.rs:82-93 for one query against normalised leaf vectors."""
    qk, qv, qn = _kv(qk, qv)
    out = np.zeros(max(1, leaves.n_rows), np.float32)
    lk = leaves.keys if len(leaves.keys) else np.zeros(1, np.uint64)
    lv = leaves.vals if len(leaves.vals) else np.zeros(1, np.float32)
    lib().myo_score_all(sim, leaves.row_off, lk, lv, leaves.n_rows, qk, qv, qn, threads, out)
    return out[:leaves.n_rows]


def topx(scores: np.ndarray, x: int, id_base: int = 0):
    scores = np.ascontiguousarray(scores, np.float32)
    ts = np.zeros(max(1, x), np.float32)
    ti = np.zeros(max(1, x), np.uint32)
    s = scores if len(scores) else np.zeros(1, np.float32)
    lib().myo_topx(s, len(scores), id_base, x, ts, ti)
    return ts[:x], ti[:x]


def add(ak, av, bk, bv):
    ak, av, an = _kv(ak, av)
    bk, bv, bn = _kv(bk, bv)
    ck = np.zeros(an + bn + 1, np.uint64)
    cv = np.zeros(an + bn + 1, np.float32)
    n = lib().myo_add(ak, av, an, bk, bv, bn, ck, cv)
    return ck[:n], cv[:n]


def centroid(csr: Csr, member_ids):
    """clustering.rs:44-53"""
    ids = np.ascontiguousarray(member_ids, np.uint64)
    total = int(sum(int(csr.row_off[i + 1] - csr.row_off[i]) for i in ids))
    ck = np.zeros(total + 1, np.uint64)
    cv = np.zeros(total + 1, np.float32)
    idp = ids if len(ids) else np.zeros(1, np.uint64)
    n = lib().myo_centroid(csr.row_off, csr.keys, np.ascontiguousarray(csr.vals, np.float32), idp,
                           len(ids), ck, cv)
    return ck[:n], cv[:n]


def invert(leaves: Csr):
    """Definition of the K3 inverted index: (ukeys, offs, post_leaf, post_val)."""
    n = leaves.n_rows
    nu = C.c_size_t(0)
    lk = leaves.keys if len(leaves.keys) else np.zeros(1, np.uint64)
    lv = np.ascontiguousarray(leaves.vals, np.float32) if len(leaves.vals) else np.zeros(1, np.float32)
    lib().myo_invert(leaves.row_off, lk, lv, n, C.byref(nu), None, None, None, None)
    U, P = nu.value, int(leaves.row_off[n])
    uk = np.zeros(max(1, U), np.uint64)
    of = np.zeros(U + 1, np.uint64)
    pl = np.zeros(max(1, P), np.uint32)
    pv = np.zeros(max(1, P), np.float32)
    lib().myo_invert(leaves.row_off, lk, lv, n, C.byref(nu), uk.ctypes.data, of.ctypes.data,
                     pl.ctypes.data, pv.ctypes.data)
    return uk[:U], of, pl[:P], pv[:P]


def cluster(counts: Csr, t2: int, init: Csr | None, expected_nb_of_clusters: int,
            eta_improvement: float = 1e-4, nb_iters_max: int = 20, silhouette: float | None = None,
            sim: int = SIM_COSINE, threads: int = 0, silhouette_sample: np.ndarray | None = None):
    """clustering.rs:55-270 from the step-1 counts. Returns (assign, n_iters, sil).
    silhouette_sample (bool per read): compute the silhouette score of these reads only and skip the
    trimming (full-size checks; see myo_set_silhouette_sample)."""
    n = counts.n_rows
    p = _ClusterParams()
    p.t2 = t2
    p.similarity = sim
    keep = []
    if init is not None and init.n_rows > 0:
        ro = np.ascontiguousarray(init.row_off, np.uint64)
        ik = np.ascontiguousarray(init.keys, np.uint64)
        iv = np.ascontiguousarray(init.vals, np.float32)
        keep += [ro, ik, iv]
        p.n_init = init.n_rows
        p.init_row_off = ro.ctypes.data
        p.init_keys = ik.ctypes.data
        p.init_vals = iv.ctypes.data
    else:
        p.n_init = 0
    p.expected_nb_of_clusters = expected_nb_of_clusters
    p.eta_improvement = eta_improvement
    p.nb_iters_max = nb_iters_max
    p.has_silhouette = 0 if silhouette is None else 1
    p.ssdcf = 0.0 if silhouette is None else silhouette
    p.threads = threads
    assign = np.zeros(max(1, n), np.int32)
    sil = np.zeros(max(1, n), np.float32)
    iters = C.c_uint32(0)
    ck = counts.keys if len(counts.keys) else np.zeros(1, np.uint64)
    cv = np.ascontiguousarray(counts.vals, np.float32) if len(counts.vals) else np.zeros(1, np.float32)
    hck = np.ascontiguousarray(counts.aux, np.uint64) if n else np.zeros(1, np.uint64)
    mask = None
    if silhouette_sample is not None:
        mask = np.ascontiguousarray(silhouette_sample, np.uint8)
        assert len(mask) == n
        lib().myo_set_silhouette_sample(mask.ctypes.data)
    try:
        e = lib().myo_cluster(counts.row_off, ck, cv, hck, n, C.byref(p), assign, C.byref(iters),
                              sil.ctypes.data)
    finally:
        if mask is not None:
            lib().myo_set_silhouette_sample(None)
    if e:
        raise ValueError(f"oracle cluster error {e}")
    return assign[:n], int(iters.value), sil[:n]


def pairwise(rows: Csr, cols: Csr, sim: int = SIM_COSINE, threads: int = 0) -> np.ndarray:
    out = np.zeros((rows.n_rows, cols.n_rows), np.float32)
    if out.size == 0:
        return out
    lib().myo_pairwise(sim, rows.row_off, rows.keys, np.ascontiguousarray(rows.vals, np.float32),
                       rows.n_rows, cols.row_off, cols.keys,
                       np.ascontiguousarray(cols.vals, np.float32), cols.n_rows, threads, out)
    return out


# ------------------------------------------------------------------ .myrtree (myrtree_oracle.c)

@dataclass
class StoreContent:
    """A TaxTreeStore as flat arrays: taxonomy in pre-order, one 4-bit IUPAC code per byte, counts per k as Csr
    (vals uint16, aux = rescale float32)."""
    gene: str
    root_rank: int
    kind: np.ndarray        # uint8 [n_nodes]  0 branch, 1 leaf
    arg: np.ndarray         # uint64 [n_nodes] children count | payload_id
    names: list             # str per node
    seq_codes: np.ndarray   # uint8 [total symbols]
    seq_off: np.ndarray     # uint64 [n_payloads + 1]
    k_precomputed: list
    counts: list            # Csr per k


def store_encode(st: StoreContent, zstd_level: int = 6, threads: int = 1) -> bytes:
    """TaxTreeStore::to_bytes (tax/store.rs:342-405)."""
    L = lib()
    enc = [n.encode() for n in st.names]
    name_off = np.zeros(len(enc) + 1, np.uint64)
    name_off[1:] = np.cumsum([len(e) for e in enc])
    names = np.frombuffer(b"".join(enc) + b"\0", np.uint8).copy()
    kind = np.ascontiguousarray(st.kind, np.uint8)
    arg = np.ascontiguousarray(st.arg, np.uint64)
    codes = np.ascontiguousarray(st.seq_codes, np.uint8)
    if len(codes) == 0:
        codes = np.zeros(1, np.uint8)
    soff = np.ascontiguousarray(st.seq_off, np.uint64)
    nk = len(st.k_precomputed)
    kp = np.ascontiguousarray(st.k_precomputed if nk else [0], np.uint64)
    keep = []
    ptrs = [(C.c_void_p * max(1, nk))() for _ in range(4)]
    for i, c in enumerate(st.counts):
        arrs = (np.ascontiguousarray(c.row_off, np.uint64), np.ascontiguousarray(c.keys if len(c.keys) else [0], np.uint64),
                np.ascontiguousarray(c.vals if len(c.vals) else [0], np.uint16),
                np.ascontiguousarray(c.aux if len(c.aux) else [0], np.float32))
        keep.append(arrs)
        for j in range(4):
            ptrs[j][i] = arrs[j].ctypes.data
    d = _StoreDesc(st.gene.encode(), st.root_rank, len(kind), kind.ctypes.data, arg.ctypes.data, name_off.ctypes.data,
                   names.ctypes.data, len(soff) - 1, codes.ctypes.data, soff.ctypes.data, nk, kp.ctypes.data,
                   C.cast(ptrs[0], C.c_void_p), C.cast(ptrs[1], C.c_void_p), C.cast(ptrs[2], C.c_void_p),
                   C.cast(ptrs[3], C.c_void_p))
    out, n = C.c_void_p(), C.c_uint64(0)
    e = L.myo_store_encode(C.byref(d), zstd_level, threads, C.byref(out), C.byref(n))
    if e:
        raise ValueError(f"oracle store_encode error {e}")
    try:
        return C.string_at(out, n.value)
    finally:
        L.myo_bytes_free(out)


def store_decode(data: bytes) -> StoreContent:
    """TaxTreeStore::from_bytes (tax/store.rs:408-454)."""
    L = lib()
    buf = np.frombuffer(data, np.uint8)
    h = C.c_void_p()
    e = L.myo_store_decode(buf.ctypes.data if len(buf) else None, len(buf), C.byref(h))
    if e:
        raise ValueError(f"oracle store_decode error {e}")
    try:
        d = _StoreDesc()
        ptrs = [(C.c_void_p * 64)() for _ in range(4)]
        L.myo_store_view(h, C.byref(d), ptrs[0], ptrs[1], ptrs[2], ptrs[3])

        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
        nn, npay, nk = int(d.n_nodes), int(d.n_payloads), int(d.n_k)
        name_off = arr(d.name_off, nn + 1, np.uint64)
        names_raw = arr(d.names, int(name_off[-1]) if nn else 0, np.uint8).tobytes()
        names = [names_raw[int(name_off[i]):int(name_off[i + 1])].decode() for i in range(nn)]
        soff = arr(d.seq_off, npay + 1, np.uint64)
        counts = []
        for i in range(nk):
            ro = arr(ptrs[0][i], npay + 1, np.uint64)
            nnz = int(ro[-1])
            counts.append(Csr(ro, arr(ptrs[1][i], nnz, np.uint64), arr(ptrs[2][i], nnz, np.uint16),
                              arr(ptrs[3][i], npay, np.float32)))
        return StoreContent(d.gene.decode(), int(d.root_rank), arr(d.kind, nn, np.uint8), arr(d.arg, nn, np.uint64), names,
                            arr(d.seq_codes, int(soff[-1]), np.uint8), soff,
                            [int(v) for v in arr(d.k_precomputed, nk, np.uint64)], counts)
    finally:
        L.myo_store_free(h)


def store_section(data: bytes, which: int) -> bytes:
    """Decompressed section of a .myrtree file: -1 header, 0 core, 1 + c payload chunk c."""
    L = lib()
    buf = np.frombuffer(data, np.uint8)
    out, n = C.c_void_p(), C.c_uint64(0)
    e = L.myo_store_section(buf.ctypes.data, len(buf), which, C.byref(out), C.byref(n))
    if e:
        raise ValueError(f"oracle store_section error {e}")
    try:
        return C.string_at(out, n.value)
    finally:
        L.myo_bytes_free(out)


# ------------------------------------------------------------------ results (results_oracle.c)

class ResultsTree:
    """TaxTreeResults of one query (tax/results.rs): same fields as myrio_b200.api.BResTree."""

    def __init__(self, handle):
        self.h = handle
        v = _ResultsView()
        lib().myo_results_view(handle, C.byref(v))
        n, npay = int(v.n_nodes), int(v.n_payloads)

        def arr(ptr, cnt, dt):
            if cnt == 0 or not ptr:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(cnt,)).copy()
        self.kind, self.arg, self.src_node = arr(v.kind, n, np.uint8), arr(v.arg, n, np.uint64), arr(v.src, n, np.uint64)
        self.nb_leaves = arr(v.nb_leaves, n, np.uint64)
        self.mean, self.max = arr(v.mean, n, np.float32), arr(v.max, n, np.float32)
        self.pool_score, self.conf = arr(v.pool_score, n, np.float32), arr(v.conf, n, np.float32)
        self.conf_state, self.force_keep = arr(v.conf_state, n, np.uint8), arr(v.force_keep, n, np.uint8)
        self.payload_score = arr(v.payload_score, npay, np.float32)
        self.payload_src_id = arr(v.payload_orig, npay, np.uint64)

    def __del__(self):
        try:
            if self.h:
                lib().myo_results_free(self.h)
            self.h = None
        except Exception:
            pass

    def cut(self, nb_best: int) -> "ResultsTree":
        h = C.c_void_p()
        e = lib().myo_results_cut(self.h, nb_best, C.byref(h))
        if e:
            raise ValueError(f"oracle results_cut error {e}")
        return ResultsTree(h)


def results_from_scores(root_rank: int, kind, arg, scores, max_amount_of_leaves_per_branch: int = 15,
                        nb_best_analysis: int = 100, lambda_leaf: float = 2.0, lambda_branch: float = 20.0,
                        mu: float = 1.1, gamma: float = 2.5, delta: float = 0.9, epsilon: float = 0.55) -> ResultsTree:
    """TaxTreeResults::from_compute_tree after the score loop (tax/results.rs:96-252)."""
    kind = np.ascontiguousarray(kind, np.uint8)
    arg = np.ascontiguousarray(arg, np.uint64)
    scores = np.ascontiguousarray(scores, np.float32)
    p = _ResultsParams(max_amount_of_leaves_per_branch, nb_best_analysis, lambda_leaf, lambda_branch, mu, gamma, delta,
                       epsilon)
    h = C.c_void_p()
    e = lib().myo_results_from_scores(root_rank, len(kind), kind, arg, scores if len(scores) else np.zeros(1, np.float32),
                                      len(scores), C.byref(p), C.byref(h))
    if e:
        raise ValueError(f"oracle results_from_scores error {e}")
    return ResultsTree(h)


def in_database_confidence(genus_off, name_id, pool_score, conf_state, conf) -> np.float32:
    """myrio-cli/src/app/process/run.rs:405-466"""
    a = (np.ascontiguousarray(genus_off, np.uint64), np.ascontiguousarray(name_id, np.uint64),
         np.ascontiguousarray(pool_score, np.float32), np.ascontiguousarray(conf_state, np.uint8),
         np.ascontiguousarray(conf, np.float32))
    return np.float32(lib().myo_in_database_confidence(len(a[0]) - 1, *a))
