"""Host-side mirror of the `myrio-core` API for the hot path (prelude.rs:1-12),
over the C ABI of include/myrio_b200.h.  Names, argument meaning and error
behaviour follow the reference; batches replace per-item calls
(`MyrSeqs.compute_kmer_counts` == `MyrSeq::compute_kmer_counts` for every read).

Every call runs on the GPU through libmyrio_b200.so; nothing here computes on
the CPU and nothing imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import MyrioError, SIM_COSINE, SIM_JACARD_TANIMOTO, SIM_OVERLAP  # noqa: F401
from .synth import SeqSet, pack_dna2, pack_iupac4, tree_to_iupac


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU + stream (myrio_ctx)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.myrio_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.myrio_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._L.myrio_ctx_sync(self.h))

    def set_kmer_bit_order(self, msb_first: bool):
        """Integer value of a k-mer key: False (default) = bio-seq's LSB-first packing, True = MSB-first."""
        _lib.check(self._L.myrio_ctx_set_kmer_bit_order(self.h, int(bool(msb_first))))

    def set_topx_bound_pruning(self, on: bool):
        """Fused top-x: finish only the scores whose upper bound reaches the query's running bound (default) or
        every score (comparison runs); the top-x is bit-identical either way."""
        _lib.check(self._L.myrio_ctx_set_topx_bound_pruning(self.h, int(bool(on))))

    @property
    def stream(self) -> int:
        """cudaStream_t of the context as an integer (torch.cuda.ExternalStream(ctx.stream))."""
        return int(self._L.myrio_ctx_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._L.myrio_ctx_launch_count(self.h))

    def prof_enable(self, on: bool = True):
        _lib.check(self._L.myrio_prof_enable(self.h, int(on)))

    def prof_reset(self):
        _lib.check(self._L.myrio_prof_reset(self.h))

    def prof_query(self, prefix: str = ""):
        ms, n = C.c_double(0), C.c_uint64(0)
        _lib.check(self._L.myrio_prof_query(self.h, prefix.encode(), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    # ---- multi-GPU (include/myrio_b200.h "multi-GPU"): the library owns its NCCL communicator
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId through the library (rank 0; ship the 128 bytes to every rank)."""
        buf = (C.c_uint8 * _lib.UNIQUE_ID_BYTES)()
        _lib.check(_lib.load().myrio_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """Collective: every rank calls it with rank 0's id (ncclCommInitRank on this context's GPU)."""
        assert len(unique_id) == _lib.UNIQUE_ID_BYTES
        buf = (C.c_uint8 * _lib.UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        _lib.check(self._L.myrio_comm_init(self.h, nranks, rank, buf))

    def comm_destroy(self):
        _lib.check(self._L.myrio_comm_destroy(self.h))

    def comm_info(self):
        r, w, v = C.c_uint32(0), C.c_uint32(1), C.c_int32(0)
        _lib.check(self._L.myrio_comm_info(self.h, C.byref(r), C.byref(w), C.byref(v)))
        return {"rank": int(r.value), "nranks": int(w.value), "nccl_version": int(v.value)}

    def clock_probe(self, slot: int):
        """Enqueue an on-device SM clock measurement (cycles per ns over ~0.5 ms) into `slot` on the ctx stream."""
        _lib.check(self._L.myrio_bench_clock_probe(self.h, int(slot)))

    def clock_read(self, n: int):
        """The first n clock-probe slots in MHz (synchronises the stream)."""
        out = (C.c_double * max(n, 1))()
        _lib.check(self._L.myrio_bench_clock_read(self.h, int(n), out))
        return [float(out[i]) for i in range(n)]

    def bench_smem_bandwidth(self):
        """Measured shared-memory read bandwidth of this GPU (the K6 roofline denominator)."""
        g, ms = C.c_double(0), C.c_double(0)
        _lib.check(self._L.myrio_bench_smem_bandwidth(self.h, C.byref(g), C.byref(ms)))
        return {"gbps": g.value, "ms": ms.value}

    def score_stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _lib.check(self._L.myrio_score_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"query_keys": int(a.value), "key_hits": int(b.value), "postings_touched": int(c.value)}


    def cluster_stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _lib.check(self._L.myrio_cluster_stats(self.h, C.byref(a), C.byref(b)))
        return {"pair_similarities": int(a.value), "terms": int(b.value)}


class SFVecs:
    """Device-resident batch of SparseVec<T> (myrio_svecs): row r is one sorted
    (usize key, value) vector -- SFVec for f32, SparseVec<u16> for leaf counts."""

    aux_is_u64 = True

    def __init__(self, ctx: Context, handle, has_aux: bool = False):
        self.ctx, self.h, self.has_aux = ctx, handle, has_aux
        n, nnz, vt = C.c_uint64(0), C.c_uint64(0), C.c_int32(0)
        _lib.check(ctx._L.myrio_svecs_info(handle, C.byref(n), C.byref(nnz), C.byref(vt)))
        self.n_rows, self.nnz, self.vtype = int(n.value), int(nnz.value), int(vt.value)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx._L.myrio_svecs_free(self.h)
            self.h = None
        except Exception:
            pass

    @classmethod
    def upload(cls, ctx: Context, row_off, keys, vals) -> "SFVecs":
        row_off = np.ascontiguousarray(row_off, np.uint64)
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = np.ascontiguousarray(vals)
        if vals.dtype == np.uint16:
            vt = _lib.VAL_U16
        else:
            vals = np.ascontiguousarray(vals, np.float32)
            vt = _lib.VAL_F32
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_svecs_upload(ctx.h, len(row_off) - 1, _p(row_off), _p(keys), _p(vals), vt,
                                             C.byref(h)))
        return cls(ctx, h)

    def download(self, aux: bool = True):
        """(row_off, keys, vals, aux) as numpy arrays."""
        row_off = np.zeros(self.n_rows + 1, np.uint64)
        keys = np.zeros(self.nnz, np.uint64)
        vals = np.zeros(self.nnz, np.float32 if self.vtype == _lib.VAL_F32 else np.uint16)
        ax = None
        if aux and self.has_aux:
            # nb_hck (u64) for read counts, rescale factor (f32) for leaf counts
            ax = np.zeros(self.n_rows, np.uint64 if self.aux_is_u64 else np.float32)
        _lib.check(self.ctx._L.myrio_svecs_download(self.ctx.h, self.h, _p(row_off), _p(keys), _p(vals),
                                                    _p(ax)))
        return row_off, keys, vals, ax

    def into_normalized_l2(self) -> "SFVecs":
        """SparseVec::into_normalized_l2 per row (data/sparse.rs:385-388)."""
        h = C.c_void_p()
        _lib.check(self.ctx._L.myrio_svecs_normalize(self.ctx.h, self.h, C.byref(h)))
        out = SFVecs(self.ctx, h, self.has_aux)
        out.aux_is_u64 = getattr(self, "aux_is_u64", True)
        return out


class MyrSeqs:
    """A batch of MyrSeq (data/myrseq.rs:21-33) resident on the GPU."""

    def __init__(self, ctx: Context, seqs: SeqSet):
        if seqs.qual is None:
            raise ValueError("reads need quality scores")
        self.ctx, self.n = ctx, seqs.n
        words, word_off = pack_dna2(seqs)
        self._upload(words, word_off, np.ascontiguousarray(seqs.base_off, np.uint64),
                     np.ascontiguousarray(seqs.qual, np.uint8))

    def _upload(self, words, word_off, base_off, qual):
        h = C.c_void_p()
        _lib.check(self.ctx._L.myrio_reads_upload(self.ctx.h, _p(words), _p(word_off), _p(base_off), _p(qual),
                                                  self.n, C.byref(h)))
        self.h = h

    @classmethod
    def from_packed(cls, ctx, words, word_off, base_off, qual) -> "MyrSeqs":
        self = cls.__new__(cls)
        self.ctx, self.n = ctx, len(base_off) - 1
        self._upload(words, word_off, base_off, qual)
        return self

    @classmethod
    def from_fastq_records(cls, ctx, seq_ascii, qual_ascii, base_off, min_length: int = 50,
                           min_mean_qual: float = 12.0, max_qual: int = 255):
        """io::read_fastq's record conversion (io/mod.rs:46-48) + MyrSeq::pre_process
        (data/myrseq.rs:113-131, defaults of myrio-cli config.rs:24-26) on the device.
        Returns (MyrSeqs of the kept reads, keep flags per record)."""
        seq_ascii = np.ascontiguousarray(seq_ascii, np.uint8)
        qual_ascii = np.ascontiguousarray(qual_ascii, np.uint8)
        base_off = np.ascontiguousarray(base_off, np.uint64)
        n = len(base_off) - 1
        keep = np.zeros(max(1, n), np.uint8)
        self = cls.__new__(cls)
        self.ctx = ctx
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_reads_from_fastq(ctx.h, _p(seq_ascii), _p(qual_ascii), _p(base_off), n, min_length,
                                                 min_mean_qual, max_qual, _p(keep), C.byref(h)))
        self.h = h
        self.n = int(keep[:n].sum())
        return self, keep[:n].astype(bool)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx._L.myrio_seqs_free(self.h)
            self.h = None
        except Exception:
            pass

    def compute_kmer_counts(self, k: int, cutoff: float) -> SFVecs:
        """MyrSeq::compute_kmer_counts for every read (data/myrseq.rs:76-111);
        the returned batch carries nb_hck as its aux column."""
        h = C.c_void_p()
        _lib.check(self.ctx._L.myrio_count_reads(self.ctx.h, self.h, k, cutoff, C.byref(h)))
        out = SFVecs(self.ctx, h, True)
        out.aux_is_u64 = True
        return out


RANKS = {"Domain": 8, "Kingdom": 7, "Phylum": 6, "Class": 5, "Order": 4, "Family": 3, "Genus": 2, "Species": 1}


class MyrTreeFile:
    """Host-side content of a `.myrtree` file (myrio_store): TaxTreeStore::to_bytes / from_bytes / save_to_file /
    load_from_file (tax/store.rs:53-93, 342-454).  The taxonomy is a flat PRE-ORDER node list: kind (0 branch,
    1 leaf), arg (branch: number of children, leaf: payload_id), names.  No GPU is needed for the byte format;
    `leaves` / `counts` / `append_counts` move the payloads between the file content and device batches."""

    def __init__(self, handle):
        self._L = _lib.load()
        self.h = handle

    def __del__(self):
        try:
            if self.h:
                self._L.myrio_store_free(self.h)
            self.h = None
        except Exception:
            pass

    @classmethod
    def new(cls, gene: str, root_rank: int, node_kind, node_arg, node_names, iupac, base_off) -> "MyrTreeFile":
        """TaxTreeCore::new over payloads without cached counts (store.rs:238-246).  iupac: one 4-bit code per byte."""
        L = _lib.load()
        kind = np.ascontiguousarray(node_kind, np.uint8)
        arg = np.ascontiguousarray(node_arg, np.uint64)
        enc = [n.encode() for n in node_names]
        name_off = np.zeros(len(enc) + 1, np.uint64)
        name_off[1:] = np.cumsum([len(e) for e in enc])
        names = b"".join(enc)
        base_off = np.ascontiguousarray(base_off, np.uint64)
        words, word_off = pack_iupac4(np.ascontiguousarray(iupac, np.uint8), base_off)
        h = C.c_void_p()
        _lib.check(L.myrio_store_new(gene.encode(), root_rank, len(kind), _p(kind), _p(arg), _p(name_off),
                                     C.c_char_p(names), len(base_off) - 1, _p(words), _p(word_off), _p(base_off),
                                     C.byref(h)))
        return cls(h)

    @classmethod
    def from_bytes(cls, data: bytes, threads: int = 0) -> "MyrTreeFile":
        L = _lib.load()
        buf = np.frombuffer(data, np.uint8)
        h = C.c_void_p()
        _lib.check(L.myrio_store_from_bytes(_p(buf) if len(buf) else None, len(buf), threads, C.byref(h)))
        return cls(h)

    @classmethod
    def load_from_file(cls, path: str, threads: int = 0) -> "MyrTreeFile":
        L = _lib.load()
        h = C.c_void_p()
        _lib.check(L.myrio_store_load(str(path).encode(), threads, C.byref(h)))
        return cls(h)

    def to_bytes(self, zstd_compression_level: int = 6, threads: int = 0) -> bytes:
        out, n = C.c_void_p(), C.c_uint64(0)
        _lib.check(self._L.myrio_store_to_bytes(self.h, zstd_compression_level, threads, C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            self._L.myrio_bytes_free(out)

    def save_to_file(self, path: str, zstd_compression_level: int = 6, threads: int = 0):
        _lib.check(self._L.myrio_store_save(self.h, str(path).encode(), zstd_compression_level, threads))

    def info(self):
        a, b, c, e, f, g = (C.c_uint64(0) for _ in range(6))
        r = C.c_uint32(0)
        _lib.check(self._L.myrio_store_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(r), C.byref(e),
                                            C.byref(f), C.byref(g)))
        return {"n_nodes": int(a.value), "n_payloads": int(b.value), "n_k": int(c.value), "root_rank": int(r.value),
                "names_bytes": int(e.value), "gene_bytes": int(f.value), "seq_words": int(g.value)}

    def tree(self):
        i = self.info()
        gene = C.create_string_buffer(i["gene_bytes"] + 1)
        kind = np.zeros(i["n_nodes"], np.uint8)
        arg = np.zeros(i["n_nodes"], np.uint64)
        off = np.zeros(i["n_nodes"] + 1, np.uint64)
        names = np.zeros(max(1, i["names_bytes"]), np.uint8)
        kp = np.zeros(max(1, i["n_k"]), np.uint64)
        _lib.check(self._L.myrio_store_tree(self.h, gene, _p(kind), _p(arg), _p(off), _p(names), _p(kp)))
        raw = names.tobytes()
        return {"gene": gene.value.decode(), "root_rank": i["root_rank"], "kind": kind, "arg": arg,
                "names": [raw[int(off[j]):int(off[j + 1])].decode() for j in range(i["n_nodes"])],
                "k_precomputed": [int(v) for v in kp[:i["n_k"]]]}

    @property
    def k_precomputed(self):
        return self.tree()["k_precomputed"]

    def seqs(self):
        """(words, word_off, base_off): the packed Seq<Iupac> of every payload."""
        i = self.info()
        words = np.zeros(max(1, i["seq_words"]), np.uint64)
        woff = np.zeros(i["n_payloads"] + 1, np.uint64)
        boff = np.zeros(i["n_payloads"] + 1, np.uint64)
        _lib.check(self._L.myrio_store_seqs(self.h, _p(words), _p(woff), _p(boff)))
        return words[:i["seq_words"]], woff, boff

    def counts_host(self, k_index: int):
        """kmer_store_counts_vec[k_index] of every payload: (row_off, keys, u16 counts, f32 rescale)."""
        n = C.c_uint64(0)
        _lib.check(self._L.myrio_store_counts_info(self.h, k_index, C.byref(n)))
        L = self.info()["n_payloads"]
        ro = np.zeros(L + 1, np.uint64)
        keys = np.zeros(max(1, n.value), np.uint64)
        vals = np.zeros(max(1, n.value), np.uint16)
        resc = np.zeros(max(1, L), np.float32)
        _lib.check(self._L.myrio_store_counts_host(self.h, k_index, _p(ro), _p(keys), _p(vals), _p(resc)))
        return ro, keys[:n.value], vals[:n.value], resc[:L]

    def counts(self, ctx: Context, k_index: int) -> "SFVecs":
        h = C.c_void_p()
        _lib.check(self._L.myrio_store_counts(ctx.h, self.h, k_index, C.byref(h)))
        out = SFVecs(ctx, h, True)
        out.aux_is_u64 = False
        return out

    def append_counts(self, ctx: Context, k: int, counts: "SFVecs"):
        _lib.check(self._L.myrio_store_append_counts(ctx.h, self.h, k, counts.h))

    def append_counts_host(self, k: int, row_off, keys, counts, rescale):
        ro = np.ascontiguousarray(row_off, np.uint64)
        kk = np.ascontiguousarray(keys, np.uint64)
        vv = np.ascontiguousarray(counts, np.uint16)
        rs = np.ascontiguousarray(rescale, np.float32)
        _lib.check(self._L.myrio_store_append_counts_host(self.h, k, _p(ro), _p(kk), _p(vv), _p(rs)))

    def overwrite_counts(self, ctx: Context, k: int, counts: "SFVecs"):
        _lib.check(self._L.myrio_store_overwrite_counts(ctx.h, self.h, k, counts.h))

    def shrink(self):
        """TaxTreeStore::shrink (store.rs:259-262)"""
        _lib.check(self._L.myrio_store_shrink(self.h))


class TaxTreeStore:
    """The leaf sequences + cached k-mer counts of a .myrtree (tax/store.rs:28-50) on the GPU.  When the store
    comes from (or is attached to) a MyrTreeFile, counts computed on the GPU are written back into it, so
    `save_to_file` produces the reference's cache file (compute.rs:86-91: --cache-counts)."""

    def __init__(self, ctx: Context, leaves: SeqSet | None = None, *, iupac=None, base_off=None,
                 gene: str = "", leaf_id_base: int = 0):
        self.ctx, self.gene, self.leaf_id_base = ctx, gene, leaf_id_base
        if leaves is not None:
            iupac, base_off = tree_to_iupac(leaves), leaves.base_off
        base_off = np.ascontiguousarray(base_off, np.uint64)
        words, word_off = pack_iupac4(np.ascontiguousarray(iupac, np.uint8), base_off)
        self.n_leaves = len(base_off) - 1
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_leaves_upload(ctx.h, _p(words), _p(word_off), _p(base_off), self.n_leaves,
                                              C.byref(h)))
        self.h = h
        self.k_precomputed: list[int] = []
        self.kmer_store_counts_vec: list[SFVecs] = []

    file: "MyrTreeFile | None" = None

    @classmethod
    def from_file(cls, ctx: Context, f: "MyrTreeFile", *, leaf_id_base: int = 0) -> "TaxTreeStore":
        """TaxTreeStore::load_from_file (store.rs:77-93), GPU half: the payload sequences and every cached count
        vector become device batches."""
        self = cls.__new__(cls)
        t = f.tree()
        self.ctx, self.gene, self.leaf_id_base, self.file = ctx, t["gene"], leaf_id_base, f
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_store_leaves(ctx.h, f.h, C.byref(h)))
        self.h = h
        self.n_leaves = f.info()["n_payloads"]
        self.k_precomputed = list(t["k_precomputed"])
        self.kmer_store_counts_vec = [f.counts(ctx, i) for i in range(len(self.k_precomputed))]
        return self

    @classmethod
    def load_from_file(cls, ctx: Context, path: str, **kw) -> "TaxTreeStore":
        return cls.from_file(ctx, MyrTreeFile.load_from_file(path), **kw)

    def save_to_file(self, path: str, zstd_compression_level: int = 6):
        """TaxTreeStore::save_to_file (store.rs:53-75)"""
        if self.file is None:
            raise MyrioError(_lib.ERR_BAD_ARG, "this store has no .myrtree content attached (MyrTreeFile)")
        self.file.save_to_file(path, zstd_compression_level)

    @classmethod
    def from_fasta_records(cls, ctx: Context, seq_ascii, base_off, *, gene: str = "", leaf_id_base: int = 0):
        """The sequence half of TaxTreeStore::load_from_fasta_string (tax/store.rs:166-195) on the device:
        `Seq::<Iupac>::from_str` of every record's concatenated sequence lines; records with an invalid
        byte are skipped (store.rs:181-187).  Returns (store of the kept records, keep flags)."""
        seq_ascii = np.ascontiguousarray(seq_ascii, np.uint8)
        base_off = np.ascontiguousarray(base_off, np.uint64)
        n = len(base_off) - 1
        keep = np.zeros(max(1, n), np.uint8)
        self = cls.__new__(cls)
        self.ctx, self.gene, self.leaf_id_base = ctx, gene, leaf_id_base
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_leaves_from_fasta(ctx.h, _p(seq_ascii), _p(base_off), n, _p(keep), C.byref(h)))
        self.h = h
        self.n_leaves = int(keep[:n].sum())
        self.k_precomputed, self.kmer_store_counts_vec = [], []
        return self, keep[:n].astype(bool)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx._L.myrio_seqs_free(self.h)
            self.h = None
        except Exception:
            pass

    def _count(self, k: int, nb_resamples: int, seed: int) -> SFVecs:
        h = C.c_void_p()
        _lib.check(self.ctx._L.myrio_count_leaves(self.ctx.h, self.h, k, nb_resamples, seed, self.leaf_id_base,
                                                  C.byref(h)))
        out = SFVecs(self.ctx, h, True)
        out.aux_is_u64 = False
        return out

    def compute_and_append_kmer_counts(self, k: int, nb_resamples: int, seed: int = 0):
        """tax/store.rs:265-285"""
        self.kmer_store_counts_vec.append(self._count(k, nb_resamples, seed))
        self.k_precomputed.append(k)
        if self.file is not None:
            self.file.append_counts(self.ctx, k, self.kmer_store_counts_vec[-1])

    def compute_and_overwrite_kmer_counts(self, k: int, nb_resamples: int, seed: int = 0):
        """tax/store.rs:287-308"""
        self.kmer_store_counts_vec = [self._count(k, nb_resamples, seed)]
        self.k_precomputed = [k]
        if self.file is not None:
            self.file.overwrite_counts(self.ctx, k, self.kmer_store_counts_vec[0])


class TaxTreeCompute:
    """tax/compute.rs:13-20 -- the normalised leaf vectors, held as a tiled inverted index."""

    def __init__(self, ctx, index_handle, n_leaves, leaf_id_base, leaf_vecs: SFVecs | None,
                 counts: SFVecs | None = None):
        self.ctx, self.h, self.n_leaves, self.leaf_id_base = ctx, index_handle, n_leaves, leaf_id_base
        self.leaf_vecs = leaf_vecs
        self._counts = counts  # the store's cached counts: leaf vectors are re-derived on demand

    def leaf_vectors(self) -> SFVecs:
        """The reference's `Box<[SFVec]>` of L2-normalised leaf vectors (tax/compute.rs:100-108),
        materialised only when a merge-join similarity (Overlap) needs them."""
        if self.leaf_vecs is None:
            if self._counts is None:
                raise MyrioError(_lib.ERR_BAD_ARG, "this TaxTreeCompute was built without its leaf counts")
            self.leaf_vecs = self._counts.into_normalized_l2()
        return self.leaf_vecs

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx._L.myrio_index_free(self.h)
            self.h = None
        except Exception:
            pass

    @classmethod
    def from_store_tree(cls, ttstore: TaxTreeStore, search_k: int, fasta_nb_resamples: int = 32, *,
                        cache_counts: bool = True, seed: int = 0, tile_leaves: int = 0,
                        keep_leaf_vecs: bool = False) -> "TaxTreeCompute":
        """tax/compute.rs:80-118: pick (or compute) the cached counts for search_k, convert
        each leaf to an L2-normalised f32 vector, and build the index the score loop runs on."""
        if search_k in ttstore.k_precomputed:
            idx = ttstore.k_precomputed.index(search_k)
        elif cache_counts:
            ttstore.compute_and_append_kmer_counts(search_k, fasta_nb_resamples, seed)
            idx = len(ttstore.k_precomputed) - 1
        else:
            ttstore.compute_and_overwrite_kmer_counts(search_k, fasta_nb_resamples, seed)
            idx = 0
        counts = ttstore.kmer_store_counts_vec[idx]
        return cls.from_counts(counts, ttstore.leaf_id_base, tile_leaves=tile_leaves,
                               keep_leaf_vecs=keep_leaf_vecs)

    @classmethod
    def from_counts(cls, counts: SFVecs, leaf_id_base: int = 0, *, tile_leaves: int = 0,
                    keep_leaf_vecs: bool = False) -> "TaxTreeCompute":
        ctx = counts.ctx
        normalized = counts.into_normalized_l2()
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_index_build(ctx.h, normalized.h, leaf_id_base, tile_leaves, C.byref(h)))
        return cls(ctx, h, counts.n_rows, leaf_id_base, normalized if keep_leaf_vecs else None, counts)

    def info(self):
        a, b, c, d, e = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        _lib.check(self.ctx._L.myrio_index_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d),
                                                C.byref(e)))
        return {"n_leaves": int(a.value), "n_tiles": int(b.value), "n_postings": int(c.value),
                "n_unique_total": int(d.value), "tile_leaves": int(e.value)}

    def export_tile(self, tile: int):
        nu, npst = C.c_uint64(0), C.c_uint64(0)
        L = self.ctx._L
        _lib.check(L.myrio_index_export_tile(self.ctx.h, self.h, tile, C.byref(nu), C.byref(npst), None, None,
                                             None, None))
        uk = np.zeros(nu.value, np.uint64)
        of = np.zeros(nu.value + 1, np.uint64)
        pl = np.zeros(npst.value, np.uint32)
        pv = np.zeros(npst.value, np.float32)
        _lib.check(L.myrio_index_export_tile(self.ctx.h, self.h, tile, C.byref(nu), C.byref(npst), _p(uk),
                                             _p(of), _p(pl), _p(pv)))
        return uk, of, pl, pv


@dataclass
class TaxTreeResults:
    """The scoring part of tax/results.rs: full score rows and/or the cut() top-x."""
    scores: np.ndarray | None      # [Q, n_leaves] in payload order (results.rs:82-93)
    top_scores: np.ndarray | None  # [Q, x]  (results.rs:310-329; score desc, payload_id asc)
    top_ids: np.ndarray | None     # [Q, x]

    @classmethod
    def from_compute_tree(cls, queries: SFVecs, ttcompute: TaxTreeCompute, similarity: int = SIM_COSINE,
                          nb_best_analysis: int = 100, *, q_first: int = 0, q_count: int | None = None,
                          full_scores: bool = False, direct: bool = False) -> "TaxTreeResults":
        ctx = ttcompute.ctx
        q_count = queries.n_rows - q_first if q_count is None else q_count
        scores = ts = ti = None
        # Overlap sums maxima over the UNION of the keys (similarity.rs:172-182): merge-join over the
        # leaf vectors, like the reference; Cosine / JacardTanimoto run on the inverted index
        direct = direct or similarity == SIM_OVERLAP
        L = ctx._L
        if full_scores:
            scores = np.zeros((q_count, ttcompute.n_leaves), np.float32)
            if direct:
                _lib.check(L.myrio_score_all_direct(ctx.h, ttcompute.leaf_vectors().h, queries.h, q_first, q_count,
                                                    similarity, _p(scores)))
            else:
                _lib.check(L.myrio_score_all(ctx.h, ttcompute.h, queries.h, q_first, q_count, similarity,
                                             _p(scores)))
        if nb_best_analysis:
            ts = np.zeros((q_count, nb_best_analysis), np.float32)
            ti = np.zeros((q_count, nb_best_analysis), np.uint32)
            if direct:
                _lib.check(L.myrio_score_topx_direct(ctx.h, ttcompute.leaf_vectors().h, ttcompute.leaf_id_base,
                                                     queries.h, q_first, q_count, nb_best_analysis, similarity,
                                                     _p(ts), _p(ti)))
            else:
                _lib.check(L.myrio_score_topx(ctx.h, ttcompute.h, queries.h, q_first, q_count,
                                              nb_best_analysis, similarity, _p(ts), _p(ti)))
        return cls(scores, ts, ti)


class Taxonomy:
    """The taxonomy of one tree prepared for the results pass (myrio_taxonomy): flat pre-order node list."""

    def __init__(self, ctx: Context, root_rank: int, node_kind, node_arg, n_payloads: int, names=None):
        self.ctx = ctx
        self.kind = np.ascontiguousarray(node_kind, np.uint8)
        self.arg = np.ascontiguousarray(node_arg, np.uint64)
        self.names, self.root_rank, self.n_payloads = names, root_rank, n_payloads
        h = C.c_void_p()
        _lib.check(ctx._L.myrio_taxonomy_create(ctx.h, root_rank, len(self.kind), _p(self.kind), _p(self.arg),
                                                n_payloads, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx._L.myrio_taxonomy_free(self.h)
            self.h = None
        except Exception:
            pass


@dataclass
class SearchParameters:
    """myrio-cli/src/app/config.rs:158-175 ([search] of config_default.toml)"""
    max_amount_of_leaves_per_branch: int = 15
    nb_best_analysis: int = 100
    lambda_leaf: float = 2.0
    lambda_branch: float = 20.0
    mu: float = 1.1
    gamma: float = 2.5
    delta: float = 0.9
    epsilon: float = 0.55
    nb_best_display: int = 7


class BResTree:
    """TaxTreeResults (tax/results.rs:59-61): TaxTreeCore<BRes, SimScore> of one query as flat pre-order arrays.
    kind / arg / src_node per node; nb_leaves, mean, max, pool_score, conf_state (0 None, 1 Some(None),
    2 Some(Some(conf))), conf, force_keep per node (branches); payload_score / payload_src_id per kept leaf."""

    def __init__(self, L, handle):
        self._L, self.h = L, handle
        n, npay, rr = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        _lib.check(L.myrio_results_info(handle, C.byref(n), C.byref(npay), C.byref(rr)))
        n, npay = int(n.value), int(npay.value)
        self.root_rank = int(rr.value)
        self.kind = np.zeros(n, np.uint8)
        self.arg = np.zeros(n, np.uint64)
        self.src_node = np.zeros(n, np.uint64)
        self.nb_leaves = np.zeros(n, np.uint64)
        self.mean, self.max, self.pool_score, self.conf = (np.zeros(n, np.float32) for _ in range(4))
        self.conf_state = np.zeros(n, np.uint8)
        self.force_keep = np.zeros(n, np.uint8)
        self.payload_score = np.zeros(npay, np.float32)
        self.payload_src_id = np.zeros(npay, np.uint64)
        _lib.check(L.myrio_results_tree(handle, _p(self.kind), _p(self.arg), _p(self.src_node), _p(self.nb_leaves),
                                        _p(self.mean), _p(self.max), _p(self.pool_score), _p(self.conf_state),
                                        _p(self.conf), _p(self.force_keep),
                                        _p(self.payload_score) if npay else None,
                                        _p(self.payload_src_id) if npay else None))

    def __del__(self):
        try:
            if self.h:
                self._L.myrio_results_free(self.h)
            self.h = None
        except Exception:
            pass

    def cut(self, nb_best: int) -> "BResTree":
        """TaxTreeResults::cut (results.rs:310-381)"""
        h = C.c_void_p()
        _lib.check(self._L.myrio_results_cut(self.h, nb_best, C.byref(h)))
        return BResTree(self._L, h)

    def depths(self) -> np.ndarray:
        d = np.zeros(len(self.kind), np.int64)
        stack = []
        for i in range(len(self.kind)):
            d[i] = len(stack)
            if stack:
                stack[-1] -= 1
            if self.kind[i] == 0:
                stack.append(int(self.arg[i]))
            while stack and stack[-1] == 0:
                stack.pop()
        return d

    def branches_at_rank(self, rank: int) -> np.ndarray:
        """gather_branches_at_rank (tax/core.rs:180-208): node indices of the branches at `rank`."""
        return np.nonzero((self.kind == 0) & (self.depths() == self.root_rank - rank))[0]


def results_from_compute_tree(queries: SFVecs, ttcompute: TaxTreeCompute, taxonomy: Taxonomy,
                              similarity: int = SIM_COSINE, params: SearchParameters | None = None, *,
                              q_first: int = 0, q_count: int | None = None) -> list:
    """TaxTreeResults::from_compute_tree (tax/results.rs:68-252), the whole of it, for a batch of queries: scores of
    every leaf, branch nb_leaves / mean / max, trim, cut, pooling, confidence.  Returns one BResTree per query."""
    ctx = ttcompute.ctx
    params = params or SearchParameters()
    q_count = queries.n_rows - q_first if q_count is None else q_count
    p = _lib.ResultsParams(params.max_amount_of_leaves_per_branch, params.nb_best_analysis, params.lambda_leaf,
                           params.lambda_branch, params.mu, params.gamma, params.delta, params.epsilon)
    out = (C.c_void_p * max(1, q_count))()
    leaf_vecs = ttcompute.leaf_vectors().h if similarity == SIM_OVERLAP else None
    _lib.check(ctx._L.myrio_results_from_compute_tree(ctx.h, ttcompute.h, leaf_vecs, taxonomy.h, queries.h, q_first,
                                                      q_count, similarity, C.byref(p), out))
    return [BResTree(ctx._L, C.c_void_p(out[i])) for i in range(q_count)]


def in_database_confidence(per_gene: list) -> float:
    """myrio-cli/src/app/process/run.rs:405-466.  per_gene: one (BResTree, names) pair per gene, `names` the node
    names of that gene's taxonomy (indexed by src_node)."""
    ids, off, name_id, pool, state, conf = {}, [0], [], [], [], []
    for tree, names in per_gene:
        for i in tree.branches_at_rank(RANKS["Genus"]):
            name_id.append(ids.setdefault(names[int(tree.src_node[i])], len(ids)))
            pool.append(tree.pool_score[i])
            state.append(tree.conf_state[i])
            conf.append(tree.conf[i])
        off.append(len(name_id))
    a = [np.ascontiguousarray(off, np.uint64), np.ascontiguousarray(name_id, np.uint64),
         np.ascontiguousarray(pool, np.float32), np.ascontiguousarray(state, np.uint8),
         np.ascontiguousarray(conf, np.float32)]
    out = C.c_float(0)
    _lib.check(_lib.load().myrio_in_database_confidence(len(per_gene), *[_p(x) for x in a], C.byref(out)))
    return float(np.float32(out.value))


@dataclass
class ClusteringParameters:
    """myrio-core/src/clustering.rs:15-36 (defaults: myrio-cli/src/app/config.rs:107-119)"""
    k: int = 6
    t1: float = 0.2
    t2: int = 20
    similarity: int = SIM_COSINE
    intial_centroids: SFVecs | None = None  # (sic) spelled as in the reference
    expected_nb_of_clusters: int = 1
    eta_improvement: float = 1e-4
    nb_iters_max: int = 20
    silhouette_trimming: float | None = None


@dataclass
class ClusteringOutput:
    """clustering.rs:38-41 as index lists: clusters[c] = read indices in input order;
    rejected = reads dropped by t2 (first) then by silhouette trimming."""
    clusters: list
    rejected: np.ndarray
    assign: np.ndarray
    n_iters: int
    silhouette: np.ndarray = field(default=None)
    exchange: dict = field(default=None)  # row-sharded runs: all-gather statistics


def cluster(myrseqs: MyrSeqs, params: ClusteringParameters, shard=None) -> ClusteringOutput:
    """clustering::cluster (myrio-core/src/clustering.rs:55-270).  `shard` = (rank, world, allgather)
    runs the row-sharded variant (myrio_cluster_sharded; see myrio_b200.dist.cluster_sharded): every
    rank passes the same reads and gets the same result."""
    ctx = myrseqs.ctx
    p = _lib.ClusterParams()
    p.k, p.t1, p.t2, p.similarity = params.k, params.t1, params.t2, params.similarity
    p.expected_nb_of_clusters = params.expected_nb_of_clusters
    p.eta_improvement, p.nb_iters_max = params.eta_improvement, params.nb_iters_max
    p.has_silhouette_trimming = int(params.silhouette_trimming is not None)
    p.silhouette_trimming = params.silhouette_trimming or 0.0
    assign = np.zeros(max(1, myrseqs.n), np.int32)
    sil = np.zeros(max(1, myrseqs.n), np.float32)
    iters = C.c_uint32(0)
    init = params.intial_centroids.h if params.intial_centroids is not None else None
    if shard is None:
        _lib.check(ctx._L.myrio_cluster(ctx.h, myrseqs.h, C.byref(p), init, _p(assign), C.byref(iters), _p(sil)))
    else:
        rank, world, allgather = shard
        if allgather is None:  # the library's own communicator (Context.comm_init): device-side exchange
            cb = _lib.ALLGATHER_FN()
        else:
            cb = allgather if isinstance(allgather, _lib.ALLGATHER_FN) else _lib.ALLGATHER_FN(allgather)
        sh = _lib.Shard(rank, world, cb, None)
        _lib.check(ctx._L.myrio_cluster_sharded(ctx.h, myrseqs.h, C.byref(p), init, C.byref(sh), _p(assign),
                                                C.byref(iters), _p(sil)))
    assign, sil = assign[:myrseqs.n], sil[:myrseqs.n]
    nc = int(assign.max()) + 1 if (assign >= 0).any() else 0
    nc = max(nc, params.expected_nb_of_clusters,
             params.intial_centroids.n_rows if params.intial_centroids is not None else 0)
    clusters = [np.nonzero(assign == c)[0] for c in range(nc)]
    rejected = np.concatenate([np.nonzero(assign == -1)[0], np.nonzero(assign == -2)[0]])
    return ClusteringOutput(clusters, rejected, assign, int(iters.value), sil)


def compute_cluster_centroid(counts: SFVecs, member_ids) -> SFVecs:
    """clustering::compute_cluster_centroid (clustering.rs:44-53) over rows of a raw count batch.
    Also `rows.sum::<SFVec>().into_normalized_l2()` for any f32 batch, and -- for a u16 batch of
    cached leaf counts -- the sum of kmer_store_counts_to_kmer_counts(row) (tax/mod.rs:55-58)."""
    ids = np.ascontiguousarray(member_ids, np.uint64)
    h = C.c_void_p()
    _lib.check(counts.ctx._L.myrio_compute_cluster_centroid(counts.ctx.h, counts.h, _p(ids), len(ids),
                                                            C.byref(h)))
    return SFVecs(counts.ctx, h)


def stack_rows(ctx: Context, vecs: list) -> SFVecs:
    """One batch from several 1-row batches (e.g. the local fingerprints of a tree)."""
    rows = [v.download(aux=False) for v in vecs]
    ro = np.zeros(len(rows) + 1, np.uint64)
    ro[1:] = np.cumsum([len(r[1]) for r in rows])
    keys = np.concatenate([r[1] for r in rows]) if rows else np.zeros(0, np.uint64)
    vals = np.concatenate([r[2] for r in rows]) if rows else np.zeros(0, np.float32)
    return SFVecs.upload(ctx, ro, keys, vals)


def reweight_fingerprints(local_fingerprints: SFVecs, naive_query: SFVecs, similarity: int = SIM_COSINE,
                          alpha: float = 2.5) -> SFVecs:
    """myrio-cli/src/app/process/run.rs:243-262: the initial centroid of the second clustering pass,
    normalize_l2(sum_i local_i * exp(1 + alpha * simfunc(local_i, naive_query)))."""
    h = C.c_void_p()
    ctx = local_fingerprints.ctx
    _lib.check(ctx._L.myrio_reweight_fingerprints(ctx.h, local_fingerprints.h, naive_query.h, similarity,
                                                  C.c_float(alpha), C.byref(h)))
    return SFVecs(ctx, h)


def pairwise(rows: SFVecs, cols: SFVecs, similarity: int = SIM_COSINE) -> np.ndarray:
    """The reads x reads / reads x centroids similarity tile of the clustering loop."""
    out = np.zeros((rows.n_rows, cols.n_rows), np.float32)
    if out.size:
        _lib.check(rows.ctx._L.myrio_pairwise(rows.ctx.h, rows.h, cols.h, similarity, _p(out)))
    return out


def fingerprints(ttstore: TaxTreeStore, cluster_k: int, branch_leaf_samples, fasta_nb_resamples: int = 32,
                 seed: int = 0):
    """The fingerprints of TaxTreeCompute::from_store_tree (tax/compute.rs:51-77) at `cluster_k`:
    one local fingerprint per entry of `branch_leaf_samples` (the leaves drawn from one branch at
    the fingerprint rank -- the reference draws them with an OS-seeded RNG, `choose_multiple`, so
    the caller supplies the sample) and the global fingerprint (the normalised sum of the local
    ones).  Returns (local: list[SFVecs], global: SFVecs)."""
    counts = ttstore._count(cluster_k, fasta_nb_resamples, seed)
    local = [compute_cluster_centroid(counts, ids) for ids in branch_leaf_samples]
    rows = [fp.download(aux=False) for fp in local]
    ro = np.zeros(len(rows) + 1, np.uint64)
    ro[1:] = np.cumsum([len(r[1]) for r in rows])
    keys = np.concatenate([r[1] for r in rows]) if rows else np.zeros(0, np.uint64)
    vals = np.concatenate([r[2] for r in rows]) if rows else np.zeros(0, np.float32)
    batch = SFVecs.upload(ttstore.ctx, ro, keys, vals)
    return local, compute_cluster_centroid(batch, np.arange(len(rows)))
