// myrtree.cu -- the `.myrtree` file: reader, writer and the cached k-mer counts' way to and from the GPU.
//
// Restates TaxTreeStore::to_bytes / from_bytes (myrio-core/src/tax/store.rs:342-454) with the manual bincode
// impls of the types inside: SparseVec (data/sparse.rs:413-436), Leaf / Branch / Node / TaxTreeCore
// (tax/core.rs:22-26, 332-410), StorePayload (tax/store.rs:28-32), ChunkInfo / TTSHeader (store.rs:314-328).
//
//   file   = MAGIC "MYRIO-\xCE\xA8" . header_len u64 LE . header . zstd(core without payloads) . zstd(chunk)*
//   header = ChunkInfo core_info {c_size, uc_size} . Vec<ChunkInfo> chunk_info_vec . Vec<usize> k_precomputed
//   core   = String gene . Node root . Rank root_rank . Box<[StorePayload]> (empty: the payloads go to chunks)
//   chunk  = Vec<StorePayload>, payload = Seq<Iupac> seq . Vec<(SparseVec<u16> {keys, values}, f32 rescale)>
//   bincode 2 `standard().with_fixed_int_encoding()`: little endian, usize -> u64, Vec/Box<[T]>/str -> u64
//   length + items, u16 / f32 / u8 as is, () -> nothing, Node -> u8 tag (0 branch, 1 leaf).
//
// ASSUMED LAYOUTS (not verifiable here: the code that decides them is not in the reference tree) -- each is
// one function below and listed in DESIGN.md:
//   * Seq<Iupac> (bincode impl of the un-vendored bio-seq fork, Cargo.toml:23): u64 symbol count, then the raw
//     BitVec<usize, Lsb0> words as Vec<usize> (u64 word count + words), the pair Seq::from_raw(len, &[usize])
//     takes (myrio-exp/src/simseq/mod.rs:106)                      -> encode_seq / decode_seq
//   * Rank (bincode-derive 2.0.1 on a #[repr(usize)] enum with explicit discriminants, tax/clade.rs:93-104): the
//     variant is written as u32 = its discriminant value (Domain = 8 ... Species = 1) -> encode_rank / decode_rank
//
// zstd is bound at run time (dlopen libzstd.so.1; the image ships the runtime library without headers).
// Chunking follows store.rs:363-367: chunk_size = len < n ? len : len / n + 1 payloads, n = worker threads.
#include <dlfcn.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

#include "myrtree.cuh"

namespace myrio {

// ---------------------------------------------------------------- zstd (run-time binding)

namespace {
struct ZstdApi {
    void *handle = nullptr;
    size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
    size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
    size_t (*compressBound)(size_t) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    const char *(*getErrorName)(size_t) = nullptr;
};

ZstdApi &zstd() {
    static ZstdApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("MYRIO_ZSTD_LIB"), "libzstd.so.1", "libzstd.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.compress = reinterpret_cast<decltype(api.compress)>(dlsym(api.handle, "ZSTD_compress"));
        api.decompress = reinterpret_cast<decltype(api.decompress)>(dlsym(api.handle, "ZSTD_decompress"));
        api.compressBound = reinterpret_cast<decltype(api.compressBound)>(dlsym(api.handle, "ZSTD_compressBound"));
        api.isError = reinterpret_cast<decltype(api.isError)>(dlsym(api.handle, "ZSTD_isError"));
        api.getErrorName = reinterpret_cast<decltype(api.getErrorName)>(dlsym(api.handle, "ZSTD_getErrorName"));
    });
    if (!api.handle || !api.compress || !api.decompress || !api.compressBound || !api.isError)
        fail(MYRIO_ERR_UNSUPPORTED, "zstd is not available (dlopen libzstd.so.1)");
    return api;
}

std::vector<uint8_t> zstd_compress(const std::vector<uint8_t> &raw, int level) {  // zstd::bulk::compress
    ZstdApi &z = zstd();
    std::vector<uint8_t> out(z.compressBound(raw.size()));
    const size_t n = z.compress(out.data(), out.size(), raw.data(), raw.size(), level);
    if (z.isError(n)) fail(MYRIO_ERR_FORMAT, "zstd compress: %s", z.getErrorName ? z.getErrorName(n) : "error");
    out.resize(n);
    return out;
}

std::vector<uint8_t> zstd_decompress(const uint8_t *src, size_t c_size, size_t uc_size) {  // zstd::bulk::decompress
    ZstdApi &z = zstd();
    std::vector<uint8_t> out(uc_size);
    const size_t n = z.decompress(out.data(), out.size(), src, c_size);
    if (z.isError(n)) fail(MYRIO_ERR_FORMAT, "zstd decompress: %s", z.getErrorName ? z.getErrorName(n) : "error");
    out.resize(n);
    return out;
}

// ---------------------------------------------------------------- bincode (fixint, little endian)

struct Writer {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void raw(const void *p, size_t n) {
        const uint8_t *q = static_cast<const uint8_t *>(p);
        b.insert(b.end(), q, q + n);
    }
    void u16(uint16_t v) { raw(&v, 2); }  // x86-64 / aarch64: little endian hosts only
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    void f32(float v) { raw(&v, 4); }
    void str(const std::string &s) {
        u64(s.size());
        raw(s.data(), s.size());
    }
};

struct Reader {  // utils.rs:93-132 BCursor + bincode's "unexpected end" error
    const uint8_t *p;
    size_t n, pos = 0;
    Reader(const uint8_t *data, size_t len) : p(data), n(len) {}
    const uint8_t *take(size_t by) {
        if (by > n - pos) fail(MYRIO_ERR_FORMAT, "truncated .myrtree data: %zu bytes wanted at offset %zu of %zu", by, pos, n);
        const uint8_t *q = p + pos;
        pos += by;
        return q;
    }
    uint8_t u8() { return *take(1); }
    uint16_t u16() { uint16_t v; memcpy(&v, take(2), 2); return v; }
    uint32_t u32() { uint32_t v; memcpy(&v, take(4), 4); return v; }
    uint64_t u64() { uint64_t v; memcpy(&v, take(8), 8); return v; }
    float f32() { float v; memcpy(&v, take(4), 4); return v; }
    // a length prefix of `elem`-byte items: must fit in what is left (guards the allocations)
    uint64_t len(size_t elem) {
        const uint64_t l = u64();
        if (elem && l > (n - pos) / elem) fail(MYRIO_ERR_FORMAT, "corrupt length prefix %llu at offset %zu", (unsigned long long)l, pos - 8);
        return l;
    }
    std::string str() {
        const uint64_t l = len(1);
        const uint8_t *q = take(l);
        return std::string(reinterpret_cast<const char *>(q), l);
    }
};

// ---- assumed layouts, one function each ----------------------------------------------------------------

// Rank as bincode-derive writes a fieldless enum with explicit discriminants: u32 = the discriminant value
void encode_rank(Writer &w, uint32_t rank) { w.u32(rank); }
uint32_t decode_rank(Reader &r) {
    const uint32_t v = r.u32();
    if (v < 1 || v > 8) fail(MYRIO_ERR_FORMAT, "invalid Rank discriminant %u", v);
    return v;
}

// Seq<Iupac>: u64 symbol count + Vec<usize> of the raw Lsb0 words (symbol i at bits 4*(i%16) of word i/16)
void encode_seq(Writer &w, const uint64_t *words, uint64_t n_sym) {
    const uint64_t nw = (n_sym + 15) / 16;
    w.u64(n_sym);
    w.u64(nw);
    w.raw(words, nw * 8);
}
// appends the words to `out` (padded to an even word count: the library's 16-byte sequence alignment)
uint64_t decode_seq(Reader &r, std::vector<uint64_t> &out) {
    const uint64_t n_sym = r.u64();
    const uint64_t nw = r.len(8);
    if (nw != (n_sym + 15) / 16) fail(MYRIO_ERR_FORMAT, "Seq<Iupac>: %llu words for %llu symbols", (unsigned long long)nw, (unsigned long long)n_sym);
    const uint8_t *q = r.take(nw * 8);
    const size_t at = out.size();
    out.resize(at + (nw + 1) / 2 * 2, 0);
    memcpy(out.data() + at, q, nw * 8);
    if (n_sym % 16) out[at + nw - 1] &= (~0ull) >> (64 - 4 * (n_sym % 16));  // bits past the end are not data
    return n_sym;
}

// ---- tree (tax/core.rs:355-410) ------------------------------------------------------------------------

// pre-order node list -> nested bincode.  Returns the index after the subtree rooted at `i`.
size_t encode_node(Writer &w, const std::vector<StoreNode> &nodes, size_t i, int depth) {
    if (i >= nodes.size()) fail(MYRIO_ERR_BAD_ARG, "tree: node list ends inside a branch");
    if (depth > 64) fail(MYRIO_ERR_BAD_ARG, "tree deeper than 64 levels");
    const StoreNode &nd = nodes[i];
    if (nd.kind == 1) {  // Node::Leaf
        w.u8(1);
        w.str(nd.name);
        w.u64(nd.arg);  // payload_id
        return i + 1;
    }
    if (nd.kind != 0) fail(MYRIO_ERR_BAD_ARG, "tree: node kind %u", nd.kind);
    w.u8(0);  // Node::Branch
    w.str(nd.name);
    w.u64(nd.arg);  // children: Box<[Node]> -> length
    size_t j = i + 1;
    for (uint64_t c = 0; c < nd.arg; ++c) j = encode_node(w, nodes, j, depth + 1);
    // extra: () -> nothing
    return j;
}

void decode_node(Reader &r, std::vector<StoreNode> &nodes, int depth) {
    if (depth > 64) fail(MYRIO_ERR_FORMAT, "tree deeper than 64 levels");
    const uint8_t tag = r.u8();
    StoreNode nd;
    if (tag == 1) {
        nd.kind = 1;
        nd.name = r.str();
        nd.arg = r.u64();
        nodes.push_back(std::move(nd));
        return;
    }
    if (tag != 0) fail(MYRIO_ERR_FORMAT, "invalid enum tag %u", tag);  // core.rs:381
    nd.kind = 0;
    nd.name = r.str();
    nd.arg = r.len(1 + 8 + 8);  // a node is at least tag + name length + one u64
    const uint64_t n_children = nd.arg;
    nodes.push_back(std::move(nd));
    for (uint64_t c = 0; c < n_children; ++c) decode_node(r, nodes, depth + 1);
}

// ---- payload chunks -----------------------------------------------------------------------------------------

void encode_payloads(Writer &w, const myrio_store &st, uint64_t lo, uint64_t hi) {
    w.u64(hi - lo);  // Vec<StorePayload> / &[StorePayload]
    for (uint64_t p = lo; p < hi; ++p) {
        encode_seq(w, st.seq_words.data() + st.seq_word_off[p], st.seq_base_off[p + 1] - st.seq_base_off[p]);
        w.u64(st.counts.size());  // Vec<(SparseVec<u16>, Float)>
        for (const StoreCounts &c : st.counts) {
            const uint64_t b = c.row_off[p], n = c.row_off[p + 1] - b;
            w.u64(n);
            w.raw(c.keys.data() + b, n * 8);  // keys: Vec<usize>
            w.u64(n);
            w.raw(c.vals.data() + b, n * 2);  // values: Vec<u16>
            w.f32(c.rescale[p]);
        }
    }
}

struct DecodedChunk {
    uint64_t n = 0;
    std::vector<uint64_t> words, word_off, base_off;  // relative to the chunk
    std::vector<StoreCounts> counts;                  // row_off relative to the chunk
};

void decode_payloads(Reader &r, DecodedChunk &d, size_t n_k_expected) {
    d.n = r.len(8 + 8 + 8);
    d.word_off.assign(1, 0);
    d.base_off.assign(1, 0);
    d.counts.assign(n_k_expected, StoreCounts());
    for (auto &c : d.counts) c.row_off.assign(1, 0);
    for (uint64_t p = 0; p < d.n; ++p) {
        const uint64_t n_sym = decode_seq(r, d.words);
        d.word_off.push_back(d.words.size());
        d.base_off.push_back(d.base_off.back() + n_sym);
        const uint64_t nk = r.len(8 + 8 + 4);
        if (nk != n_k_expected)
            fail(MYRIO_ERR_FORMAT, "payload holds %llu count vectors, k_precomputed lists %zu", (unsigned long long)nk, n_k_expected);
        for (uint64_t i = 0; i < nk; ++i) {
            StoreCounts &c = d.counts[i];
            const uint64_t n = r.len(8);
            const uint8_t *kq = r.take(n * 8);
            const uint64_t nv = r.len(2);
            if (nv != n) fail(MYRIO_ERR_FORMAT, "SparseVec with %llu keys and %llu values", (unsigned long long)n, (unsigned long long)nv);
            const uint8_t *vq = r.take(n * 2);
            const size_t at = c.keys.size();
            c.keys.resize(at + n);
            c.vals.resize(at + n);
            memcpy(c.keys.data() + at, kq, n * 8);
            memcpy(c.vals.data() + at, vq, n * 2);
            c.rescale.push_back(r.f32());
            c.row_off.push_back(c.keys.size());
        }
    }
}

void parallel_for(size_t n, unsigned threads, const std::function<void(size_t)> &f) {
    if (n == 0) return;
    threads = (unsigned)std::min<size_t>(std::max(1u, threads), n);
    std::atomic<size_t> next{0};
    std::vector<Failure> errs(threads, Failure{MYRIO_OK, ""});
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            try {
                for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i);
            } catch (const Failure &e) {
                errs[t] = e;
            } catch (const std::exception &e) {
                errs[t] = Failure{MYRIO_ERR_BAD_ARG, e.what()};
            }
        });
    for (auto &th : pool) th.join();
    for (auto &e : errs)
        if (e.code != MYRIO_OK) throw e;
}

void check_store(const myrio_store &st) {
    const uint64_t L = st.n_payloads;
    if (st.seq_word_off.size() != L + 1 || st.seq_base_off.size() != L + 1) fail(MYRIO_ERR_BAD_ARG, "store: sequence offsets");
    if (st.k_precomputed.size() != st.counts.size()) fail(MYRIO_ERR_BAD_ARG, "store: k_precomputed vs counts");
    for (const StoreCounts &c : st.counts)
        if (c.row_off.size() != L + 1 || c.rescale.size() != L || c.keys.size() != c.vals.size() ||
            c.row_off.back() != c.keys.size())
            fail(MYRIO_ERR_BAD_ARG, "store: malformed cached counts");
}
}  // namespace

static const uint8_t MYRTREE_MAGIC[8] = {77, 89, 82, 73, 79, 45, 206, 168};  // "MYRIO-Ψ", store.rs:339-340

unsigned store_default_threads() {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
}

// TaxTreeStore::to_bytes (store.rs:342-405)
std::vector<uint8_t> store_to_bytes(const myrio_store &st, int zstd_level, unsigned threads) {
    check_store(st);
    if (threads == 0) threads = store_default_threads();
    // the core without its payloads (:347-359)
    Writer core;
    core.str(st.gene);
    if (encode_node(core, st.nodes, 0, 0) != st.nodes.size()) fail(MYRIO_ERR_BAD_ARG, "tree: nodes after the root's subtree");
    encode_rank(core, st.root_rank);
    core.u64(0);  // payloads: Box<[StorePayload]> swapped for an empty one
    const std::vector<uint8_t> core_c = zstd_compress(core.b, zstd_level);
    // payload chunks (:363-385)
    const uint64_t len = st.n_payloads;
    const uint64_t chunk_size = len < threads ? len : len / threads + 1;
    if (chunk_size == 0) fail(MYRIO_ERR_EMPTY, "a store without payloads cannot be serialised (par_chunks(0) panics, store.rs:369)");
    const uint64_t n_chunks = (len + chunk_size - 1) / chunk_size;
    std::vector<std::vector<uint8_t>> chunks(n_chunks);
    std::vector<uint64_t> uc(n_chunks);
    parallel_for(n_chunks, threads, [&](size_t c) {
        Writer w;
        const uint64_t lo = c * chunk_size, hi = std::min(len, lo + chunk_size);
        encode_payloads(w, st, lo, hi);
        uc[c] = w.b.size();
        chunks[c] = zstd_compress(w.b, zstd_level);
    });
    // header (:390-391)
    Writer h;
    h.u64(core_c.size());
    h.u64(core.b.size());
    h.u64(n_chunks);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        h.u64(chunks[c].size());
        h.u64(uc[c]);
    }
    h.u64(st.k_precomputed.size());
    for (uint64_t k : st.k_precomputed) h.u64(k);
    // :393-400
    Writer out;
    out.raw(MYRTREE_MAGIC, 8);
    out.u64(h.b.size());
    out.raw(h.b.data(), h.b.size());
    out.raw(core_c.data(), core_c.size());
    for (auto &c : chunks) out.raw(c.data(), c.size());
    return std::move(out.b);
}

// TaxTreeStore::from_bytes (store.rs:408-454)
myrio_store *store_from_bytes(const uint8_t *data, uint64_t len, unsigned threads) {
    if (threads == 0) threads = store_default_threads();
    Reader cur(data, len);
    if (memcmp(cur.take(8), MYRTREE_MAGIC, 8) != 0) fail(MYRIO_ERR_FORMAT, "not a .myrtree file (magic number)");  // NotATaxTreeStore
    const uint64_t header_size = cur.u64();
    const uint8_t *hp = cur.take(header_size);
    Reader h(hp, header_size);
    const uint64_t core_c = h.u64(), core_uc = h.u64();
    const uint64_t n_chunks = h.len(16);
    std::vector<std::pair<uint64_t, uint64_t>> ci(n_chunks);
    for (auto &c : ci) {
        c.first = h.u64();
        c.second = h.u64();
    }
    auto st = std::make_unique<myrio_store>();
    const uint64_t nk = h.len(8);
    for (uint64_t i = 0; i < nk; ++i) st->k_precomputed.push_back(h.u64());
    // core
    {
        if (core_uc > (1ull << 40)) fail(MYRIO_ERR_FORMAT, "core size %llu", (unsigned long long)core_uc);
        const std::vector<uint8_t> raw = zstd_decompress(cur.take(core_c), core_c, core_uc);
        Reader r(raw.data(), raw.size());
        st->gene = r.str();
        decode_node(r, st->nodes, 0);
        st->root_rank = decode_rank(r);
        const uint64_t inline_payloads = r.u64();
        if (inline_payloads != 0) fail(MYRIO_ERR_FORMAT, "core carries %llu inline payloads (to_bytes writes none)", (unsigned long long)inline_payloads);
    }
    // chunks, decoded in parallel like the reference, concatenated in file order
    std::vector<const uint8_t *> cp(n_chunks);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        if (ci[c].second > (1ull << 40)) fail(MYRIO_ERR_FORMAT, "chunk size %llu", (unsigned long long)ci[c].second);
        cp[c] = cur.take(ci[c].first);
    }
    std::vector<DecodedChunk> dec(n_chunks);
    parallel_for(n_chunks, threads, [&](size_t c) {
        const std::vector<uint8_t> raw = zstd_decompress(cp[c], ci[c].first, ci[c].second);
        Reader r(raw.data(), raw.size());
        decode_payloads(r, dec[c], nk);
    });
    st->counts.assign(nk, StoreCounts());
    for (auto &c : st->counts) c.row_off.assign(1, 0);
    st->seq_word_off.assign(1, 0);
    st->seq_base_off.assign(1, 0);
    for (auto &d : dec) {
        const uint64_t w0 = st->seq_words.size(), b0 = st->seq_base_off.back();
        st->seq_words.insert(st->seq_words.end(), d.words.begin(), d.words.end());
        for (uint64_t p = 1; p <= d.n; ++p) {
            st->seq_word_off.push_back(w0 + d.word_off[p]);
            st->seq_base_off.push_back(b0 + d.base_off[p]);
        }
        for (uint64_t i = 0; i < nk; ++i) {
            StoreCounts &dst = st->counts[i], &src = d.counts[i];
            const uint64_t k0 = dst.keys.size();
            dst.keys.insert(dst.keys.end(), src.keys.begin(), src.keys.end());
            dst.vals.insert(dst.vals.end(), src.vals.begin(), src.vals.end());
            dst.rescale.insert(dst.rescale.end(), src.rescale.begin(), src.rescale.end());
            for (uint64_t p = 1; p <= d.n; ++p) dst.row_off.push_back(k0 + src.row_off[p]);
        }
        st->n_payloads += d.n;
        d = DecodedChunk();
    }
    return st.release();
}

// ---- GPU side: the cached counts and the leaf sequences move between the store and device batches ----------

myrio_seqs *store_leaves_to_device(myrio_ctx *ctx, const myrio_store &st) {
    auto s = std::make_unique<myrio_seqs>();
    const uint64_t L = st.n_payloads;
    s->kind = 1;
    s->n = L;
    s->h_base_off = st.seq_base_off;
    s->h_aux.assign(L, 0);
    for (uint64_t r = 0; r < L; ++r) {  // ambiguity codes per leaf (scratch sizing of K2), as myrio_leaves_upload
        const uint64_t nw = (st.seq_base_off[r + 1] - st.seq_base_off[r] + 15) / 16;
        uint32_t amb = 0;
        for (uint64_t w = 0; w < nw; ++w) {
            const uint64_t a = st.seq_words[st.seq_word_off[r] + w];
            uint64_t t = (a & (a >> 1) & 0x7777777777777777ull) | (a & (a >> 2) & 0x3333333333333333ull) |
                         (a & (a >> 3) & 0x1111111111111111ull);
            t |= t >> 1;
            t |= t >> 2;
            amb += (uint32_t)__builtin_popcountll(t & 0x1111111111111111ull);
        }
        s->h_aux[r] = amb;
    }
    const uint64_t total_w = st.seq_words.size();
    s->words.alloc(total_w + 4);
    s->word_off.alloc(L + 1);
    s->base_off.alloc(L + 1);
    MYRIO_CK(cudaMemsetAsync(s->words.p + total_w, 0, 4 * sizeof(uint64_t), ctx->stream));
    h2d(ctx, s->words.p, st.seq_words.data(), total_w);
    h2d(ctx, s->word_off.p, st.seq_word_off.data(), L + 1);
    h2d(ctx, s->base_off.p, st.seq_base_off.data(), L + 1);
    sync(ctx);
    return s.release();
}

myrio_svecs *store_counts_to_device(myrio_ctx *ctx, const myrio_store &st, uint64_t k_index) {
    if (k_index >= st.counts.size()) fail(MYRIO_ERR_BAD_ARG, "k index %llu of %zu", (unsigned long long)k_index, st.counts.size());
    const StoreCounts &c = st.counts[k_index];
    auto v = std::make_unique<myrio_svecs>();
    v->n_rows = st.n_payloads;
    v->nnz = c.keys.size();
    v->vtype = MYRIO_VAL_U16;
    v->aux_kind = 2;
    v->row_off.alloc(v->n_rows + 1);
    v->keys.alloc(std::max<uint64_t>(v->nnz, 1));
    v->vals_u16.alloc(std::max<uint64_t>(v->nnz, 1));
    v->aux_f32.alloc(std::max<uint64_t>(v->n_rows, 1));
    h2d(ctx, v->row_off.p, c.row_off.data(), v->n_rows + 1);
    h2d(ctx, v->keys.p, c.keys.data(), v->nnz);
    h2d(ctx, v->vals_u16.p, c.vals.data(), v->nnz);
    h2d(ctx, v->aux_f32.p, c.rescale.data(), v->n_rows);
    sync(ctx);
    return v.release();
}

// compute_and_append_kmer_counts / _overwrite_ (store.rs:265-308): K2's device batch lands in the store
void store_put_counts(myrio_ctx *ctx, myrio_store &st, uint64_t k, const myrio_svecs *counts, bool overwrite) {
    if (counts->vtype != MYRIO_VAL_U16 || counts->aux_kind != 2)
        fail(MYRIO_ERR_BAD_ARG, "cached counts must be a u16 batch with its rescale column (myrio_count_leaves)");
    if (counts->n_rows != st.n_payloads)
        fail(MYRIO_ERR_BAD_ARG, "%llu count rows for %llu payloads", (unsigned long long)counts->n_rows, (unsigned long long)st.n_payloads);
    StoreCounts c;
    c.row_off.resize(counts->n_rows + 1);
    c.keys.resize(counts->nnz);
    c.vals.resize(counts->nnz);
    c.rescale.resize(counts->n_rows);
    d2h(ctx, c.row_off.data(), counts->row_off.p, counts->n_rows + 1);
    d2h(ctx, c.keys.data(), counts->keys.p, counts->nnz);
    d2h(ctx, c.vals.data(), counts->vals_u16.p, counts->nnz);
    d2h(ctx, c.rescale.data(), counts->aux_f32.p, counts->n_rows);
    sync(ctx);
    if (overwrite) {
        st.counts.clear();
        st.k_precomputed.clear();
    }
    st.counts.push_back(std::move(c));
    st.k_precomputed.push_back(k);
}

}  // namespace myrio
