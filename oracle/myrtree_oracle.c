/*
 * myrtree_oracle.c -- CPU restatement of the `.myrtree` byte format.  TEST INFRASTRUCTURE ONLY (see
 * myrio_oracle.h): tests/ compare the library's reader / writer (myrio_b200/csrc/myrtree.cu) with it.
 *
 * Follows, line by line (paths relative to the reference root, Myrio v0.3.0):
 *   myrio-core/src/tax/store.rs:314-328   ChunkInfo {c_size, uc_size}, TTSHeader {core_info, chunk_info_vec, k_precomputed}
 *   myrio-core/src/tax/store.rs:336-340   bincode standard().with_fixed_int_encoding(); MAGIC "MYRIO-Ψ"
 *   myrio-core/src/tax/store.rs:342-405   to_bytes          :408-454 from_bytes
 *   myrio-core/src/tax/store.rs:28-32     StorePayload {seq, kmer_store_counts_vec: Vec<(SparseVec<u16>, Float)>}
 *   myrio-core/src/data/sparse.rs:413-436 SparseVec: keys then values
 *   myrio-core/src/tax/core.rs:22-26      Leaf {name, payload_id}      :332-343 Branch {name, children, extra}
 *   myrio-core/src/tax/core.rs:355-372    Node: u8 tag 0 = Branch, 1 = Leaf   :384-396 TaxTreeCore {gene, root, root_rank, payloads}
 *
 * bincode 2.0.1 fixed-int little endian: usize -> u64; String / Box<str> / Vec<T> / Box<[T]> / &[T] -> u64 length
 * then the items; u8, u16, u32, f32 verbatim; () -> nothing; tuples and structs field by field.
 *
 * PARITY UNPINNED for two encodings that live outside the reference tree (no golden bytes exist: the
 * reference's only format test is a self round trip, store.rs:465-506, restated in tests/test_myrtree.py):
 *   - Seq<Iupac>: the bincode impl of the un-vendored bio-seq fork.  ASSUMED: u64 number of symbols, then the
 *     raw BitVec<usize, Lsb0> words as Vec<usize> -- what Seq::from_raw(len, &[usize]) consumes
 *     (myrio-exp/src/simseq/mod.rs:106).                                   -> put_seq / get_seq
 *   - Rank: #[derive(Encode)] on `#[repr(usize)] enum Rank { Domain = 8, ... Species = 1 }` (tax/clade.rs:93-104).
 *     ASSUMED: bincode-derive writes the explicit discriminant as u32.     -> put_rank / get_rank
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "myrio_oracle.h"

/* ------------------------------------------------------------------ zstd::bulk::{compress, decompress} */

typedef size_t (*zstd_compress_fn)(void *, size_t, const void *, size_t, int);
typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef size_t (*zstd_bound_fn)(size_t);
typedef unsigned (*zstd_iserr_fn)(size_t);
static zstd_compress_fn z_compress;
static zstd_decompress_fn z_decompress;
static zstd_bound_fn z_bound;
static zstd_iserr_fn z_iserr;

static int zstd_load(void) {
    if (z_compress) return 0;
    void *h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!h) return -1;
    z_compress = (zstd_compress_fn)dlsym(h, "ZSTD_compress");
    z_decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
    z_bound = (zstd_bound_fn)dlsym(h, "ZSTD_compressBound");
    z_iserr = (zstd_iserr_fn)dlsym(h, "ZSTD_isError");
    return (z_compress && z_decompress && z_bound && z_iserr) ? 0 : -1;
}

/* ------------------------------------------------------------------ byte buffer */

typedef struct {
    uint8_t *p;
    size_t n, cap;
} buf_t;

static void put(buf_t *b, const void *src, size_t n) {
    if (b->n + n > b->cap) {
        size_t c = b->cap ? b->cap : 256;
        while (c < b->n + n) c *= 2;
        b->p = (uint8_t *)realloc(b->p, c);
        b->cap = c;
    }
    if (n) memcpy(b->p + b->n, src, n);
    b->n += n;
}
static void put_u8(buf_t *b, uint8_t v) { put(b, &v, 1); }
static void put_u32(buf_t *b, uint32_t v) { put(b, &v, 4); } /* little-endian hosts */
static void put_u64(buf_t *b, uint64_t v) { put(b, &v, 8); }
static void put_str(buf_t *b, const char *s, uint64_t n) {
    put_u64(b, n);
    put(b, s, n);
}

/* ASSUMED: Rank -> u32 discriminant */
static void put_rank(buf_t *b, uint32_t rank) { put_u32(b, rank); }

/* ASSUMED: Seq<Iupac> -> u64 symbols, Vec<usize> raw words; symbol i at bits 4*(i%16) of word i/16 */
static void put_seq(buf_t *b, const uint8_t *codes4, uint64_t n) {
    uint64_t nw = (n + 15) / 16;
    put_u64(b, n);
    put_u64(b, nw);
    for (uint64_t w = 0; w < nw; ++w) {
        uint64_t word = 0;
        for (uint64_t i = w * 16; i < n && i < w * 16 + 16; ++i) word |= (uint64_t)(codes4[i] & 15u) << (4 * (i - w * 16));
        put_u64(b, word);
    }
}

/* Node (core.rs:355-372) from the pre-order list; returns the index after the subtree, 0 on error */
static uint64_t put_node(buf_t *b, const myo_store_desc *d, uint64_t i) {
    if (i >= d->n_nodes) return 0;
    const char *name = d->names + d->name_off[i];
    uint64_t name_len = d->name_off[i + 1] - d->name_off[i];
    if (d->kind[i] == 1) { /* Node::Leaf => 1u8, Leaf {name, payload_id} */
        put_u8(b, 1);
        put_str(b, name, name_len);
        put_u64(b, d->arg[i]);
        return i + 1;
    }
    put_u8(b, 0); /* Node::Branch => 0u8, Branch {name, children, extra: ()} */
    put_str(b, name, name_len);
    put_u64(b, d->arg[i]);
    uint64_t j = i + 1;
    for (uint64_t c = 0; c < d->arg[i]; ++c) {
        j = put_node(b, d, j);
        if (!j) return 0;
    }
    return j;
}

static int compress_into(const buf_t *raw, int level, buf_t *out) {
    size_t cap = z_bound(raw->n);
    uint8_t *tmp = (uint8_t *)malloc(cap ? cap : 1);
    size_t n = z_compress(tmp, cap, raw->p, raw->n, level);
    if (z_iserr(n)) {
        free(tmp);
        return -1;
    }
    put(out, tmp, n);
    free(tmp);
    return 0;
}

/* store.rs:342-405 */
int myo_store_encode(const myo_store_desc *d, int zstd_level, unsigned threads, uint8_t **out, uint64_t *out_len) {
    if (zstd_load()) return -20;
    if (threads == 0) threads = 1;
    /* :347-359 the core with its payloads swapped for an empty boxed slice */
    buf_t core = {0};
    put_str(&core, d->gene, strlen(d->gene));
    if (put_node(&core, d, 0) != d->n_nodes) {
        free(core.p);
        return -21;
    }
    put_rank(&core, d->root_rank);
    put_u64(&core, 0);
    buf_t core_c = {0};
    if (compress_into(&core, zstd_level, &core_c)) return -22;
    /* :363-367 */
    uint64_t len = d->n_payloads;
    uint64_t chunk_size = len < threads ? len : len / threads + 1;
    if (chunk_size == 0) { /* par_chunks(0) panics */
        free(core.p);
        free(core_c.p);
        return -23;
    }
    uint64_t n_chunks = (len + chunk_size - 1) / chunk_size;
    buf_t body = {0}, header = {0};
    put_u64(&header, core_c.n); /* core_info */
    put_u64(&header, core.n);
    put_u64(&header, n_chunks); /* chunk_info_vec */
    for (uint64_t c = 0; c < n_chunks; ++c) {
        uint64_t lo = c * chunk_size, hi = lo + chunk_size < len ? lo + chunk_size : len;
        buf_t raw = {0};
        put_u64(&raw, hi - lo);
        for (uint64_t p = lo; p < hi; ++p) {
            put_seq(&raw, d->seq_codes + d->seq_off[p], d->seq_off[p + 1] - d->seq_off[p]);
            put_u64(&raw, d->n_k);
            for (uint64_t i = 0; i < d->n_k; ++i) {
                uint64_t b = d->row_off[i][p], n = d->row_off[i][p + 1] - b;
                put_u64(&raw, n); /* SparseVec.keys */
                put(&raw, d->keys[i] + b, n * 8);
                put_u64(&raw, n); /* SparseVec.values */
                put(&raw, d->vals[i] + b, n * 2);
                put(&raw, &d->rescale[i][p], 4);
            }
        }
        size_t before = body.n;
        if (compress_into(&raw, zstd_level, &body)) return -22;
        put_u64(&header, body.n - before); /* ChunkInfo {c_size, uc_size} */
        put_u64(&header, raw.n);
        free(raw.p);
    }
    put_u64(&header, d->n_k); /* k_precomputed */
    for (uint64_t i = 0; i < d->n_k; ++i) put_u64(&header, d->k_precomputed[i]);
    /* :393-400 */
    static const uint8_t magic[8] = {77, 89, 82, 73, 79, 45, 206, 168};
    buf_t f = {0};
    put(&f, magic, 8);
    put_u64(&f, header.n);
    put(&f, header.p, header.n);
    put(&f, core_c.p, core_c.n);
    put(&f, body.p, body.n);
    free(core.p);
    free(core_c.p);
    free(body.p);
    free(header.p);
    *out = f.p;
    *out_len = f.n;
    return 0;
}

/* ------------------------------------------------------------------ from_bytes (store.rs:408-454) */

typedef struct {
    const uint8_t *p;
    size_t n, pos;
    int bad;
} cur_t;

static const uint8_t *take(cur_t *c, size_t by) { /* BCursor::try_capture */
    if (c->bad || by > c->n - c->pos) {
        c->bad = 1;
        return NULL;
    }
    const uint8_t *q = c->p + c->pos;
    c->pos += by;
    return q;
}
static uint64_t get_u64(cur_t *c) {
    const uint8_t *q = take(c, 8);
    uint64_t v = 0;
    if (q) memcpy(&v, q, 8);
    return v;
}
static uint32_t get_u32(cur_t *c) {
    const uint8_t *q = take(c, 4);
    uint32_t v = 0;
    if (q) memcpy(&v, q, 4);
    return v;
}
static uint64_t get_len(cur_t *c, size_t elem) {
    uint64_t l = get_u64(c);
    if (!c->bad && elem && l > (c->n - c->pos) / elem) c->bad = 1;
    return c->bad ? 0 : l;
}

struct myo_store_owned {
    char *gene;
    uint32_t root_rank;
    uint64_t n_nodes, cap_nodes;
    uint8_t *kind;
    uint64_t *arg, *name_off;
    buf_t names;
    uint64_t n_payloads;
    buf_t seq_codes;
    buf_t seq_off; /* u64 */
    uint64_t n_k;
    uint64_t *k_precomputed;
    buf_t *row_off, *keys, *vals, *rescale;
};

static void push_node(myo_store_owned *s, uint8_t kind, uint64_t arg, const uint8_t *name, uint64_t nlen) {
    if (s->n_nodes == s->cap_nodes) {
        s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 64;
        s->kind = (uint8_t *)realloc(s->kind, s->cap_nodes);
        s->arg = (uint64_t *)realloc(s->arg, s->cap_nodes * 8);
        s->name_off = (uint64_t *)realloc(s->name_off, (s->cap_nodes + 1) * 8);
    }
    s->kind[s->n_nodes] = kind;
    s->arg[s->n_nodes] = arg;
    s->name_off[s->n_nodes] = s->names.n;
    put(&s->names, name, nlen);
    s->n_nodes++;
    s->name_off[s->n_nodes] = s->names.n;
}

static void get_node(cur_t *c, myo_store_owned *s, int depth) {
    if (depth > 64) c->bad = 1;
    const uint8_t *t = take(c, 1);
    if (!t) return;
    uint64_t nlen = get_len(c, 1);
    const uint8_t *name = take(c, nlen);
    if (c->bad) return;
    if (*t == 1) {
        push_node(s, 1, get_u64(c), name, nlen);
    } else if (*t == 0) {
        uint64_t nch = get_len(c, 17);
        push_node(s, 0, nch, name, nlen);
        for (uint64_t i = 0; i < nch && !c->bad; ++i) get_node(c, s, depth + 1);
    } else {
        c->bad = 1; /* "invalid enum tag" */
    }
}

static uint32_t get_rank(cur_t *c) { /* ASSUMED, see header */
    uint32_t v = get_u32(c);
    if (v < 1 || v > 8) c->bad = 1;
    return v;
}

static void get_seq(cur_t *c, myo_store_owned *s) { /* ASSUMED, see header */
    uint64_t n = get_u64(c);
    uint64_t nw = get_len(c, 8);
    if (nw != (n + 15) / 16) c->bad = 1;
    const uint8_t *q = take(c, nw * 8);
    if (c->bad) return;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t word;
        memcpy(&word, q + (i / 16) * 8, 8);
        uint8_t code = (uint8_t)((word >> (4 * (i % 16))) & 15u);
        put(&s->seq_codes, &code, 1);
    }
    uint64_t end = s->seq_codes.n;
    put(&s->seq_off, &end, 8);
}

void myo_store_free(myo_store_owned *s) {
    if (!s) return;
    free(s->gene);
    free(s->kind);
    free(s->arg);
    free(s->name_off);
    free(s->names.p);
    free(s->seq_codes.p);
    free(s->seq_off.p);
    free(s->k_precomputed);
    for (uint64_t i = 0; i < s->n_k; ++i) {
        if (s->row_off) free(s->row_off[i].p);
        if (s->keys) free(s->keys[i].p);
        if (s->vals) free(s->vals[i].p);
        if (s->rescale) free(s->rescale[i].p);
    }
    free(s->row_off);
    free(s->keys);
    free(s->vals);
    free(s->rescale);
    free(s);
}

int myo_store_decode(const uint8_t *data, uint64_t len, myo_store_owned **out) {
    if (zstd_load()) return -20;
    static const uint8_t magic[8] = {77, 89, 82, 73, 79, 45, 206, 168};
    cur_t cur = {data, len, 0, 0};
    const uint8_t *m = take(&cur, 8);
    if (!m || memcmp(m, magic, 8) != 0) return -30; /* NotATaxTreeStore */
    uint64_t hsize = get_u64(&cur);
    const uint8_t *hp = take(&cur, hsize);
    if (cur.bad) return -31;
    cur_t h = {hp, hsize, 0, 0};
    uint64_t core_c = get_u64(&h), core_uc = get_u64(&h);
    uint64_t n_chunks = get_len(&h, 16);
    uint64_t *ci = (uint64_t *)malloc((n_chunks * 2 + 1) * 8);
    for (uint64_t i = 0; i < n_chunks * 2; ++i) ci[i] = get_u64(&h);
    myo_store_owned *s = (myo_store_owned *)calloc(1, sizeof(*s));
    s->n_k = get_len(&h, 8);
    s->k_precomputed = (uint64_t *)malloc((s->n_k + 1) * 8);
    for (uint64_t i = 0; i < s->n_k; ++i) s->k_precomputed[i] = get_u64(&h);
    int rc = 0;
    if (h.bad || core_uc > ((uint64_t)1 << 40)) {
        rc = -31;
        goto done;
    }
    s->row_off = (buf_t *)calloc(s->n_k + 1, sizeof(buf_t));
    s->keys = (buf_t *)calloc(s->n_k + 1, sizeof(buf_t));
    s->vals = (buf_t *)calloc(s->n_k + 1, sizeof(buf_t));
    s->rescale = (buf_t *)calloc(s->n_k + 1, sizeof(buf_t));
    uint64_t zero = 0;
    put(&s->seq_off, &zero, 8);
    for (uint64_t i = 0; i < s->n_k; ++i) put(&s->row_off[i], &zero, 8);
    { /* :421-429 core */
        const uint8_t *cp = take(&cur, core_c);
        if (!cp) {
            rc = -31;
            goto done;
        }
        uint8_t *raw = (uint8_t *)malloc(core_uc ? core_uc : 1);
        size_t n = z_decompress(raw, core_uc, cp, core_c);
        if (z_iserr(n)) {
            free(raw);
            rc = -32;
            goto done;
        }
        cur_t r = {raw, n, 0, 0};
        uint64_t glen = get_len(&r, 1);
        const uint8_t *g = take(&r, glen);
        if (g) {
            s->gene = (char *)malloc(glen + 1);
            memcpy(s->gene, g, glen);
            s->gene[glen] = 0;
        }
        get_node(&r, s, 0);
        s->root_rank = get_rank(&r);
        if (get_u64(&r) != 0) r.bad = 1; /* to_bytes writes no inline payloads */
        free(raw);
        if (r.bad) {
            rc = -33;
            goto done;
        }
    }
    for (uint64_t c = 0; c < n_chunks; ++c) { /* :431-446, in file order */
        uint64_t c_size = ci[2 * c], uc_size = ci[2 * c + 1];
        const uint8_t *cp = take(&cur, c_size);
        if (!cp || uc_size > ((uint64_t)1 << 40)) {
            rc = -31;
            goto done;
        }
        uint8_t *raw = (uint8_t *)malloc(uc_size ? uc_size : 1);
        size_t n = z_decompress(raw, uc_size, cp, c_size);
        if (z_iserr(n)) {
            free(raw);
            rc = -32;
            goto done;
        }
        cur_t r = {raw, n, 0, 0};
        uint64_t np = get_len(&r, 24);
        for (uint64_t p = 0; p < np && !r.bad; ++p) {
            get_seq(&r, s);
            uint64_t nk = get_len(&r, 20);
            if (nk != s->n_k) r.bad = 1;
            for (uint64_t i = 0; i < nk && !r.bad; ++i) {
                uint64_t nkeys = get_len(&r, 8);
                const uint8_t *kq = take(&r, nkeys * 8);
                uint64_t nvals = get_len(&r, 2);
                const uint8_t *vq = take(&r, nvals * 2);
                const uint8_t *fq = take(&r, 4);
                if (r.bad || nkeys != nvals) {
                    r.bad = 1;
                    break;
                }
                put(&s->keys[i], kq, nkeys * 8);
                put(&s->vals[i], vq, nvals * 2);
                put(&s->rescale[i], fq, 4);
                uint64_t end = s->keys[i].n / 8;
                put(&s->row_off[i], &end, 8);
            }
            s->n_payloads++;
        }
        free(raw);
        if (r.bad) {
            rc = -33;
            goto done;
        }
    }
done:
    free(ci);
    if (rc) {
        myo_store_free(s);
        return rc;
    }
    *out = s;
    return 0;
}

void myo_store_view(const myo_store_owned *s, myo_store_desc *d, const uint64_t **row_off, const uint64_t **keys,
                    const uint16_t **vals, const float **rescale) {
    d->gene = s->gene;
    d->root_rank = s->root_rank;
    d->n_nodes = s->n_nodes;
    d->kind = s->kind;
    d->arg = s->arg;
    d->name_off = s->name_off;
    d->names = (const char *)s->names.p;
    d->n_payloads = s->n_payloads;
    d->seq_codes = s->seq_codes.p;
    d->seq_off = (const uint64_t *)s->seq_off.p;
    d->n_k = s->n_k;
    d->k_precomputed = s->k_precomputed;
    for (uint64_t i = 0; i < s->n_k; ++i) {
        row_off[i] = (const uint64_t *)s->row_off[i].p;
        keys[i] = (const uint64_t *)s->keys[i].p;
        vals[i] = (const uint16_t *)s->vals[i].p;
        rescale[i] = (const float *)s->rescale[i].p;
    }
    d->row_off = row_off;
    d->keys = keys;
    d->vals = vals;
    d->rescale = rescale;
}

/* one section of a .myrtree file, decompressed: which = -1 header (as stored), 0 core, 1 + c payload chunk c.
 * returns a malloc'd buffer (free with myo_bytes_free). Used by the tests to compare the bincode bytes with a
 * hand-assembled expectation independently of the compressor. */
int myo_store_section(const uint8_t *data, uint64_t len, int which, uint8_t **out, uint64_t *out_len) {
    if (zstd_load()) return -20;
    cur_t cur = {data, len, 0, 0};
    if (!take(&cur, 8)) return -31;
    uint64_t hsize = get_u64(&cur);
    const uint8_t *hp = take(&cur, hsize);
    if (cur.bad) return -31;
    if (which < 0) {
        *out = (uint8_t *)malloc(hsize ? hsize : 1);
        memcpy(*out, hp, hsize);
        *out_len = hsize;
        return 0;
    }
    cur_t h = {hp, hsize, 0, 0};
    uint64_t c_size = get_u64(&h), uc_size = get_u64(&h);
    uint64_t n_chunks = get_len(&h, 16);
    const uint8_t *sp = take(&cur, c_size);
    for (int c = 0; c < which; ++c) {
        if ((uint64_t)c >= n_chunks) return -34;
        c_size = get_u64(&h);
        uc_size = get_u64(&h);
        sp = take(&cur, c_size);
    }
    if (cur.bad || h.bad || !sp) return -31;
    uint8_t *raw = (uint8_t *)malloc(uc_size ? uc_size : 1);
    size_t n = z_decompress(raw, uc_size, sp, c_size);
    if (z_iserr(n)) {
        free(raw);
        return -32;
    }
    *out = raw;
    *out_len = n;
    return 0;
}

void myo_bytes_free(uint8_t *p) { free(p); }
