// capi_store.cu -- extern "C" entry points of the .myrtree reader / writer (myrtree.cu)
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "myrtree.cuh"

using namespace myrio;

extern "C" {

int32_t myrio_store_new(const char *gene, uint32_t root_rank, uint64_t n_nodes, const uint8_t *node_kind,
                        const uint64_t *node_arg, const uint64_t *name_off, const char *names,
                        uint64_t n_payloads, const uint64_t *seq_words, const uint64_t *seq_word_off,
                        const uint64_t *seq_base_off, myrio_store **out) {
    return guarded([&] {
        if (!out || !gene || (n_nodes && (!node_kind || !node_arg || !name_off)) ||
            (n_payloads && (!seq_words || !seq_word_off || !seq_base_off)))
            fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        if (root_rank < 1 || root_rank > 8) fail(MYRIO_ERR_BAD_ARG, "root_rank %u is not a Rank (1..8)", root_rank);
        if (n_nodes == 0) fail(MYRIO_ERR_BAD_ARG, "a tree needs a root node");
        auto st = std::make_unique<myrio_store>();
        st->gene = gene;
        st->root_rank = root_rank;
        st->nodes.resize(n_nodes);
        uint64_t n_leaf = 0;
        for (uint64_t i = 0; i < n_nodes; ++i) {
            if (node_kind[i] > 1) fail(MYRIO_ERR_BAD_ARG, "node %llu: kind %u", (unsigned long long)i, node_kind[i]);
            if (name_off[i + 1] < name_off[i]) fail(MYRIO_ERR_BAD_ARG, "name_off is not ascending");
            st->nodes[i].kind = node_kind[i];
            st->nodes[i].arg = node_arg[i];
            st->nodes[i].name.assign(names + name_off[i], names + name_off[i + 1]);
            if (node_kind[i] == 1) {
                ++n_leaf;
                if (node_arg[i] >= n_payloads) fail(MYRIO_ERR_BAD_ARG, "leaf node %llu: payload_id out of range", (unsigned long long)i);
            }
        }
        st->n_payloads = n_payloads;
        st->seq_word_off.assign(seq_word_off, seq_word_off + n_payloads + 1);
        st->seq_base_off.assign(seq_base_off, seq_base_off + n_payloads + 1);
        for (uint64_t p = 0; p < n_payloads; ++p) {
            if (seq_word_off[p] & 1ull) fail(MYRIO_ERR_BAD_ARG, "sequence %llu does not start on an even word", (unsigned long long)p);
            if ((seq_word_off[p + 1] - seq_word_off[p]) * 16 < seq_base_off[p + 1] - seq_base_off[p])
                fail(MYRIO_ERR_BAD_ARG, "sequence %llu: word range too short", (unsigned long long)p);
        }
        st->seq_words.assign(seq_words, seq_words + (n_payloads ? seq_word_off[n_payloads] : 0));
        *out = st.release();
    });
}

int32_t myrio_store_from_bytes(const uint8_t *data, uint64_t len, uint32_t threads, myrio_store **out) {
    return guarded([&] {
        if (!out || (!data && len)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        *out = store_from_bytes(data, len, threads);
    });
}

int32_t myrio_store_to_bytes(const myrio_store *st, int32_t zstd_level, uint32_t threads, uint8_t **out, uint64_t *len) {
    return guarded([&] {
        if (!st || !out || !len) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        std::vector<uint8_t> b = store_to_bytes(*st, zstd_level, threads);
        uint8_t *p = static_cast<uint8_t *>(malloc(std::max<size_t>(b.size(), 1)));
        if (!p) fail(MYRIO_ERR_OOM, "out of host memory");
        memcpy(p, b.data(), b.size());
        *out = p;
        *len = b.size();
    });
}

void myrio_bytes_free(uint8_t *p) { free(p); }

int32_t myrio_store_save(const myrio_store *st, const char *path, int32_t zstd_level, uint32_t threads) {
    return guarded([&] {
        if (!st || !path) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        std::vector<uint8_t> b = store_to_bytes(*st, zstd_level, threads);
        FILE *f = fopen(path, "wb");  // store.rs:66-71: create, write, truncate
        if (!f) fail(MYRIO_ERR_BAD_ARG, "cannot open '%s' for writing", path);
        const size_t w = fwrite(b.data(), 1, b.size(), f);
        const int rc = fclose(f);
        if (w != b.size() || rc != 0) fail(MYRIO_ERR_BAD_ARG, "short write to '%s'", path);
    });
}

int32_t myrio_store_load(const char *path, uint32_t threads, myrio_store **out) {
    return guarded([&] {
        if (!path || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        FILE *f = fopen(path, "rb");
        if (!f) fail(MYRIO_ERR_BAD_ARG, "cannot open '%s'", path);
        std::vector<uint8_t> b;
        uint8_t buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) b.insert(b.end(), buf, buf + n);
        fclose(f);
        *out = store_from_bytes(b.data(), b.size(), threads);
    });
}

int32_t myrio_store_info(const myrio_store *st, uint64_t *n_nodes, uint64_t *n_payloads, uint64_t *n_k,
                         uint32_t *root_rank, uint64_t *names_bytes, uint64_t *gene_bytes, uint64_t *seq_words_total) {
    return guarded([&] {
        if (!st) fail(MYRIO_ERR_BAD_ARG, "NULL store");
        if (n_nodes) *n_nodes = st->nodes.size();
        if (n_payloads) *n_payloads = st->n_payloads;
        if (n_k) *n_k = st->k_precomputed.size();
        if (root_rank) *root_rank = st->root_rank;
        if (names_bytes) {
            uint64_t t = 0;
            for (auto &nd : st->nodes) t += nd.name.size();
            *names_bytes = t;
        }
        if (gene_bytes) *gene_bytes = st->gene.size();
        if (seq_words_total) *seq_words_total = st->seq_words.size();
    });
}

int32_t myrio_store_tree(const myrio_store *st, char *gene, uint8_t *node_kind, uint64_t *node_arg,
                         uint64_t *name_off, char *names, uint64_t *k_precomputed) {
    return guarded([&] {
        if (!st) fail(MYRIO_ERR_BAD_ARG, "NULL store");
        if (gene) {
            memcpy(gene, st->gene.data(), st->gene.size());
            gene[st->gene.size()] = 0;
        }
        uint64_t off = 0;
        for (size_t i = 0; i < st->nodes.size(); ++i) {
            const StoreNode &nd = st->nodes[i];
            if (node_kind) node_kind[i] = nd.kind;
            if (node_arg) node_arg[i] = nd.arg;
            if (name_off) name_off[i] = off;
            if (names) memcpy(names + off, nd.name.data(), nd.name.size());
            off += nd.name.size();
        }
        if (name_off) name_off[st->nodes.size()] = off;
        if (k_precomputed)
            for (size_t i = 0; i < st->k_precomputed.size(); ++i) k_precomputed[i] = st->k_precomputed[i];
    });
}

int32_t myrio_store_seqs(const myrio_store *st, uint64_t *seq_words, uint64_t *seq_word_off, uint64_t *seq_base_off) {
    return guarded([&] {
        if (!st) fail(MYRIO_ERR_BAD_ARG, "NULL store");
        if (seq_words) memcpy(seq_words, st->seq_words.data(), st->seq_words.size() * 8);
        if (seq_word_off) memcpy(seq_word_off, st->seq_word_off.data(), st->seq_word_off.size() * 8);
        if (seq_base_off) memcpy(seq_base_off, st->seq_base_off.data(), st->seq_base_off.size() * 8);
    });
}

int32_t myrio_store_counts_info(const myrio_store *st, uint64_t k_index, uint64_t *nnz) {
    return guarded([&] {
        if (!st || !nnz) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        if (k_index >= st->counts.size()) fail(MYRIO_ERR_BAD_ARG, "k index out of range");
        *nnz = st->counts[k_index].keys.size();
    });
}

int32_t myrio_store_counts_host(const myrio_store *st, uint64_t k_index, uint64_t *row_off, uint64_t *keys,
                                uint16_t *counts, float *rescale) {
    return guarded([&] {
        if (!st) fail(MYRIO_ERR_BAD_ARG, "NULL store");
        if (k_index >= st->counts.size()) fail(MYRIO_ERR_BAD_ARG, "k index out of range");
        const StoreCounts &c = st->counts[k_index];
        if (row_off) memcpy(row_off, c.row_off.data(), c.row_off.size() * 8);
        if (keys) memcpy(keys, c.keys.data(), c.keys.size() * 8);
        if (counts) memcpy(counts, c.vals.data(), c.vals.size() * 2);
        if (rescale) memcpy(rescale, c.rescale.data(), c.rescale.size() * 4);
    });
}

int32_t myrio_store_leaves(myrio_ctx *ctx, const myrio_store *st, myrio_seqs **out) {
    return guarded([&] {
        if (!ctx || !st || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = store_leaves_to_device(ctx, *st);
    });
}

int32_t myrio_store_counts(myrio_ctx *ctx, const myrio_store *st, uint64_t k_index, myrio_svecs **out) {
    return guarded([&] {
        if (!ctx || !st || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = store_counts_to_device(ctx, *st, k_index);
    });
}

int32_t myrio_store_append_counts(myrio_ctx *ctx, myrio_store *st, uint64_t k, const myrio_svecs *counts) {
    return guarded([&] {
        if (!ctx || !st || !counts) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        if (k < 2 || k > 32) fail(MYRIO_ERR_INVALID_K, "k=%llu: only k in 2..32 is supported", (unsigned long long)k);
        ScopedCtx sc(ctx);
        store_put_counts(ctx, *st, k, counts, false);
    });
}

int32_t myrio_store_overwrite_counts(myrio_ctx *ctx, myrio_store *st, uint64_t k, const myrio_svecs *counts) {
    return guarded([&] {
        if (!ctx || !st || !counts) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        if (k < 2 || k > 32) fail(MYRIO_ERR_INVALID_K, "k=%llu: only k in 2..32 is supported", (unsigned long long)k);
        ScopedCtx sc(ctx);
        store_put_counts(ctx, *st, k, counts, true);
    });
}

int32_t myrio_store_append_counts_host(myrio_store *st, uint64_t k, const uint64_t *row_off, const uint64_t *keys,
                                       const uint16_t *counts, const float *rescale) {
    return guarded([&] {
        // k is metadata here (the reference's own format test stores k_precomputed = [1, 2, 3, 4], store.rs:491)
        if (!st || !row_off || (st->n_payloads && !rescale)) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        const uint64_t L = st->n_payloads, nnz = row_off[L];
        if (nnz && (!keys || !counts)) fail(MYRIO_ERR_BAD_ARG, "NULL keys/counts");
        for (uint64_t p = 0; p < L; ++p)
            if (row_off[p + 1] < row_off[p]) fail(MYRIO_ERR_BAD_ARG, "row_off is not ascending");
        StoreCounts c;
        c.row_off.assign(row_off, row_off + L + 1);
        c.keys.assign(keys, keys + nnz);
        c.vals.assign(counts, counts + nnz);
        c.rescale.assign(rescale, rescale + L);
        st->counts.push_back(std::move(c));
        st->k_precomputed.push_back(k);
    });
}

int32_t myrio_store_shrink(myrio_store *st) {
    return guarded([&] {
        if (!st) fail(MYRIO_ERR_BAD_ARG, "NULL store");
        st->counts.clear();
        st->k_precomputed.clear();
    });
}

void myrio_store_free(myrio_store *st) { delete st; }

}  // extern "C"
