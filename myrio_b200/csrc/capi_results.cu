// capi_results.cu -- extern "C" entry points of the results pass (results.cu)
#include <memory>

#include "results.cuh"

using namespace myrio;

extern "C" {

int32_t myrio_taxonomy_create(myrio_ctx *ctx, uint32_t root_rank, uint64_t n_nodes, const uint8_t *node_kind,
                              const uint64_t *node_arg, uint64_t n_payloads, myrio_taxonomy **out) {
    return guarded([&] {
        if (!ctx || !out || !node_kind || !node_arg) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        *out = taxonomy_create(ctx, root_rank, n_nodes, node_kind, node_arg, n_payloads);
    });
}

void myrio_taxonomy_free(myrio_taxonomy *t) { delete t; }

int32_t myrio_results_from_compute_tree(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *leaf_vecs,
                                        myrio_taxonomy *taxonomy, const myrio_svecs *queries, uint64_t q_first,
                                        uint64_t q_count, int32_t similarity, const myrio_results_params *params,
                                        myrio_results **out) {
    return guarded([&] {
        if (!ctx || (!ix && !leaf_vecs) || !taxonomy || !queries || !params || (!out && q_count))
            fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        ScopedCtx sc(ctx);
        results_from_compute_tree(ctx, ix, leaf_vecs, taxonomy, queries, q_first, q_count, similarity, *params, out);
    });
}

int32_t myrio_results_cut(const myrio_results *r, uint64_t nb_best, myrio_results **out) {
    return guarded([&] {
        if (!r || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        *out = results_cut(r, nb_best);
    });
}

int32_t myrio_results_info(const myrio_results *r, uint64_t *n_nodes, uint64_t *n_payloads, uint32_t *root_rank) {
    return guarded([&] {
        if (!r) fail(MYRIO_ERR_BAD_ARG, "NULL results");
        if (n_nodes) *n_nodes = r->tree.kind.size();
        if (n_payloads) *n_payloads = r->tree.pscore.size();
        if (root_rank) *root_rank = r->root_rank;
    });
}

int32_t myrio_results_tree(const myrio_results *r, uint8_t *node_kind, uint64_t *node_arg, uint64_t *src_node,
                           uint64_t *nb_leaves, float *mean, float *max, float *pool_score, uint8_t *conf_state,
                           float *conf, uint8_t *force_keep, float *payload_score, uint64_t *payload_src_id) {
    return guarded([&] {
        if (!r) fail(MYRIO_ERR_BAD_ARG, "NULL results");
        const ResTree &t = r->tree;
        auto cp = [](auto *dst, const auto &v) {
            if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
        };
        cp(node_kind, t.kind);
        cp(node_arg, t.arg);
        cp(src_node, t.src);
        cp(nb_leaves, t.nb_leaves);
        cp(mean, t.mean);
        cp(max, t.max);
        cp(pool_score, t.pool);
        cp(conf_state, t.conf_state);
        cp(conf, t.conf);
        cp(force_keep, t.force);
        cp(payload_score, t.pscore);
        cp(payload_src_id, t.porig);
    });
}

void myrio_results_free(myrio_results *r) { delete r; }

int32_t myrio_in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                                     const float *pool_score, const uint8_t *conf_state, const float *conf,
                                     float *out) {
    return guarded([&] {
        if (!genus_off || !name_id || !pool_score || !conf_state || !conf || !out) fail(MYRIO_ERR_BAD_ARG, "NULL argument");
        *out = in_database_confidence(n_genes, genus_off, name_id, pool_score, conf_state, conf);
    });
}

}  // extern "C"
