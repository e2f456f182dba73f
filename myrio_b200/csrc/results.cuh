// results.cuh -- taxonomy plan and results trees of the post-scoring pass (results.cu)
#pragma once
#include <vector>

#include "index.cuh"

namespace myrio {
struct ResTree {  // TaxTreeCore<BRes, SimScore> (tax/results.rs:22-61) as a flat PRE-ORDER node list
    std::vector<uint8_t> kind;    // 0 branch, 1 leaf
    std::vector<uint64_t> arg;    // branch: number of children; leaf: index into pscore / porig
    std::vector<uint64_t> src;    // node index in the taxonomy the results were computed over (names live there)
    std::vector<uint64_t> nb_leaves;
    std::vector<float> mean, max, pool, conf;  // BRes (branches)
    std::vector<uint8_t> conf_state;           // 0 None, 1 Some(None), 2 Some(Some(conf))
    std::vector<uint8_t> force;
    std::vector<float> pscore;    // payloads: SimScore of every kept leaf
    std::vector<uint64_t> porig;  // its payload_id in the scored tree
};
}  // namespace myrio

struct myrio_taxonomy {
    uint32_t root_rank = 1;
    uint64_t n_payloads = 0;
    std::vector<uint8_t> kind;
    std::vector<uint64_t> arg;
    std::vector<uint32_t> node_branch, branch_node, depth, leaf_node;
    std::vector<uint32_t> ch_off, ch, lf_off, lf_pid, level_order, level_off, leafy;
    std::vector<uint64_t> nb_leaves;
    myrio::DevBuf<uint32_t> d_ch_off, d_ch, d_lf_off, d_lf_pid, d_level_order, d_leafy, d_nb;
};

struct myrio_results {
    uint32_t root_rank = 1;
    myrio::ResTree tree;
};

namespace myrio {
myrio_taxonomy *taxonomy_create(myrio_ctx *ctx, uint32_t root_rank, uint64_t n_nodes, const uint8_t *kind,
                                const uint64_t *arg, uint64_t n_payloads);
void results_from_compute_tree(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *leaf_vecs,
                               myrio_taxonomy *tx, const myrio_svecs *queries, uint64_t q_first, uint64_t q_count,
                               int32_t sim, const myrio_results_params &p, myrio_results **out);
myrio_results *results_cut(const myrio_results *r, uint64_t nb_best);
float in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                             const float *pool_score, const uint8_t *conf_state, const float *conf);
}  // namespace myrio
