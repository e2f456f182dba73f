// results.cu -- what TaxTreeResults::from_compute_tree does AFTER the score loop
// (myrio-core/src/tax/results.rs:96-252): per-branch nb_leaves / mean / max over ALL leaves, trim(n) (the n
// best direct leaves of every branch), cut(x) (global top-x of the trimmed leaves, tree rebuilt), softmax
// pooling per rank, confidence, force_keep; and the in-database confidence of a run
// (myrio-cli/src/app/process/run.rs:405-466).
//
// Device (per batch of queries, score rows never leave HBM):
//   K8a branch_stats_kernel  one thread per (branch, query), level by level from the deepest branches up:
//       sum = ((0 + child_1) + child_2) + ... in CHILD ORDER, a leaf child contributing its score and a branch
//       child its own sum -- the nesting of dive_recursive (results.rs:96-143), so mean = sum / nb_leaves has
//       the reference's f32 bits; max likewise.
//   K8b branch_trim_kernel   one warp per (branch with direct leaves, query): the n best of those leaves by
//       (score desc, child order asc), n = max_amount_of_leaves_per_branch (results.rs:286-293).
// Host (library C++, flat pre-order arrays, f32 in the reference's order, expf / powf from libm like Rust's
// f32::exp / f32::powf): the trimmed tree, cut, pooling and confidence.
//
// Ties: k_largest_by_key (itertools 0.14.0) leaves the order of equal keys unspecified and the reference's
// child order comes from a HashMap; here equal scores keep the input order of the taxonomy (documented contract).
#include <algorithm>
#include <cmath>
#include <memory>

#include "results.cuh"
#include "score.cuh"

namespace myrio {

static constexpr uint32_t BR_FLAG = 0x80000000u;

// ------------------------------------------------------------------ taxonomy plan

myrio_taxonomy *taxonomy_create(myrio_ctx *ctx, uint32_t root_rank, uint64_t n_nodes, const uint8_t *kind,
                                const uint64_t *arg, uint64_t n_payloads) {
    if (root_rank < 1 || root_rank > 8) fail(MYRIO_ERR_BAD_ARG, "root_rank %u is not a Rank (1..8)", root_rank);
    if (n_nodes == 0 || kind[0] != 0) fail(MYRIO_ERR_BAD_ARG, "the root of a taxonomy must be a branch");
    if (n_nodes >= BR_FLAG || n_payloads >= BR_FLAG) fail(MYRIO_ERR_BAD_ARG, "taxonomy too large");
    auto t = std::make_unique<myrio_taxonomy>();
    t->root_rank = root_rank;
    t->n_payloads = n_payloads;
    t->kind.assign(kind, kind + n_nodes);
    t->arg.assign(arg, arg + n_nodes);
    t->node_branch.assign(n_nodes, 0xFFFFFFFFu);
    // one pass over the pre-order list with a stack of open branches
    struct Open {
        uint32_t b;
        uint64_t left;
    };
    std::vector<Open> stack;
    std::vector<std::vector<uint32_t>> children;  // per branch, in child order
    std::vector<uint8_t> seen(n_payloads, 0);
    for (uint64_t i = 0; i < n_nodes; ++i) {
        if (i > 0 && stack.empty()) fail(MYRIO_ERR_BAD_ARG, "taxonomy: node %llu lies outside the root's subtree", (unsigned long long)i);
        uint32_t entry;
        if (kind[i] == 1) {
            if (arg[i] >= n_payloads) fail(MYRIO_ERR_BAD_ARG, "taxonomy: leaf node %llu has payload_id %llu >= %llu", (unsigned long long)i, (unsigned long long)arg[i], (unsigned long long)n_payloads);
            if (seen[arg[i]]) fail(MYRIO_ERR_BAD_ARG, "taxonomy: payload_id %llu occurs twice", (unsigned long long)arg[i]);
            seen[arg[i]] = 1;
            entry = (uint32_t)arg[i];
        } else if (kind[i] == 0) {
            const uint32_t b = (uint32_t)t->branch_node.size();
            t->node_branch[i] = b;
            t->branch_node.push_back((uint32_t)i);
            t->depth.push_back((uint32_t)stack.size());
            children.emplace_back();
            entry = BR_FLAG | b;
        } else {
            fail(MYRIO_ERR_BAD_ARG, "taxonomy: node kind %u", kind[i]);
        }
        if (!stack.empty()) {
            children[stack.back().b].push_back(entry);
            --stack.back().left;
        }
        if (kind[i] == 0) stack.push_back({(uint32_t)(t->branch_node.size() - 1), arg[i]});
        while (!stack.empty() && stack.back().left == 0) stack.pop_back();
        if (stack.size() > 64) fail(MYRIO_ERR_BAD_ARG, "taxonomy deeper than 64 levels");
    }
    if (!stack.empty()) fail(MYRIO_ERR_BAD_ARG, "taxonomy: the node list ends inside a branch");
    const uint32_t B = (uint32_t)t->branch_node.size();
    t->ch_off.assign(B + 1, 0);
    t->lf_off.assign(B + 1, 0);
    t->nb_leaves.assign(B, 0);
    std::vector<uint32_t> height(B, 0);
    for (uint32_t b = 0; b < B; ++b) {
        t->ch_off[b + 1] = t->ch_off[b] + (uint32_t)children[b].size();
        for (uint32_t e : children[b]) {
            t->ch.push_back(e);
            if (!(e & BR_FLAG)) t->lf_pid.push_back(e);
        }
        t->lf_off[b + 1] = (uint32_t)t->lf_pid.size();
    }
    for (uint32_t b = B; b-- > 0;) {  // children follow their parent in pre-order: reverse = bottom-up
        uint64_t nb = 0;
        uint32_t h = 0;
        for (uint32_t e : children[b]) {
            if (e & BR_FLAG) {
                nb += t->nb_leaves[e & ~BR_FLAG];
                h = std::max(h, height[e & ~BR_FLAG] + 1);
            } else {
                nb += 1;
            }
        }
        t->nb_leaves[b] = nb;
        height[b] = h;
    }
    uint32_t max_h = 0;
    for (uint32_t b = 0; b < B; ++b) max_h = std::max(max_h, height[b]);
    t->level_off.assign(max_h + 2, 0);
    for (uint32_t b = 0; b < B; ++b) t->level_off[height[b] + 1]++;
    for (uint32_t l = 0; l <= max_h; ++l) t->level_off[l + 1] += t->level_off[l];
    t->level_order.assign(B, 0);
    {
        std::vector<uint32_t> cur(t->level_off.begin(), t->level_off.end() - 1);
        for (uint32_t b = 0; b < B; ++b) t->level_order[cur[height[b]]++] = b;
    }
    for (uint32_t b = 0; b < B; ++b)
        if (t->lf_off[b + 1] > t->lf_off[b]) t->leafy.push_back(b);
    t->d_ch_off.alloc(B + 1);
    t->d_ch.alloc(std::max<size_t>(t->ch.size(), 1));
    t->d_lf_off.alloc(B + 1);
    t->d_lf_pid.alloc(std::max<size_t>(t->lf_pid.size(), 1));
    t->d_level_order.alloc(B);
    t->d_leafy.alloc(std::max<size_t>(t->leafy.size(), 1));
    t->d_nb.alloc(B);
    std::vector<uint32_t> nb32(B);
    for (uint32_t b = 0; b < B; ++b) nb32[b] = (uint32_t)t->nb_leaves[b];
    h2d(ctx, t->d_ch_off.p, t->ch_off.data(), B + 1);
    h2d(ctx, t->d_ch.p, t->ch.data(), t->ch.size());
    h2d(ctx, t->d_lf_off.p, t->lf_off.data(), B + 1);
    h2d(ctx, t->d_lf_pid.p, t->lf_pid.data(), t->lf_pid.size());
    h2d(ctx, t->d_level_order.p, t->level_order.data(), B);
    h2d(ctx, t->d_leafy.p, t->leafy.data(), t->leafy.size());
    h2d(ctx, t->d_nb.p, nb32.data(), B);
    sync(ctx);
    return t.release();
}

// ------------------------------------------------------------------ kernels

// K8a: branches of one height level; thread = (branch of the level, query)
__global__ void __launch_bounds__(128) branch_stats_kernel(const uint32_t *__restrict__ level_order, uint32_t n_level,
                                                           const uint32_t *__restrict__ ch_off,
                                                           const uint32_t *__restrict__ ch,
                                                           const uint32_t *__restrict__ nb,
                                                           const float *__restrict__ scores, uint64_t ld, uint32_t B,
                                                           float *__restrict__ bsum, float *__restrict__ bmean,
                                                           float *__restrict__ bmax) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = blockIdx.y;
    if (i >= n_level) return;
    const uint32_t b = level_order[i];
    const float *row = scores + (size_t)q * ld;
    float sum = 0.0f, mx = 0.0f;  // results.rs:108-110
    for (uint32_t e = ch_off[b]; e < ch_off[b + 1]; ++e) {
        const uint32_t c = ch[e];
        float s, m;
        if (c & BR_FLAG) {
            s = bsum[(size_t)q * B + (c & ~BR_FLAG)];
            m = bmax[(size_t)q * B + (c & ~BR_FLAG)];
        } else {
            s = m = row[c];
        }
        sum = __fadd_rn(sum, s);  // sum += sum_i (:117)
        mx = fmaxf(mx, m);        // max = max.max(max_i) (:118)
    }
    bsum[(size_t)q * B + b] = sum;
    bmean[(size_t)q * B + b] = __fdiv_rn(sum, (float)nb[b]);  // sum / nb_leaves as Float (:126)
    bmax[(size_t)q * B + b] = mx;
}

// K8b: warp = (branch with direct leaves, query).  Round j picks the best composite (score bits, earlier child
// first) below the previous pick: the n best in descending order, ties in child order.
__global__ void __launch_bounds__(256) branch_trim_kernel(const uint32_t *__restrict__ leafy, uint32_t n_leafy,
                                                          const uint32_t *__restrict__ lf_off,
                                                          const uint32_t *__restrict__ lf_pid,
                                                          const uint32_t *__restrict__ tr_off, uint32_t T,
                                                          uint32_t max_keep, const float *__restrict__ scores,
                                                          uint64_t ld, float *__restrict__ out_score,
                                                          uint32_t *__restrict__ out_pid) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t q = blockIdx.y, lane = threadIdx.x & 31;
    if (w >= n_leafy) return;
    const uint32_t b = leafy[w];
    const uint32_t lo = lf_off[b], n = lf_off[b + 1] - lo;
    const uint32_t keep = min(n, max_keep);
    const float *row = scores + (size_t)q * ld;
    unsigned long long prev = ~0ull;
    for (uint32_t j = 0; j < keep; ++j) {
        unsigned long long best = 0;
        for (uint32_t i = lane; i < n; i += 32) {
            const unsigned long long c =
                ((unsigned long long)__float_as_uint(row[lf_pid[lo + i]]) << 32) | (unsigned long long)(0xFFFFFFFEu - i);
            if (c < prev && c > best) best = c;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o > best ? o : best;
        }
        prev = best;
        if (lane == 0) {
            const uint32_t i = 0xFFFFFFFEu - (uint32_t)best;
            out_score[(size_t)q * T + tr_off[w] + j] = __uint_as_float((uint32_t)(best >> 32));
            out_pid[(size_t)q * T + tr_off[w] + j] = lf_pid[lo + i];
        }
    }
}

// ------------------------------------------------------------------ host: flat results trees

namespace {

void tree_push(ResTree &t, uint8_t kind, uint64_t arg, uint64_t src) {
    t.kind.push_back(kind);
    t.arg.push_back(arg);
    t.src.push_back(src);
    t.nb_leaves.push_back(0);
    t.mean.push_back(0.0f);
    t.max.push_back(0.0f);
    t.pool.push_back(0.0f);
    t.conf.push_back(0.0f);
    t.conf_state.push_back(0);
    t.force.push_back(0);
}

// parent of every node of a pre-order list (root: ~0), depth of every node
void tree_links(const ResTree &t, std::vector<uint64_t> &parent, std::vector<uint32_t> &depth) {
    const size_t n = t.kind.size();
    parent.assign(n, ~0ull);
    depth.assign(n, 0);
    std::vector<std::pair<uint64_t, uint64_t>> stack;  // (node, children left)
    for (size_t i = 0; i < n; ++i) {
        if (!stack.empty()) {
            parent[i] = stack.back().first;
            depth[i] = (uint32_t)stack.size();
            --stack.back().second;
        }
        if (t.kind[i] == 0) stack.emplace_back(i, t.arg[i]);
        while (!stack.empty() && stack.back().second == 0) stack.pop_back();
    }
}

// TaxTreeResults::cut (results.rs:310-381)
ResTree tree_cut(const ResTree &t, uint64_t nb_best) {
    const size_t n = t.kind.size();
    // gather_leaves order == pre-order; sorted_by_key(-score) is stable; take(nb_best)
    std::vector<uint64_t> leaves;
    for (size_t i = 0; i < n; ++i)
        if (t.kind[i] == 1) leaves.push_back(i);
    std::vector<uint64_t> order(leaves);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint64_t a, uint64_t b) { return t.pscore[t.arg[a]] > t.pscore[t.arg[b]]; });
    std::vector<uint8_t> keep(n, 0);
    for (size_t i = 0; i < order.size() && i < nb_best; ++i) keep[order[i]] = 1;
    std::vector<uint64_t> parent;
    std::vector<uint32_t> depth;
    tree_links(t, parent, depth);
    // a node survives iff a kept leaf or a force_keep branch sits at or below it (:338-346)
    for (size_t i = n; i-- > 0;) {
        if (t.kind[i] == 0 && t.force[i]) keep[i] = 1;
        if (keep[i] && parent[i] != ~0ull) keep[parent[i]] = 1;
    }
    ResTree out;
    if (n == 0 || !keep[0]) return out;  // extract_singlevec of an empty vector panics in the reference
    std::vector<uint64_t> new_index(n, ~0ull);
    for (size_t i = 0; i < n; ++i) {
        if (!keep[i]) continue;
        new_index[i] = out.kind.size();
        if (t.kind[i] == 1) {
            tree_push(out, 1, out.pscore.size(), t.src[i]);
            out.pscore.push_back(t.pscore[t.arg[i]]);
            out.porig.push_back(t.porig[t.arg[i]]);
        } else {
            tree_push(out, 0, 0, t.src[i]);
            const size_t o = out.kind.size() - 1;
            out.nb_leaves[o] = t.nb_leaves[i];
            out.mean[o] = t.mean[i];
            out.max[o] = t.max[i];
            out.pool[o] = t.pool[i];
            out.conf[o] = t.conf[i];
            out.conf_state[o] = t.conf_state[i];
            out.force[o] = t.force[i];
        }
        if (parent[i] != ~0ull) out.arg[new_index[parent[i]]]++;
    }
    return out;
}

float softmax_pooling(const std::vector<float> &scores, float lambda) {  // results.rs:160-173
    float numerator = 0.0f, denominator = 0.0f;
    for (float s : scores) {
        numerator = numerator + s * expf(lambda * s);
        denominator = denominator + expf(lambda * s);
    }
    return numerator / denominator;
}

// results.rs:175-248
void pool_and_confidence(ResTree &t, uint32_t root_rank, const myrio_results_params &p) {
    const size_t n = t.kind.size();
    std::vector<uint64_t> parent;
    std::vector<uint32_t> depth;
    tree_links(t, parent, depth);
    // children lists (child order) of every branch
    std::vector<std::vector<uint64_t>> kids(n);
    for (size_t i = 1; i < n; ++i) kids[parent[i]].push_back(i);
    for (uint32_t rank = 2; rank <= root_rank; ++rank) {  // Genus ..= root_rank
        const uint32_t want = root_rank - rank;            // gather_branches_at_rank: height -> depth
        std::vector<uint64_t> at;
        for (size_t i = 0; i < n; ++i)
            if (t.kind[i] == 0 && depth[i] == want) at.push_back(i);
        for (uint64_t b : at) {
            std::vector<float> sims, pools;
            for (uint64_t c : kids[b]) {
                if (t.kind[c] == 1)
                    sims.push_back(t.pscore[t.arg[c]]);
                else
                    pools.push_back(t.pool[c]);
            }
            if (!sims.empty()) {
                const float nf = (float)sims.size();
                const float penalty = nf / (nf + p.mu);
                pools.push_back(softmax_pooling(sims, p.lambda_leaf) * penalty);
            }
            t.pool[b] = softmax_pooling(pools, p.lambda_branch);
        }
        if (at.empty()) continue;
        if (at.size() == 1) {
            if (rank != root_rank) t.conf_state[at[0]] = 1;  // Some(None)
            continue;
        }
        const float nf = (float)at.size();
        float sum = 0.0f;
        for (uint64_t b : at) sum = sum + t.pool[b];
        const float mean = sum / nf;
        float sq = 0.0f;
        for (uint64_t b : at) {
            const float d = t.pool[b] - mean;
            sq = sq + d * d;
        }
        const float sd = fmaxf(sqrtf((1.0f / nf) * sq), 1E-3f);
        size_t first = ~(size_t)0, second = ~(size_t)0;  // k_largest_by_key(2): ties -> earlier branch first
        for (size_t j = 0; j < at.size(); ++j) {
            const float v = t.pool[at[j]];
            if (!std::isfinite(v)) fail(MYRIO_ERR_NONFINITE, "non-finite pooling score (SimScore::try_new(..).unwrap() panics, results.rs:223)");
            if (first == ~(size_t)0 || v > t.pool[at[first]]) {
                second = first;
                first = j;
            } else if (second == ~(size_t)0 || v > t.pool[at[second]]) {
                second = j;
            }
        }
        const uint64_t f = at[first], s = at[second];
        const float x = (t.pool[f] - t.pool[s]) / sd;
        const float conf = powf(t.pool[f], p.epsilon) * powf(x, p.gamma) / (p.delta + powf(x, p.gamma));
        t.conf_state[f] = 2;
        t.conf[f] = conf;
        t.force[f] = 1;
        if (t.pool[s] > 0.2f && conf < 0.3f) t.force[s] = 1;
    }
}

// the trimmed tree (results.rs:257-307) of one query from the device results
void emit_trimmed(const myrio_taxonomy &tx, uint32_t b, const std::vector<uint32_t> &tr_slot, const float *bmean,
                  const float *bmax, const float *tr_score, const uint32_t *tr_pid, const uint32_t *tr_off,
                  uint32_t max_keep, ResTree &out) {
    tree_push(out, 0, 0, tx.branch_node[b]);
    const size_t me = out.kind.size() - 1;
    out.nb_leaves[me] = tx.nb_leaves[b];
    out.mean[me] = bmean[b];
    out.max[me] = bmax[b];
    uint64_t n_children = 0;
    for (uint32_t e = tx.ch_off[b]; e < tx.ch_off[b + 1]; ++e)  // branch children first, in child order (:280-282)
        if (tx.ch[e] & BR_FLAG) {
            emit_trimmed(tx, tx.ch[e] & ~BR_FLAG, tr_slot, bmean, bmax, tr_score, tr_pid, tr_off, max_keep, out);
            ++n_children;
        }
    const uint32_t n_leaf = tx.lf_off[b + 1] - tx.lf_off[b];
    if (n_leaf) {  // then the kept leaves, best first (:286-293)
        const uint32_t w = tr_slot[b], keep = std::min(n_leaf, max_keep);
        for (uint32_t j = 0; j < keep; ++j) {
            const uint32_t pid = tr_pid[tr_off[w] + j];
            tree_push(out, 1, out.pscore.size(), tx.leaf_node.empty() ? 0 : tx.leaf_node[pid]);
            out.pscore.push_back(tr_score[tr_off[w] + j]);
            out.porig.push_back(pid);
            ++n_children;
        }
    }
    out.arg[me] = n_children;
}

}  // namespace

// TaxTreeResults::from_compute_tree for queries [q_first, q_first + q_count): score rows (K4 over the index, or the
// merge-join path when leaf_vecs is given), K8a / K8b, then the host part per query
void results_from_compute_tree(myrio_ctx *ctx, const myrio_index *ix, const myrio_svecs *leaf_vecs,
                               myrio_taxonomy *tx, const myrio_svecs *queries, uint64_t q_first, uint64_t q_count,
                               int32_t sim, const myrio_results_params &p, myrio_results **out) {
    const uint64_t L = leaf_vecs ? leaf_vecs->n_rows : ix->n_leaves;
    if (tx->n_payloads != L) fail(MYRIO_ERR_BAD_ARG, "taxonomy of %llu payloads over %llu leaf vectors", (unsigned long long)tx->n_payloads, (unsigned long long)L);
    if (!leaf_vecs && ix->leaf_id_base != 0) fail(MYRIO_ERR_BAD_ARG, "the results pass needs the index of the whole tree (leaf_id_base 0)");
    if (p.max_amount_of_leaves_per_branch >= BR_FLAG) fail(MYRIO_ERR_BAD_ARG, "max_amount_of_leaves_per_branch too large");
    const uint32_t B = (uint32_t)tx->branch_node.size();
    const uint32_t max_keep = (uint32_t)p.max_amount_of_leaves_per_branch;
    if (tx->leaf_node.empty()) {  // payload id -> node index (names of the kept leaves)
        tx->leaf_node.assign(tx->n_payloads, 0);
        for (size_t i = 0; i < tx->kind.size(); ++i)
            if (tx->kind[i] == 1) tx->leaf_node[tx->arg[i]] = (uint32_t)i;
    }
    // output layout of the trim pass: tr_off[w] over the branches with direct leaves
    const uint32_t n_leafy = (uint32_t)tx->leafy.size();
    std::vector<uint32_t> tr_off(n_leafy + 1, 0), tr_slot(B, 0xFFFFFFFFu);
    for (uint32_t w = 0; w < n_leafy; ++w) {
        const uint32_t b = tx->leafy[w];
        tr_slot[b] = w;
        tr_off[w + 1] = tr_off[w] + std::min(tx->lf_off[b + 1] - tx->lf_off[b], max_keep);
    }
    const uint32_t T = tr_off[n_leafy];
    DevBuf<uint32_t> d_tr_off(n_leafy + 1);
    h2d(ctx, d_tr_off.p, tr_off.data(), n_leafy + 1);
    sync(ctx);
    std::vector<std::unique_ptr<myrio_results>> done;
    for (uint64_t q0 = 0; q0 < q_count;) {
        const uint64_t qn = leaf_vecs ? score_rows_direct_device(ctx, leaf_vecs, queries, q_first + q0, q_count - q0, sim)
                                      : score_rows_device(ctx, ix, queries, q_first + q0, q_count - q0, sim);
        const float *rows = ctx->score_rows.p;
        DevBuf<float> bsum((size_t)qn * B), bmean((size_t)qn * B), bmax((size_t)qn * B);
        DevBuf<float> t_score(std::max<size_t>((size_t)qn * T, 1));
        DevBuf<uint32_t> t_pid(std::max<size_t>((size_t)qn * T, 1));
        for (size_t l = 0; l + 1 < tx->level_off.size(); ++l) {
            const uint32_t n_level = tx->level_off[l + 1] - tx->level_off[l];
            if (!n_level) continue;
            for (uint64_t qq = 0; qq < qn; qq += 65535) {  // gridDim.y limit
                const uint32_t qc = (uint32_t)std::min<uint64_t>(65535, qn - qq);
                MYRIO_LAUNCH(ctx, "k8a_branch_stats", branch_stats_kernel, dim3((n_level + 127) / 128, qc), 128, 0,
                             tx->d_level_order.p + tx->level_off[l], n_level, tx->d_ch_off.p, tx->d_ch.p, tx->d_nb.p,
                             rows + qq * L, L, B, bsum.p + qq * B, bmean.p + qq * B, bmax.p + qq * B);
            }
        }
        if (n_leafy && T)
            for (uint64_t qq = 0; qq < qn; qq += 65535) {
                const uint32_t qc = (uint32_t)std::min<uint64_t>(65535, qn - qq);
                MYRIO_LAUNCH(ctx, "k8b_branch_trim", branch_trim_kernel, dim3((n_leafy + 7) / 8, qc), 256, 0,
                             tx->d_leafy.p, n_leafy, tx->d_lf_off.p, tx->d_lf_pid.p, d_tr_off.p, T, max_keep,
                             rows + qq * L, L, t_score.p + qq * T, t_pid.p + qq * T);
            }
        std::vector<float> h_mean((size_t)qn * B), h_max((size_t)qn * B), h_ts((size_t)qn * T);
        std::vector<uint32_t> h_tp((size_t)qn * T);
        d2h(ctx, h_mean.data(), bmean.p, (size_t)qn * B);
        d2h(ctx, h_max.data(), bmax.p, (size_t)qn * B);
        d2h(ctx, h_ts.data(), t_score.p, (size_t)qn * T);
        d2h(ctx, h_tp.data(), t_pid.p, (size_t)qn * T);
        sync(ctx);
        for (uint64_t q = 0; q < qn; ++q) {
            ResTree trimmed;
            emit_trimmed(*tx, 0, tr_slot, h_mean.data() + q * B, h_max.data() + q * B, h_ts.data() + q * T,
                         h_tp.data() + q * T, tr_off.data(), max_keep, trimmed);
            auto r = std::make_unique<myrio_results>();
            r->root_rank = tx->root_rank;
            r->tree = tree_cut(trimmed, p.nb_best_analysis);
            if (r->tree.kind.empty()) fail(MYRIO_ERR_EMPTY, "the results tree is empty (nb_best_analysis = 0 or a tree without leaves)");
            pool_and_confidence(r->tree, tx->root_rank, p);
            done.push_back(std::move(r));
        }
        q0 += qn;
    }
    for (uint64_t q = 0; q < q_count; ++q) out[q] = done[q].release();
}

myrio_results *results_cut(const myrio_results *r, uint64_t nb_best) {
    auto o = std::make_unique<myrio_results>();
    o->root_rank = r->root_rank;
    o->tree = tree_cut(r->tree, nb_best);
    if (o->tree.kind.empty()) fail(MYRIO_ERR_EMPTY, "cut(%llu) leaves nothing", (unsigned long long)nb_best);
    return o.release();
}

// run.rs:405-466
float in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                             const float *pool_score, const uint8_t *conf_state, const float *conf) {
    constexpr uint64_t TAKE = 8;
    if (n_genes == 0) fail(MYRIO_ERR_EMPTY, "no gene");
    std::vector<std::vector<uint64_t>> ranked(n_genes);
    for (uint64_t g = 0; g < n_genes; ++g) {
        if (genus_off[g + 1] <= genus_off[g]) fail(MYRIO_ERR_EMPTY, "gene %llu has no genus branch (index [0] panics, run.rs:424)", (unsigned long long)g);
        for (uint64_t i = genus_off[g]; i < genus_off[g + 1]; ++i) {
            if (!std::isfinite(pool_score[i])) fail(MYRIO_ERR_NONFINITE, "non-finite pooling score");
            ranked[g].push_back(i);
        }
        std::stable_sort(ranked[g].begin(), ranked[g].end(), [&](uint64_t a, uint64_t b) { return pool_score[a] > pool_score[b]; });
    }
    float mismatch = 0.0f;
    std::vector<float> top_conf;
    for (uint64_t i = 0; i < n_genes; ++i) {
        const uint64_t top = ranked[i][0];
        if (conf_state[top] == 0) fail(MYRIO_ERR_BAD_ARG, "gene %llu: the top genus has no confidence (unwrap on None panics, run.rs:425)", (unsigned long long)i);
        const float c = conf_state[top] == 2 ? conf[top] : 0.5f;
        top_conf.push_back(c);
        for (uint64_t o = i + 1; o < n_genes; ++o) {
            uint64_t posn = TAKE;
            for (uint64_t t = 0; t < TAKE && t < ranked[o].size(); ++t)
                if (name_id[ranked[o][t]] == name_id[top]) {
                    posn = t;
                    break;
                }
            const float penalty = (float)posn / (float)TAKE;
            mismatch = mismatch + penalty * fmaxf(powf(c, 0.7f), 0.12f);
        }
    }
    mismatch = mismatch / (float)n_genes;
    float csum = 0.0f;
    for (float c : top_conf) csum = csum + c;
    const float lambda = 0.5f + 2.0f * csum;
    const float conf_score = 1.0f - softmax_pooling(top_conf, lambda);
    const float x = mismatch, y = powf(conf_score, 0.8f);
    const float unknown = 1.0f / (1.0f + expf(-8.0f * ((x + y) / 2.0f - 0.45f)));
    return 1.0f - unknown;
}

}  // namespace myrio
